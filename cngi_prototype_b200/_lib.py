"""ctypes binding of libcngi_b200.so (include/cngi_b200.h).

There is NO CPU fallback: if the library is missing or no sm_100 device is present, calls raise.
"""
import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CNGI_B200_LIB") or os.path.join(HERE, "csrc", "libcngi_b200.so")   # env: tuning builds only

F32, F64 = 0, 1
CHAN_GENERAL, CHAN_CUBE, CHAN_CONTINUUM = 0, 1, 2
ALGO_AUTO, ALGO_NAIVE, ALGO_TRACK, ALGO_SHIFT, ALGO_WINDOW = 0, 1, 2, 3, 4
ELEM_F32, ELEM_F64, ELEM_C64, ELEM_C128 = 0, 1, 2, 3

i64, i32, f64, vp = C.c_int64, C.c_int32, C.c_double, C.c_void_p


class StdGridArgs(C.Structure):
    _fields_ = [
        ("n_time", i64), ("n_baseline", i64), ("n_chan", i64), ("n_pol", i64),
        ("n_imag_chan", i64), ("n_imag_pol", i64), ("n_u", i64), ("n_v", i64),
        ("vis", vp), ("weight", vp), ("flag", vp), ("uvw", vp), ("freq_chan", vp),
        ("chan_map", vp), ("pol_map", vp), ("cgk_1D", vp), ("grid", vp), ("sum_weight", vp),
        ("delta_lm", f64 * 2),
        ("support", i32), ("oversampling", i32), ("precision", i32), ("do_psf", i32),
        ("complex_grid", i32), ("chan_mode", i32), ("algorithm", i32), ("chan_group", i32),
        ("time_segment", i32), ("reserved", i32),
    ]


class IwFusedArgs(C.Structure):
    _fields_ = [
        ("density", vp), ("density_stride", i64 * 4), ("briggs_factors", vp), ("imaging_weight", vp),
        ("n_u", i64), ("n_v", i64), ("delta_lm", f64 * 2), ("pol_shared", i32), ("reserved", i32),
    ]


class IwGridArgs(C.Structure):
    _fields_ = [
        ("n_time", i64), ("n_baseline", i64), ("n_chan", i64), ("n_pol", i64),
        ("n_imag_chan", i64), ("n_imag_pol", i64), ("n_u", i64), ("n_v", i64),
        ("weight", vp), ("uvw", vp), ("freq_chan", vp), ("chan_map", vp), ("pol_map", vp),
        ("density", vp), ("sum_weight", vp),
        ("delta_lm", f64 * 2),
        ("precision", i32), ("chan_mode", i32), ("first_pol_only", i32), ("reserved", i32),
    ]


class IwDegridArgs(C.Structure):
    _fields_ = [
        ("n_time", i64), ("n_baseline", i64), ("n_chan", i64), ("n_pol", i64),
        ("n_imag_chan", i64), ("n_imag_pol", i64), ("n_u", i64), ("n_v", i64),
        ("natural_weight", vp), ("uvw", vp), ("freq_chan", vp), ("chan_map", vp), ("pol_map", vp),
        ("density", vp), ("density_stride", i64 * 4), ("briggs_factors", vp), ("imaging_weight", vp),
        ("delta_lm", f64 * 2),
        ("precision", i32), ("chan_mode", i32), ("pol_shared", i32), ("reserved", i32),
    ]


class ApertureGridArgs(C.Structure):
    _fields_ = [
        ("n_time", i64), ("n_baseline", i64), ("n_chan", i64), ("n_pol", i64),
        ("n_imag_chan", i64), ("n_imag_pol", i64), ("n_u", i64), ("n_v", i64),
        ("vis", vp), ("weight", vp), ("flag", vp), ("uvw", vp), ("freq_chan", vp),
        ("chan_map", vp), ("pol_map", vp), ("field", vp), ("field_id", vp),
        ("cf_baseline_map", vp), ("cf_chan_map", vp), ("cf_pol_map", vp),
        ("conv_kernel", vp), ("weight_support", vp), ("phase_gradient", vp),
        ("grid", vp), ("sum_weight", vp),
        ("delta_lm", f64 * 2),
        ("n_field", i64), ("n_cfb", i64), ("n_cfc", i64), ("n_cfp", i64), ("n_cu", i64), ("n_cv", i64),
        ("oversampling", i32 * 2), ("max_support", i32), ("precision", i32), ("do_psf", i32),
        ("chan_mode", i32),
    ]


class StdDegridArgs(C.Structure):
    _fields_ = [
        ("n_time", i64), ("n_baseline", i64), ("n_chan", i64), ("n_pol", i64),
        ("n_imag_chan", i64), ("n_imag_pol", i64), ("n_u", i64), ("n_v", i64),
        ("model_grid", vp), ("uvw", vp), ("freq_chan", vp), ("chan_map", vp), ("pol_map", vp),
        ("cgk_1D", vp), ("vis", vp),
        ("delta_lm", f64 * 2),
        ("support", i32), ("oversampling", i32), ("precision", i32), ("chan_mode", i32),
        ("normalize", i32), ("algorithm", i32),
    ]


class GridToImageArgs(C.Structure):
    _fields_ = [
        ("n_planes", i64), ("n_u", i64), ("n_v", i64), ("image_size", i64 * 2),
        ("grid", vp), ("grid_is_complex", i32), ("precision", i32),
        ("sum_weight", vp), ("corr_u", vp), ("corr_v", vp),
        ("norm_image", vp), ("norm_image_planes", i64),
        ("pb_image", vp), ("pb_image_planes", i64),
        ("pb_limit", f64),
        ("divide_by_centre", i32), ("single_precision_roundtrip", i32),
        ("image", vp), ("centre_pixel", i64 * 2),
    ]


class ImageToGridArgs(C.Structure):
    _fields_ = [
        ("n_planes", i64), ("n_u", i64), ("n_v", i64), ("image_size", i64 * 2),
        ("image", vp), ("corr_u", vp), ("corr_v", vp), ("precision", i32), ("reserved", i32), ("grid", vp),
    ]


class DirectionRotateArgs(C.Structure):
    _fields_ = [
        ("n_time", i64), ("n_baseline", i64), ("n_chan", i64), ("n_pol", i64),
        ("vis", vp), ("vis_rot", vp), ("uvw", vp), ("uvw_rot", vp), ("field", vp), ("freq_chan", vp),
        ("uvw_rotmat", vp), ("phase_rotation", vp), ("rot_field_id", vp),
        ("n_field", i64), ("status", vp),
        ("common_tangent_reprojection", i32), ("single_precision", i32), ("precision", i32),
    ]


class GcfArgs(C.Structure):
    _fields_ = [
        ("n_pad", i64 * 2), ("conv_size", i64 * 2), ("pb_cell", f64 * 2),
        ("oversampling", i32 * 2), ("max_support", i32 * 2), ("function", i32), ("reserved", i32),
        ("n_dish", i64), ("dish_diameter_host", vp), ("blockage_diameter_host", vp),
        ("n_pair", i64), ("ant_pairs_host", vp), ("n_freq", i64), ("pb_freq_host", vp),
        ("support_cut_level", f64),
        ("conv_kernel", vp), ("weight_conv_kernel", vp), ("support", vp), ("status", vp),
    ]


class PbArgs(C.Structure):
    _fields_ = [
        ("image_size", i64 * 2), ("image_center", i64 * 2), ("cell_size", f64 * 2), ("function", i32), ("ipower", i32),
        ("n_chan", i64), ("freq_chan_host", vp), ("n_pol", i64), ("n_dish", i64),
        ("dish_diameter_host", vp), ("blockage_diameter_host", vp), ("pb", vp),
    ]


class ZarrChunkJob(C.Structure):
    _fields_ = [("path", C.c_char_p), ("chunk_shape", i64 * 8), ("src_start", i64 * 8), ("dst_start", i64 * 8),
                ("extent", i64 * 8)]


ZARR_RAW, ZARR_ZLIB, ZARR_BLOSC = 0, 1, 2

# every symbol include/cngi_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "cngi_b200_abi_version", "cngi_b200_last_error", "cngi_b200_check_device",
    "cngi_b200_standard_grid", "cngi_b200_imaging_weight_grid", "cngi_b200_briggs_factors",
    "cngi_b200_imaging_weight_degrid", "cngi_b200_aperture_grid", "cngi_b200_aperture_weight_grid",
    "cngi_b200_standard_degrid", "cngi_b200_fft_plan_create", "cngi_b200_fft_plan_destroy",
    "cngi_b200_grid_to_image", "cngi_b200_standard_grid_host", "cngi_b200_microbench_red",
    "cngi_b200_microbench_smem_atomics", "cngi_b200_direction_rotate", "cngi_b200_make_gcf",
    "cngi_b200_phase_gradient", "cngi_b200_image_to_grid", "cngi_b200_standard_grid_image_psf", "cngi_b200_make_pb",
    "cngi_b200_apply_flags", "cngi_b200_zarr_read_chunks", "cngi_b200_standard_grid_weighted",
    "cngi_b200_multimem_reduce_f32", "cngi_b200_multimem_allreduce_f64",
]

_lib = None
_lib_lock = threading.Lock()


class CngiError(RuntimeError):
    pass


def lib():
    """Loads libcngi_b200.so.  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise CngiError("%s not found: run `python -m cngi_prototype_b200.build` (needs nvcc). "
                            "cngi_prototype_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.cngi_b200_last_error.restype = C.c_char_p
        L.cngi_b200_abi_version.restype = C.c_int
        for name in EXPORTS:
            getattr(L, name)  # AttributeError here means header and library disagree
        L.cngi_b200_briggs_factors.argtypes = [vp, vp, vp, i64, i64, f64, i32, vp]
        L.cngi_b200_fft_plan_create.argtypes = [C.POINTER(vp), i64, i64, i64, i32]
        L.cngi_b200_fft_plan_destroy.argtypes = [vp]
        L.cngi_b200_grid_to_image.argtypes = [vp, C.POINTER(GridToImageArgs), vp]
        L.cngi_b200_standard_grid.argtypes = [C.POINTER(StdGridArgs), vp]
        L.cngi_b200_standard_grid_image_psf.argtypes = [C.POINTER(StdGridArgs), vp, vp, vp]
        L.cngi_b200_standard_grid_weighted.argtypes = [C.POINTER(StdGridArgs), C.POINTER(IwFusedArgs), vp]
        L.cngi_b200_standard_grid_host.argtypes = [C.POINTER(StdGridArgs), i64]
        L.cngi_b200_imaging_weight_grid.argtypes = [C.POINTER(IwGridArgs), vp]
        L.cngi_b200_imaging_weight_degrid.argtypes = [C.POINTER(IwDegridArgs), vp]
        L.cngi_b200_aperture_grid.argtypes = [C.POINTER(ApertureGridArgs), vp]
        L.cngi_b200_aperture_weight_grid.argtypes = [C.POINTER(ApertureGridArgs), vp]
        L.cngi_b200_standard_degrid.argtypes = [C.POINTER(StdDegridArgs), vp]
        L.cngi_b200_direction_rotate.argtypes = [C.POINTER(DirectionRotateArgs), vp]
        L.cngi_b200_image_to_grid.argtypes = [vp, C.POINTER(ImageToGridArgs), vp]
        L.cngi_b200_make_pb.argtypes = [C.POINTER(PbArgs), vp]
        L.cngi_b200_make_gcf.argtypes = [C.POINTER(GcfArgs), vp]
        L.cngi_b200_phase_gradient.argtypes = [vp, i64, i64, i64, vp, vp]
        L.cngi_b200_zarr_read_chunks.argtypes = [C.POINTER(ZarrChunkJob), i64, vp, C.POINTER(i64), i32, i32, i32, vp, i32]
        L.cngi_b200_apply_flags.argtypes = [vp, vp, vp, i64, i32, vp, vp]
        L.cngi_b200_multimem_reduce_f32.argtypes = [vp, vp, i64, i32, vp]
        L.cngi_b200_multimem_allreduce_f64.argtypes = [vp, i64, i32, i32, i32, vp]
        L.cngi_b200_microbench_red.argtypes = [vp, i64, i32, i32, i32, vp]
        L.cngi_b200_microbench_smem_atomics.argtypes = [vp, i32, i32, vp]
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().cngi_b200_last_error()
        raise CngiError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else "?"))


_device_ok = False


def require_device():
    """Raises unless an sm_100 device is usable (checked once per process; there is no fallback)."""
    global _device_ok
    if not _device_ok:
        check(lib().cngi_b200_check_device(), "cngi_b200_check_device")
        _device_ok = True
