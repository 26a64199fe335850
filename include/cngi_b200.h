/*
 * cngi_b200.h -- C ABI of libcngi_b200.so: B200 (sm_100a) convolutional gridding for ngcasa.
 *
 * Drop-in boundary for ONE hot path of casangi/cngi_prototype: the per-chunk gridding operators
 * that the reference's dask graphs call (paths relative to the reference tree):
 *
 *   ngcasa/imaging/_imaging_utils/_standard_grid.py:123   _standard_grid_numpy_wrap        -> cngi_b200_standard_grid
 *   ngcasa/imaging/_imaging_utils/_standard_grid.py:180   _standard_grid_psf_numpy_wrap    -> cngi_b200_standard_grid (do_psf)
 *                                                         (do_imaging_weight, support 1)   -> cngi_b200_imaging_weight_grid
 *   ngcasa/imaging/make_imaging_weight.py:198             calculate_briggs_parms           -> cngi_b200_briggs_factors
 *   ngcasa/imaging/_imaging_utils/_standard_grid.py:443   _standard_imaging_weight_degrid_numpy_wrap
 *                                                                                          -> cngi_b200_imaging_weight_degrid
 *   ngcasa/imaging/_imaging_utils/_aperture_grid.py:294   _aperture_grid_numpy_wrap        -> cngi_b200_aperture_grid
 *   ngcasa/imaging/_imaging_utils/_aperture_grid.py:333   _aperture_psf_grid_numpy_wrap    -> cngi_b200_aperture_grid (do_psf)
 *   ngcasa/imaging/_imaging_utils/_aperture_grid.py:146   _aperture_weight_grid_numpy_wrap -> cngi_b200_aperture_weight_grid
 *   ngcasa/imaging/predict_modelvis_image.py:20 (stub)    degrid predict                   -> cngi_b200_standard_degrid
 *   ngcasa/imaging/make_image.py:116-130                  ifft2 + crop + correct_image     -> cngi_b200_grid_to_image
 *   ngcasa/imaging/_imaging_utils/_normalize.py:39-89     normalize_image                  -> cngi_b200_grid_to_image (pb/sinc)
 *   ngcasa/imaging/direction_rotate.py:190-248            apply_rotation_matrix/apply_phasor -> cngi_b200_direction_rotate
 *   ngcasa/imaging/make_gridding_convolution_function.py:161-457  a_term GCF            -> cngi_b200_make_gcf, cngi_b200_phase_gradient
 *   cngi/vis/apply_flags.py:53                            apply_flags (where FLAG == 0)    -> cngi_b200_apply_flags
 *   cngi/dio/read_vis.py:186-197                          zarr chunk reads (host)          -> cngi_b200_zarr_read_chunks
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless its name ends in _host.  Arrays are C-order and
 *     contiguous.  Complex values are interleaved (re, im).
 *   - Sample arrays are (n_time, n_baseline, n_chan, n_pol); uvw is (n_time, n_baseline, 3) float64
 *     metres; grids are kernel-side (n_imag_chan, n_imag_pol, n_u, n_v), v fastest -- the layout the
 *     reference's jit functions use before the moveaxis at _standard_grid.py:101-104.
 *   - Gridding entry points ACCUMULATE into caller-owned grid / sum_weight buffers (the reference's
 *     wrappers allocate zeroed outputs; zero them yourself or with cudaMemsetAsync).
 *   - precision: CNGI_F32 = vis complex64, weight float32, grid complex64/float32;
 *                CNGI_F64 = vis complex128, weight float64, grid complex128/float64.
 *     Cell indices, oversampling offsets and masks are always computed in IEEE fp64 with the
 *     reference's operation order (no FMA contraction), so they are bit-exact in both precisions.
 *     sum_weight is always float64.
 *   - All functions are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream), re-entrant, and return 0 on success.  They never throw.  On failure the message is
 *     available from cngi_b200_last_error() (thread-local).
 *   - Bad samples are silently skipped exactly as the reference does (NaN u/v, stamp leaving the grid,
 *     NaN or zero weighted data, field < 0); that is not an error.
 */
#ifndef CNGI_B200_H
#define CNGI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNGI_B200_ABI_VERSION 1

enum { CNGI_OK = 0, CNGI_ERR_INVALID = 1, CNGI_ERR_CUDA = 2, CNGI_ERR_UNSUPPORTED = 3, CNGI_ERR_NO_DEVICE = 4 };
enum { CNGI_F32 = 0, CNGI_F64 = 1 };
/* chan_map shortcut: CUBE = identity map (a_chan = i_chan), CONTINUUM = all zero
   (_standard_grid.py:151-156); GENERAL reads the chan_map array. */
enum { CNGI_CHAN_GENERAL = 0, CNGI_CHAN_CUBE = 1, CNGI_CHAN_CONTINUUM = 2 };
/* kernel selection for the standard gridder */
enum { CNGI_ALGO_AUTO = 0, CNGI_ALGO_NAIVE = 1, CNGI_ALGO_TRACK = 2, CNGI_ALGO_SHIFT = 3, CNGI_ALGO_WINDOW = 4 };

int cngi_b200_abi_version(void);
/* Measurement aid, not on the product path: issues blocks*256*per_thread REDG.E.ADD.F32x2 reductions into
   buf[0 .. n_cells) (8-byte cells).  pattern 0 = every lane its own 32-byte sector, 1 = groups of 8 lanes on 64
   contiguous bytes, 2 = a warp on 256 contiguous bytes.  bench.py times it to obtain the reduction ("atomic")
   roofline the gridders' flush traffic is compared with (no published peak exists). */
int cngi_b200_microbench_red(void *buf, int64_t n_cells, int32_t pattern, int32_t blocks, int32_t per_thread, void *stream);
/* Measurement aid: the alternative design (per-block shared-memory subgrid + fp atomics) in isolation.  blocks x 256
   threads each add per_thread 7x7 complex stamps into a 32x32 shared-memory subgrid with atomicAdd(float) and flush it
   once; sink holds >= 1024 complex64.  tools/red_peak.py reports tap updates/s next to the product kernel's. */
int cngi_b200_microbench_smem_atomics(void *sink, int32_t blocks, int32_t per_thread, void *stream);
const char *cngi_b200_last_error(void);
/* 0 if a compute-capability 10.x device is current/available; CNGI_ERR_NO_DEVICE otherwise. */
int cngi_b200_check_device(void);

/* ------------------------------------------------------------------------------------------------
 * A1/A2  standard (prolate-spheroidal, separable taps) gridder.   _standard_grid.py:242-371
 * ---------------------------------------------------------------------------------------------- */
typedef struct cngi_std_grid_args {
    int64_t n_time, n_baseline, n_chan, n_pol;      /* sample array shape                           */
    int64_t n_imag_chan, n_imag_pol, n_u, n_v;      /* grid shape                                   */
    const void *vis;            /* complex [n_time,n_baseline,n_chan,n_pol]; ignored when do_psf      */
    const void *weight;         /* real, same shape                                                  */
    const uint8_t *flag;        /* optional (may be NULL), same shape; non-zero => vis treated as NaN */
    const double *uvw;          /* [n_time,n_baseline,3]                                             */
    const double *freq_chan;    /* [n_chan] Hz                                                       */
    const int64_t *chan_map;    /* [n_chan]; may be NULL unless chan_mode == CNGI_CHAN_GENERAL       */
    const int64_t *pol_map;     /* [n_pol];  NULL = identity                                         */
    const double *cgk_1D;       /* [oversampling*(support/2+1)] half tap table (always float64)      */
    void *grid;                 /* complex or real [n_imag_chan,n_imag_pol,n_u,n_v], accumulated      */
    double *sum_weight;         /* [n_imag_chan,n_imag_pol], accumulated                             */
    double delta_lm[2];         /* cell size in radians, x already negated (_check_imaging_parms:39)  */
    int32_t support, oversampling;
    int32_t precision;          /* CNGI_F32 / CNGI_F64                                               */
    int32_t do_psf;             /* grid weights only (real data)                                     */
    int32_t complex_grid;       /* grid cell type; image mode requires 1                             */
    int32_t chan_mode;          /* CNGI_CHAN_*                                                       */
    int32_t algorithm;          /* CNGI_ALGO_*                                                       */
    int32_t chan_group;         /* track kernel: channels walked per work item (0 = auto)            */
    int32_t time_segment;       /* track kernel: time steps per work item (0 = auto)                 */
    int32_t reserved;
} cngi_std_grid_args;

int cngi_b200_standard_grid(const cngi_std_grid_args *args, void *stream);
/* N1 (fused per-channel pipeline, synthesis_imaging_cube.py:195-211): ONE pass over uvw / weight / vis accumulates the
   complex image grid exactly as cngi_b200_standard_grid does in image mode (args->grid, args->sum_weight) AND the real
   psf grid + its sum_weight exactly as the do_psf mode does (psf_grid [n_imag_chan,n_imag_pol,n_u,n_v] real,
   psf_sum_weight [n_imag_chan,n_imag_pol]); the two share every cell index and tap.  args must describe the image pass
   (do_psf 0, complex_grid 1).  Support 7 only (make_image.py:106-107); CNGI_ERR_UNSUPPORTED otherwise -- call
   cngi_b200_standard_grid twice then. */
int cngi_b200_standard_grid_image_psf(const cngi_std_grid_args *args, void *psf_grid, double *psf_sum_weight, void *stream);

/* A4 folded into A1: make_imaging_weight's degrid (_standard_imaging_weight_degrid_jit, _standard_grid.py:466-518) and
   make_grid / make_image's gridding (_standard_grid_jit :242-371) in ONE pass over the samples.  args describes the image
   pass (do_psf 0, complex_grid 1, support 7) but args->weight holds the NATURAL weights (the `natural_imaging_weight` of
   :443); per sample the kernel forms  w_img = (n_pol == 2 ? (w0 + w1) / 2 : w) / (f0 * rho[cell] + f1)  with exactly the
   masks and arithmetic of cngi_b200_imaging_weight_degrid -- 0 off the density grid or for NaN uv, undivided where the
   natural weight or rho is 0 / NaN -- and grids vis * w_img.  The density gather is software-pipelined one round ahead
   (cp.async), so the pass costs about what the gridding alone costs, and the imaging weights are neither written nor
   re-read (imaging_weight may be NULL; give a buffer to receive them as cngi_b200_imaging_weight_degrid would write
   them).  The density grid has its own geometry (make_imaging_weight does not pad, make_imaging_weight.py:153). */
typedef struct cngi_iw_fused_args {
    const double *density;      /* float64 density grid (output of cngi_b200_imaging_weight_grid)     */
    int64_t density_stride[4];  /* element strides of (u, v, imaging chan, imaging pol), as in cngi_iw_degrid_args */
    const double *briggs_factors; /* [2, n_imag_chan, n_imag_pol] float64, contiguous                   */
    void *imaging_weight;       /* optional out: real [n_time,n_baseline,n_chan,n_pol]; NULL = not written */
    int64_t n_u, n_v;           /* density grid size                                                   */
    double delta_lm[2];         /* its cell size (radians, x negated)                                  */
    int32_t pol_shared;         /* the caller GUARANTEES that all pol planes of density and all pol columns of
                                   briggs_factors are identical (true for what cngi_b200_imaging_weight_grid and
                                   cngi_b200_briggs_factors produce when n_pol >= 2: the weights are pol-averaged,
                                   _standard_grid.py:328-330): one gather and one division per sample instead of one per pol */
    int32_t reserved;
} cngi_iw_fused_args;

int cngi_b200_standard_grid_weighted(const cngi_std_grid_args *args, const cngi_iw_fused_args *iw, void *stream);

/* ------------------------------------------------------------------------------------------------
 * A2  imaging-weight density grid: support 1, nearest cell + conjugate cell, pol-averaged weight when
 *     n_pol >= 2, sum_weight doubled.   _standard_grid.py:306-318,328-330,362-369 as called from
 *     make_imaging_weight.py:153-161.  density / sum_weight are float64 in both precisions; the
 *     conjugate cell IS bounds checked here (the reference does not).
 * ---------------------------------------------------------------------------------------------- */
typedef struct cngi_iw_grid_args {
    int64_t n_time, n_baseline, n_chan, n_pol;
    int64_t n_imag_chan, n_imag_pol, n_u, n_v;
    const void *weight;         /* real [n_time,n_baseline,n_chan,n_pol] (precision selects f32/f64)  */
    const double *uvw;
    const double *freq_chan;
    const int64_t *chan_map;
    const int64_t *pol_map;
    double *density;            /* float64 [n_imag_chan,n_imag_pol,n_u,n_v], accumulated              */
    double *sum_weight;         /* float64 [n_imag_chan,n_imag_pol], accumulated                      */
    double delta_lm[2];
    int32_t precision;
    int32_t chan_mode;
    int32_t first_pol_only;     /* n_pol >= 2 only: every pol plane receives the same (pol-averaged) weights, so update
                                   plane pol_map[0] and sum_weight[.., pol_map[0]] only; the caller replicates the plane
                                   (halves the reductions here and the bytes of a multi-GPU all-reduce)             */
    int32_t reserved;
} cngi_iw_grid_args;

int cngi_b200_imaging_weight_grid(const cngi_iw_grid_args *args, void *stream);

/* A3  briggs_factors[0] = (5*10^-robust)^2 / (sum(density^2)/sum_weight), [1] = 1  (weighting 0 = briggs)
 *     or [0] = 1, [1] = 0 (weighting 1 = uniform).   make_imaging_weight.py:198-213
 *     density [n_planes, n_u*n_v] float64, sum_weight [n_planes], briggs_factors [2, n_planes] float64. */
int cngi_b200_briggs_factors(const double *density, const double *sum_weight, double *briggs_factors,
                             int64_t n_planes, int64_t n_cells, double robust, int32_t weighting, void *stream);

/* A4  imaging_weight = (pol-averaged) natural weight / (f0*density[cell] + f1).  _standard_grid.py:466-518
 *     density is addressed through element strides so that both the kernel-side (chan,pol,u,v) and the
 *     API-side (u,v,chan,pol) layouts work without a transpose. */
typedef struct cngi_iw_degrid_args {
    int64_t n_time, n_baseline, n_chan, n_pol;
    int64_t n_imag_chan, n_imag_pol, n_u, n_v;
    const void *natural_weight; /* real [n_time,n_baseline,n_chan,n_pol]                              */
    const double *uvw;
    const double *freq_chan;
    const int64_t *chan_map;
    const int64_t *pol_map;
    const double *density;      /* float64                                                           */
    int64_t density_stride[4];  /* element strides for (u, v, chan, pol)                              */
    const double *briggs_factors; /* [2, n_imag_chan, n_imag_pol]                                     */
    void *imaging_weight;       /* real out [n_time,n_baseline,n_chan,n_pol], fully overwritten       */
    double delta_lm[2];
    int32_t precision;
    int32_t chan_mode;
    int32_t pol_shared;         /* n_pol == 2 only: the caller GUARANTEES identical density planes and Briggs factors for
                                   both pols (pol-averaged weights, _standard_grid.py:328-330): one gather and one
                                   division per sample, bit-identical results                                          */
    int32_t reserved;
} cngi_iw_degrid_args;

int cngi_b200_imaging_weight_degrid(const cngi_iw_degrid_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------
 * A5/A6  aperture (A-projection / mosaic) gridders.   _aperture_grid.py:376-513 and :180-291
 * ---------------------------------------------------------------------------------------------- */
typedef struct cngi_aperture_grid_args {
    int64_t n_time, n_baseline, n_chan, n_pol;
    int64_t n_imag_chan, n_imag_pol, n_u, n_v;
    const void *vis;                 /* complex; ignored when do_psf or grid_weights                 */
    const void *weight;              /* imaging weight, real                                        */
    const uint8_t *flag;             /* optional                                                    */
    const double *uvw;
    const double *freq_chan;
    const int64_t *chan_map, *pol_map;
    const int64_t *field;            /* [n_time,n_baseline] FIELD_ID column; <= -1 => row skipped     */
    const int64_t *field_id;         /* [n_field] ids in phase_gradient order                        */
    const int64_t *cf_baseline_map;  /* [n_baseline] */
    const int64_t *cf_chan_map;      /* [n_chan]     */
    const int64_t *cf_pol_map;       /* [n_pol]      */
    const double *conv_kernel;       /* float64 [n_cfb,n_cfc,n_cfp,n_cu,n_cv] (CONV_KERNEL or WEIGHT_CONV_KERNEL) */
    const int64_t *weight_support;   /* [n_cfb,n_cfc,n_cfp,2]                                        */
    const double *phase_gradient;    /* complex128 [n_field,n_cu,n_cv]                               */
    void *grid;                      /* complex [n_imag_chan,n_imag_pol,n_u,n_v], accumulated         */
    double *sum_weight;
    double delta_lm[2];
    int64_t n_field, n_cfb, n_cfc, n_cfp, n_cu, n_cv;
    int32_t oversampling[2];
    int32_t max_support;             /* max over weight_support (host knows it; bounds test :397,447) */
    int32_t precision;
    int32_t do_psf;
    int32_t chan_mode;
} cngi_aperture_grid_args;

int cngi_b200_aperture_grid(const cngi_aperture_grid_args *args, void *stream);
/* A6: stamps at the grid centre with unshifted CF indices; `vis`, `flag`, `do_psf` ignored. */
int cngi_b200_aperture_weight_grid(const cngi_aperture_grid_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------
 * A7  degridding predict = adjoint of A1 (no reference implementation; predict_modelvis_image.py:20-40
 *     is a stub).  vis[t,b,c,p] = sum_taps cgk*cgk*model_grid[chan_map[c], pol_map[p], ...]; samples the
 *     gridder would skip get 0.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cngi_std_degrid_args {
    int64_t n_time, n_baseline, n_chan, n_pol;
    int64_t n_imag_chan, n_imag_pol, n_u, n_v;
    const void *model_grid;     /* complex [n_imag_chan,n_imag_pol,n_u,n_v]                           */
    const double *uvw;
    const double *freq_chan;
    const int64_t *chan_map, *pol_map;
    const double *cgk_1D;
    void *vis;                  /* complex out [n_time,n_baseline,n_chan,n_pol], fully overwritten    */
    double delta_lm[2];
    int32_t support, oversampling;
    int32_t precision;
    int32_t chan_mode;
    int32_t normalize;          /* 1: divide each sample by its tap sum (sum_u cgk * sum_v cgk), the same
                                   normalisation the imaging side applies through sum_weight; 0: raw adjoint */
    int32_t algorithm;          /* 0 auto, 1 gather kernel (any support / pol count), 2 register-window kernel       */
} cngi_std_degrid_args;

int cngi_b200_standard_degrid(const cngi_std_degrid_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------
 * A9/A10  grid -> image: fftshift(ifft2(ifftshift(G))) (cuFFT, unnormalised inverse == numpy ifft2 * N),
 *         centre crop, real part, / sum_weight (0 -> 1), / correcting image [* sinc * pb], pb_limit mask.
 *         make_image.py:116-130, _remove_padding.py:20-31, _normalize.py:39-89
 * ---------------------------------------------------------------------------------------------- */
typedef struct cngi_fft_plan cngi_fft_plan; /* opaque: cuFFT plan + work buffer for n_planes of n_u x n_v */

int cngi_b200_fft_plan_create(cngi_fft_plan **plan, int64_t n_u, int64_t n_v, int64_t max_planes,
                              int32_t precision);
int cngi_b200_fft_plan_destroy(cngi_fft_plan *plan);

typedef struct cngi_grid_to_image_args {
    int64_t n_planes;           /* n_imag_chan * n_imag_pol                                          */
    int64_t n_u, n_v;           /* padded grid size                                                  */
    int64_t image_size[2];      /* cropped image size (l, m)                                         */
    const void *grid;           /* complex or real [n_planes,n_u,n_v]; not modified                   */
    int32_t grid_is_complex;
    int32_t precision;
    const double *sum_weight;   /* [n_planes] or NULL (no division)                                  */
    const double *corr_u;       /* [image_size[0]] separable correcting function along l, or NULL     */
    const double *corr_v;       /* [image_size[1]] along m; image is divided by corr_u[i]*corr_v[j]   */
    const void *norm_image;     /* optional real [n_planes or 1, l, m] extra divisor (PB / WEIGHT_PB)  */
    int64_t norm_image_planes;  /* 1 = broadcast over planes                                         */
    const void *pb_image;       /* optional real [n_planes or 1, l, m]: pixels with pb < pb_limit -> 0 */
    int64_t pb_image_planes;
    double pb_limit;
    int32_t divide_by_centre;   /* make_psf_with_gcf.py:140: divide plane by its centre pixel; 1 = pixel (l/2, m/2),
                                   2 = pixel centre_pixel[] (grid_parms['image_center']); a plane whose centre value is
                                   0 or not finite is left undivided (the reference would fill it with inf / NaN)       */
    int32_t single_precision_roundtrip; /* _normalize.py:86-87                                        */
    void *image;                /* real out [n_planes, l, m] (kernel-side plane order)                */
    int64_t centre_pixel[2];    /* used when divide_by_centre == 2                                      */
} cngi_grid_to_image_args;

int cngi_b200_grid_to_image(cngi_fft_plan *plan, const cngi_grid_to_image_args *args, void *stream);

/* image -> uv-grid, the inverse of cngi_b200_grid_to_image's transform, for the degridding predict:
   G = fftshift(fft2(ifftshift(pad(image / (corr_u x corr_v))))) (the inverse of make_image.py:116-130; the reference's
   predict_modelvis_image.py:20-40 is a stub that lists "fourier_transform, _degrid").  Unnormalised forward DFT. */
typedef struct cngi_image_to_grid_args {
    int64_t n_planes;           /* n_imag_chan * n_imag_pol                                          */
    int64_t n_u, n_v;           /* padded grid size                                                  */
    int64_t image_size[2];      /* image size (l, m) <= (n_u, n_v); centred in the padded plane       */
    const void *image;          /* real [n_planes, l, m]                                             */
    const double *corr_u;       /* [image_size[0]] or NULL: image is divided by corr_u[l]*corr_v[m]   */
    const double *corr_v;
    int32_t precision;
    int32_t reserved;
    void *grid;                 /* complex out [n_planes, n_u, n_v]                                   */
} cngi_image_to_grid_args;

int cngi_b200_image_to_grid(cngi_fft_plan *plan, const cngi_image_to_grid_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------
 * N3  direction_rotate: rotate uvw to a new phase centre and phase-rotate the visibilities (the step in
 *     front of the mosaic gridders).  ngcasa/imaging/direction_rotate.py:190-213 (apply_rotation_matrix),
 *     :217-248 (apply_phasor).  The per-field matrices come from calc_rotation_mats (:127-175), a host-side
 *     n_field x 3 x 3 computation (cngi_prototype_b200/direction_rotate.py).
 *       uvw_rot[t,b,:] = uvw[t,b,:] @ uvw_rotmat[field(t)]
 *       vis_rot        = vis * exp(i ((2 pi d) f) (1/c)),  d = uvw_rot[0:end] . phase_rotation[field(t), 0:end]
 *     field(t) = index in rot_field_id of the one FIELD_ID value of integration t (ids > -1 for the rotation,
 *     != INT_NAN for the phasor, as the reference); the reference asserts it is constant over baseline --
 *     here *status is set to 1 instead (outputs of that integration: uvw_rot 0, vis_rot NaN).
 * ---------------------------------------------------------------------------------------------- */
typedef struct cngi_direction_rotate_args {
    int64_t n_time, n_baseline, n_chan, n_pol;
    const void *vis;            /* complex [n_time,n_baseline,n_chan,n_pol], or NULL: rotate uvw only       */
    void *vis_rot;              /* complex out, same shape; may alias vis                                  */
    const double *uvw;          /* [n_time,n_baseline,3]                                                   */
    double *uvw_rot;            /* out, may alias uvw, may be NULL                                         */
    const int64_t *field;       /* FIELD_ID [n_time,n_baseline]                                            */
    const double *freq_chan;    /* [n_chan] Hz                                                             */
    const double *uvw_rotmat;   /* [n_field,3,3]                                                           */
    const double *phase_rotation; /* [n_field,3]                                                           */
    const int64_t *rot_field_id;  /* [n_field]                                                             */
    int64_t n_field;
    int32_t *status;            /* optional device int, set to 1 on a field lookup failure (never cleared)  */
    int32_t common_tangent_reprojection; /* 1: phase uses u,v only (:219-222)                              */
    int32_t single_precision;   /* CNGI_F64 only: round the result through complex64 (:244-245)            */
    int32_t precision;          /* CNGI_F32: complex64 in/out (arithmetic is fp64 either way)              */
} cngi_direction_rotate_args;

int cngi_b200_direction_rotate(const cngi_direction_rotate_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------
 * N2  A-term gridding convolution functions (the producer of conv_kernel / weight_conv_kernel / weight_support /
 *     phase_gradient that cngi_b200_aperture_grid consumes).  a_term branch of
 *     ngcasa/imaging/make_gridding_convolution_function.py:161-311: Airy voltage patterns per antenna type
 *     (_make_pb_symmetric.py:135-235) -> per antenna-type pair products (:394-412) -> fft (:246-247) -> support
 *     search (:414-457) -> crop + normalise (:361-392).  Outputs are float64 like the reference's.
 *     status (device int, OR-ed, never cleared): 1 = min >= cut level (:423), 2 = walk left the image (:429,:440),
 *     4 = support >= max_support (:447-448) -- the reference's asserts.
 * ---------------------------------------------------------------------------------------------- */
enum { CNGI_PB_AIRY = 0, CNGI_PB_CASA_AIRY = 1 };
typedef struct cngi_gcf_args {
    int64_t n_pad[2];               /* grid_parms['image_size_padded']                                  */
    int64_t conv_size[2];           /* resize_conv_size = (max_support + 1) * oversampling  (:134)       */
    double pb_cell[2];              /* cell_size * oversampling, radians (:398)                          */
    int32_t oversampling[2];
    int32_t max_support[2];
    int32_t function;               /* CNGI_PB_AIRY / CNGI_PB_CASA_AIRY                                  */
    int32_t reserved;
    int64_t n_dish;                 /* unique antenna types                                             */
    const double *dish_diameter_host, *blockage_diameter_host;   /* HOST [n_dish], metres                */
    int64_t n_pair;
    const int64_t *ant_pairs_host;  /* HOST [n_pair,2] antenna-type pairs (create_cf_baseline_map :512)  */
    int64_t n_freq;
    const double *pb_freq_host;     /* HOST [n_freq] PB frequencies (create_cf_chan_map :536)            */
    double support_cut_level;
    double *conv_kernel;            /* out [n_pair,n_freq,1,conv_size[0],conv_size[1]]                  */
    double *weight_conv_kernel;     /* out, same shape                                                  */
    int64_t *support;               /* out [n_pair,n_freq,1,2]                                          */
    int32_t *status;                /* device int, see above                                            */
} cngi_gcf_args;

int cngi_b200_make_gcf(const cngi_gcf_args *args, void *stream);
/* make_phase_gradient :351-358: out[f,i,j] = exp(i ((i - cu/2) pix[f,0] + (j - cv/2) pix[f,1])), complex128.
   pix [n_field,2] (device) = -(SIN world2pix offset) * 2 pi / (n_pad * oversampling), computed by the host mirror. */
int cngi_b200_phase_gradient(const double *pix, int64_t n_field, int64_t cu, int64_t cv, void *phase_gradient,
                             void *stream);

/* make_pb: primary-beam images of each dish type, pb[l, m, chan, pol, dish] = (Airy voltage pattern)^ipower
   (ngcasa/imaging/make_pb.py:95-118 with _airy_disk / _casa_airy_disk, _imaging_utils/_make_pb_symmetric.py:26-132;
   make_pb uses ipower 2, synthesis_imaging_cube.py:277 too).  cell_size in radians as given (x negative). */
typedef struct cngi_pb_args {
    int64_t image_size[2];
    int64_t image_center[2];
    double cell_size[2];
    int32_t function;               /* CNGI_PB_AIRY / CNGI_PB_CASA_AIRY                                  */
    int32_t ipower;                 /* 1 voltage pattern, 2 primary beam                                */
    int64_t n_chan;
    const double *freq_chan_host;   /* HOST [n_chan] Hz                                                  */
    int64_t n_pol;                  /* the pattern is replicated over pol (np.tile, :66,:130)            */
    int64_t n_dish;                 /* <= 8                                                             */
    const double *dish_diameter_host, *blockage_diameter_host;   /* HOST [n_dish]                        */
    double *pb;                     /* out float64 [l, m, chan, pol, dish]                              */
} cngi_pb_args;

int cngi_b200_make_pb(const cngi_pb_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------
 * N4  apply_flags (cngi/vis/apply_flags.py:53; the in-place form is synthesis_imaging_cube.py:180):
 *       out[i] = flag[i] ? NaN : data[i]
 *     for one data variable with FLAG's dims, both flattened to n_elem.  NaN is what xarray's where() + astype
 *     produce (fill value of xarray core/dtypes.py maybe_promote): the quiet NaN for reals, (NaN, NaN) for complex
 *     -- bit-identical to numpy.where(flag == 0, data, fill).astype(dtype).
 *     out may be data itself (in place: only the flag bytes are read and only flagged elements are written) or a
 *     disjoint buffer.  n_flagged (optional device counter, added to, never cleared) receives the number of flagged
 *     elements.  Integer variables are refused (the reference's NaN -> int cast is undefined behaviour).
 * ---------------------------------------------------------------------------------------------- */
enum { CNGI_ELEM_F32 = 0, CNGI_ELEM_F64 = 1, CNGI_ELEM_C64 = 2, CNGI_ELEM_C128 = 3 };
int cngi_b200_apply_flags(const void *data, void *out, const uint8_t *flag, int64_t n_elem, int32_t elem_kind,
                          uint64_t *n_flagged, void *stream);

/* N4, host side (no device work): decode zarr v2 chunk files into a box of a C-order HOST array (normally a pinned
 * staging buffer of the chunk stream) on n_threads native threads -- the per-chunk reads xarray.open_zarr + dask issue in
 * the reference (cngi/dio/read_vis.py:186-197; chunks written by cngi/dio/append_xds.py:69 with blosc/zstd).
 * Each job names one chunk file (stored at the FULL chunk shape; NULL or a missing file = fill element) and the box to
 * copy: decoded_chunk[src_start : src_start + extent] -> dst[dst_start : dst_start + extent].  compressor: RAW, ZLIB, or
 * BLOSC (c-blosc 1.x frames; zstd / lz4 / zlib codecs, byte shuffle, split and unsplit blocks; bit shuffle, blosclz and
 * snappy are refused).  Synchronous; all pointers are HOST pointers. */
enum { CNGI_ZARR_RAW = 0, CNGI_ZARR_ZLIB = 1, CNGI_ZARR_BLOSC = 2 };
typedef struct cngi_zarr_chunk_job {
    const char *path;
    int64_t chunk_shape[8];
    int64_t src_start[8], dst_start[8], extent[8];
} cngi_zarr_chunk_job;
int cngi_b200_zarr_read_chunks(const cngi_zarr_chunk_job *jobs, int64_t n_jobs, void *dst_host, const int64_t *dst_shape,
                               int32_t ndim, int32_t elem_bytes, int32_t compressor, const void *fill_elem,
                               int32_t n_threads);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer entry point: what a ctypes / cgo-style binding calls with numpy-like HOST arrays.
 * Streams the sample arrays through pinned staging buffers in time chunks (H2D overlapped with the
 * gridding kernels on a second stream), accumulates on the device, and copies grid + sum_weight back.
 * All pointers in `args` are HOST pointers here; grid/sum_weight are overwritten (allocate-and-return
 * semantics of _standard_grid_numpy_wrap).  Synchronous.
 * ---------------------------------------------------------------------------------------------- */
int cngi_b200_standard_grid_host(const cngi_std_grid_args *args_host, int64_t time_chunk);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU sums of uv-grids through an NVSwitch multicast mapping (NVLS), replacing the reference's
 * tree sum over chunks (_standard_grid.py:109-120) for partial grids that live on different GPUs.
 * `multicast_ptr` is the multicast address of a buffer that every rank allocated symmetrically (e.g.
 * torch.distributed._symmetric_memory); the caller orders the ranks (barrier before: all partial grids
 * complete; barrier after: nobody overwrites a buffer that is still being read).
 *   reduce:     dst[i] = sum over ranks of buffer[i]      (run by the root only; 16-byte aligned, n % 4 == 0)
 *   allreduce:  buffer[i] = sum over ranks, on every rank (each rank handles its 1/world_size slice)
 * ---------------------------------------------------------------------------------------------- */
int cngi_b200_multimem_reduce_f32(const void *multicast_ptr, void *dst, int64_t n_floats, int32_t n_blocks, void *stream);
int cngi_b200_multimem_allreduce_f64(void *multicast_ptr, int64_t n_doubles, int32_t rank, int32_t world_size,
                                     int32_t n_blocks, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CNGI_B200_H */
