"""Seeded synthetic visibility sets shaped like BASELINE.json's configs.

Pure numpy, host side.  The same arrays feed the CUDA path, the parity oracle and the CPU
baseline.  Geometry: antennas ~ N(0, sigma) on a plane (z squashed), earth-rotation synthesis
tracks at a fixed declination; the cell size is chosen so that the longest projected baseline
at the highest frequency lands at 1/1.15 of the grid half width, i.e. every sample is well
inside the grid (the reference's conjugate-cell write is not bounds checked,
_standard_grid.py:317-318,363-364, so inputs must stay clear of the edge).
"""
import numpy as np

C_LIGHT = 299792458.0
EARTH_RATE = 7.2921150e-5  # rad / s


def antenna_layout(n_ant, sigma_m, rng):
    xyz = rng.normal(0.0, sigma_m, size=(n_ant, 3))
    xyz[:, 2] *= 0.05
    return xyz


def baseline_vectors(xyz):
    i, j = np.triu_indices(len(xyz), k=1)
    return xyz[j] - xyz[i]


def uvw_tracks(bl_xyz, hour_angle, dec):
    """(n_time, n_baseline, 3) metres from equatorial baseline vectors."""
    H = np.asarray(hour_angle)[:, None]
    lx, ly, lz = bl_xyz[None, :, 0], bl_xyz[None, :, 1], bl_xyz[None, :, 2]
    sh, ch, sd, cd = np.sin(H), np.cos(H), np.sin(dec), np.cos(dec)
    u = sh * lx + ch * ly
    v = -sd * ch * lx + sd * sh * ly + cd * lz
    w = cd * ch * lx - cd * sh * ly + sd * lz
    return np.ascontiguousarray(np.stack([u, v, w], axis=-1))


def pick_cell_size(uvw, freq, margin=1.15):
    """Cell (radians) such that max |u|,|v| (wavelengths) sits at N/(2*margin) cells."""
    uv_max_m = np.nanmax(np.abs(uvw[..., :2]))
    uv_max_lambda = uv_max_m * np.max(freq) / C_LIGHT
    return 1.0 / (2.0 * uv_max_lambda * margin)


def make_vis_set(n_ant, n_time, n_chan, n_pol, freq_lo, freq_hi, sigma_m, integration_s, dec_deg=34.0,
                 seed=1234, flag_frac=0.02, complex_data=True, bad_rows=True, dtype="f64", layout_seed=None,
                 time_offset=None):
    """Returns dict(vis, uvw, weight, freq_chan, cell, n_baseline).

    vis ~ CN(0,1) with flag_frac of samples NaN (apply_flags semantics, cngi/vis/apply_flags.py:53),
    weight ~ U(0.5,1.5) with a few exact zeros and NaNs, a few NaN uvw rows (bad_rows).
    layout_seed / time_offset: shards of ONE observation (same array, consecutive time ranges, different
    noise) for the multi-GPU runs -- then the cell size comes from the baseline lengths, not from the tracks,
    so that every shard uses the same grid.
    """
    rng = np.random.default_rng(seed)
    sharded = layout_seed is not None or time_offset is not None
    xyz = antenna_layout(n_ant, sigma_m, np.random.default_rng(layout_seed) if layout_seed is not None else rng)
    bl = baseline_vectors(xyz)
    n_bl = len(bl)
    span = EARTH_RATE * integration_s * n_time
    ha = np.linspace(-span / 2, span / 2, n_time)
    if time_offset is not None:
        ha = EARTH_RATE * integration_s * (time_offset + np.arange(n_time))
    uvw = uvw_tracks(bl, ha, np.deg2rad(dec_deg))
    freq = np.linspace(freq_lo, freq_hi, n_chan)
    if sharded:
        cell = 1.0 / (2.0 * np.max(np.linalg.norm(bl, axis=1)) * np.max(freq) / C_LIGHT * 1.15)
    else:
        cell = pick_cell_size(uvw, freq)
    shape = (n_time, n_bl, n_chan, n_pol)
    fdt = np.float32 if dtype == "f32" else np.float64
    weight = rng.uniform(0.5, 1.5, size=shape).astype(fdt)
    vis = None
    if complex_data:
        cdt = np.complex64 if dtype == "f32" else np.complex128
        vis = np.empty(shape, dtype=cdt)
        vis.real = rng.standard_normal(size=shape, dtype=fdt)
        vis.imag = rng.standard_normal(size=shape, dtype=fdt)
        if flag_frac > 0:
            n_flag = int(flag_frac * vis.size)
            idx = rng.integers(0, vis.size, size=n_flag)
            vis.reshape(-1)[idx] = np.nan
    if bad_rows:
        n_bad = max(1, weight.size // 5000)
        weight.reshape(-1)[rng.integers(0, weight.size, size=n_bad)] = 0.0
        weight.reshape(-1)[rng.integers(0, weight.size, size=n_bad)] = np.nan
        n_bad_uvw = max(1, (n_time * n_bl) // 2000)
        rows = rng.integers(0, n_time * n_bl, size=n_bad_uvw)
        uvw.reshape(-1, 3)[rows, rng.integers(0, 2, size=n_bad_uvw)] = np.nan
    return dict(vis=vis, uvw=uvw, weight=weight, freq_chan=freq, cell=cell, n_baseline=n_bl)


def grid_parms_for(n_uv, cell, chan_mode="cube", support=7, oversampling=100, do_psf=False,
                   complex_grid=True, do_imaging_weight=False):
    """grid_parms as the reference's wrappers read them (_standard_grid.py:151-172): cell_size in
    radians with the x axis negated (_check_imaging_parms.py:38-39), image_size_padded an int array."""
    return dict(chan_mode=chan_mode, image_size_padded=np.array([n_uv, n_uv], dtype=np.int64),
                image_size=np.array([n_uv, n_uv], dtype=np.int64),
                cell_size=np.array([-cell, cell], dtype=np.float64), oversampling=oversampling,
                support=support, do_psf=do_psf, complex_grid=complex_grid,
                do_imaging_weight=do_imaging_weight)


# ---- BASELINE.json configs ------------------------------------------------------------------
def config_c1(n_time=1000, n_chan=64, seed=1234, dtype="f64"):
    """VLA-like: 27 antennas (351 baselines) x n_time x 64 chan x 2 pol, 1.0-1.128 GHz, HA in [-1,1] rad."""
    integration = 2.0 / EARTH_RATE / 1000.0  # 1000 steps span 2 rad
    return make_vis_set(27, n_time, n_chan, 2, 1.0e9, 1.128e9, 350.0, integration, seed=seed, dtype=dtype)


def config_c2(n_time=500, n_chan=128, seed=4321, dtype="f32", shard=None):
    """ALMA-like: 43 antennas (903 baselines) x 500 x 6 s x 128 chan x 2 pol, 345-347 GHz.
    shard=r gives the r-th consecutive 500-integration block of one long observation (multi-GPU weak scaling)."""
    if shard is None:
        return make_vis_set(43, n_time, n_chan, 2, 345.0e9, 347.0e9, 300.0, 6.0, dec_deg=-23.0, seed=seed,
                            dtype=dtype)
    return make_vis_set(43, n_time, n_chan, 2, 345.0e9, 347.0e9, 300.0, 6.0, dec_deg=-23.0, seed=seed + 1000 * shard,
                        dtype=dtype, layout_seed=seed, time_offset=shard * n_time - 2000)


def config_c4(n_time=1000, n_chan=64, seed=99, dtype="f64"):
    """27-antenna set for the degrid predict (no flags, no data)."""
    integration = 2.0 / EARTH_RATE / 1000.0
    return make_vis_set(27, n_time, n_chan, 2, 1.0e9, 1.128e9, 350.0, integration, seed=seed,
                        complex_data=False, bad_rows=False, dtype=dtype)


def make_mosaic_gcf(n_bl, n_chan, n_pol, n_field=7, n_cf_baseline=3, n_cf_chan=2, n_cf_pol=1,
                    oversampling=(10, 10), max_support=(15, 15), seed=7):
    """Synthetic gcf_dataset contents for the aperture gridders (A5/A6).

    Shapes follow make_gridding_convolution_function.py:161-311: CONV_KERNEL / WEIGHT_CONV_KERNEL
    (n_cfb, n_cfc, n_cfp, (max_support+1)*oversampling, ...) real, SUPPORT (n_cfb,n_cfc,n_cfp,2) odd ints,
    PHASE_GRADIENT (n_field, cu, cv) unit-modulus complex, CF_*_MAP index arrays, field_id.  The kernels are
    smooth bumps (Gaussian x cosine ripple) rather than FFTs of Airy patterns: the gridders only index them.
    """
    rng = np.random.default_rng(seed)
    os_u, os_v = oversampling
    cu = (max_support[0] + 1) * os_u
    cv = (max_support[1] + 1) * os_v
    x = (np.arange(cu) - cu // 2) / os_u
    y = (np.arange(cv) - cv // 2) / os_v
    conv = np.zeros((n_cf_baseline, n_cf_chan, n_cf_pol, cu, cv))
    wconv = np.zeros_like(conv)
    support = np.zeros((n_cf_baseline, n_cf_chan, n_cf_pol, 2), dtype=np.int64)
    for b in range(n_cf_baseline):
        for c in range(n_cf_chan):
            for p in range(n_cf_pol):
                width = 1.6 + 0.7 * b + 0.3 * c
                gx = np.exp(-0.5 * (x / width) ** 2) * (1 + 0.1 * np.cos(1.3 * x))
                gy = np.exp(-0.5 * (y / width) ** 2) * (1 + 0.1 * np.cos(0.9 * y))
                k = np.outer(gx, gy)
                conv[b, c, p] = k / k.sum() * (os_u * os_v)
                k2 = np.outer(gx ** 2, gy ** 2)
                wconv[b, c, p] = k2 / k2.sum() * (os_u * os_v)
                s = min(max_support[0], 2 * int(np.ceil(2.2 * width)) + 1)
                support[b, c, p] = (s, s)
    field_id = np.arange(n_field, dtype=np.int64) + 3  # ids need not start at 0
    px = rng.uniform(-0.3, 0.3, size=n_field)
    py = rng.uniform(-0.3, 0.3, size=n_field)
    X, Y = np.meshgrid(np.arange(cu) - cu // 2, np.arange(cv) - cv // 2, indexing="ij")
    pg = np.exp(1j * (X[None] * px[:, None, None] + Y[None] * py[:, None, None]))
    return dict(conv_kernel=conv, weight_conv_kernel=wconv, weight_support=support, phase_gradient=pg,
                cf_baseline_map=rng.integers(0, n_cf_baseline, size=n_bl).astype(np.int64),
                cf_chan_map=(np.arange(n_chan) * n_cf_chan // max(n_chan, 1)).astype(np.int64),
                cf_pol_map=np.zeros(n_pol, dtype=np.int64) if n_cf_pol == 1 else
                (np.arange(n_pol) % n_cf_pol).astype(np.int64),
                field_id=field_id, oversampling=np.array(oversampling, dtype=np.int64))


def mosaic_field_column(n_time, n_bl, field_id, seed=11, frac_unset=0.01):
    """FIELD_ID (n_time, n_baseline): cycles through the pointings per time step, constant over baseline
    (direction_rotate.py:199,228 assert that), with a few rows set to -1 (skipped, _aperture_grid.py:422)."""
    rng = np.random.default_rng(seed)
    f = np.repeat(field_id[np.arange(n_time) % len(field_id)][:, None], n_bl, axis=1).astype(np.int64)
    n_unset = int(frac_unset * f.size)
    if n_unset:
        f.reshape(-1)[rng.integers(0, f.size, size=n_unset)] = -1
    return f
