/*
 * cngi_oracle.c -- CPU restatement of the ngcasa convolutional-gridding hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * cngi_prototype_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path never does.
 *
 * Every function restates one numba loop of the reference (paths relative to
 * /root/reference/ngcasa/imaging/_imaging_utils/), in plain C, fp64 throughout,
 * with the same operation order, so that (compiled with -ffp-contract=off) its
 * output is bit-identical to the reference's numba output.  That claim is pinned
 * by tests/golden/ (vectors generated from the reference itself by
 * tests/golden/make_golden.py) and by oracle/check_against_reference.py.
 *
 * Arrays are C-order.  Complex arrays are interleaved (re, im) doubles.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include <sys/mman.h>

#define C_LIGHT 299792458.0

typedef int64_t i64;

/* numba's int(x) on a float64: truncation toward zero (C cast does the same). */
static inline i64 trunc_to_int(double x) { return (i64)x; }

/* uv_scale[k][c] = -(freq[c] * delta_lm[k] * n_uv[k]) / c   (_standard_grid.py:273-276) */
static void fill_uv_scale(double *us, double *vs, const double *freq, i64 n_chan,
                          const double *delta_lm, i64 n_u, i64 n_v)
{
    for (i64 c = 0; c < n_chan; ++c) {
        us[c] = -(freq[c] * delta_lm[0] * (double)n_u) / C_LIGHT;
        vs[c] = -(freq[c] * delta_lm[1] * (double)n_v) / C_LIGHT;
    }
}

/*
 * A1 / A2 : _standard_grid_jit  (_standard_grid.py:242-371)
 *   grid        (n_imag_chan, n_imag_pol, n_u, n_v)  complex128 if complex_grid else float64
 *   sum_weight  (n_imag_chan, n_imag_pol)
 *   vis         (n_time, n_baseline, n_chan, n_pol) complex128 (ignored when do_psf)
 *   uvw         (n_time, n_baseline, 3)
 *   weight      (n_time, n_baseline, n_chan, n_pol)
 * t0/t1 and c0/c1 restrict the loops to a (time, chan) window of the arrays --
 * that is how the multi-threaded driver below mirrors the reference's dask
 * chunking (_standard_grid.py:58-92) without copying the inputs.
 */
void oracle_standard_grid_window(double *grid, double *sum_weight, int do_psf, int do_imaging_weight,
                                 int complex_grid, const double *vis, const double *uvw,
                                 const double *freq_chan, const i64 *chan_map, const i64 *pol_map,
                                 const double *weight, const double *cgk_1D, i64 n_time, i64 n_baseline,
                                 i64 n_chan, i64 n_pol, i64 n_imag_pol, i64 n_u, i64 n_v,
                                 const double *delta_lm, i64 support, i64 oversampling, i64 t0, i64 t1,
                                 i64 c0, i64 c1)
{
    double *us = (double *)malloc(sizeof(double) * (size_t)n_chan);
    double *vs = (double *)malloc(sizeof(double) * (size_t)n_chan);
    fill_uv_scale(us, vs, freq_chan, n_chan, delta_lm, n_u, n_v);

    const i64 support_center = support / 2;
    const i64 u_mid = n_u / 2, v_mid = n_v / 2;
    const i64 start_support = -support_center;
    const i64 end_support = support - support_center;
    (void)n_time;

    for (i64 it = t0; it < t1; ++it)
        for (i64 ib = 0; ib < n_baseline; ++ib) {
            const double *p_uvw = uvw + (it * n_baseline + ib) * 3;
            for (i64 ic = c0; ic < c1; ++ic) {
                const i64 a_chan = chan_map[ic];
                const double u = p_uvw[0] * us[ic];
                const double v = p_uvw[1] * vs[ic];
                if (isnan(u) || isnan(v)) continue;                      /* :302 */
                const double u_pos = u + (double)u_mid;
                const double v_pos = v + (double)v_mid;
                const double u_pos_conj = -u + (double)u_mid;
                const double v_pos_conj = -v + (double)v_mid;
                const i64 uc = trunc_to_int(u_pos + 0.5);                /* :314-315 */
                const i64 vc = trunc_to_int(v_pos + 0.5);
                const i64 uc_conj = trunc_to_int(u_pos_conj + 0.5);
                const i64 vc_conj = trunc_to_int(v_pos_conj + 0.5);
                if (!((uc + support_center < n_u) && (vc + support_center < n_v) &&
                      (uc - support_center >= 0) && (vc - support_center >= 0)))
                    continue;                                            /* :320 */
                const double u_offset = (double)uc - u_pos;
                const i64 u_off_idx = (i64)floor(u_offset * (double)oversampling + 0.5);
                const double v_offset = (double)vc - v_pos;
                const i64 v_off_idx = (i64)floor(v_offset * (double)oversampling + 0.5);

                const i64 sample = ((it * n_baseline + ib) * n_chan + ic) * n_pol;
                for (i64 ip = 0; ip < n_pol; ++ip) {
                    double sel_weight, wd_re, wd_im = 0.0;
                    if (do_psf) {
                        if (n_pol >= 2 && do_imaging_weight) {           /* :328-330 */
                            wd_re = (weight[sample + 0] + weight[sample + 1]) / 2.0;
                            sel_weight = wd_re;
                        } else {
                            sel_weight = weight[sample + ip];
                            wd_re = sel_weight;
                        }
                    } else {
                        /* complex128 * float64: numba promotes the weight to (w + 0j)   :336 */
                        sel_weight = weight[sample + ip];
                        const double a = vis[2 * (sample + ip)], b = vis[2 * (sample + ip) + 1];
                        wd_re = a * sel_weight - b * 0.0;
                        wd_im = a * 0.0 + b * sel_weight;
                    }
                    if (isnan(wd_re) || isnan(wd_im)) continue;          /* :340 */
                    if (wd_re == 0.0 && wd_im == 0.0) continue;
                    const i64 a_pol = pol_map[ip];
                    const i64 plane = (a_chan * n_imag_pol + a_pol) * n_u;
                    double norm = 0.0;
                    for (i64 iv = start_support; iv < end_support; ++iv) {
                        const i64 v_indx = vc + iv;
                        const double conv_v = cgk_1D[llabs(oversampling * iv + v_off_idx)];
                        const i64 v_indx_conj = vc_conj + iv;
                        for (i64 iu = start_support; iu < end_support; ++iu) {
                            const i64 u_indx = uc + iu;
                            const double conv_u = cgk_1D[llabs(oversampling * iu + u_off_idx)];
                            const double conv = conv_u * conv_v;
                            const i64 cell = (plane + u_indx) * n_v + v_indx;
                            if (complex_grid) {
                                /* float64 * complex128 -> (conv + 0j) * wd              :357 */
                                grid[2 * cell] += conv * wd_re - 0.0 * wd_im;
                                grid[2 * cell + 1] += conv * wd_im + 0.0 * wd_re;
                            } else {
                                grid[cell] += conv * wd_re;
                            }
                            norm = norm + conv;
                            if (do_imaging_weight) {                     /* :362-364 */
                                const i64 u_indx_conj = uc_conj + iu;
                                /* The reference does not bounds-check the conjugate cell
                                   (numba wraps negatives / corrupts memory); the oracle
                                   refuses to write out of range so it stays well defined. */
                                if (u_indx_conj >= 0 && u_indx_conj < n_u && v_indx_conj >= 0 &&
                                    v_indx_conj < n_v) {
                                    const i64 cc = (plane + u_indx_conj) * n_v + v_indx_conj;
                                    if (complex_grid) {
                                        grid[2 * cc] += conv * wd_re - 0.0 * wd_im;
                                        grid[2 * cc + 1] += conv * wd_im + 0.0 * wd_re;
                                    } else {
                                        grid[cc] += conv * wd_re;
                                    }
                                }
                            }
                        }
                    }
                    double *sw = sum_weight + a_chan * n_imag_pol + a_pol;
                    *sw = *sw + sel_weight * norm;                        /* :366 */
                    if (do_imaging_weight) *sw = *sw + sel_weight * norm; /* :368-369 */
                }
            }
        }
    free(us);
    free(vs);
}

void oracle_standard_grid(double *grid, double *sum_weight, int do_psf, int do_imaging_weight,
                          int complex_grid, const double *vis, const double *uvw,
                          const double *freq_chan, const i64 *chan_map, const i64 *pol_map,
                          const double *weight, const double *cgk_1D, i64 n_time, i64 n_baseline,
                          i64 n_chan, i64 n_pol, i64 n_imag_pol, i64 n_u, i64 n_v,
                          const double *delta_lm, i64 support, i64 oversampling)
{
    oracle_standard_grid_window(grid, sum_weight, do_psf, do_imaging_weight, complex_grid, vis, uvw,
                                freq_chan, chan_map, pol_map, weight, cgk_1D, n_time, n_baseline, n_chan,
                                n_pol, n_imag_pol, n_u, n_v, delta_lm, support, oversampling, 0, n_time, 0,
                                n_chan);
}

/*
 * Multi-threaded driver used as the CPU baseline.  It mirrors how the reference
 * parallelises (_standard_grid.py:58-98): cube mode = one task per channel chunk,
 * each writing its own image channels; continuum = one task per time chunk with a
 * PRIVATE full grid, combined afterwards by a pairwise sum (_tree_sum_list :109-120).
 * cube_mode != 0 requires chan_map to be the identity.  Plain pthreads (one thread
 * per task) so the build needs nothing beyond libc.
 */
typedef struct {
    double *grid, *sum_weight;
    int do_psf, do_imaging_weight, complex_grid;
    const double *vis, *uvw, *freq_chan;
    const i64 *chan_map, *pol_map;
    const double *weight, *cgk_1D;
    i64 n_time, n_baseline, n_chan, n_pol, n_imag_pol, n_u, n_v;
    const double *delta_lm;
    i64 support, oversampling, t0, t1, c0, c1;
    size_t n_priv_chan; /* image channels of a private grid (continuum tasks) */
} grid_task;

static void *grid_task_run(void *p)
{
    grid_task *k = (grid_task *)p;
    oracle_standard_grid_window(k->grid, k->sum_weight, k->do_psf, k->do_imaging_weight, k->complex_grid,
                                k->vis, k->uvw, k->freq_chan, k->chan_map, k->pol_map, k->weight, k->cgk_1D,
                                k->n_time, k->n_baseline, k->n_chan, k->n_pol, k->n_imag_pol, k->n_u, k->n_v,
                                k->delta_lm, k->support, k->oversampling, k->t0, k->t1, k->c0, k->c1);
    return NULL;
}

/* Private grids of the continuum tasks: 2 MiB-aligned and advised as huge pages, so that
 * first touch costs one fault per 2 MiB instead of per 4 KiB; each worker zero-fills its own
 * grid (parallel, first touch on the worker's node) before gridding into it. */
#define ORACLE_HUGE ((size_t)2 << 20)
static double *private_grid_alloc(size_t n_doubles)
{
    size_t bytes = (n_doubles * sizeof(double) + ORACLE_HUGE - 1) / ORACLE_HUGE * ORACLE_HUGE;
    void *p = NULL;
    if (posix_memalign(&p, ORACLE_HUGE, bytes) != 0) return NULL;
#ifdef MADV_HUGEPAGE
    madvise(p, bytes, MADV_HUGEPAGE);
#endif
    return (double *)p;
}

static void *grid_task_zero_run(void *p)
{
    grid_task *k = (grid_task *)p;
    size_t cells = (size_t)k->n_u * (size_t)k->n_v * (k->complex_grid ? 2 : 1);
    /* continuum: one image channel (chan_map is all 0 there), n_imag_pol planes */
    memset(k->grid, 0, cells * (size_t)k->n_imag_pol * k->n_priv_chan * sizeof(double));
    return grid_task_run(p);
}

/* Sum of the private grids into the caller's grid, one slice of the cells per thread.  Inside a
 * slice the private grids are combined by the same pairwise order as _tree_sum_list (:109-120),
 * block by block so that the tree runs out of cache: the result is the tree sum bit for bit. */
typedef struct {
    double **src;      /* n_src private grids (modified) */
    i64 n_src;
    double *dst;
    size_t lo, hi;
} sum_task;

static void *sum_task_run(void *p)
{
    sum_task *k = (sum_task *)p;
    const size_t B = 2048;
    for (size_t b0 = k->lo; b0 < k->hi; b0 += B) {
        size_t b1 = b0 + B < k->hi ? b0 + B : k->hi;
        for (i64 stride = 1; stride < k->n_src; stride *= 2)
            for (i64 j = 0; j + stride < k->n_src; j += 2 * stride) {
                double *a = k->src[j];
                const double *b = k->src[j + stride];
                for (size_t i = b0; i < b1; ++i) a[i] += b[i];
            }
        const double *a0 = k->src[0];
        for (size_t i = b0; i < b1; ++i) k->dst[i] += a0[i];
    }
    return NULL;
}

int oracle_standard_grid_mt(double *grid, double *sum_weight, int do_psf, int do_imaging_weight,
                            int complex_grid, const double *vis, const double *uvw,
                            const double *freq_chan, const i64 *chan_map, const i64 *pol_map,
                            const double *weight, const double *cgk_1D, i64 n_time, i64 n_baseline,
                            i64 n_chan, i64 n_pol, i64 n_imag_chan, i64 n_imag_pol, i64 n_u, i64 n_v,
                            const double *delta_lm, i64 support, i64 oversampling, int cube_mode,
                            int n_threads)
{
    if (n_threads <= 1) {
        oracle_standard_grid(grid, sum_weight, do_psf, do_imaging_weight, complex_grid, vis, uvw, freq_chan,
                             chan_map, pol_map, weight, cgk_1D, n_time, n_baseline, n_chan, n_pol,
                             n_imag_pol, n_u, n_v, delta_lm, support, oversampling);
        return 1;
    }
    i64 n_tasks = cube_mode ? (n_chan < n_threads ? n_chan : n_threads)
                            : (n_time < n_threads ? n_time : n_threads);
    if (n_tasks < 1) return 0;
    grid_task *tasks = (grid_task *)calloc((size_t)n_tasks, sizeof(grid_task));
    pthread_t *th = (pthread_t *)calloc((size_t)n_tasks, sizeof(pthread_t));
    size_t grid_doubles = (size_t)n_imag_chan * n_imag_pol * n_u * n_v * (complex_grid ? 2 : 1);
    size_t sw_doubles = (size_t)n_imag_chan * n_imag_pol;
    int ok = 1;
    for (i64 k = 0; k < n_tasks; ++k) {
        grid_task t = {grid, sum_weight, do_psf, do_imaging_weight, complex_grid, vis, uvw, freq_chan,
                       chan_map, pol_map, weight, cgk_1D, n_time, n_baseline, n_chan, n_pol, n_imag_pol,
                       n_u, n_v, delta_lm, support, oversampling, 0, n_time, 0, n_chan, (size_t)n_imag_chan};
        if (cube_mode) {
            t.c0 = (n_chan * k) / n_tasks;
            t.c1 = (n_chan * (k + 1)) / n_tasks;
        } else {
            t.t0 = (n_time * k) / n_tasks;
            t.t1 = (n_time * (k + 1)) / n_tasks;
            t.grid = private_grid_alloc(grid_doubles); /* zero-filled by its worker */
            t.sum_weight = (double *)calloc(sw_doubles, sizeof(double));
            if (!t.grid || !t.sum_weight) ok = 0;
        }
        tasks[k] = t;
    }
    if (ok) {
        void *(*run)(void *) = cube_mode ? grid_task_run : grid_task_zero_run;
        for (i64 k = 0; k < n_tasks; ++k) pthread_create(&th[k], NULL, run, &tasks[k]);
        for (i64 k = 0; k < n_tasks; ++k) pthread_join(th[k], NULL);
        if (!cube_mode) {
            double **src = (double **)calloc((size_t)n_tasks, sizeof(double *));
            sum_task *sums = (sum_task *)calloc((size_t)n_tasks, sizeof(sum_task));
            for (i64 k = 0; k < n_tasks; ++k) src[k] = tasks[k].grid;
            for (i64 k = 0; k < n_tasks; ++k) {
                sum_task s = {src, n_tasks, grid, grid_doubles * (size_t)k / (size_t)n_tasks,
                              grid_doubles * (size_t)(k + 1) / (size_t)n_tasks};
                sums[k] = s;
                pthread_create(&th[k], NULL, sum_task_run, &sums[k]);
            }
            for (i64 k = 0; k < n_tasks; ++k) pthread_join(th[k], NULL);
            free(sums);
            free(src);
            /* sum_weight: the same pairwise order */
            for (i64 stride = 1; stride < n_tasks; stride *= 2)
                for (i64 k = 0; k + stride < n_tasks; k += 2 * stride)
                    for (size_t i = 0; i < sw_doubles; ++i)
                        tasks[k].sum_weight[i] += tasks[k + stride].sum_weight[i];
            for (size_t i = 0; i < sw_doubles; ++i) sum_weight[i] += tasks[0].sum_weight[i];
        }
    }
    if (!cube_mode)
        for (i64 k = 0; k < n_tasks; ++k) {
            free(tasks[k].grid);
            free(tasks[k].sum_weight);
        }
    free(tasks);
    free(th);
    return ok ? (int)n_tasks : -1;
}

/*
 * A4 : _standard_imaging_weight_degrid_jit  (_standard_grid.py:466-518)
 *   imaging_weight (out)      (n_time, n_baseline, n_chan, n_pol), caller zero-fills (:460)
 *   grid_imaging_weight       (n_u, n_v, n_imag_chan, n_imag_pol)   <- API-side layout (:514)
 *   briggs_factors            (2, n_imag_chan, n_imag_pol)
 */
void oracle_imaging_weight_degrid(double *imaging_weight, const double *grid_imaging_weight,
                                  const double *briggs_factors, const double *uvw, const double *freq_chan,
                                  const i64 *chan_map, const i64 *pol_map, const double *natural, i64 n_time,
                                  i64 n_baseline, i64 n_chan, i64 n_pol, i64 n_imag_chan, i64 n_imag_pol,
                                  i64 n_u, i64 n_v, const double *delta_lm)
{
    double *us = (double *)malloc(sizeof(double) * (size_t)n_chan);
    double *vs = (double *)malloc(sizeof(double) * (size_t)n_chan);
    fill_uv_scale(us, vs, freq_chan, n_chan, delta_lm, n_u, n_v);
    const i64 u_mid = n_u / 2, v_mid = n_v / 2;
    for (i64 it = 0; it < n_time; ++it)
        for (i64 ib = 0; ib < n_baseline; ++ib) {
            const double *p_uvw = uvw + (it * n_baseline + ib) * 3;
            for (i64 ic = 0; ic < n_chan; ++ic) {
                const i64 a_chan = chan_map[ic];
                const double u = p_uvw[0] * us[ic];
                const double v = p_uvw[1] * vs[ic];
                if (isnan(u) || isnan(v)) continue;
                const double u_pos = u + (double)u_mid;
                const double v_pos = v + (double)v_mid;
                const i64 uc = trunc_to_int(u_pos + 0.5);
                const i64 vc = trunc_to_int(v_pos + 0.5);
                if (!((uc < n_u) && (vc < n_v) && (uc >= 0) && (vc >= 0))) continue;     /* :502 */
                const i64 sample = ((it * n_baseline + ib) * n_chan + ic) * n_pol;
                for (i64 ip = 0; ip < n_pol; ++ip) {
                    const i64 a_pol = pol_map[ip];
                    double iw;
                    if (n_pol == 2)                                                         /* :508 */
                        iw = (natural[sample + 0] + natural[sample + 1]) / 2.0;
                    else
                        iw = natural[sample + ip];
                    const double nat = natural[sample + ip];
                    if (!isnan(nat) && nat != 0.0) {
                        const double rho =
                            grid_imaging_weight[((uc * n_v + vc) * n_imag_chan + a_chan) * n_imag_pol + a_pol];
                        if (!isnan(rho) && rho != 0.0) {
                            const double f0 = briggs_factors[a_chan * n_imag_pol + a_pol];
                            const double f1 = briggs_factors[(n_imag_chan + a_chan) * n_imag_pol + a_pol];
                            const double d = f0 * rho + f1;                                 /* :515 */
                            iw = iw / d;
                        }
                    }
                    imaging_weight[sample + ip] = iw;
                }
            }
        }
    free(us);
    free(vs);
}

/* A4 over time chunks on n_threads pthreads (the reference maps the chunk function over the dask
 * chunks, _standard_grid.py:417-437): samples are independent, each task writes its own rows. */
typedef struct {
    double *imaging_weight;
    const double *grid_imaging_weight, *briggs_factors, *uvw, *freq_chan;
    const i64 *chan_map, *pol_map;
    const double *natural;
    i64 n_time, n_baseline, n_chan, n_pol, n_imag_chan, n_imag_pol, n_u, n_v;
    const double *delta_lm;
} iw_degrid_task;

static void *iw_degrid_task_run(void *p)
{
    iw_degrid_task *k = (iw_degrid_task *)p;
    oracle_imaging_weight_degrid(k->imaging_weight, k->grid_imaging_weight, k->briggs_factors, k->uvw,
                                 k->freq_chan, k->chan_map, k->pol_map, k->natural, k->n_time, k->n_baseline,
                                 k->n_chan, k->n_pol, k->n_imag_chan, k->n_imag_pol, k->n_u, k->n_v, k->delta_lm);
    return NULL;
}

int oracle_imaging_weight_degrid_mt(double *imaging_weight, const double *grid_imaging_weight,
                                    const double *briggs_factors, const double *uvw, const double *freq_chan,
                                    const i64 *chan_map, const i64 *pol_map, const double *natural, i64 n_time,
                                    i64 n_baseline, i64 n_chan, i64 n_pol, i64 n_imag_chan, i64 n_imag_pol,
                                    i64 n_u, i64 n_v, const double *delta_lm, int n_threads)
{
    i64 n_tasks = n_time < n_threads ? n_time : n_threads;
    if (n_tasks <= 1) {
        oracle_imaging_weight_degrid(imaging_weight, grid_imaging_weight, briggs_factors, uvw, freq_chan,
                                     chan_map, pol_map, natural, n_time, n_baseline, n_chan, n_pol,
                                     n_imag_chan, n_imag_pol, n_u, n_v, delta_lm);
        return 1;
    }
    iw_degrid_task *tasks = (iw_degrid_task *)calloc((size_t)n_tasks, sizeof(iw_degrid_task));
    pthread_t *th = (pthread_t *)calloc((size_t)n_tasks, sizeof(pthread_t));
    const i64 row = n_baseline * n_chan * n_pol;
    for (i64 k = 0; k < n_tasks; ++k) {
        const i64 t0 = (n_time * k) / n_tasks, t1 = (n_time * (k + 1)) / n_tasks;
        iw_degrid_task t = {imaging_weight + t0 * row, grid_imaging_weight, briggs_factors,
                            uvw + t0 * n_baseline * 3, freq_chan, chan_map, pol_map, natural + t0 * row,
                            t1 - t0, n_baseline, n_chan, n_pol, n_imag_chan, n_imag_pol, n_u, n_v, delta_lm};
        tasks[k] = t;
        pthread_create(&th[k], NULL, iw_degrid_task_run, &tasks[k]);
    }
    for (i64 k = 0; k < n_tasks; ++k) pthread_join(th[k], NULL);
    free(tasks);
    free(th);
    return (int)n_tasks;
}

/* first index i with field_id[i] == f, or -1 (np.where(...)[0][0], _aperture_grid.py:423) */
static i64 find_field(const i64 *field_id, i64 n_field, i64 f)
{
    for (i64 i = 0; i < n_field; ++i)
        if (field_id[i] == f) return i;
    return -1;
}

static i64 max_support(const i64 *weight_support, i64 n)
{
    i64 m = weight_support[0];
    for (i64 i = 1; i < n; ++i)
        if (weight_support[i] > m) m = weight_support[i];
    return m;
}

/*
 * A5 : _aperture_grid_jit  (_aperture_grid.py:376-513)
 *   grid            (n_imag_chan, n_imag_pol, n_u, n_v) complex128
 *   conv_kernel     (n_cfb, n_cfc, n_cfp, n_cu, n_cv) float64
 *   weight_support  (n_cfb, n_cfc, n_cfp, 2) int64
 *   phase_gradient  (n_field, n_cu, n_cv) complex128
 *   field           (n_time, n_baseline) int64 ; field_id (n_field) int64
 * The reference materialises conv_kernel*phase_gradient[field] for the whole stack
 * (:428-430); element-wise that is (k + 0j)*(pr + i pi) = (k*pr - 0*pi) + i(k*pi + 0*pr),
 * evaluated here on the fly for the one element that is read.
 */
void oracle_aperture_grid(double *grid, double *sum_weight, int do_psf, const double *vis, const double *uvw,
                          const double *freq_chan, const i64 *chan_map, const i64 *pol_map,
                          const i64 *cf_baseline_map, const i64 *cf_chan_map, const i64 *cf_pol_map,
                          const double *imaging_weight, const double *conv_kernel, const i64 *weight_support,
                          const double *phase_gradient, const i64 *field, const i64 *field_id, i64 n_field,
                          i64 n_time, i64 n_baseline, i64 n_chan, i64 n_pol, i64 n_imag_pol, i64 n_u,
                          i64 n_v, const double *delta_lm, const i64 *oversampling, i64 n_cfb, i64 n_cfc,
                          i64 n_cfp, i64 n_cu, i64 n_cv)
{
    double *us = (double *)malloc(sizeof(double) * (size_t)n_chan);
    double *vs = (double *)malloc(sizeof(double) * (size_t)n_chan);
    fill_uv_scale(us, vs, freq_chan, n_chan, delta_lm, n_u, n_v);
    const i64 u_mid = n_u / 2, v_mid = n_v / 2;
    const i64 msc = max_support(weight_support, n_cfb * n_cfc * n_cfp * 2);   /* :397 */
    const i64 conv_u_center = n_cu / 2, conv_v_center = n_cv / 2;

    for (i64 it = 0; it < n_time; ++it)
        for (i64 ib = 0; ib < n_baseline; ++ib) {
            const i64 f = field[it * n_baseline + ib];
            if (!(f > -1)) continue;                                          /* :422 */
            const i64 field_indx = find_field(field_id, n_field, f);
            if (field_indx < 0) continue; /* reference would raise IndexError */
            const double *pg = phase_gradient + 2 * field_indx * n_cu * n_cv;
            const i64 cf_b = cf_baseline_map[ib];
            const double *p_uvw = uvw + (it * n_baseline + ib) * 3;
            for (i64 ic = 0; ic < n_chan; ++ic) {
                const i64 cf_c = cf_chan_map[ic];
                const i64 a_chan = chan_map[ic];
                const double u = p_uvw[0] * us[ic];
                const double v = p_uvw[1] * vs[ic];
                if (isnan(u) || isnan(v)) continue;
                const double u_pos = u + (double)u_mid;
                const double v_pos = v + (double)v_mid;
                const i64 uc = trunc_to_int(u_pos + 0.5);
                const i64 vc = trunc_to_int(v_pos + 0.5);
                if (!((uc + msc < n_u) && (vc + msc < n_v) && (uc - msc >= 0) && (vc - msc >= 0)))
                    continue;                                                 /* :447 */
                const double u_offset = (double)uc - u_pos;
                const i64 u_off = (i64)floor(u_offset * (double)oversampling[0] + 0.5) + conv_u_center;
                const double v_offset = (double)vc - v_pos;
                const i64 v_off = (i64)floor(v_offset * (double)oversampling[1] + 0.5) + conv_v_center;
                const i64 sample = ((it * n_baseline + ib) * n_chan + ic) * n_pol;
                for (i64 ip = 0; ip < n_pol; ++ip) {
                    const double w = imaging_weight[sample + ip];
                    double wd_re, wd_im;
                    if (do_psf) {
                        wd_re = w;
                        wd_im = 0.0;
                    } else {
                        const double a = vis[2 * (sample + ip)], b = vis[2 * (sample + ip) + 1];
                        wd_re = a * w - b * 0.0;
                        wd_im = a * 0.0 + b * w;
                    }
                    if (isnan(wd_re) || isnan(wd_im)) continue;
                    if (wd_re == 0.0 && wd_im == 0.0) continue;
                    const i64 cf_p = cf_pol_map[ip];
                    const i64 a_pol = pol_map[ip];
                    const i64 cf = (cf_b * n_cfc + cf_c) * n_cfp + cf_p;
                    const i64 su = weight_support[cf * 2 + 0], sv = weight_support[cf * 2 + 1];
                    const i64 su_c = su / 2, sv_c = sv / 2;
                    const double *ck = conv_kernel + cf * n_cu * n_cv;
                    const i64 plane = (a_chan * n_imag_pol + a_pol) * n_u;
                    double norm_re = 0.0, norm_im = 0.0;
                    for (i64 iv = -sv_c; iv < sv - sv_c; ++iv) {
                        const i64 v_indx = vc + iv;
                        const i64 cf_v = oversampling[1] * iv + v_off;
                        for (i64 iu = -su_c; iu < su - su_c; ++iu) {
                            const i64 u_indx = uc + iu;
                            const i64 cf_u = oversampling[0] * iu + u_off;
                            const double k = ck[cf_u * n_cv + cf_v];
                            const double pr = pg[2 * (cf_u * n_cv + cf_v)];
                            const double pi = pg[2 * (cf_u * n_cv + cf_v) + 1];
                            const double cr = k * pr - 0.0 * pi;
                            const double ci = k * pi + 0.0 * pr;
                            const i64 cell = (plane + u_indx) * n_v + v_indx;
                            if (do_psf) { /* complex * float64 -> complex * (w + 0j) */
                                grid[2 * cell] += cr * wd_re - ci * 0.0;
                                grid[2 * cell + 1] += cr * 0.0 + ci * wd_re;
                            } else {
                                grid[2 * cell] += cr * wd_re - ci * wd_im;
                                grid[2 * cell + 1] += cr * wd_im + ci * wd_re;
                            }
                            norm_re += cr;
                            norm_im += ci;
                        }
                    }
                    double *sw = sum_weight + a_chan * n_imag_pol + a_pol;
                    if (do_psf)
                        *sw = *sw + w * norm_re;                              /* :509 */
                    else                                                      /* Re(norm**2) :511 */
                        *sw = *sw + w * (norm_re * norm_re - norm_im * norm_im);
                }
            }
        }
    free(us);
    free(vs);
}

/*
 * A6 : _aperture_weight_grid_jit  (_aperture_grid.py:180-291)
 * Stamps weight * (weight_conv_kernel * phase_gradient[field]) at the GRID CENTRE with
 * unshifted CF indices (:276-287); the per-sample bounds test still applies (:240).
 */
void oracle_aperture_weight_grid(double *grid, double *sum_weight, const double *uvw, const double *freq_chan,
                                 const i64 *chan_map, const i64 *pol_map, const i64 *cf_baseline_map,
                                 const i64 *cf_chan_map, const i64 *cf_pol_map, const double *imaging_weight,
                                 const double *weight_conv_kernel, const i64 *weight_support,
                                 const double *phase_gradient, const i64 *field, const i64 *field_id,
                                 i64 n_field, i64 n_time, i64 n_baseline, i64 n_chan, i64 n_pol,
                                 i64 n_imag_pol, i64 n_u, i64 n_v, const double *delta_lm,
                                 const i64 *oversampling, i64 n_cfb, i64 n_cfc, i64 n_cfp, i64 n_cu, i64 n_cv)
{
    double *us = (double *)malloc(sizeof(double) * (size_t)n_chan);
    double *vs = (double *)malloc(sizeof(double) * (size_t)n_chan);
    fill_uv_scale(us, vs, freq_chan, n_chan, delta_lm, n_u, n_v);
    const i64 u_mid = n_u / 2, v_mid = n_v / 2;
    const i64 msc = max_support(weight_support, n_cfb * n_cfc * n_cfp * 2);
    const i64 conv_u_center = n_cu / 2, conv_v_center = n_cv / 2;

    for (i64 it = 0; it < n_time; ++it)
        for (i64 ib = 0; ib < n_baseline; ++ib) {
            const i64 f = field[it * n_baseline + ib];
            if (!(f > -1)) continue;
            const i64 field_indx = find_field(field_id, n_field, f);
            if (field_indx < 0) continue;
            const double *pg = phase_gradient + 2 * field_indx * n_cu * n_cv;
            const i64 cf_b = cf_baseline_map[ib];
            const double *p_uvw = uvw + (it * n_baseline + ib) * 3;
            for (i64 ic = 0; ic < n_chan; ++ic) {
                const i64 cf_c = cf_chan_map[ic];
                const i64 a_chan = chan_map[ic];
                const double u = p_uvw[0] * us[ic];
                const double v = p_uvw[1] * vs[ic];
                if (isnan(u) || isnan(v)) continue;
                const double u_pos = u + (double)u_mid;
                const double v_pos = v + (double)v_mid;
                const i64 uc = trunc_to_int(u_pos + 0.5);
                const i64 vc = trunc_to_int(v_pos + 0.5);
                if (!((uc + msc < n_u) && (vc + msc < n_v) && (uc - msc >= 0) && (vc - msc >= 0)))
                    continue;
                const i64 sample = ((it * n_baseline + ib) * n_chan + ic) * n_pol;
                for (i64 ip = 0; ip < n_pol; ++ip) {
                    const double w = imaging_weight[sample + ip];
                    if (isnan(w) || w == 0.0) continue;
                    const i64 cf_p = cf_pol_map[ip];
                    const i64 a_pol = pol_map[ip];
                    const i64 cf = (cf_b * n_cfc + cf_c) * n_cfp + cf_p;
                    const i64 su = weight_support[cf * 2 + 0], sv = weight_support[cf * 2 + 1];
                    const i64 su_c = su / 2, sv_c = sv / 2;
                    const double *ck = weight_conv_kernel + cf * n_cu * n_cv;
                    const i64 plane = (a_chan * n_imag_pol + a_pol) * n_u;
                    double norm_re = 0.0;
                    for (i64 iv = -sv_c; iv < sv - sv_c; ++iv) {
                        const i64 v_indx = v_mid + iv;
                        const i64 cf_v = oversampling[1] * iv + conv_v_center;
                        for (i64 iu = -su_c; iu < su - su_c; ++iu) {
                            const i64 u_indx = u_mid + iu;
                            const i64 cf_u = oversampling[0] * iu + conv_u_center;
                            const double k = ck[cf_u * n_cv + cf_v];
                            const double pr = pg[2 * (cf_u * n_cv + cf_v)];
                            const double pi = pg[2 * (cf_u * n_cv + cf_v) + 1];
                            const double cr = k * pr - 0.0 * pi;
                            const double ci = k * pi + 0.0 * pr;
                            const i64 cell = (plane + u_indx) * n_v + v_indx;
                            grid[2 * cell] += cr * w - ci * 0.0;
                            grid[2 * cell + 1] += cr * 0.0 + ci * w;
                            norm_re += cr;
                        }
                    }
                    double *sw = sum_weight + a_chan * n_imag_pol + a_pol;
                    *sw = *sw + w * norm_re;                                   /* :289 */
                }
            }
        }
    free(us);
    free(vs);
}

/*
 * A7 : degridding predict.  NO REFERENCE IMPLEMENTATION EXISTS
 * (predict_modelvis_image.py:20-40 is a stub; _standard_grid.py:418-430 prints
 * "still needs to be implemented").  PARITY UNPINNED: this is the exact adjoint of
 * oracle_standard_grid (same index math, same taps, gather instead of scatter),
 * checked by <grid(x), y> == <x, degrid(y)> in tests/.
 *   model_grid (n_imag_chan, n_imag_pol, n_u, n_v) complex128 ; vis out (n_t,n_b,n_c,n_p) complex128
 * Samples the gridder would skip (NaN uv, stamp off grid) yield 0.
 */
void oracle_standard_degrid(double *vis, const double *model_grid, const double *uvw, const double *freq_chan,
                            const i64 *chan_map, const i64 *pol_map, const double *cgk_1D, i64 n_time,
                            i64 n_baseline, i64 n_chan, i64 n_pol, i64 n_imag_pol, i64 n_u, i64 n_v,
                            const double *delta_lm, i64 support, i64 oversampling, int normalize)
{
    double *us = (double *)malloc(sizeof(double) * (size_t)n_chan);
    double *vs = (double *)malloc(sizeof(double) * (size_t)n_chan);
    fill_uv_scale(us, vs, freq_chan, n_chan, delta_lm, n_u, n_v);
    const i64 sc = support / 2;
    const i64 u_mid = n_u / 2, v_mid = n_v / 2;
    for (i64 it = 0; it < n_time; ++it)
        for (i64 ib = 0; ib < n_baseline; ++ib) {
            const double *p_uvw = uvw + (it * n_baseline + ib) * 3;
            for (i64 ic = 0; ic < n_chan; ++ic) {
                const i64 sample = ((it * n_baseline + ib) * n_chan + ic) * n_pol;
                for (i64 ip = 0; ip < n_pol; ++ip) vis[2 * (sample + ip)] = vis[2 * (sample + ip) + 1] = 0.0;
                const i64 a_chan = chan_map[ic];
                const double u = p_uvw[0] * us[ic];
                const double v = p_uvw[1] * vs[ic];
                if (isnan(u) || isnan(v)) continue;
                const double u_pos = u + (double)u_mid;
                const double v_pos = v + (double)v_mid;
                const i64 uc = trunc_to_int(u_pos + 0.5);
                const i64 vc = trunc_to_int(v_pos + 0.5);
                if (!((uc + sc < n_u) && (vc + sc < n_v) && (uc - sc >= 0) && (vc - sc >= 0))) continue;
                const i64 u_off_idx = (i64)floor(((double)uc - u_pos) * (double)oversampling + 0.5);
                const i64 v_off_idx = (i64)floor(((double)vc - v_pos) * (double)oversampling + 0.5);
                for (i64 ip = 0; ip < n_pol; ++ip) {
                    const i64 plane = (a_chan * n_imag_pol + pol_map[ip]) * n_u;
                    double acc_re = 0.0, acc_im = 0.0, norm = 0.0;
                    for (i64 iv = -sc; iv < support - sc; ++iv) {
                        const double conv_v = cgk_1D[llabs(oversampling * iv + v_off_idx)];
                        for (i64 iu = -sc; iu < support - sc; ++iu) {
                            const double conv = cgk_1D[llabs(oversampling * iu + u_off_idx)] * conv_v;
                            const i64 cell = (plane + uc + iu) * n_v + vc + iv;
                            acc_re += conv * model_grid[2 * cell];
                            acc_im += conv * model_grid[2 * cell + 1];
                            norm += conv;
                        }
                    }
                    if (normalize) { /* same normalisation the imaging side applies via sum_weight (:366) */
                        acc_re /= norm;
                        acc_im /= norm;
                    }
                    vis[2 * (sample + ip)] = acc_re;
                    vis[2 * (sample + ip) + 1] = acc_im;
                }
            }
        }
    free(us);
    free(vs);
}

int oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
