/*
 * bdl_oracle.c -- plain-C CPU restatement of BayesDLL's sampler update rules and of the
 * Philox4x32-10 + Box-Muller noise stream.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Used by tests/ as a checker at sizes where
 * the numpy oracle is slow, and by bench.py's cpu_baseline / --impl reference leg as the timed
 * CPU port of the reference path (OpenMP over all host cores).  The product (bayesdll_b200)
 * never links or calls this file.
 *
 * Same struct layouts as include/bdl.h so tests can drive both sides with identical arguments;
 * all pointers here are HOST pointers.
 *
 * Reference sites restated (relative to the reference checkout):
 *   SGLD        methods/sgld.py:469-484 + torch SGD step :226
 *   SGHMC       methods/sghmc.py:482-510 + SGD(momentum=0) :229
 *   cSGHMC      methods/csghmc.py:747-778
 *   Adam-SGHMC  methods/adam_sghmc.py:507-553 + SGD(momentum) :233
 *   Adam-cSGHMC methods/adam_csghmc.py:814-861 + SGD(momentum=0) :322
 * Philox4x32-10: Salmon et al., SC'11 (Random123 v1.14 philox.h); pinned by the Random123
 * known-answer vectors in tests/test_philox_cpu.py.
 *
 * Compile with -ffp-contract=off: every fp32 op must round once, the only fused op is the
 * explicit fmaf() that restates torch's add_(x, alpha=-lr).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "bdl.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* Thread control for the timed CPU baseline: torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently
 * turn the "all host cores" baseline into a single-threaded one. */
int bdl_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                 uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void bdl_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

static inline void box_muller(uint32_t ra, uint32_t rb, float* z0, float* z1) {
    const float u = fmaf((float)ra, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float a = (float)rb * 2.3283064365386963e-10f;
    const float r = sqrtf(-1.3862943611198906f * log2f(u));
    const float ang = 6.2831853071795865f * a;
    *z0 = r * cosf(ang);
    *z1 = r * sinf(ang);
}

static inline void normal4(uint64_t seed, uint32_t stream_id, uint64_t subseq, uint64_t q, float z[4]) {
    uint32_t r[4];
    philox4x32_10((uint32_t)q, stream_id, (uint32_t)subseq, (uint32_t)(subseq >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    box_muller(r[0], r[1], &z[0], &z[1]);
    box_muller(r[2], r[3], &z[2], &z[3]);
}

int bdl_oracle_philox_normal(float* out, uint64_t n, uint64_t seed, uint32_t stream_id, uint64_t subseq) {
    if (n % 4) return BDL_ERR_INVALID;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)(n / 4); ++q) normal4(seed, stream_id, subseq, (uint64_t)q, out + 4 * q);
    return BDL_OK;
}

/* ------------------------------------------------------------------------------------------ */
static inline float div_s(float x, float s, float inv_s, int mode) { return mode == BDL_DIV_IEEE ? x / s : x * inv_s; }
/* integer-valued divisors (sample counts): reciprocal mode = fp32(1.0 / s) taken in double, as torch CUDA does */
static inline float scalar_reciprocal(float s, int mode) { return mode == BDL_DIV_RECIP ? (float)(1.0 / (double)s) : 1.0f / s; }

typedef struct {
    float lr[2], c[2];
    float oma, sig2, inv_sig2, N, inv_N, mu, b1, omb1, b2, omb2, bc1, inv_bc1, bc2, inv_bc2, eps, two_alpha, nd, T, inv_T;
    int first_step, add_noise, div;
    int clip;      /* 0 none; 1 norm pass (clipq receives p.grad, caller stores nothing); 2 p.grad scaled by coef */
    float coef;
} sc_t;

static inline float prior_term(const sc_t* p, float th, float th0) {
    float d = th - th0;
    d = div_s(d, p->sig2, p->inv_sig2, p->div);
    d = div_s(d, p->N, p->inv_N, p->div);
    return d;
}

static inline float sgd_apply(const sc_t* p, int has_buf, float th, float gp, float lr, float* b) {
    float d = gp;
    if (has_buf) {
        *b = p->first_step ? gp : (*b * p->mu) + gp;
        d = *b;
    }
    return fmaf(d, -lr, th);
}

/* clipq: torch.nn.utils.clip_grad_norm_ between Model.forward and optimizer.step() (methods/csgld.py:250-251,
 * methods/adam_csghmc.py:319-320) acts on p.grad = g' (SGLD family) / v (Adam-cSGHMC). */
static inline void update_one(int variant, int has_buf, const sc_t* p, uint32_t cls, float* th, float g, float th0,
                              float* v, float* m, float* s, float* b, float xi, float* clipq) {
    const int h = cls & BDL_CLS_HEAD;
    const int prior = (cls & BDL_CLS_PRIOR) != 0;
    const float lr = p->lr[h];
    if (variant == BDL_SGLD) {
        const float noise = p->c[h] * xi;
        const float add = prior ? prior_term(p, *th, th0) + noise : noise;
        float gp = g + add;
        if (p->clip == 1) *clipq = gp;
        if (p->clip == 2) gp = gp * p->coef;
        *th = sgd_apply(p, has_buf, *th, gp, lr, b);
    } else if (variant == BDL_SGHMC) {
        const float gU = prior ? g + prior_term(p, *th, th0) : g;
        const float noise = p->c[h] * xi;
        *v = ((*v * p->oma) + (lr * gU)) + noise;
        const float gp = g + *v;
        *th = fmaf(gp, -lr, *th);
    } else if (variant == BDL_CSGHMC) {
        const float gU = g + (p->sig2 * *th);
        float vn = (*v * p->oma) - (lr * gU);
        if (p->add_noise) vn = vn + (p->c[h] * xi);
        *v = vn;
        *th = *th + vn;
    } else {
        const int cyc = variant == BDL_ADAM_CSGHMC;
        const float gl = cyc ? div_s(g, p->T, p->inv_T, p->div) : g;
        const float gU = prior ? gl + prior_term(p, *th, th0) : gl;
        *m = (p->b1 * *m) + (p->omb1 * gU);
        *s = (p->b2 * *s) + (p->omb2 * (gU * gU));
        const float mh = div_s(*m, p->bc1, p->inv_bc1, p->div);
        const float sh = div_s(*s, p->bc2, p->inv_bc2, p->div);
        const float den = sqrtf(sh) + p->eps;
        const float pg = mh / den;
        const float pre = 1.0f / den;
        const float ns = p->nd * sqrtf(div_s(p->two_alpha * pre, p->N, p->inv_N, p->div));
        const float noise = ns * xi;
        *v = ((*v * p->oma) + (lr * pg)) + noise;
        if (cyc) {
            if (p->clip == 1) *clipq = *v;
            *th = fmaf(p->clip == 2 ? *v * p->coef : *v, -lr, *th);
        } else {
            const float gp = g + *v;
            *th = sgd_apply(p, has_buf, *th, gp, lr, b);
        }
    }
}

/* Same contract as bdl_step() / bdl_step_gradnorm() / bdl_step_clipped() in include/bdl.h with HOST pointers (stream
 * ignored).  clip 0: plain step; 1: *sumsq += sum of squares of p.grad over the real elements of tensors with a gradient,
 * state untouched; 2: p.grad scaled by coef. */
static int step_impl(int variant, float* theta, const float* g, const float* theta0, float* v, float* m, float* s,
                     float* buf, uint64_t n, const bdl_run* runs, uint32_t nruns, const bdl_scalars* sc,
                     const bdl_noise* nz, int clip, float coef, double* sumsq) {
    if (variant < BDL_SGLD || variant > BDL_ADAM_CSGHMC || !theta || !runs || !sc || !nz || n % 4) return BDL_ERR_INVALID;
    if (clip && variant != BDL_SGLD && variant != BDL_ADAM_CSGHMC) return BDL_ERR_UNSUPPORTED;
    sc_t p;
    p.clip = clip; p.coef = coef;
    double total = 0.0;
    for (int h = 0; h < 2; ++h) { p.lr[h] = sc->lr[h]; p.c[h] = sc->noise_scale[h]; }
    /* reciprocal mode multiplies by the host-supplied fp32(1.0 / s_double) (what torch CUDA uses); 0 = not supplied */
#define INV_OF(s_f, host_inv) ((sc->div_mode == BDL_DIV_RECIP && (host_inv) != 0.0f) ? (host_inv) : 1.0f / (s_f))
    p.oma = sc->one_minus_alpha; p.sig2 = sc->sig2; p.inv_sig2 = INV_OF(sc->sig2, sc->inv_sig2); p.N = sc->N; p.inv_N = INV_OF(sc->N, sc->inv_N);
    p.mu = sc->mu; p.b1 = sc->beta1; p.omb1 = sc->one_minus_beta1; p.b2 = sc->beta2; p.omb2 = sc->one_minus_beta2;
    p.bc1 = sc->bias_corr1; p.inv_bc1 = INV_OF(sc->bias_corr1, sc->inv_bias_corr1); p.bc2 = sc->bias_corr2; p.inv_bc2 = INV_OF(sc->bias_corr2, sc->inv_bias_corr2);
    p.eps = sc->eps; p.two_alpha = sc->two_alpha; p.nd = sc->nd; p.T = sc->temperature; p.inv_T = INV_OF(sc->temperature, sc->inv_temperature);
#undef INV_OF
    p.first_step = sc->first_step; p.add_noise = sc->add_noise; p.div = sc->div_mode;
    const int has_buf = (variant == BDL_SGLD || variant == BDL_ADAM_SGHMC) && sc->mu != 0.0f;
    const float* xi = (const float*)(uintptr_t)nz->xi_dev;
    const uint64_t seed = nz->seed, subseq = nz->subseq;
    const uint32_t sid = nz->stream_id;
    for (uint32_t r = 0; r < nruns; ++r) {
        const bdl_run run = runs[r];
        if (run.cls & BDL_CLS_SKIP) continue;
        const int64_t q0 = (int64_t)(run.begin / 4), q1 = (int64_t)(run.end / 4);
#pragma omp parallel for schedule(static) reduction(+ : total)
        for (int64_t q = q0; q < q1; ++q) {
            float z[4];
            if (xi) memcpy(z, xi + 4 * q, sizeof z);
            else normal4(seed, sid, subseq, (uint64_t)q, z);
            for (int k = 0; k < 4; ++k) {
                const uint64_t i = 4 * (uint64_t)q + k;
                float gi;
                if (run.g_dev) gi = i < run.valid_end ? run.g_dev[i - run.begin] : 0.0f;
                else gi = g[i];
                float dv = 0, dm = 0, ds = 0, db = 0, cq = 0;
                if (clip == 1) {                         /* norm pass: work on copies, store nothing */
                    float th = theta[i], vv = v ? v[i] : 0, mm = m ? m[i] : 0, ss = s ? s[i] : 0, bb = buf ? buf[i] : 0;
                    update_one(variant, has_buf, &p, run.cls, &th, gi, theta0 ? theta0[i] : 0.0f, &vv, &mm, &ss, &bb, z[k], &cq);
                    if (i < run.valid_end) total += (double)cq * (double)cq;
                    continue;
                }
                update_one(variant, has_buf, &p, run.cls, &theta[i], gi, theta0 ? theta0[i] : 0.0f, v ? &v[i] : &dv,
                           m ? &m[i] : &dm, s ? &s[i] : &ds, buf ? &buf[i] : &db, z[k], &cq);
            }
        }
    }
    if (clip == 1 && sumsq) *sumsq += total;
    return BDL_OK;
}

int bdl_oracle_step(int variant, float* theta, const float* g, const float* theta0, float* v, float* m, float* s,
                    float* buf, uint64_t n, const bdl_run* runs, uint32_t nruns, const bdl_scalars* sc,
                    const bdl_noise* nz) {
    return step_impl(variant, theta, g, theta0, v, m, s, buf, n, runs, nruns, sc, nz, 0, 1.0f, NULL);
}

/* per-tensor runs required (valid_end delimits the real elements) */
int bdl_oracle_step_gradnorm(int variant, const float* theta, const float* g, const float* theta0, const float* v,
                             const float* m, const float* s, const float* buf, uint64_t n, const bdl_run* runs,
                             uint32_t nruns, const bdl_scalars* sc, const bdl_noise* nz, double* sumsq) {
    return step_impl(variant, (float*)theta, g, theta0, (float*)v, (float*)m, (float*)s, (float*)buf, n, runs, nruns, sc, nz,
                     1, 1.0f, sumsq);
}

/* total_norm = fp32(sqrt(sumsq)); coef = min(1, (total_norm + 1e-6).reciprocal() * max_norm) -- clip_grad_norm_'s statements */
float bdl_oracle_clip_coef(double sumsq, float max_norm, float* total_norm) {
    const float tn = (float)sqrt(sumsq);
    const float c = (1.0f / (tn + 1e-6f)) * max_norm;
    if (total_norm) *total_norm = tn;
    return c < 1.0f ? c : 1.0f;
}

int bdl_oracle_step_clipped(int variant, float* theta, const float* g, const float* theta0, float* v, float* m, float* s,
                            float* buf, uint64_t n, const bdl_run* runs, uint32_t nruns, const bdl_scalars* sc,
                            const bdl_noise* nz, float coef) {
    return step_impl(variant, theta, g, theta0, v, m, s, buf, n, runs, nruns, sc, nz, 2, coef, NULL);
}

/* posterior draw, same contract as bdl_draw() */
int bdl_oracle_draw(const float* mean, const float* second, const float* center, float* out, uint64_t n, int var_mode,
                    float scale, int div_mode, const bdl_noise* nz) {
    if (n % 4) return BDL_ERR_INVALID;
    const float* xi = (const float*)(uintptr_t)nz->xi_dev;
    const float inv = scalar_reciprocal(scale, div_mode);
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)(n / 4); ++q) {
        float z[4];
        if (xi) memcpy(z, xi + 4 * q, sizeof z);
        else normal4(nz->seed, nz->stream_id, nz->subseq, (uint64_t)q, z);
        for (int k = 0; k < 4; ++k) {
            const uint64_t i = 4 * (uint64_t)q + k;
            float var;
            if (var_mode == 0) { var = scale * (second[i] - (mean[i] * mean[i])); var = fmaxf(var, 1e-12f); }
            else if (var_mode == 1) var = fmaxf(div_s(second[i], scale, inv, div_mode), 1e-12f);
            else if (var_mode == 2) var = 1e-12f;
            else var = second[i];
            if (var_mode == 4) { out[i] = (center ? center[i] : mean[i]) + (fmaxf(second[i], 1e-8f) * z[k]); continue; } /* vi.py:402-406 */
            out[i] = (center ? center[i] : mean[i]) + (sqrtf(var) * z[k]);
        }
    }
    return BDL_OK;
}

/* MC-Dropout reparameterisation draw (methods/mc_dropout.py:378-394), same contract as bdl_dropout_mix() with host
 * pointers: z = (u > p_drop), theta = z*m + (1-z)*theta0; runs flagged BDL_CLS_NODROP keep z = 1.  In-kernel uniforms:
 * the 4 Philox outputs of group q (same counter layout as the Gaussian stream), 24 bits each, u = (r >> 8) * 2^-24. */
int bdl_oracle_dropout_mix(const float* m, const float* theta0, float* out, float* z_out, uint64_t n, const bdl_run* runs,
                           uint32_t nruns, float p_drop, const bdl_noise* nz) {
    if (n % 4) return BDL_ERR_INVALID;
    const float* u_in = (const float*)(uintptr_t)nz->xi_dev;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)(n / 4); ++q) {
        float u[4];
        if (u_in) memcpy(u, u_in + 4 * q, sizeof u);
        else {
            uint32_t r[4];
            philox4x32_10((uint32_t)q, nz->stream_id, (uint32_t)nz->subseq, (uint32_t)(nz->subseq >> 32), (uint32_t)nz->seed,
                          (uint32_t)(nz->seed >> 32), r);
            for (int k = 0; k < 4; ++k) u[k] = (float)(r[k] >> 8) * 5.9604644775390625e-08f;
        }
        uint32_t cls = 0;
        for (uint32_t j = 0; j < nruns; ++j)
            if ((uint64_t)(4 * q) < runs[j].end) { cls = runs[j].cls; break; }
        for (int k = 0; k < 4; ++k) {
            const uint64_t i = 4 * (uint64_t)q + k;
            const float z = ((cls & BDL_CLS_NODROP) || u[k] > p_drop) ? 1.0f : 0.0f;
            out[i] = (z * m[i]) + ((1.0f - z) * theta0[i]);
            if (z_out) z_out[i] = z;
        }
    }
    return BDL_OK;
}

int bdl_oracle_moments_avg(const float* theta, float* mom1, float* mom2, uint64_t n, float cnt, float cntp1, int init,
                           int div_mode) {
    const float inv = scalar_reciprocal(cntp1, div_mode);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const float t = theta[i];
        if (init) {
            mom1[i] = t * 1.0f;
            if (mom2) mom2[i] = t * t;
        } else {
            mom1[i] = div_s(t + (cnt * mom1[i]), cntp1, inv, div_mode);
            if (mom2) mom2[i] = div_s((t * t) + (cnt * mom2[i]), cntp1, inv, div_mode);
        }
    }
    return BDL_OK;
}

int bdl_oracle_moments_welford(const float* theta, float* mean, float* M2, uint64_t n, float nf, int init, int div_mode) {
    const float inv = scalar_reciprocal(nf, div_mode);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const float t = theta[i];
        if (init) { mean[i] = t; M2[i] = 0.0f; }
        else {
            const float d = t - mean[i];
            mean[i] = mean[i] + div_s(d, nf, inv, div_mode);
            const float d2 = t - mean[i];
            M2[i] = M2[i] + (d * d2);
        }
    }
    return BDL_OK;
}
