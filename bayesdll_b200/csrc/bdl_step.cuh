// bdl_step.cuh -- kernels and launch templates of the fused SG-MCMC sampler update.
//
// One pass over the padded flat state replaces the reference's per-tensor Python loop
// (methods/sghmc.py:482-510 etc.) *and* the torch.optim.SGD step that follows it
// (methods/sghmc.py:229): prior pull, friction/momentum, Adam moments, Gaussian noise (injected
// or in-kernel Philox), SGD momentum buffer and the parameter write, with every element read
// once and written once.
//
// Roofline: HBM bandwidth.  Algorithmic bytes per element (fp32):
//   SGLD mu!=0 24 (R theta,g,theta0,b; W b,theta) | SGLD mu=0 16 | SGHMC 24 | cSGHMC 20
//   Adam-SGHMC mu!=0 48 / mu=0 40 | Adam-cSGHMC 40 | +4 with injected noise.
//
// Arithmetic: every operation is an explicitly rounded fp32 intrinsic in the reference's
// operation order so that, with injected noise, results are bit-identical
// to oracle/ (and hence to the reference's eager ops).  The only fused multiply-add is the one
// torch's `add_(x, alpha=-lr)` performs.
//
// Mapping: ONE tile per CTA, CTAs dispatched in address order (grid = #tiles); 64-256 threads by variant, each thread
// owns kU (default 1) float4 groups, consecutive threads touch consecutive 16-byte groups (512 B per warp per
// stream), all loads of a tile are issued before the first dependent instruction.  Three builds of the same
// arithmetic: step_kernel<kFast> (flat gradient, body | head table in the kernel arguments: the benchmarked launch),
// step_table_kernel (device run table with per-tensor gradient pointers: the launch Runner.train() makes; 4
// consecutive tiles per CTA) and the generic step_kernel (capped / persistent grids via bdl_set_launch_config,
// chunked host-buffer steps, > 2^31 tiles).  A persistent grid-stride grid measured 8 % slower.
#pragma once
#include <cstdlib>

#include "bdl_common.cuh"

#ifndef BDL_ADAM_OPT
#define BDL_ADAM_OPT 1      // Adam variants: optimistic fast-path sqrt / rcp / quotient, one cold branch per element
#endif

namespace bdl {

constexpr int kInlineRuns = 8;

struct PTable;

struct StepParams {
    float* theta;
    const float* g;
    const float* theta0;
    float* v;
    float* m;
    float* s;
    float* buf;
    const float* xi;
    const bdl_run* runs;
    uint32_t nruns;
    uint32_t flat_g;           // 1: no run carries its own gradient pointer (known from the host table): g loads do not wait for the lookup
    uint32_t inl_n;            // > 0: the run table is small and has no gradient pointers -> inlined below (constant bank)
    uint32_t inl_end4[kInlineRuns];
    uint32_t inl_cls[kInlineRuns];
    uint32_t n4;       // one past the last float4 group to process
    uint32_t q_begin;  // first float4 group to process (0 except for chunked host-buffer steps)
    uint32_t tpc;      // > 1: every CTA walks this many CONSECUTIVE tiles (device run tables: one table search per CTA, then a cursor)
    // scalars (already rounded to fp32 by the host)
    float lr[2], neg_lr[2], c[2];
    float oma, sig2, inv_sig2, N, inv_N, mu;
    float b1, omb1, b2, omb2, bc1, inv_bc1, bc2, inv_bc2, eps, two_alpha, nd, T, inv_T;
    int first_step, add_noise;
    // fused sample capture (kCap != 0): running moments of the NEW theta, same arithmetic as bdl_capture.cu
    float* cap1;               // mom1 (avg) / mean (Welford)
    float* cap2;               // mom2 (avg, may be null: nst == 0) / M2 (Welford)
    float cap_a, cap_b, cap_inv;   // avg: cnt, cnt+1, 1/(cnt+1);  Welford: n, -, 1/n
    int cap_init, cap_kind;
    NoiseKey key;
    const PTable* ptab;        // HOST pointer (launch plumbing only, never dereferenced on the device): table for step_ptable_kernel
};

// Run table carried INSIDE the kernel arguments (constant bank; CUDA >= 12.1 allows 32 KB of parameters on sm_70+): the
// launch Runner.train() makes has <= 512 rows for every backbone of the reference (ViT-L/32: 296, ResNet-101: 314), so
// rows and a coarse directory (first run of every 2^dir_shift groups) fit: the per-CTA lookup becomes a few uniform
// constant loads instead of a 32-ary search through L1/L2, and no table has to be staged to the device at all.
constexpr int kPRows = 512;
constexpr int kPDir = 5120;
struct PTable {
    uint32_t nrows, dir_shift;
    uint32_t end4[kPRows];         // run end, in float4 groups
    uint32_t valid_end[kPRows];    // one past the tensor's last real element (elements; [valid_end, 4 * end4) is padding)
    uint32_t tail_q[kPRows];       // group holding the tensor's last real elements when numel % 4 != 0 (own gradient), else 0xFFFFFFFF
    uint64_t gbase[kPRows];        // address A such that the gradient of flat element i is ((const float*)A)[i]
    uint8_t cls[kPRows];
    uint8_t tail_n[kPRows];        // real elements in the tail group (1..3)
    uint16_t dir[kPDir];           // dir[j] = first run whose end is beyond group (j << dir_shift)
};
static_assert(sizeof(PTable) < 24 * 1024, "PTable + StepParams must stay well inside the 32 KB kernel-parameter space");

struct RunCursor {
    uint32_t idx;
    uint32_t end4;        // end / 4 of the current run
    uint32_t cls;
    const float* gbase;   // address of gradient element for flat index i is gbase + i
    uint64_t valid_end;
    bool own_g;
};

__device__ __forceinline__ void cursor_load(RunCursor& c, const StepParams& p, uint32_t idx) {
    const bdl_run* r = p.runs + idx;
    c.idx = idx;
    c.end4 = static_cast<uint32_t>(__ldg(&r->end) >> 2);
    c.cls = __ldg(&r->cls);
    const float* gr = reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(&r->g_dev)));
    c.own_g = gr != nullptr;
    c.valid_end = __ldg(&r->valid_end);
    c.gbase = c.own_g ? gr - __ldg(&r->begin) : p.g;
}

__device__ __forceinline__ void cursor_seek(RunCursor& c, const StepParams& p, uint32_t q) {
    // runs are sorted and contiguous; q only ever increases within a thread
    while (q >= c.end4 && c.idx + 1 < p.nruns) cursor_load(c, p, c.idx + 1);
}

// -------------------------------------------------------------------------------------------
// per-element update rules
// -------------------------------------------------------------------------------------------
template <int kDiv>
__device__ __forceinline__ float prior_term(const StepParams& p, float th, float th0) {
    float d = __fsub_rn(th, th0);                       // (p - p0)
    d = div_scalar<kDiv>(d, p.sig2, p.inv_sig2);        //   / prior_sig**2
    d = div_scalar<kDiv>(d, p.N, p.inv_N);              //   / N
    return d;
}

template <bool kHasBuf>
__device__ __forceinline__ float sgd_apply(const StepParams& p, float th, float gp, float neg_lr, float& b) {
    float d = gp;
    if constexpr (kHasBuf) {
        b = p.first_step ? gp : __fadd_rn(__fmul_rn(b, p.mu), gp);
        d = b;
    }
    return __fmaf_rn(d, neg_lr, th);                    // param.add_(d, alpha=-lr)
}

// kClip (gradient-norm clipping between Model.forward and optimizer.step(), methods/csgld.py:250-251): 0 = none;
// 1 = norm pass: ``clipq`` receives what the reference holds in p.grad at that point (nothing is stored by the caller);
// 2 = apply pass: that quantity is scaled by ``coef`` = min(1, max_norm / (||p.grad|| + 1e-6)) before the SGD step.
template <int kVariant, bool kHasBuf, int kDiv, int kClip = 0>
__device__ __forceinline__ void update_one(const StepParams& p, uint32_t cls, float& th, float g, float th0,
                                           float& v, float& m, float& s, float& b, float xi, float coef = 1.0f,
                                           float* clipq = nullptr) {
    const int h = cls & BDL_CLS_HEAD;
    const bool prior = (cls & BDL_CLS_PRIOR) != 0;
    const float lr = p.lr[h], neg_lr = p.neg_lr[h];
    if constexpr (kVariant == BDL_SGLD) {
        const float noise = __fmul_rn(p.c[h], xi);
        const float add = prior ? __fadd_rn(prior_term<kDiv>(p, th, th0), noise) : noise;
        float gp = __fadd_rn(g, add);
        if constexpr (kClip == 1) *clipq = gp;
        if constexpr (kClip == 2) gp = __fmul_rn(gp, coef);         // g.mul_(clip_coef_clamped)
        th = sgd_apply<kHasBuf>(p, th, gp, neg_lr, b);
    } else if constexpr (kVariant == BDL_SGHMC) {
        const float gU = prior ? __fadd_rn(g, prior_term<kDiv>(p, th, th0)) : g;
        const float noise = __fmul_rn(p.c[h], xi);
        v = __fadd_rn(__fadd_rn(__fmul_rn(v, p.oma), __fmul_rn(lr, gU)), noise);
        const float gp = __fadd_rn(g, v);               // p.grad = p.grad + v
        th = __fmaf_rn(gp, neg_lr, th);                 // SGD(momentum=0)
    } else if constexpr (kVariant == BDL_CSGHMC) {
        const float gU = __fadd_rn(g, __fmul_rn(p.sig2, th));     // g + prior_sig * theta
        float vn = __fsub_rn(__fmul_rn(v, p.oma), __fmul_rn(lr, gU));
        if (p.add_noise) vn = __fadd_rn(vn, __fmul_rn(p.c[h], xi));
        v = vn;
        th = __fadd_rn(th, vn);                         // p.data.add_(v)
    } else {
        constexpr bool kCyc = (kVariant == BDL_ADAM_CSGHMC);
        const float gl = kCyc ? div_scalar<kDiv>(g, p.T, p.inv_T) : g;
        const float gU = prior ? __fadd_rn(gl, prior_term<kDiv>(p, th, th0)) : gl;
        m = __fadd_rn(__fmul_rn(p.b1, m), __fmul_rn(p.omb1, gU));
        s = __fadd_rn(__fmul_rn(p.b2, s), __fmul_rn(p.omb2, __fmul_rn(gU, gU)));
        const float mh = div_scalar<kDiv>(m, p.bc1, p.inv_bc1);
        const float sh = div_scalar<kDiv>(s, p.bc2, p.inv_bc2);
        float pg, ns;
#if BDL_ADAM_OPT
        // the three correctly rounded sqrt / reciprocal / quotient steps, optimistically on their fast paths with ONE
        // cold branch to the library intrinsics per element (bdl_common.cuh: same bits)
        bool slow = false;
        const float den_f = __fadd_rn(sqrt_rn_opt(sh, slow), p.eps);
        const float pre_f = rcp_rn_fast(den_f);             // range: div_by_rcp_opt's divisor window (2^-30, 2^30) is inside rcp's
        pg = div_by_rcp_opt(mh, den_f, pre_f, slow);
        ns = __fmul_rn(p.nd, sqrt_rn_opt(div_scalar<kDiv>(__fmul_rn(p.two_alpha, pre_f), p.N, p.inv_N), slow));
        if (slow)
#endif
        {
            const float den = __fadd_rn(__fsqrt_rn(sh), p.eps);
            const float pre = __frcp_rn(den);               // 1.0 / den, correctly rounded
            pg = div_by_rcp(mh, den, pre);                  // m_hat / den, correctly rounded (shares the reciprocal)
            ns = __fmul_rn(p.nd, __fsqrt_rn(div_scalar<kDiv>(__fmul_rn(p.two_alpha, pre), p.N, p.inv_N)));
        }
        const float noise = __fmul_rn(ns, xi);
        v = __fadd_rn(__fadd_rn(__fmul_rn(v, p.oma), __fmul_rn(lr, pg)), noise);
        if constexpr (kCyc) {
            if constexpr (kClip == 1) *clipq = v;
            th = __fmaf_rn(kClip == 2 ? __fmul_rn(v, coef) : v, neg_lr, th);   // p.grad = v [* clip coef] ; SGD(momentum=0)
        } else {
            const float gp = __fadd_rn(g, v);           // p.grad = p.grad + v
            th = sgd_apply<kHasBuf>(p, th, gp, neg_lr, b);
        }
    }
}

template <int kVariant>
struct Uses {
    static constexpr bool theta0 = (kVariant != BDL_CSGHMC);
    static constexpr bool v = (kVariant != BDL_SGLD);
    static constexpr bool adam = (kVariant == BDL_ADAM_SGHMC || kVariant == BDL_ADAM_CSGHMC);
};

// Fold the (new) theta of one float4 group into the running moments: the arithmetic of bdl_moments_avg (kCap == 1) /
// bdl_moments_welford (kCap == 2) in bdl_capture.cu, on values still in registers.
template <int kCap, int kDiv>
__device__ __forceinline__ void capture_fold(const StepParams& p, uint64_t i, const float4& th, const float4& c1, const float4& c2) {
    const float t[4] = {th.x, th.y, th.z, th.w};
    float a[4] = {c1.x, c1.y, c1.z, c1.w};
    float bb[4] = {c2.x, c2.y, c2.z, c2.w};
    const bool has2 = kCap == 2 || p.cap2 != nullptr;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if constexpr (kCap == 1) {
            if (p.cap_init) {
                a[k] = __fmul_rn(t[k], 1.0f);                      // theta_vec*1.0
                bb[k] = __fmul_rn(t[k], t[k]);                     // theta_vec**2
            } else {                                               // (theta^k + cnt*mom) / (cnt+1)
                a[k] = div_scalar<kDiv>(__fadd_rn(t[k], __fmul_rn(p.cap_a, a[k])), p.cap_b, p.cap_inv);
                if (has2)
                    bb[k] = div_scalar<kDiv>(__fadd_rn(__fmul_rn(t[k], t[k]), __fmul_rn(p.cap_a, bb[k])), p.cap_b, p.cap_inv);
            }
        } else {
            if (p.cap_init) {
                a[k] = t[k];                                       // mean = theta.clone()
                bb[k] = 0.0f;                                      // M2 = zeros_like
            } else {
                const float d = __fsub_rn(t[k], a[k]);             // delta
                a[k] = __fadd_rn(a[k], div_scalar<kDiv>(d, p.cap_a, p.cap_inv));
                const float d2 = __fsub_rn(t[k], a[k]);            // delta2
                bb[k] = __fadd_rn(bb[k], __fmul_rn(d, d2));
            }
        }
    }
    st_stream(p.cap1 + i, make_float4(a[0], a[1], a[2], a[3]));
    if (has2) st_stream(p.cap2 + i, make_float4(bb[0], bb[1], bb[2], bb[3]));
}

constexpr int kDefaultUnroll = 1;
constexpr long kDefaultTableTpc = 4;   // profiles/r01_ab_table_tpc.log: 1.169 -> 1.080 ms (SGHMC), 2.111 -> 1.970 ms (Adam-cSGHMC) at ViT-L/32 size

// Resident CTAs per SM the kernel is compiled for: 16 data registers per stream per unroll step plus ~28 registers of
// addressing / Philox state, rounded to the allocation granule, against the 64K-entry register file.
template <int kVariant, bool kHasBuf, bool kPhilox, int kU, int kT, int kCap = 0>
constexpr int min_blocks() {
    int streams = (kVariant == BDL_SGLD) ? 3 : (kVariant == BDL_SGHMC) ? 4 : (kVariant == BDL_CSGHMC) ? 3 : 6;
    streams += (kHasBuf ? 1 : 0) + (kPhilox ? 0 : 1) + (kCap ? 2 : 0);
    int regs = 4 * kU * streams + 28 + (kCap ? 12 : 0);   // capture: moment arithmetic temporaries live across the update
    regs = (regs + 7) / 8 * 8;
    int blocks = 65536 / (kT * regs);
    const int cap = 2048 / kT > 32 ? 32 : 2048 / kT;   // 64 warps and 32 CTAs per SM
    blocks = blocks > cap ? cap : blocks;
    return blocks < 1 ? 1 : blocks;
}

__device__ __forceinline__ uint32_t cursor_find_warp(const StepParams& p, uint32_t q) { return run_find_warp(p.runs, p.nruns, q); }

// Launch shape.  Default: ONE tile per CTA (grid = #tiles), CTAs dispatched in address order.  Measured on B200 this
// beats a persistent grid-stride grid by ~8 % (6.87 vs 6.35 TB/s on the SGHMC step): in-order dispatch keeps the set
// of DRAM pages being streamed compact, whereas persistent CTAs drift apart and scatter the access window.  The
// tile loop remains for capped grids (bdl_set_launch_config) and for > 2^31 tiles.
// kCap: 0 = plain step; 1 = also fold the new theta into running moments (bdl_moments_avg arithmetic); 2 = Welford
// (bdl_moments_welford arithmetic).  Fusing saves the capture kernel's re-read of theta: 40 instead of 44 B/param for
// SGHMC + moments, which is every step after burn-in when thin = 1 (BASELINE.json configs[2]).
// kFast: the launch every BASELINE config makes -- grid == #tiles (one tile per CTA, no tile loop), the two-run
// body | head table inside the kernel arguments, a flat gradient buffer (with or without capture).  Same arithmetic, ~10 % fewer
// instructions (no tile-loop bookkeeping, no table-kind dispatch); matters when the box's power cap pulls the SM clock
// down and the kernel turns issue-sensitive.
template <int kVariant, bool kHasBuf, bool kPhilox, int kDiv, int kU, int kT, int kCap = 0, bool kFast = false>
__global__ void __launch_bounds__(kT, (min_blocks<kVariant, kHasBuf, kPhilox, kU, kT, kCap>()))
step_kernel(const StepParams p) {
    using U = Uses<kVariant>;
    constexpr uint32_t tile_groups = kT * kU;
    uint32_t tile = blockIdx.x;
    uint32_t ntiles = 0, tile_step = 0;
    if constexpr (!kFast) {
        ntiles = (p.n4 - p.q_begin + tile_groups - 1) / tile_groups;
        tile_step = gridDim.x;                             // capped grid: grid-stride over the tiles
        if (p.tpc > 1) {                                   // consecutive tiles: still dispatched in address order
            tile = blockIdx.x * p.tpc;
            tile_step = 1;
            if (ntiles > tile + p.tpc) ntiles = tile + p.tpc;
        }
        if (tile >= ntiles) return;
    }
    RunCursor cur;
    bool have_cursor = false;

    for (;;) {
        const uint32_t q0 = p.q_begin + tile * tile_groups + threadIdx.x;
        float4 th[kU], g[kU], th0[kU], v[kU], m[kU], s[kU], b[kU], xi[kU], c1[kU], c2[kU];
        uint32_t cls[kU];
        bool act[kU], inr[kU];
        // ---- 1. every load that does not depend on the run table (kU * #streams independent 128-bit requests) ----
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const uint32_t q = q0 + u * kT;
            act[u] = q < p.n4;
            inr[u] = act[u];
            if (act[u]) {
                const uint64_t i = static_cast<uint64_t>(q) << 2;
                th[u] = ld_stream(p.theta + i);
                if constexpr (kCap != 0) {
                    if (!p.cap_init) {
                        c1[u] = ld_stream(p.cap1 + i);
                        if (kCap == 2 || p.cap2) c2[u] = ld_stream(p.cap2 + i);
                    }
                }
                if constexpr (U::theta0) th0[u] = ld_stream(p.theta0 + i);
                if constexpr (U::v) v[u] = ld_stream(p.v + i);
                if constexpr (U::adam) {
                    m[u] = ld_stream(p.m + i);
                    s[u] = ld_stream(p.s + i);
                }
                if constexpr (kHasBuf) b[u] = ld_stream(p.buf + i);
                if constexpr (!kPhilox) xi[u] = ld_stream(p.xi + i);
                if (kFast || p.flat_g) g[u] = ld_stream(p.g + i);     // no per-run gradient pointers: the load need not wait for the class lookup
            }
        }
        // ---- 2. element class / gradient pointer, then the gradient loads ----
        if constexpr (kFast) {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const uint32_t q = q0 + u * kT;
                const uint32_t c = q >= p.inl_end4[0] ? p.inl_cls[1] : p.inl_cls[0];
                cls[u] = c;
                act[u] = act[u] && (c & BDL_CLS_SKIP) == 0;
            }
        } else if (p.inl_n) {
            // small merged table (e.g. body | head): classes come from the kernel arguments, no table loads at all
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const uint32_t q = q0 + u * kT;
                if (act[u]) {
                    uint32_t c = p.inl_cls[0];
                    if (p.inl_n == 2) {                    // body | head, the layout of every BASELINE config: one compare
                        if (q >= p.inl_end4[0]) c = p.inl_cls[1];
                    } else {
#pragma unroll
                        for (int r = 1; r < kInlineRuns; ++r)
                            if (r < static_cast<int>(p.inl_n) && q >= p.inl_end4[r - 1]) c = p.inl_cls[r];
                    }
                    cls[u] = c;
                    act[u] = (c & BDL_CLS_SKIP) == 0;
                }
            }
        } else {
        // run table in device memory (L1-resident after the first CTA of an SM touched it)
        if (!have_cursor) {
            // ONE search per CTA (warp 0, for the CTA's first group), published through shared memory.  The probes of a
            // 32-ary search touch ~40 separate L1 sectors, more than the warp's own 24 data sectors; every thread then walks
            // forward from the CTA's run (uniform addresses: one broadcast sector per field).  A/B against a search per
            // warp (profiles/r01_ab_search_mode.log): equal at 64 threads, 1-2 % faster at 128 / 256.  have_cursor and
            // the tile loop are CTA-uniform, so the barrier is safe.
            __shared__ uint32_t run0_sh;
            if (threadIdx.x < 32) {
                const uint32_t r = cursor_find_warp(p, p.q_begin + tile * tile_groups);   // < n4 for every launched tile
                if (threadIdx.x == 0) run0_sh = r;
            }
            __syncthreads();
            cursor_load(cur, p, run0_sh);
            have_cursor = true;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const uint32_t q = q0 + u * kT;
            if (act[u]) {
                cursor_seek(cur, p, q);
                cls[u] = cur.cls;
                const uint64_t i = static_cast<uint64_t>(q) << 2;
                const bool skipped = (cur.cls & BDL_CLS_SKIP) != 0;       // p.grad is None: no gradient to read (g may be null)
                if (!p.flat_g && !skipped) g[u] = ld_stream(cur.gbase + i);
                if (cur.own_g && i + 4 > cur.valid_end) {  // tail group of a per-run gradient: zero the padding lanes
                    if (i + 0 >= cur.valid_end) g[u].x = 0.f;
                    if (i + 1 >= cur.valid_end) g[u].y = 0.f;
                    if (i + 2 >= cur.valid_end) g[u].z = 0.f;
                    if (i + 3 >= cur.valid_end) g[u].w = 0.f;
                }
                act[u] = (cur.cls & BDL_CLS_SKIP) == 0;     // p.grad is None -> tensor left untouched
            }
        }
        }
        // ---- 3. compute + store ----
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const uint32_t q = q0 + u * kT;
            if (act[u]) {
                const uint64_t i = static_cast<uint64_t>(q) << 2;
                if constexpr (kPhilox) xi[u] = philox_normal4(p.key, q);
                update_one<kVariant, kHasBuf, kDiv>(p, cls[u], th[u].x, g[u].x, th0[u].x, v[u].x, m[u].x, s[u].x, b[u].x, xi[u].x);
                update_one<kVariant, kHasBuf, kDiv>(p, cls[u], th[u].y, g[u].y, th0[u].y, v[u].y, m[u].y, s[u].y, b[u].y, xi[u].y);
                update_one<kVariant, kHasBuf, kDiv>(p, cls[u], th[u].z, g[u].z, th0[u].z, v[u].z, m[u].z, s[u].z, b[u].z, xi[u].z);
                update_one<kVariant, kHasBuf, kDiv>(p, cls[u], th[u].w, g[u].w, th0[u].w, v[u].w, m[u].w, s[u].w, b[u].w, xi[u].w);
                if constexpr (U::v) st_stream(p.v + i, v[u]);
                if constexpr (U::adam) {
                    st_stream(p.m + i, m[u]);
                    st_stream(p.s + i, s[u]);
                }
                if constexpr (kHasBuf) st_stream(p.buf + i, b[u]);
                st_stream(p.theta + i, th[u]);
            }
            if constexpr (kCap != 0) {
                if (inr[u])                                // also for skipped tensors: their (unchanged) theta is a sample too
                    capture_fold<kCap, kDiv>(p, static_cast<uint64_t>(q) << 2, th[u], c1[u], c2[u]);
            }
        }
        if constexpr (kFast) {
            break;
        } else {
            tile += tile_step;
            if (tile >= ntiles) break;
        }
    }
}

// -------------------------------------------------------------------------------------------
// Lean build of the training-loop launch: a device run table with one row per tensor,
// each row carrying the address of that tensor's own autograd gradient (p.grad read in place, no gather pass).
// Same arithmetic as step_kernel (update_one / capture_fold), different control:
//   * a CTA owns kTableTpc CONSECUTIVE tiles (still dispatched in address order): ONE table search per CTA (warp 0,
//     32-ary ballot search, published through shared memory), issued after the first tile's table-independent loads;
//   * CTA-uniform fast path: when the CTA's whole span lies inside the run found (every CTA of a large tensor), class,
//     gradient base and "tensor has a gradient" are CTA-uniform -- no per-thread cursor, no tail test, no bounds test,
//     tiles fully unrolled: the per-tile instruction count is the kFast kernel's;
//   * otherwise (span crosses a tensor boundary, holds a tensor's 16-byte tail group, or is the last CTA): per-thread
//     register cursor {end4, cls, gbase, tail group (32-bit), tail lanes}, advanced monotonically.
// -------------------------------------------------------------------------------------------
#ifndef BDL_TABLE_TPC_K
#define BDL_TABLE_TPC_K 4
#endif
#ifndef BDL_TABLE_PREFETCH
#define BDL_TABLE_PREFETCH 0    // 1: uniform path issues tile t+1's loads before tile t's arithmetic (A/B knob)
#endif
constexpr int kTableTpc = BDL_TABLE_TPC_K;   // profiles/r01_ab_table_tpc.log: 2 tiles 1.104, 4: 1.080, 8: 1.101, 16: 1.115 ms (SGHMC, ViT-L/32)

struct TableCursor {
    uint32_t idx, end4, cls;
    uint32_t tail_q;      // group holding the tensor's last real elements when numel % 4 != 0, else 0xFFFFFFFF
    uint32_t tail_n;      // number of real elements in that group (1..3)
    const float* gbase;   // gradient element for flat index i is gbase[i]
};

__device__ __forceinline__ void tcursor_load(TableCursor& c, const StepParams& p, uint32_t idx) {
    const bdl_run* r = p.runs + idx;
    c.idx = idx;
    c.end4 = static_cast<uint32_t>(__ldg(&r->end) >> 2);
    c.cls = __ldg(&r->cls);
    const float* gr = reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(&r->g_dev)));
    const uint64_t ve = __ldg(&r->valid_end);
    c.tail_n = gr ? static_cast<uint32_t>(ve) & 3u : 0u;     // the flat gradient buffer carries its own (zero) padding
    c.tail_q = c.tail_n ? static_cast<uint32_t>(ve >> 2) : 0xFFFFFFFFu;
    c.gbase = gr ? gr - __ldg(&r->begin) : p.g;
}

template <int kVariant, bool kHasBuf, bool kPhilox, int kCap>
struct TileRegs {
    float4 th, g, th0, v, m, s, b, xi, c1, c2;
};

template <int kVariant, bool kHasBuf, bool kPhilox, int kCap>
__device__ __forceinline__ void tile_load(TileRegs<kVariant, kHasBuf, kPhilox, kCap>& r, const StepParams& p, uint64_t i) {
    using U = Uses<kVariant>;
    r.th = ld_stream(p.theta + i);
    if constexpr (kCap != 0) {
        if (!p.cap_init) {
            r.c1 = ld_stream(p.cap1 + i);
            if (kCap == 2 || p.cap2) r.c2 = ld_stream(p.cap2 + i);
        }
    }
    if constexpr (U::theta0) r.th0 = ld_stream(p.theta0 + i);
    if constexpr (U::v) r.v = ld_stream(p.v + i);
    if constexpr (U::adam) {
        r.m = ld_stream(p.m + i);
        r.s = ld_stream(p.s + i);
    }
    if constexpr (kHasBuf) r.b = ld_stream(p.buf + i);
    if constexpr (!kPhilox) r.xi = ld_stream(p.xi + i);
}

template <int kVariant, bool kHasBuf, bool kPhilox, int kDiv, int kCap>
__device__ __forceinline__ void tile_update_store(TileRegs<kVariant, kHasBuf, kPhilox, kCap>& r, const StepParams& p,
                                                  uint32_t q, uint64_t i, uint32_t cls) {
    using U = Uses<kVariant>;
    if constexpr (kPhilox) r.xi = philox_normal4(p.key, q);
    update_one<kVariant, kHasBuf, kDiv>(p, cls, r.th.x, r.g.x, r.th0.x, r.v.x, r.m.x, r.s.x, r.b.x, r.xi.x);
    update_one<kVariant, kHasBuf, kDiv>(p, cls, r.th.y, r.g.y, r.th0.y, r.v.y, r.m.y, r.s.y, r.b.y, r.xi.y);
    update_one<kVariant, kHasBuf, kDiv>(p, cls, r.th.z, r.g.z, r.th0.z, r.v.z, r.m.z, r.s.z, r.b.z, r.xi.z);
    update_one<kVariant, kHasBuf, kDiv>(p, cls, r.th.w, r.g.w, r.th0.w, r.v.w, r.m.w, r.s.w, r.b.w, r.xi.w);
    if constexpr (U::v) st_stream(p.v + i, r.v);
    if constexpr (U::adam) {
        st_stream(p.m + i, r.m);
        st_stream(p.s + i, r.s);
    }
    if constexpr (kHasBuf) st_stream(p.buf + i, r.b);
    st_stream(p.theta + i, r.th);
}

#ifdef BDL_TABLE_MINB
#define BDL_TABLE_BLOCKS(...) BDL_TABLE_MINB
#else
#define BDL_TABLE_BLOCKS(...) (min_blocks<__VA_ARGS__>())
#endif
template <int kVariant, bool kHasBuf, bool kPhilox, int kDiv, int kT, int kCap = 0>
__global__ void __launch_bounds__(kT, BDL_TABLE_BLOCKS(kVariant, kHasBuf, kPhilox, 1, kT, kCap))
step_table_kernel(const StepParams p) {
    constexpr uint32_t cta_groups = kT * kTableTpc;
    const uint32_t q_cta = blockIdx.x * cta_groups;          // < n4 for every launched CTA (whole range, q_begin == 0)
    uint32_t q = q_cta + threadIdx.x;
    TileRegs<kVariant, kHasBuf, kPhilox, kCap> r;
    // tile 0: the loads that do not depend on the table go out before the search
    if (q < p.n4) tile_load(r, p, static_cast<uint64_t>(q) << 2);
    __shared__ uint32_t run0_sh;
    if (threadIdx.x < 32) {
        const uint32_t r0 = run_find_warp(p.runs, p.nruns, q_cta);
        if (threadIdx.x == 0) run0_sh = r0;
    }
    __syncthreads();
    TableCursor cur;
    tcursor_load(cur, p, run0_sh);
    const uint32_t q_last = q_cta + cta_groups;              // one past the CTA's span
    if (q_last <= p.n4 && (q_last < cur.end4 || (q_last == cur.end4 && cur.tail_n == 0))) {
        // ---- CTA-uniform: the whole span is real elements of ONE tensor ----
        const uint32_t cls = cur.cls;
        const bool live = (cls & BDL_CLS_SKIP) == 0;         // p.grad is None -> tensor left untouched (still captured)
        const float* gbase = cur.gbase;
#if BDL_TABLE_PREFETCH
        if (live) r.g = ld_stream(gbase + (static_cast<uint64_t>(q) << 2));
#pragma unroll
        for (int t = 0; t < kTableTpc; ++t, q += kT) {
            const uint64_t i = static_cast<uint64_t>(q) << 2;
            TileRegs<kVariant, kHasBuf, kPhilox, kCap> nx;
            if (t + 1 < kTableTpc) {                         // next tile's loads go out before this tile's arithmetic
                const uint64_t i2 = static_cast<uint64_t>(q + kT) << 2;
                tile_load(nx, p, i2);
                if (live) nx.g = ld_stream(gbase + i2);
            }
            if (live) tile_update_store<kVariant, kHasBuf, kPhilox, kDiv, kCap>(r, p, q, i, cls);
            if constexpr (kCap != 0) capture_fold<kCap, kDiv>(p, i, r.th, r.c1, r.c2);
            if (t + 1 < kTableTpc) r = nx;
        }
#else
#ifdef BDL_TABLE_UNROLL1
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int t = 0; t < kTableTpc; ++t, q += kT) {
            const uint64_t i = static_cast<uint64_t>(q) << 2;
            if (t > 0) tile_load(r, p, i);
            if (live) {
                r.g = ld_stream(gbase + i);
                tile_update_store<kVariant, kHasBuf, kPhilox, kDiv, kCap>(r, p, q, i, cls);
            }
            if constexpr (kCap != 0) capture_fold<kCap, kDiv>(p, i, r.th, r.c1, r.c2);
        }
#endif
        return;
    }
    // ---- general: per-thread cursor ----
#pragma unroll 1
    for (int t = 0; t < kTableTpc; ++t, q += kT) {
        if (q >= p.n4) break;
        const uint64_t i = static_cast<uint64_t>(q) << 2;
        if (t > 0) tile_load(r, p, i);
        while (q >= cur.end4 && cur.idx + 1 < p.nruns) tcursor_load(cur, p, cur.idx + 1);   // runs are sorted and contiguous
        if ((cur.cls & BDL_CLS_SKIP) == 0) {
            r.g = ld_stream(cur.gbase + i);
            if (q == cur.tail_q) {                           // tail group of a tensor: lanes past its end are padding (g = 0)
                if (cur.tail_n < 2) r.g.y = 0.f;
                if (cur.tail_n < 3) r.g.z = 0.f;
                r.g.w = 0.f;
            }
            tile_update_store<kVariant, kHasBuf, kPhilox, kDiv, kCap>(r, p, q, i, cur.cls);
        }
        if constexpr (kCap != 0) capture_fold<kCap, kDiv>(p, i, r.th, r.c1, r.c2);
    }
}

// -------------------------------------------------------------------------------------------
// The same launch with the run table in the kernel arguments (PTable): preferred whenever the host copy of the table is
// at hand and fits (every backbone of the reference).  Control flow as step_table_kernel, lookup = directory entry +
// forward scan over uniform constant loads; no shared memory, no barrier.
// -------------------------------------------------------------------------------------------
#ifndef BDL_PTABLE_TPC
#define BDL_PTABLE_TPC 1   // profiles/r02_ab_builds_ptable.log: 1 tile per CTA 1.038 / 1.773 ms (= the flat launch), 2: 1.047 / 1.796, 4: 1.056 / 1.809 (SGHMC / Adam-cSGHMC)
#endif
constexpr int kPTableTpc = BDL_PTABLE_TPC;

template <int kVariant, bool kHasBuf, bool kPhilox, int kDiv, int kT, int kCap = 0>
__global__ void __launch_bounds__(kT, BDL_TABLE_BLOCKS(kVariant, kHasBuf, kPhilox, 1, kT, kCap))
step_ptable_kernel(const __grid_constant__ StepParams p, const __grid_constant__ PTable t) {
    constexpr uint32_t cta_groups = kT * kPTableTpc;
    const uint32_t q_cta = blockIdx.x * cta_groups;          // < n4 for every launched CTA (whole range, q_begin == 0)
    uint32_t q = q_cta + threadIdx.x;
    TileRegs<kVariant, kHasBuf, kPhilox, kCap> r;
    if (q < p.n4) tile_load(r, p, static_cast<uint64_t>(q) << 2);
    uint32_t idx = t.dir[q_cta >> t.dir_shift];
    while (q_cta >= t.end4[idx]) ++idx;                      // the last run ends at n4 > q_cta
    const uint32_t end4 = t.end4[idx];
    const uint32_t q_last = q_cta + cta_groups;
    if (q_last <= p.n4 && (q_last < end4 || (q_last == end4 && t.tail_q[idx] == 0xFFFFFFFFu))) {
        // ---- CTA-uniform: the whole span is real elements of ONE tensor ----
        const uint32_t cls = t.cls[idx];
        const bool live = (cls & BDL_CLS_SKIP) == 0;
        const float* gbase = reinterpret_cast<const float*>(t.gbase[idx]);
#pragma unroll
        for (int k = 0; k < kPTableTpc; ++k, q += kT) {
            const uint64_t i = static_cast<uint64_t>(q) << 2;
            if (k > 0) tile_load(r, p, i);
            if (live) {
                r.g = ld_stream(gbase + i);
                tile_update_store<kVariant, kHasBuf, kPhilox, kDiv, kCap>(r, p, q, i, cls);
            }
            if constexpr (kCap != 0) capture_fold<kCap, kDiv>(p, i, r.th, r.c1, r.c2);
        }
        return;
    }
    // ---- general: per-thread row index, advanced monotonically ----
#pragma unroll 1
    for (int k = 0; k < kPTableTpc; ++k, q += kT) {
        if (q >= p.n4) break;
        const uint64_t i = static_cast<uint64_t>(q) << 2;
        if (k > 0) tile_load(r, p, i);
        while (q >= t.end4[idx] && idx + 1 < t.nrows) ++idx;
        const uint32_t cls = t.cls[idx];
        if ((cls & BDL_CLS_SKIP) == 0) {
            r.g = ld_stream(reinterpret_cast<const float*>(t.gbase[idx]) + i);
            if (q == t.tail_q[idx]) {                        // tail group of a tensor: lanes past its end are padding (g = 0)
                const uint32_t tn = t.tail_n[idx];
                if (tn < 2) r.g.y = 0.f;
                if (tn < 3) r.g.z = 0.f;
                r.g.w = 0.f;
            }
            tile_update_store<kVariant, kHasBuf, kPhilox, kDiv, kCap>(r, p, q, i, cls);
        }
        if constexpr (kCap != 0) capture_fold<kCap, kDiv>(p, i, r.th, r.c1, r.c2);
    }
}

// Host: PTable from the host copy of a run table; false when it does not fit (then the device-table kernel runs).
inline bool build_ptable(PTable& t, const bdl_run* rows, uint32_t nruns, const float* g_flat, uint64_t n) {
    if (nruns > static_cast<uint32_t>(kPRows) || n >= 0xFFFFFFFFull) return false;
    const uint32_t n4 = static_cast<uint32_t>(n >> 2);
    uint32_t shift = 10;
    while (((n4 - 1) >> shift) >= static_cast<uint32_t>(kPDir)) ++shift;
    t.nrows = nruns;
    t.dir_shift = shift;
    for (uint32_t r = 0; r < nruns; ++r) {
        t.end4[r] = static_cast<uint32_t>(rows[r].end >> 2);
        t.valid_end[r] = static_cast<uint32_t>(rows[r].valid_end);
        t.cls[r] = static_cast<uint8_t>(rows[r].cls);
        const bool own = rows[r].g_dev != nullptr;
        const uint32_t tn = own ? static_cast<uint32_t>(rows[r].valid_end & 3u) : 0u;   // the flat buffer carries its own (zero) padding
        t.tail_n[r] = static_cast<uint8_t>(tn);
        t.tail_q[r] = tn ? static_cast<uint32_t>(rows[r].valid_end >> 2) : 0xFFFFFFFFu;
        t.gbase[r] = own ? reinterpret_cast<uint64_t>(rows[r].g_dev) - 4ull * rows[r].begin : reinterpret_cast<uint64_t>(g_flat);
    }
    if (t.end4[nruns - 1] != n4) return false;               // rows must cover [0, n)
    uint32_t r = 0;
    const uint32_t ndir = ((n4 - 1) >> shift) + 1;
    for (uint32_t j = 0; j < ndir; ++j) {
        const uint32_t q = j << shift;
        while (q >= t.end4[r]) ++r;
        t.dir[j] = static_cast<uint16_t>(r);
    }
    return true;
}

// -------------------------------------------------------------------------------------------
// Gradient-norm clipping (args.clip_grad; methods/csgld.py:250-251, methods/adam_csghmc.py:319-320): the reference calls
// torch.nn.utils.clip_grad_norm_ on p.grad = g' (SGLD family) / v (Adam-cSGHMC) between Model.forward and
// optimizer.step().  Fused here as TWO passes over the state with the table in the kernel arguments:
//   pass 1 (kPass == 1)  recompute that quantity (counter-based noise: the same draw both times), sum of squares over the
//                        real elements of every tensor with a gradient -> *sumsq (fp64 atomics, one per warp);
//   pass 2 (kPass == 2)  the ordinary update with the quantity scaled by *coef (bdl_clip_coef, device scalar: no host sync).
// Rare path (no driver of the reference sets args.clip_grad): one straightforward kernel, per-thread row lookup.
// -------------------------------------------------------------------------------------------
constexpr int kClipThreads = 128;

template <int kVariant, bool kHasBuf, bool kPhilox, int kDiv, int kPass>
__global__ void __launch_bounds__(kClipThreads)
step_clip_kernel(const __grid_constant__ StepParams p, const __grid_constant__ PTable t, double* __restrict__ sumsq,
                 const float* __restrict__ coef_dev) {
    const uint32_t q = blockIdx.x * kClipThreads + threadIdx.x;
    double acc = 0.0;
    if (q < p.n4) {
        const uint64_t i = static_cast<uint64_t>(q) << 2;
        TileRegs<kVariant, kHasBuf, kPhilox, 0> r;
        tile_load(r, p, i);
        uint32_t idx = t.dir[q >> t.dir_shift];
        while (q >= t.end4[idx]) ++idx;
        const uint32_t cls = t.cls[idx];
        if ((cls & BDL_CLS_SKIP) == 0) {                         // p.grad is None: not part of the norm, not updated
            r.g = ld_stream(reinterpret_cast<const float*>(t.gbase[idx]) + i);
            if (q == t.tail_q[idx]) {
                const uint32_t tn = t.tail_n[idx];
                if (tn < 2) r.g.y = 0.f;
                if (tn < 3) r.g.z = 0.f;
                r.g.w = 0.f;
            }
            if constexpr (kPhilox) r.xi = philox_normal4(p.key, q);
            if constexpr (kPass == 1) {
                float qv[4];
                update_one<kVariant, kHasBuf, kDiv, 1>(p, cls, r.th.x, r.g.x, r.th0.x, r.v.x, r.m.x, r.s.x, r.b.x, r.xi.x, 1.0f, &qv[0]);
                update_one<kVariant, kHasBuf, kDiv, 1>(p, cls, r.th.y, r.g.y, r.th0.y, r.v.y, r.m.y, r.s.y, r.b.y, r.xi.y, 1.0f, &qv[1]);
                update_one<kVariant, kHasBuf, kDiv, 1>(p, cls, r.th.z, r.g.z, r.th0.z, r.v.z, r.m.z, r.s.z, r.b.z, r.xi.z, 1.0f, &qv[2]);
                update_one<kVariant, kHasBuf, kDiv, 1>(p, cls, r.th.w, r.g.w, r.th0.w, r.v.w, r.m.w, r.s.w, r.b.w, r.xi.w, 1.0f, &qv[3]);
                const uint32_t valid = t.valid_end[idx];
                const uint32_t e0 = q << 2;                      // n < 2^32 (build_ptable)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (e0 + k < valid) acc += static_cast<double>(qv[k]) * static_cast<double>(qv[k]);
            } else {
                using U = Uses<kVariant>;
                const float coef = __ldg(coef_dev);
                update_one<kVariant, kHasBuf, kDiv, 2>(p, cls, r.th.x, r.g.x, r.th0.x, r.v.x, r.m.x, r.s.x, r.b.x, r.xi.x, coef);
                update_one<kVariant, kHasBuf, kDiv, 2>(p, cls, r.th.y, r.g.y, r.th0.y, r.v.y, r.m.y, r.s.y, r.b.y, r.xi.y, coef);
                update_one<kVariant, kHasBuf, kDiv, 2>(p, cls, r.th.z, r.g.z, r.th0.z, r.v.z, r.m.z, r.s.z, r.b.z, r.xi.z, coef);
                update_one<kVariant, kHasBuf, kDiv, 2>(p, cls, r.th.w, r.g.w, r.th0.w, r.v.w, r.m.w, r.s.w, r.b.w, r.xi.w, coef);
                if constexpr (U::v) st_stream(p.v + i, r.v);
                if constexpr (U::adam) {
                    st_stream(p.m + i, r.m);
                    st_stream(p.s + i, r.s);
                }
                if constexpr (kHasBuf) st_stream(p.buf + i, r.b);
                st_stream(p.theta + i, r.th);
            }
        }
    }
    if constexpr (kPass == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(sumsq, acc);
    }
}

// total_norm = sqrt(sum of squares) in fp32, clip_coef = max_norm / (total_norm + 1e-6) -- which torch evaluates as
// (total_norm + 1e-6).reciprocal() * max_norm (Tensor.__rtruediv__) --, clamped to <= 1: the statements of
// torch.nn.utils.clip_grad_norm_ (fp32 tensors; the only difference is the summation order of the squares).
static __global__ void clip_coef_kernel(const double* __restrict__ sumsq, float max_norm, float* __restrict__ coef, float* __restrict__ total_norm) {
    const float tn = static_cast<float>(sqrt(*sumsq));
    const float c = __fmul_rn(__frcp_rn(__fadd_rn(tn, 1e-6f)), max_norm);    // float / Tensor = Tensor.reciprocal() * float (two roundings)
    *coef = fminf(c, 1.0f);
    if (total_norm) *total_norm = tn;
}

template <int kVariant, bool kHasBuf, bool kPhilox, int kDiv>
static int launch_clip(const StepParams& p, int pass, double* sumsq, const float* coef, cudaStream_t st) {
    const uint32_t grid = (p.n4 + kClipThreads - 1) / kClipThreads;
    if (pass == 1)
        step_clip_kernel<kVariant, kHasBuf, kPhilox, kDiv, 1><<<grid, kClipThreads, 0, st>>>(p, *p.ptab, sumsq, coef);
    else
        step_clip_kernel<kVariant, kHasBuf, kPhilox, kDiv, 2><<<grid, kClipThreads, 0, st>>>(p, *p.ptab, sumsq, coef);
    return check_cuda(cudaGetLastError(), "step_clip_kernel launch");
}

template <int kVariant, bool kHasBuf>
int launch_clip_nd(const StepParams& p, bool philox, int div, int pass, double* sumsq, const float* coef, cudaStream_t st) {
    if (philox)
        return div == BDL_DIV_IEEE ? launch_clip<kVariant, kHasBuf, true, BDL_DIV_IEEE>(p, pass, sumsq, coef, st)
                                   : launch_clip<kVariant, kHasBuf, true, BDL_DIV_RECIP>(p, pass, sumsq, coef, st);
    return div == BDL_DIV_IEEE ? launch_clip<kVariant, kHasBuf, false, BDL_DIV_IEEE>(p, pass, sumsq, coef, st)
                               : launch_clip<kVariant, kHasBuf, false, BDL_DIV_RECIP>(p, pass, sumsq, coef, st);
}

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
// Launch-shape overrides (bdl_set_launch_config): per calling thread, so a sweep on one thread / device never changes
// what another thread launches.
extern thread_local int g_ctas_per_sm;   // 0 = one tile per CTA (default); > 0 = persistent grid of #SM * ctas_per_sm CTAs
extern thread_local int g_unroll;        // 0 = default
extern thread_local int g_threads;       // 0 = default        (defined in bdl_step.cu)

// Tiles per CTA for launches whose run table carries gradient pointers (experiment knob: BDL_TABLE_TPC, read once).
inline uint32_t table_tiles_per_cta() {
    static const uint32_t v = [] {
        const char* e = getenv("BDL_TABLE_TPC");
        const long x = e ? atol(e) : kDefaultTableTpc;
        return static_cast<uint32_t>(x < 1 ? 1 : (x > 4096 ? 4096 : x));
    }();
    return v;
}

template <int kVariant, bool kHasBuf, bool kPhilox, int kDiv, int kU, int kT, int kCap = 0, bool kAllowFast = false>
static int launch_shape(const StepParams& p, cudaStream_t st) {
    constexpr uint32_t tile_groups = kT * kU;
    const uint32_t ntiles = (p.n4 - p.q_begin + tile_groups - 1) / tile_groups;
    uint64_t grid = ntiles;
    if (g_ctas_per_sm > 0) {
        const uint64_t cap = static_cast<uint64_t>(num_sms()) * g_ctas_per_sm;
        if (grid > cap) grid = cap;
    }
    if (grid > 0x7FFFFFFFull) grid = 0x7FFFFFFFull;
    if (grid == 0) return BDL_OK;
    if constexpr (kAllowFast && kU == 1) {
        if (p.ptab != nullptr && grid == ntiles) {               // per-tensor gradient pointers, table in the kernel arguments
            constexpr uint32_t span = kT * kPTableTpc;
            grid = (static_cast<uint64_t>(p.n4) + span - 1) / span;
            step_ptable_kernel<kVariant, kHasBuf, kPhilox, kDiv, kT, kCap><<<static_cast<uint32_t>(grid), kT, 0, st>>>(p, *p.ptab);
            return check_cuda(cudaGetLastError(), "step_ptable_kernel launch");
        }
    }
    BDL_REQUIRE(p.runs != nullptr || p.inl_n != 0, BDL_ERR_INVALID, "bdl_step: this launch shape needs the run table in device memory");
    if (grid == ntiles && p.inl_n == 0 && !p.flat_g && table_tiles_per_cta() > 1) {
        // run table with per-tensor gradient pointers (the training-loop launch): the gradient load depends on the table
        // lookup.  A CTA that walks a few consecutive tiles pays the search and that late first load once.  Tables
        // without gradient pointers (bias=uninformative) lose nothing to the lookup and stay at one tile per CTA
        // (1.044 vs 1.053 ms).
        if constexpr (kAllowFast && kU == 1) {
            if (table_tiles_per_cta() == static_cast<uint32_t>(kTableTpc) && p.q_begin == 0) {   // the lean build
#ifdef BDL_TABLE_T
                constexpr int kTT = BDL_TABLE_T;                 // A/B knob: CTA size of the table kernel
#else
                constexpr int kTT = kT;
#endif
                constexpr uint32_t span = kTT * kTableTpc;
                grid = (static_cast<uint64_t>(p.n4) + span - 1) / span;
                step_table_kernel<kVariant, kHasBuf, kPhilox, kDiv, kTT, kCap><<<static_cast<uint32_t>(grid), kTT, 0, st>>>(p);
                return check_cuda(cudaGetLastError(), "step_table_kernel launch");
            }
        }
        StepParams pc = p;
        pc.tpc = table_tiles_per_cta();
        grid = (ntiles + pc.tpc - 1) / pc.tpc;
        step_kernel<kVariant, kHasBuf, kPhilox, kDiv, kU, kT, kCap><<<static_cast<uint32_t>(grid), kT, 0, st>>>(pc);
        return check_cuda(cudaGetLastError(), "step_kernel launch");
    }
    if constexpr (kAllowFast) {
        if (grid == ntiles && p.inl_n == 2 && p.flat_g) {
            step_kernel<kVariant, kHasBuf, kPhilox, kDiv, kU, kT, kCap, true><<<static_cast<uint32_t>(grid), kT, 0, st>>>(p);
            return check_cuda(cudaGetLastError(), "step_kernel launch");
        }
    }
    step_kernel<kVariant, kHasBuf, kPhilox, kDiv, kU, kT, kCap><<<static_cast<uint32_t>(grid), kT, 0, st>>>(p);
    return check_cuda(cudaGetLastError(), "step_kernel launch");
}

template <int kVariant, bool kHasBuf, bool kPhilox, int kDiv>
static int launch_u(const StepParams& p, cudaStream_t st) {
    // Default block size, from the ViT-L/32 sweep (profiles/r01_sweep_final.log): the more streams a variant moves per
    // element the smaller the CTA that keeps the in-order streaming window tight; the 16 B/param SGLD (mu = 0) kernel is
    // issue-bound and prefers fewer, larger CTAs.
    constexpr int kStreams = ((kVariant == BDL_SGLD) ? 3 : (kVariant == BDL_SGHMC) ? 4 : (kVariant == BDL_CSGHMC) ? 3 : 6) +
                             (kHasBuf ? 1 : 0);
    constexpr int kAutoThreads = kStreams >= 4 ? 64 : (kVariant == BDL_SGLD ? 256 : 128);
    if (p.cap1) {
        // fused capture: two more streams per element -> always >= 5, i.e. the 64-thread shape; the launch-shape knobs
        // of bdl_set_launch_config do not apply (one instantiation per variant keeps the binary small)
        return p.cap_kind == BDL_CAPTURE_WELFORD ? launch_shape<kVariant, kHasBuf, kPhilox, kDiv, 1, 64, 2, true>(p, st)
                               : launch_shape<kVariant, kHasBuf, kPhilox, kDiv, 1, 64, 1, true>(p, st);
    }
    const int unroll = g_unroll ? g_unroll : kDefaultUnroll;
    const int threads = g_threads ? g_threads : kAutoThreads;
    if (g_unroll == 0 && g_threads == 0)       // library defaults: this shape also has the fast-path build (an explicit
        return launch_shape<kVariant, kHasBuf, kPhilox, kDiv, kDefaultUnroll, kAutoThreads, 0, true>(p, st);   // shape request runs the generic one)
#define BDL_SHAPE(UU, TT) if (unroll == UU && threads == TT) return launch_shape<kVariant, kHasBuf, kPhilox, kDiv, UU, TT>(p, st)
    BDL_SHAPE(1, 64);
#ifndef BDL_AB_SLIM
    BDL_SHAPE(1, 128); BDL_SHAPE(1, 256); BDL_SHAPE(1, 512);
    BDL_SHAPE(2, 64); BDL_SHAPE(2, 128); BDL_SHAPE(2, 256); BDL_SHAPE(2, 512);
#endif
#undef BDL_SHAPE
    set_error("bdl_step: unsupported launch shape unroll=%d threads=%d (unroll 1|2, threads 64|128|256|512)", unroll, threads);
    return BDL_ERR_INVALID;
}

template <int kVariant, bool kHasBuf>
int launch_nd(const StepParams& p, bool philox, int div, cudaStream_t st) {
    if (philox) {
        return div == BDL_DIV_IEEE ? launch_u<kVariant, kHasBuf, true, BDL_DIV_IEEE>(p, st)
                                   : launch_u<kVariant, kHasBuf, true, BDL_DIV_RECIP>(p, st);
    }
    return div == BDL_DIV_IEEE ? launch_u<kVariant, kHasBuf, false, BDL_DIV_IEEE>(p, st)
                               : launch_u<kVariant, kHasBuf, false, BDL_DIV_RECIP>(p, st);
}

// One translation unit per (variant, SGD-momentum-buffer) pair instantiates the kernels of that pair (bdl_step_inst_*.cu,
// compiled in parallel by build.py); bdl_step.cu only dispatches.
#define BDL_STEP_INSTANCES(X) \
    X(BDL_SGLD, false) X(BDL_SGLD, true) X(BDL_SGHMC, false) X(BDL_CSGHMC, false) X(BDL_ADAM_SGHMC, false) \
    X(BDL_ADAM_SGHMC, true) X(BDL_ADAM_CSGHMC, false)
#define BDL_CLIP_INSTANCES(X) X(BDL_SGLD, false) X(BDL_SGLD, true) X(BDL_ADAM_CSGHMC, false)
#ifndef BDL_STEP_INSTANTIATE
#define BDL_DECL_STEP(V, B) extern template int launch_nd<V, B>(const StepParams&, bool, int, cudaStream_t);
#define BDL_DECL_CLIP(V, B) extern template int launch_clip_nd<V, B>(const StepParams&, bool, int, int, double*, const float*, cudaStream_t);
BDL_STEP_INSTANCES(BDL_DECL_STEP)
BDL_CLIP_INSTANCES(BDL_DECL_CLIP)
#undef BDL_DECL_STEP
#undef BDL_DECL_CLIP
#endif

}  // namespace bdl
