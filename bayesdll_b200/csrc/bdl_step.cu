// bdl_step.cu -- C entry points of the fused SG-MCMC sampler update: argument
// validation, scalar preparation and the dispatch to the per-variant instantiation units (bdl_step_inst_*.cu).  The
// kernels themselves live in bdl_step.cuh.
#include "bdl_step.cuh"

namespace bdl {
// Launch-shape overrides (bdl_set_launch_config): per calling thread, shared by the instantiation units
thread_local int g_ctas_per_sm = 0;
thread_local int g_unroll = 0;
thread_local int g_threads = 0;
}  // namespace bdl

extern "C" int bdl_set_launch_config(int ctas_per_sm, int unroll, int threads) {
    using namespace bdl;
    BDL_REQUIRE(ctas_per_sm >= 0 && ctas_per_sm <= 65536, BDL_ERR_INVALID, "ctas_per_sm out of range");
    BDL_REQUIRE(unroll == 0 || unroll == 1 || unroll == 2, BDL_ERR_INVALID, "unroll must be 0, 1 or 2");
    BDL_REQUIRE(threads == 0 || threads == 64 || threads == 128 || threads == 256 || threads == 512, BDL_ERR_INVALID,
                "threads must be 0, 64, 128, 256 or 512");
    g_ctas_per_sm = ctas_per_sm;
    g_unroll = unroll;
    g_threads = threads;
    return BDL_OK;
}

namespace bdl {

// Update the float4 groups [q_begin, q_end) of the flat state; all pointers are the bases of the full buffers,
// so Philox counters and the run table keep their absolute indexing (results do not depend on chunking).
int step_range(int variant, float* theta, const float* g, const float* theta0, float* v, float* m, float* s, float* buf,
               uint64_t n, uint64_t q_begin, uint64_t q_end, const bdl_run* runs, uint32_t nruns, const bdl_run* runs_host,
               const bdl_scalars* sc, const bdl_noise* nz, const bdl_capture* cap, cudaStream_t st, int clip_pass,
               double* clip_sumsq, const float* clip_coef) {
    BDL_REQUIRE(variant >= BDL_SGLD && variant <= BDL_ADAM_CSGHMC, BDL_ERR_INVALID, "bdl_step: unknown variant %d", variant);
    if (n == 0 || q_begin >= q_end) return BDL_OK;   // empty state / empty range: nothing to do (pointers may be null)
    BDL_REQUIRE(theta && (runs || runs_host) && sc && nz, BDL_ERR_INVALID, "bdl_step: null theta/runs/scalars/noise");
    BDL_REQUIRE(runs || runs_host, BDL_ERR_INVALID, "bdl_step: run table required");
    BDL_REQUIRE(n % 4 == 0, BDL_ERR_INVALID, "bdl_step: n=%llu is not a multiple of 4", (unsigned long long)n);
    BDL_REQUIRE((n >> 2) < 0xFFFFFFFFull, BDL_ERR_INVALID, "bdl_step: n too large for 32-bit group index");
    BDL_REQUIRE(q_end <= (n >> 2), BDL_ERR_INVALID, "bdl_step: range end beyond n");
    BDL_REQUIRE(nruns >= 1 && nruns <= BDL_MAX_RUNS, BDL_ERR_INVALID, "bdl_step: nruns=%u out of range", nruns);
    const bool adam = variant == BDL_ADAM_SGHMC || variant == BDL_ADAM_CSGHMC;
    const bool has_buf = (variant == BDL_SGLD || variant == BDL_ADAM_SGHMC) && sc->mu != 0.0f;
    BDL_REQUIRE(variant == BDL_CSGHMC || theta0, BDL_ERR_INVALID, "bdl_step: theta0 required");
    BDL_REQUIRE(variant == BDL_SGLD || v, BDL_ERR_INVALID, "bdl_step: momentum buffer v required");
    BDL_REQUIRE(!adam || (m && s), BDL_ERR_INVALID, "bdl_step: Adam moments m,s required");
    BDL_REQUIRE(!has_buf || buf, BDL_ERR_INVALID, "bdl_step: SGD momentum buffer required when mu != 0");
    BDL_REQUIRE(sc->div_mode == BDL_DIV_IEEE || sc->div_mode == BDL_DIV_RECIP, BDL_ERR_INVALID, "bdl_step: bad div_mode");
    const bool capture = cap != nullptr && cap->kind != BDL_CAPTURE_NONE;
    if (capture) {
        BDL_REQUIRE(cap->kind == BDL_CAPTURE_AVG || cap->kind == BDL_CAPTURE_WELFORD, BDL_ERR_INVALID,
                    "bdl_step_capture: unknown capture kind %d", cap->kind);
        BDL_REQUIRE(cap->first_dev, BDL_ERR_INVALID, "bdl_step_capture: first-moment buffer required");
        BDL_REQUIRE(cap->kind == BDL_CAPTURE_AVG || cap->second_dev, BDL_ERR_INVALID, "bdl_step_capture: Welford needs the M2 buffer");
        BDL_REQUIRE(q_begin == 0 && q_end == (n >> 2), BDL_ERR_UNSUPPORTED, "bdl_step_capture: capture needs the whole range");
    }
    const void* ptrs[] = {theta, g, theta0, v, m, s, buf, nz->xi_dev, capture ? cap->first_dev : nullptr,
                          capture ? cap->second_dev : nullptr};
    for (const void* q : ptrs) BDL_REQUIRE(aligned16(q), BDL_ERR_ALIGN, "bdl_step: pointer %p is not 16-byte aligned", q);

    StepParams p{};
    p.theta = theta; p.g = g; p.theta0 = theta0; p.v = v; p.m = m; p.s = s; p.buf = buf;
    p.xi = nz->xi_dev; p.runs = runs; p.nruns = nruns;
    p.inl_n = 0;
    p.flat_g = 0;
    if (runs_host && g) {
        bool plain = true;
        for (uint32_t r = 0; r < nruns; ++r) plain = plain && runs_host[r].g_dev == nullptr;
        p.flat_g = plain ? 1u : 0u;
        if (plain && nruns <= static_cast<uint32_t>(kInlineRuns)) {
            p.inl_n = nruns;
            for (uint32_t r = 0; r < nruns; ++r) {
                p.inl_end4[r] = static_cast<uint32_t>(runs_host[r].end >> 2);
                p.inl_cls[r] = runs_host[r].cls;
            }
        }
    }
    // per-tensor gradient pointers + host copy of the table: carry the table in the kernel arguments (step_ptable_kernel)
    static thread_local PTable ptab_storage;
    p.ptab = nullptr;
#ifndef BDL_NO_PTABLE
    if (runs_host && !p.flat_g && q_begin == 0 && q_end == (n >> 2) && table_tiles_per_cta() == static_cast<uint32_t>(kDefaultTableTpc) &&
        build_ptable(ptab_storage, runs_host, nruns, g, n))
        p.ptab = &ptab_storage;
#endif
    BDL_REQUIRE(runs || p.ptab || p.inl_n, BDL_ERR_INVALID,
                "bdl_step: a device run table is required (more than %d runs without a host copy that fits the kernel arguments)", kInlineRuns);
    p.n4 = static_cast<uint32_t>(q_end); p.q_begin = static_cast<uint32_t>(q_begin);
    for (int h = 0; h < 2; ++h) {
        p.lr[h] = sc->lr[h];
        p.neg_lr[h] = -sc->lr[h];
        p.c[h] = sc->noise_scale[h];
    }
    // Reciprocals.  IEEE mode needs RN(1 / fp32(s)) for the FMA-corrected division; reciprocal mode multiplies by the
    // factor torch CUDA uses, fp32(1.0 / s_double), supplied by the host (0 -> not supplied -> fp32 reciprocal).
    const bool recip = sc->div_mode == BDL_DIV_RECIP;
    auto inv_of = [recip](float s_f, float host_inv) { return (recip && host_inv != 0.0f) ? host_inv : 1.0f / s_f; };
    p.oma = sc->one_minus_alpha;
    p.sig2 = sc->sig2; p.inv_sig2 = inv_of(sc->sig2, sc->inv_sig2);
    p.N = sc->N; p.inv_N = inv_of(sc->N, sc->inv_N);
    p.mu = sc->mu;
    p.b1 = sc->beta1; p.omb1 = sc->one_minus_beta1; p.b2 = sc->beta2; p.omb2 = sc->one_minus_beta2;
    p.bc1 = sc->bias_corr1; p.inv_bc1 = inv_of(sc->bias_corr1, sc->inv_bias_corr1);
    p.bc2 = sc->bias_corr2; p.inv_bc2 = inv_of(sc->bias_corr2, sc->inv_bias_corr2);
    p.eps = sc->eps; p.two_alpha = sc->two_alpha; p.nd = sc->nd;
    p.T = sc->temperature; p.inv_T = inv_of(sc->temperature, sc->inv_temperature);
    p.first_step = sc->first_step; p.add_noise = sc->add_noise;
    if (capture) {
        p.cap1 = cap->first_dev; p.cap2 = cap->second_dev; p.cap_kind = cap->kind; p.cap_init = cap->init != 0;
        p.cap_a = cap->cnt;
        p.cap_b = cap->cnt_plus_1;
        p.cap_inv = scalar_reciprocal(cap->kind == BDL_CAPTURE_AVG ? cap->cnt_plus_1 : cap->cnt, sc->div_mode);
    }
    p.key = host_noise_key(nz->seed, nz->stream_id, nz->subseq);

    const bool philox = nz->xi_dev == nullptr;
    const int d = sc->div_mode;
#ifndef BDL_AB_SLIM
    if (clip_pass != 0) {
        BDL_REQUIRE(clip_pass == 1 ? clip_sumsq != nullptr : clip_coef != nullptr, BDL_ERR_INVALID, "bdl_step_gradnorm / bdl_step_clipped: null scalar");
        BDL_REQUIRE(variant == BDL_SGLD || variant == BDL_ADAM_CSGHMC, BDL_ERR_UNSUPPORTED,
                    "gradient-norm clipping exists for the SGLD family and Adam-cSGHMC (the runners that consult args.clip_grad "
                    "and step on p.grad: methods/csgld.py:250, methods/adam_csghmc.py:319); variant %d", variant);
        BDL_REQUIRE(!capture, BDL_ERR_UNSUPPORTED, "clipping and fused capture cannot be combined: capture with bdl_moments_* after the step");
        if (p.ptab == nullptr) {                                 // also tables WITHOUT gradient pointers: per-tensor valid ranges are needed
            BDL_REQUIRE(runs_host && q_begin == 0 && q_end == (n >> 2) && build_ptable(ptab_storage, runs_host, nruns, g, n),
                        BDL_ERR_UNSUPPORTED, "clipping needs the HOST copy of a per-tensor run table of <= %d rows covering [0, n)", kPRows);
            p.ptab = &ptab_storage;
        }
        if (variant == BDL_SGLD)
            return has_buf ? launch_clip_nd<BDL_SGLD, true>(p, philox, d, clip_pass, clip_sumsq, clip_coef, st)
                           : launch_clip_nd<BDL_SGLD, false>(p, philox, d, clip_pass, clip_sumsq, clip_coef, st);
        return launch_clip_nd<BDL_ADAM_CSGHMC, false>(p, philox, d, clip_pass, clip_sumsq, clip_coef, st);
    }
#endif
#ifdef BDL_AB_SLIM
    // A/B builds (tools/ab_builds.py): only what the sweep launches, so that a build takes seconds
    BDL_REQUIRE(philox && d == BDL_DIV_RECIP && !has_buf && (variant == BDL_SGHMC || variant == BDL_ADAM_CSGHMC),
                BDL_ERR_UNSUPPORTED, "slim A/B build: SGHMC / Adam-cSGHMC with Philox noise and reciprocal division only");
    return variant == BDL_SGHMC ? launch_u<BDL_SGHMC, false, true, BDL_DIV_RECIP>(p, st)
                                : launch_u<BDL_ADAM_CSGHMC, false, true, BDL_DIV_RECIP>(p, st);
#else
    switch (variant) {
        case BDL_SGLD:
            return has_buf ? launch_nd<BDL_SGLD, true>(p, philox, d, st) : launch_nd<BDL_SGLD, false>(p, philox, d, st);
        case BDL_SGHMC:
            return launch_nd<BDL_SGHMC, false>(p, philox, d, st);
        case BDL_CSGHMC:
            return launch_nd<BDL_CSGHMC, false>(p, philox, d, st);
        case BDL_ADAM_SGHMC:
            return has_buf ? launch_nd<BDL_ADAM_SGHMC, true>(p, philox, d, st)
                           : launch_nd<BDL_ADAM_SGHMC, false>(p, philox, d, st);
        default:
            return launch_nd<BDL_ADAM_CSGHMC, false>(p, philox, d, st);
    }
#endif
}

}  // namespace bdl

extern "C" int bdl_step(int variant, float* theta, const float* g, const float* theta0, float* v, float* m,
                        float* s, float* buf, uint64_t n, const bdl_run* runs, uint32_t nruns, const bdl_run* runs_host,
                        const bdl_scalars* sc, const bdl_noise* nz, void* stream) {
    return bdl::step_range(variant, theta, g, theta0, v, m, s, buf, n, 0, n >> 2, runs, nruns, runs_host, sc, nz, nullptr,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int bdl_step_capture(int variant, float* theta, const float* g, const float* theta0, float* v, float* m,
                                float* s, float* buf, uint64_t n, const bdl_run* runs, uint32_t nruns,
                                const bdl_run* runs_host, const bdl_scalars* sc, const bdl_noise* nz,
                                const bdl_capture* capture, void* stream) {
    return bdl::step_range(variant, theta, g, theta0, v, m, s, buf, n, 0, n >> 2, runs, nruns, runs_host, sc, nz, capture,
                           static_cast<cudaStream_t>(stream));
}

/* Gradient-norm clipping (see step_clip_kernel). */
extern "C" int bdl_step_gradnorm(int variant, const float* theta, const float* g, const float* theta0, const float* v,
                                 const float* m, const float* s, const float* buf, uint64_t n, const bdl_run* runs_host,
                                 uint32_t nruns, const bdl_scalars* sc, const bdl_noise* nz, double* sumsq_dev, void* stream) {
    // pass 1 stores nothing: the state pointers are only read
    return bdl::step_range(variant, const_cast<float*>(theta), g, theta0, const_cast<float*>(v), const_cast<float*>(m),
                           const_cast<float*>(s), const_cast<float*>(buf), n, 0, n >> 2, nullptr, nruns, runs_host, sc, nz, nullptr,
                           static_cast<cudaStream_t>(stream), 1, sumsq_dev, nullptr);
}

extern "C" int bdl_clip_coef(const double* sumsq_dev, float max_norm, float* coef_dev, float* total_norm_dev, void* stream) {
    using namespace bdl;
    BDL_REQUIRE(sumsq_dev && coef_dev, BDL_ERR_INVALID, "bdl_clip_coef: null argument");
    clip_coef_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(sumsq_dev, max_norm, coef_dev, total_norm_dev);
    return check_cuda(cudaGetLastError(), "clip_coef_kernel launch");
}

extern "C" int bdl_step_clipped(int variant, float* theta, const float* g, const float* theta0, float* v, float* m, float* s,
                                float* buf, uint64_t n, const bdl_run* runs_host, uint32_t nruns, const bdl_scalars* sc,
                                const bdl_noise* nz, const float* coef_dev, void* stream) {
    return bdl::step_range(variant, theta, g, theta0, v, m, s, buf, n, 0, n >> 2, nullptr, nruns, runs_host, sc, nz, nullptr,
                           static_cast<cudaStream_t>(stream), 2, nullptr, coef_dev);
}
