// bdl_host.cu -- host-buffer entry points: a chain whose sampler state lives in HBM, stepped with the
// gradient arriving in HOST memory and the new parameters returned to HOST memory.
//
// This is the form a CPU-resident caller binds (the reference keeps every tensor of the update in host RAM when
// args.device is the CPU): theta0 / v / m / s / momentum stay resident on the device across steps -- only what
// changes hands every step crosses PCIe: the gradient in (4 B/param) and theta out (4 B/param).  The step is
// pipelined in chunks over three streams so the H2D copy of chunk c+1, the fused update of chunk c and the D2H
// copy of chunk c-1 overlap; Philox counters and run tables keep absolute indexing, so the result is independent
// of the chunking.
#include <vector>

#include "bdl_common.cuh"

struct bdl_chain {
    uint64_t n = 0;
    int variant = 0;
    float* buf[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // BDL_BUF_* + gradient
    bdl_run* runs_dev = nullptr;
    cudaStream_t s_h2d = nullptr, s_cmp = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev_h2d, ev_cmp;
    uint64_t chunk = 0;   // elements per steady-state chunk (multiple of 4)
    std::vector<uint64_t> bounds;   // chunk boundaries (elements): ramp-up / ramp-down chunks at both ends shorten the
                                    // pipeline's fill (first H2D before any compute) and drain (last D2H after all compute)
};

namespace {
constexpr int kGrad = 6;
constexpr uint64_t kDefaultChunk = 8ull << 20;    // 8 Mi elements = 32 MiB per direction per chunk (best of the sweep in profiles/r01_e2e_chunks.log)

bool needs(int variant, int which, bool has_mu) {
    switch (which) {
        case BDL_BUF_THETA: return true;
        case BDL_BUF_THETA0: return variant != BDL_CSGHMC;
        case BDL_BUF_V: return variant != BDL_SGLD;
        case BDL_BUF_M:
        case BDL_BUF_S: return variant == BDL_ADAM_SGHMC || variant == BDL_ADAM_CSGHMC;
        case BDL_BUF_SGD: return has_mu && (variant == BDL_SGLD || variant == BDL_ADAM_SGHMC);
        default: return true;
    }
}
}  // namespace

extern "C" int bdl_chain_destroy(bdl_chain* c) {
    if (!c) return BDL_OK;
    for (float*& p : c->buf) if (p) cudaFree(p);
    if (c->runs_dev) cudaFree(c->runs_dev);
    for (cudaEvent_t e : c->ev_h2d) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_cmp) cudaEventDestroy(e);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_cmp) cudaStreamDestroy(c->s_cmp);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    delete c;
    return BDL_OK;
}

extern "C" int bdl_chain_create(uint64_t n, int variant, int with_sgd_momentum, uint64_t chunk_elems, bdl_chain** out) {
    using namespace bdl;
    BDL_REQUIRE(out, BDL_ERR_INVALID, "bdl_chain_create: null out");
    BDL_REQUIRE(variant >= BDL_SGLD && variant <= BDL_ADAM_CSGHMC, BDL_ERR_INVALID, "bdl_chain_create: unknown variant");
    BDL_REQUIRE(n > 0 && n % 4 == 0 && (n >> 2) < 0xFFFFFFFFull, BDL_ERR_INVALID, "bdl_chain_create: bad n");
    bdl_chain* c = new bdl_chain();
    c->n = n;
    c->variant = variant;
    c->chunk = chunk_elems ? (chunk_elems + 3) / 4 * 4 : kDefaultChunk;
    int rc = BDL_OK;
    for (int w = 0; w <= kGrad && rc == BDL_OK; ++w) {
        if (!needs(variant, w, with_sgd_momentum != 0)) continue;
        rc = check_cuda(cudaMalloc(&c->buf[w], n * sizeof(float)), "cudaMalloc(chain state)");
        if (rc == BDL_OK) rc = check_cuda(cudaMemset(c->buf[w], 0, n * sizeof(float)), "cudaMemset(chain state)");
    }
    if (rc == BDL_OK) rc = check_cuda(cudaMalloc(&c->runs_dev, BDL_MAX_RUNS * sizeof(bdl_run)), "cudaMalloc(runs)");
    if (rc == BDL_OK) rc = check_cuda(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking), "stream");
    if (rc == BDL_OK) rc = check_cuda(cudaStreamCreateWithFlags(&c->s_cmp, cudaStreamNonBlocking), "stream");
    if (rc == BDL_OK) rc = check_cuda(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking), "stream");
    {   // boundaries: chunk/8, chunk/4, chunk/2, then full chunks, then chunk/2, chunk/4, chunk/8 (all multiples of 4)
        const uint64_t ramp[3] = {c->chunk / 8 / 4 * 4, c->chunk / 4 / 4 * 4, c->chunk / 2 / 4 * 4};
        uint64_t head = 0, tail = 0;
        for (uint64_t r : ramp) { head += r; tail += r; }
        c->bounds.push_back(0);
        if (ramp[0] >= 4 && n > 2 * (head + tail)) {
            uint64_t pos = 0;
            for (int i = 0; i < 3; ++i) { pos += ramp[i]; c->bounds.push_back(pos); }
            const uint64_t mid_end = n - tail;
            while (pos + c->chunk < mid_end) { pos += c->chunk; c->bounds.push_back(pos); }
            if (pos < mid_end) { pos = mid_end; c->bounds.push_back(pos); }
            for (int i = 2; i >= 0; --i) { pos += ramp[i]; c->bounds.push_back(pos); }
        } else {
            for (uint64_t pos = c->chunk; pos < n; pos += c->chunk) c->bounds.push_back(pos);
            c->bounds.push_back(n);
        }
    }
    const uint64_t nchunks = c->bounds.size() - 1;
    for (uint64_t i = 0; i < nchunks && rc == BDL_OK; ++i) {
        cudaEvent_t a, b;
        rc = check_cuda(cudaEventCreateWithFlags(&a, cudaEventDisableTiming), "event");
        if (rc == BDL_OK) { c->ev_h2d.push_back(a); rc = check_cuda(cudaEventCreateWithFlags(&b, cudaEventDisableTiming), "event"); }
        if (rc == BDL_OK) c->ev_cmp.push_back(b);
    }
    if (rc != BDL_OK) {
        bdl_chain_destroy(c);
        return rc;
    }
    *out = c;
    return BDL_OK;
}

extern "C" int bdl_chain_upload(bdl_chain* c, int which, const float* host) {
    using namespace bdl;
    BDL_REQUIRE(c && host && which >= 0 && which < kGrad && c->buf[which], BDL_ERR_INVALID, "bdl_chain_upload: bad buffer %d", which);
    BDL_CUDA(cudaMemcpy(c->buf[which], host, c->n * sizeof(float), cudaMemcpyHostToDevice));
    return BDL_OK;
}

extern "C" int bdl_chain_download(bdl_chain* c, int which, float* host) {
    using namespace bdl;
    BDL_REQUIRE(c && host && which >= 0 && which < kGrad && c->buf[which], BDL_ERR_INVALID, "bdl_chain_download: bad buffer %d", which);
    BDL_CUDA(cudaMemcpy(host, c->buf[which], c->n * sizeof(float), cudaMemcpyDeviceToHost));
    return BDL_OK;
}

extern "C" int bdl_chain_device_ptr(bdl_chain* c, int which, float** out) {
    using namespace bdl;
    BDL_REQUIRE(c && out && which >= 0 && which <= kGrad, BDL_ERR_INVALID, "bdl_chain_device_ptr: bad argument");
    *out = c->buf[which];
    return BDL_OK;
}

extern "C" int bdl_chain_step_host(bdl_chain* c, const float* g_host, float* theta_out_host, const bdl_run* runs_host,
                                   uint32_t nruns, const bdl_scalars* sc, const bdl_noise* nz) {
    using namespace bdl;
    BDL_REQUIRE(c && g_host && theta_out_host && runs_host && sc && nz, BDL_ERR_INVALID, "bdl_chain_step_host: null argument");
    BDL_REQUIRE(nruns >= 1 && nruns <= BDL_MAX_RUNS, BDL_ERR_INVALID, "bdl_chain_step_host: nruns out of range");
    BDL_REQUIRE(nz->xi_dev == nullptr, BDL_ERR_UNSUPPORTED, "bdl_chain_step_host: injected noise is not supported here");
    for (uint32_t r = 0; r < nruns; ++r)
        BDL_REQUIRE(runs_host[r].g_dev == nullptr, BDL_ERR_UNSUPPORTED, "bdl_chain_step_host: per-run gradient pointers not supported");
    BDL_CUDA(cudaMemcpyAsync(c->runs_dev, runs_host, nruns * sizeof(bdl_run), cudaMemcpyHostToDevice, c->s_cmp));
    const uint64_t nchunks = c->bounds.size() - 1;
    for (uint64_t k = 0; k < nchunks; ++k) {
        const uint64_t off = c->bounds[k];
        const uint64_t len = c->bounds[k + 1] - off;
        BDL_CUDA(cudaMemcpyAsync(c->buf[kGrad] + off, g_host + off, len * sizeof(float), cudaMemcpyHostToDevice, c->s_h2d));
        BDL_CUDA(cudaEventRecord(c->ev_h2d[k], c->s_h2d));
        BDL_CUDA(cudaStreamWaitEvent(c->s_cmp, c->ev_h2d[k], 0));
        const int rc = step_range(c->variant, c->buf[BDL_BUF_THETA], c->buf[kGrad], c->buf[BDL_BUF_THETA0], c->buf[BDL_BUF_V],
                                  c->buf[BDL_BUF_M], c->buf[BDL_BUF_S], c->buf[BDL_BUF_SGD], c->n, off >> 2, (off + len) >> 2,
                                  c->runs_dev, nruns, runs_host, sc, nz, nullptr, c->s_cmp);
        if (rc != BDL_OK) return rc;
        BDL_CUDA(cudaEventRecord(c->ev_cmp[k], c->s_cmp));
        BDL_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_cmp[k], 0));
        BDL_CUDA(cudaMemcpyAsync(theta_out_host + off, c->buf[BDL_BUF_THETA] + off, len * sizeof(float), cudaMemcpyDeviceToHost,
                                 c->s_d2h));
    }
    BDL_CUDA(cudaStreamSynchronize(c->s_d2h));     // host-buffer contract: theta_out is complete on return
    return BDL_OK;
}
