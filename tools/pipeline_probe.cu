// pipeline_probe.cu -- where does the host-buffer chain step (bdl_chain_step_host, csrc/bdl_host.cu) lose time against
// the copy-engine duplex ceiling?  Stand-alone (no libbdl): pinned host gradient in, pinned host theta out, a streaming
// kernel with the SGHMC step's traffic (4 reads, 2 writes per element) in between, chunk-pipelined in several ways.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_ab/pipeline_probe tools/pipeline_probe.cu
//   tools/_ab/pipeline_probe [n_elems]           (run under gpurun; prints one line per (mode, chunk))
//
// modes:  chain     H2D(k) -> kernel(k) -> D2H(k) on three streams (what bdl_host.cu does)
//         nokernel  H2D(k) -> D2H(k): the same event chain without the kernel (device-to-host of the gradient buffer)
//         alt2      kernel(k) and D2H(k) share stream k%2 (one event less per chunk, two D2H lanes)
//         mapped    H2D(k) -> kernel(k) that ALSO stores theta to the mapped pinned host buffer: no D2H copies at all
//         capped    the chain with a throttled kernel: G CTAs of a persistent grid (HBM left mostly to the copy engines)
//         mono      one H2D, one kernel, one D2H (no overlap) and the two monolithic copies at once (duplex ceiling)
// `--timeline` prints, for the chain mode at the default chunk, when each chunk's H2D / kernel / D2H finished.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <bool kMirror>
__global__ void __launch_bounds__(128) step_like(float* __restrict__ theta, const float* __restrict__ g, const float* __restrict__ theta0,
                                                 float* __restrict__ v, float* __restrict__ mirror, size_t q0, size_t q1) {
    const size_t q = q0 + blockIdx.x * size_t(128) + threadIdx.x;
    if (q >= q1) return;
    const float4 t = reinterpret_cast<const float4*>(theta)[q], gg = reinterpret_cast<const float4*>(g)[q];
    const float4 t0 = reinterpret_cast<const float4*>(theta0)[q], vv = reinterpret_cast<const float4*>(v)[q];
    float4 nv, nt;
    nv.x = vv.x * 0.82f + 1e-4f * (gg.x + (t.x - t0.x) * 1e-3f); nt.x = t.x - 1e-4f * (gg.x + nv.x);
    nv.y = vv.y * 0.82f + 1e-4f * (gg.y + (t.y - t0.y) * 1e-3f); nt.y = t.y - 1e-4f * (gg.y + nv.y);
    nv.z = vv.z * 0.82f + 1e-4f * (gg.z + (t.z - t0.z) * 1e-3f); nt.z = t.z - 1e-4f * (gg.z + nv.z);
    nv.w = vv.w * 0.82f + 1e-4f * (gg.w + (t.w - t0.w) * 1e-3f); nt.w = t.w - 1e-4f * (gg.w + nv.w);
    reinterpret_cast<float4*>(v)[q] = nv;
    reinterpret_cast<float4*>(theta)[q] = nt;
    if (kMirror) __stcs(reinterpret_cast<float4*>(mirror) + q, nt);
}

// the same update as a capped persistent grid: `gridDim.x` CTAs walk the range with a grid stride, two groups per thread in
// flight -- a kernel that draws a few hundred GB/s from HBM for most of a chunk period instead of 7 TB/s for 35 us
__global__ void __launch_bounds__(256) step_like_capped(float* __restrict__ theta, const float* __restrict__ g,
                                                        const float* __restrict__ theta0, float* __restrict__ v, size_t q0, size_t q1) {
    for (size_t q = q0 + blockIdx.x * size_t(256) + threadIdx.x; q < q1; q += size_t(gridDim.x) * 256) {
        const float4 t = reinterpret_cast<const float4*>(theta)[q], gg = reinterpret_cast<const float4*>(g)[q];
        const float4 t0 = reinterpret_cast<const float4*>(theta0)[q], vv = reinterpret_cast<const float4*>(v)[q];
        float4 nv, nt;
        nv.x = vv.x * 0.82f + 1e-4f * (gg.x + (t.x - t0.x) * 1e-3f); nt.x = t.x - 1e-4f * (gg.x + nv.x);
        nv.y = vv.y * 0.82f + 1e-4f * (gg.y + (t.y - t0.y) * 1e-3f); nt.y = t.y - 1e-4f * (gg.y + nv.y);
        nv.z = vv.z * 0.82f + 1e-4f * (gg.z + (t.z - t0.z) * 1e-3f); nt.z = t.z - 1e-4f * (gg.z + nv.z);
        nv.w = vv.w * 0.82f + 1e-4f * (gg.w + (t.w - t0.w) * 1e-3f); nt.w = t.w - 1e-4f * (gg.w + nv.w);
        reinterpret_cast<float4*>(v)[q] = nv;
        reinterpret_cast<float4*>(theta)[q] = nt;
    }
}

struct Ctx {
    size_t n;
    float *g_h, *out_h, *theta, *g, *theta0, *v;
    cudaStream_t s_h2d, s_cmp[2], s_d2h;
    int cap = 0;          // mode 4: CTAs of the capped kernel
};

static void launch(const Ctx& c, size_t off, size_t len, cudaStream_t st, bool mirror) {
    const size_t q0 = off / 4, q1 = (off + len) / 4;
    const unsigned grid = unsigned((q1 - q0 + 127) / 128);
    if (mirror) step_like<true><<<grid, 128, 0, st>>>(c.theta, c.g, c.theta0, c.v, c.out_h, q0, q1);
    else step_like<false><<<grid, 128, 0, st>>>(c.theta, c.g, c.theta0, c.v, nullptr, q0, q1);
}

// one pipelined step; ev arrays need >= nchunks entries.  tl != nullptr: timing events are recorded for the timeline
static void step(const Ctx& c, int mode, size_t chunk, std::vector<cudaEvent_t>& eh, std::vector<cudaEvent_t>& ec,
                 std::vector<cudaEvent_t>* ed) {
    size_t k = 0;
    for (size_t off = 0; off < c.n; off += chunk, ++k) {
        const size_t len = off + chunk <= c.n ? chunk : c.n - off;
        CK(cudaMemcpyAsync(c.g + off, c.g_h + off, len * 4, cudaMemcpyHostToDevice, c.s_h2d));
        CK(cudaEventRecord(eh[k], c.s_h2d));
        if (mode == 1) {                                  // nokernel
            CK(cudaStreamWaitEvent(c.s_d2h, eh[k], 0));
            CK(cudaMemcpyAsync(c.out_h + off, c.g + off, len * 4, cudaMemcpyDeviceToHost, c.s_d2h));
        } else if (mode == 0) {                           // chain
            CK(cudaStreamWaitEvent(c.s_cmp[0], eh[k], 0));
            launch(c, off, len, c.s_cmp[0], false);
            CK(cudaEventRecord(ec[k], c.s_cmp[0]));
            CK(cudaStreamWaitEvent(c.s_d2h, ec[k], 0));
            CK(cudaMemcpyAsync(c.out_h + off, c.theta + off, len * 4, cudaMemcpyDeviceToHost, c.s_d2h));
            if (ed) CK(cudaEventRecord((*ed)[k], c.s_d2h));
        } else if (mode == 2) {                           // alt2
            cudaStream_t s = c.s_cmp[k & 1];
            CK(cudaStreamWaitEvent(s, eh[k], 0));
            launch(c, off, len, s, false);
            CK(cudaMemcpyAsync(c.out_h + off, c.theta + off, len * 4, cudaMemcpyDeviceToHost, s));
        } else if (mode == 4) {                           // capped: the chain with a throttled kernel
            CK(cudaStreamWaitEvent(c.s_cmp[0], eh[k], 0));
            step_like_capped<<<c.cap, 256, 0, c.s_cmp[0]>>>(c.theta, c.g, c.theta0, c.v, off / 4, (off + len) / 4);
            CK(cudaEventRecord(ec[k], c.s_cmp[0]));
            CK(cudaStreamWaitEvent(c.s_d2h, ec[k], 0));
            CK(cudaMemcpyAsync(c.out_h + off, c.theta + off, len * 4, cudaMemcpyDeviceToHost, c.s_d2h));
        } else if (mode == 3) {                           // mapped
            CK(cudaStreamWaitEvent(c.s_cmp[0], eh[k], 0));
            launch(c, off, len, c.s_cmp[0], true);
        }
    }
    CK(cudaStreamSynchronize(c.s_d2h));
    CK(cudaStreamSynchronize(c.s_cmp[0]));
    CK(cudaStreamSynchronize(c.s_cmp[1]));
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    size_t n = 305548328;
    bool timeline = false, wc = false, quick = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--timeline")) timeline = true;
        else if (!strcmp(argv[i], "--wc")) wc = true;          // gradient buffer in WRITE-COMBINED pinned memory (the host only writes it)
        else if (!strcmp(argv[i], "--quick")) quick = true;    // ceilings + chain / nokernel only
        else n = strtoull(argv[i], nullptr, 10) / 4 * 4;
    }
    Ctx c{};
    c.n = n;
    CK(cudaHostAlloc(&c.g_h, n * 4, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    printf("gradient host buffer: %s pinned memory\n", wc ? "write-combined" : "default");
    CK(cudaHostAlloc(&c.out_h, n * 4, cudaHostAllocMapped));
    for (size_t i = 0; i < n; ++i) c.g_h[i] = 1e-2f * float((i * 2654435761u) >> 8 & 0xffff) / 65536.f;
    memset(c.out_h, 0, n * 4);
    for (float** p : {&c.theta, &c.g, &c.theta0, &c.v}) { CK(cudaMalloc(p, n * 4)); CK(cudaMemset(*p, 0, n * 4)); }
    CK(cudaStreamCreateWithFlags(&c.s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c.s_cmp[0], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c.s_cmp[1], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c.s_d2h, cudaStreamNonBlocking));
    const size_t max_chunks = n / (512 << 10) + 2;
    std::vector<cudaEvent_t> eh(max_chunks), ec(max_chunks), ed(max_chunks);
    for (size_t i = 0; i < max_chunks; ++i) {
        CK(cudaEventCreateWithFlags(&eh[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ec[i], cudaEventDisableTiming));
    }
    const int reps = 6;
    auto timeit = [&](auto&& fn) {
        fn(); fn();
        CK(cudaDeviceSynchronize());
        const double t0 = now();
        for (int r = 0; r < reps; ++r) fn();
        CK(cudaDeviceSynchronize());
        return (now() - t0) / reps * 1e3;
    };
    // ceilings
    {
        const double h2d = timeit([&] { CK(cudaMemcpyAsync(c.g, c.g_h, n * 4, cudaMemcpyHostToDevice, c.s_h2d)); CK(cudaStreamSynchronize(c.s_h2d)); });
        const double d2h = timeit([&] { CK(cudaMemcpyAsync(c.out_h, c.theta, n * 4, cudaMemcpyDeviceToHost, c.s_d2h)); CK(cudaStreamSynchronize(c.s_d2h)); });
        const double both = timeit([&] {
            CK(cudaMemcpyAsync(c.g, c.g_h, n * 4, cudaMemcpyHostToDevice, c.s_h2d));
            CK(cudaMemcpyAsync(c.out_h, c.theta, n * 4, cudaMemcpyDeviceToHost, c.s_d2h));
            CK(cudaStreamSynchronize(c.s_h2d)); CK(cudaStreamSynchronize(c.s_d2h)); });
        const double kern = timeit([&] { launch(c, 0, n, c.s_cmp[0], false); CK(cudaStreamSynchronize(c.s_cmp[0])); });
        const double kmap = timeit([&] { launch(c, 0, n, c.s_cmp[0], true); CK(cudaStreamSynchronize(c.s_cmp[0])); });
        printf("n %zu  H2D alone %.2f ms (%.1f GB/s)  D2H alone %.2f ms (%.1f GB/s)  both at once %.2f ms (%.1f GB/s each way)\n",
               n, h2d, n * 4 / h2d / 1e6, d2h, n * 4 / d2h / 1e6, both, n * 4 / both / 1e6);
        printf("kernel over all n: %.3f ms (%.0f GB/s);  kernel that also stores theta to mapped host memory: %.2f ms (%.1f GB/s over PCIe)\n",
               kern, n * 24 / kern / 1e6, kmap, n * 4 / kmap / 1e6);
        fflush(stdout);
    }
    const char* names[4] = {"chain", "nokernel", "alt2", "mapped"};
    const size_t chunks[] = {1u << 20, 2u << 20, 4u << 20, 8u << 20, 16u << 20};
    for (int mode = 0; mode < (quick ? 2 : 4); ++mode)
        for (size_t chunk : chunks) {
            const double ms = timeit([&] { step(c, mode, chunk, eh, ec, nullptr); });
            printf("%-9s chunk %3zu Mi elems (%4zu chunks): %7.2f ms  %5.1f GB/s each way\n", names[mode], chunk >> 20,
                   (n + chunk - 1) / chunk, ms, n * 4 / ms / 1e6);
            fflush(stdout);
        }
    if (!quick) for (int cap : {8, 16, 32, 64, 148, 592})
        for (size_t chunk : {size_t(1) << 20, size_t(2) << 20, size_t(4) << 20, size_t(8) << 20}) {
            c.cap = cap;
            const double k1 = timeit([&] { step_like_capped<<<cap, 256, 0, c.s_cmp[0]>>>(c.theta, c.g, c.theta0, c.v, 0, chunk / 4); CK(cudaStreamSynchronize(c.s_cmp[0])); });
            const double ms = timeit([&] { step(c, 4, chunk, eh, ec, nullptr); });
            printf("capped %3d CTAs  chunk %3zu Mi elems: %7.2f ms  %5.1f GB/s each way   (one chunk's kernel alone: %.3f ms = %.0f GB/s)\n", cap,
                   chunk >> 20, ms, n * 4 / ms / 1e6, k1, chunk * 24 / k1 / 1e6);
            fflush(stdout);
        }
    if (timeline) {
        const size_t chunk = 8u << 20, nch = (n + chunk - 1) / chunk;
        std::vector<cudaEvent_t> th(nch), tc(nch), td(nch);
        cudaEvent_t t0;
        CK(cudaEventCreate(&t0));
        for (size_t i = 0; i < nch; ++i) { CK(cudaEventCreate(&th[i])); CK(cudaEventCreate(&tc[i])); CK(cudaEventCreate(&td[i])); }
        step(c, 0, chunk, th, tc, &td);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(t0, c.s_h2d));
        step(c, 0, chunk, th, tc, &td);
        CK(cudaDeviceSynchronize());
        printf("timeline (chain, 8 Mi chunks): chunk  h2d_done  kernel_done  d2h_done  [ms since the step's first H2D was queued]\n");
        for (size_t i = 0; i < nch; ++i) {
            float a, b, d;
            CK(cudaEventElapsedTime(&a, t0, th[i])); CK(cudaEventElapsedTime(&b, t0, tc[i])); CK(cudaEventElapsedTime(&d, t0, td[i]));
            printf("  %3zu  %8.3f  %8.3f  %8.3f\n", i, a, b, d);
        }
    }
    return 0;
}
