// ld256_probe.cu -- do Blackwell's 256-bit global loads / stores (LDG.E.256 / STG.E.256, PTX ld.global.v8.f32, sm_100+)
// buy HBM bandwidth for the sampler's streaming shapes?  One tile per CTA in address order like the product kernels;
// SGHMC traffic (4 reads, 2 writes per element), the draw's (2R:1W) and a copy (1R:1W), each as
//   v4   one 128-bit group per thread          (what libbdl does)
//   v4x2 two 128-bit groups per thread, strided by the CTA size
//   v8   one 256-bit access per thread and stream (two ADJACENT groups)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_ab/ld256_probe tools/ld256_probe.cu && tools/_ab/ld256_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

struct __align__(32) f8 { float v[8]; };
__device__ __forceinline__ f8 ld256(const float* p) {
    f8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(float* p, const f8& r) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                 "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}
__device__ __forceinline__ f8 ld128x2(const float* p) {   // the same 8 floats as two 128-bit loads
    f8 r;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

__device__ __forceinline__ void upd(float& t, float g, float t0, float& v) {
    v = v * 0.82f + 1e-4f * (g + (t - t0) * 1e-3f);
    t = t - 1e-4f * (g + v);
}

// kMode 0: v4 (kU groups per thread strided by blockDim), 1: v8 (256-bit)   kStreams: 6 = SGHMC, 3 = draw, 2 = copy
template <int kMode, int kU, int kStreams>
__global__ void stream_kernel(float* __restrict__ theta, const float* __restrict__ g, const float* __restrict__ theta0,
                              float* __restrict__ v, size_t n4) {
    if (kMode == 0) {
        const size_t q0 = blockIdx.x * size_t(blockDim.x) * kU + threadIdx.x;
        float4 t[kU], gg[kU], t0[kU], vv[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t q = q0 + u * blockDim.x;
            if (q < n4) {
                gg[u] = reinterpret_cast<const float4*>(g)[q];
                if (kStreams >= 3) t0[u] = reinterpret_cast<const float4*>(theta0)[q];
                if (kStreams == 6) { t[u] = reinterpret_cast<const float4*>(theta)[q]; vv[u] = reinterpret_cast<const float4*>(v)[q]; }
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t q = q0 + u * blockDim.x;
            if (q < n4) {
                if (kStreams == 6) {
                    upd(t[u].x, gg[u].x, t0[u].x, vv[u].x); upd(t[u].y, gg[u].y, t0[u].y, vv[u].y);
                    upd(t[u].z, gg[u].z, t0[u].z, vv[u].z); upd(t[u].w, gg[u].w, t0[u].w, vv[u].w);
                    __stcs(reinterpret_cast<float4*>(v) + q, vv[u]);
                    __stcs(reinterpret_cast<float4*>(theta) + q, t[u]);
                } else if (kStreams == 3) {
                    __stcs(reinterpret_cast<float4*>(theta) + q, make_float4(gg[u].x + t0[u].x, gg[u].y + t0[u].y, gg[u].z + t0[u].z, gg[u].w + t0[u].w));
                } else {
                    __stcs(reinterpret_cast<float4*>(theta) + q, gg[u]);
                }
            }
        }
    } else {
        const size_t q = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) * 2;     // two adjacent groups
        if (q + 1 < n4) {
            const size_t i = q * 4;
            f8 gg = ld256(g + i), t0, t, vv;
            if (kStreams >= 3) t0 = ld256(theta0 + i);
            if (kStreams == 6) { t = ld256(theta + i); vv = ld256(v + i); }
            if (kStreams == 6) {
#pragma unroll
                for (int j = 0; j < 8; ++j) upd(t.v[j], gg.v[j], t0.v[j], vv.v[j]);
                st256(v + i, vv);
                st256(theta + i, t);
            } else if (kStreams == 3) {
#pragma unroll
                for (int j = 0; j < 8; ++j) gg.v[j] += t0.v[j];
                st256(theta + i, gg);
            } else {
                st256(theta + i, gg);
            }
        }
    }
}

template <int kMode, int kU, int kStreams>
static float run(float* theta, float* g, float* theta0, float* v, size_t n4, int threads, int reps) {
    const size_t per_cta = size_t(threads) * (kMode == 0 ? kU : 2);
    const unsigned grid = unsigned((n4 + per_cta - 1) / per_cta);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) stream_kernel<kMode, kU, kStreams><<<grid, threads>>>(theta, g, theta0, v, n4);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) stream_kernel<kMode, kU, kStreams><<<grid, threads>>>(theta, g, theta0, v, n4);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    const size_t n = 305548328, n4 = n / 4;
    float *theta, *g, *theta0, *v;
    for (float** p : {&theta, &g, &theta0, &v}) { CK(cudaMalloc(p, n * 4)); CK(cudaMemset(*p, 0, n * 4)); }
    const int reps = 100;
    for (int round = 0; round < 3; ++round) {
        for (int threads : {64, 128, 256}) {
            const float a = run<0, 1, 6>(theta, g, theta0, v, n4, threads, reps), b = run<0, 2, 6>(theta, g, theta0, v, n4, threads, reps),
                        c = run<1, 1, 6>(theta, g, theta0, v, n4, threads, reps);
            printf("round %d T=%3d  SGHMC-like 4R:2W  v4 %.4f ms (%.0f GB/s)  v4x2 %.4f (%.0f)  v8 %.4f (%.0f)\n", round, threads, a,
                   n * 24 / a / 1e6, b, n * 24 / b / 1e6, c, n * 24 / c / 1e6);
            const float d = run<0, 1, 3>(theta, g, theta0, v, n4, threads, reps), e = run<0, 2, 3>(theta, g, theta0, v, n4, threads, reps),
                        f = run<1, 1, 3>(theta, g, theta0, v, n4, threads, reps);
            printf("round %d T=%3d  draw-like  2R:1W  v4 %.4f ms (%.0f GB/s)  v4x2 %.4f (%.0f)  v8 %.4f (%.0f)\n", round, threads, d,
                   n * 12 / d / 1e6, e, n * 12 / e / 1e6, f, n * 12 / f / 1e6);
            const float h = run<0, 1, 2>(theta, g, theta0, v, n4, threads, reps), i = run<0, 2, 2>(theta, g, theta0, v, n4, threads, reps),
                        j = run<1, 1, 2>(theta, g, theta0, v, n4, threads, reps);
            printf("round %d T=%3d  copy       1R:1W  v4 %.4f ms (%.0f GB/s)  v4x2 %.4f (%.0f)  v8 %.4f (%.0f)\n", round, threads, h,
                   n * 8 / h / 1e6, i, n * 8 / i / 1e6, j, n * 8 / j / 1e6);
            fflush(stdout);
        }
    }
    return 0;
}
