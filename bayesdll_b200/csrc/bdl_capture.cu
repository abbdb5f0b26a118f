// bdl_capture.cu -- burn-in / thinning sample capture (north star (b)).
//
//   bdl_moments_avg      running first/second moments   methods/sgld.py:95-102,239-246; csgld.py:276-293
//   bdl_moments_welford  cSGHMC Welford update          methods/csghmc.py:327-348
//   bdl_capture_ring     raw-sample store               methods/csgld.py:278-279 (args.full_sample)
//
// Roofline: HBM.  Algorithmic bytes per element: moments 20 (R theta,mom1,mom2; W mom1,mom2; 12 on the
// initialising call), ring copy 8.
#include "bdl_common.cuh"

namespace bdl {

#ifndef BDL_CAP_THREADS
#define BDL_CAP_THREADS 128
#endif
constexpr int kCapThreads = BDL_CAP_THREADS;
constexpr int kCapU = 1;

template <int kDiv, bool kInit, bool kHasMom2>
__global__ void __launch_bounds__(kCapThreads, 1024 / kCapThreads)
moments_avg_kernel(const float* __restrict__ theta, float* __restrict__ mom1, float* __restrict__ mom2, uint32_t n4,
                   float cnt, float cntp1, float inv_cntp1) {
    const uint32_t tile_groups = kCapThreads * kCapU;
    const uint32_t ntiles = (n4 + tile_groups - 1) / tile_groups;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t q0 = tile * tile_groups + threadIdx.x;
        float4 th[kCapU], a[kCapU], b[kCapU];
#pragma unroll
        for (int u = 0; u < kCapU; ++u) {
            const uint32_t q = q0 + u * kCapThreads;
            if (q < n4) {
                const uint64_t i = static_cast<uint64_t>(q) << 2;
                th[u] = ld_stream(theta + i);
                if constexpr (!kInit) {
                    a[u] = ld_stream(mom1 + i);
                    if constexpr (kHasMom2) b[u] = ld_stream(mom2 + i);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kCapU; ++u) {
            const uint32_t q = q0 + u * kCapThreads;
            if (q < n4) {
                const uint64_t i = static_cast<uint64_t>(q) << 2;
                const float t[4] = {th[u].x, th[u].y, th[u].z, th[u].w};
                float m1[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
                float m2[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if constexpr (kInit) {
                        m1[k] = __fmul_rn(t[k], 1.0f);                 // theta_vec*1.0
                        m2[k] = __fmul_rn(t[k], t[k]);                 // theta_vec**2
                    } else {
                        // (theta + cnt*mom1) / (cnt+1) ; (theta**2 + cnt*mom2) / (cnt+1)
                        m1[k] = div_scalar<kDiv>(__fadd_rn(t[k], __fmul_rn(cnt, m1[k])), cntp1, inv_cntp1);
                        if constexpr (kHasMom2)
                            m2[k] = div_scalar<kDiv>(__fadd_rn(__fmul_rn(t[k], t[k]), __fmul_rn(cnt, m2[k])), cntp1, inv_cntp1);
                    }
                }
                st_stream(mom1 + i, make_float4(m1[0], m1[1], m1[2], m1[3]));
                if constexpr (kHasMom2) st_stream(mom2 + i, make_float4(m2[0], m2[1], m2[2], m2[3]));
            }
        }
    }
}

template <int kDiv, bool kInit>
__global__ void __launch_bounds__(kCapThreads, 1024 / kCapThreads)
moments_welford_kernel(const float* __restrict__ theta, float* __restrict__ mean, float* __restrict__ M2, uint32_t n4,
                       float nf, float inv_nf) {
    const uint32_t tile_groups = kCapThreads * kCapU;
    const uint32_t ntiles = (n4 + tile_groups - 1) / tile_groups;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t q0 = tile * tile_groups + threadIdx.x;
        float4 th[kCapU], a[kCapU], b[kCapU];
#pragma unroll
        for (int u = 0; u < kCapU; ++u) {
            const uint32_t q = q0 + u * kCapThreads;
            if (q < n4) {
                const uint64_t i = static_cast<uint64_t>(q) << 2;
                th[u] = ld_stream(theta + i);
                if constexpr (!kInit) {
                    a[u] = ld_stream(mean + i);
                    b[u] = ld_stream(M2 + i);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kCapU; ++u) {
            const uint32_t q = q0 + u * kCapThreads;
            if (q < n4) {
                const uint64_t i = static_cast<uint64_t>(q) << 2;
                const float t[4] = {th[u].x, th[u].y, th[u].z, th[u].w};
                float mu[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
                float m2[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if constexpr (kInit) {
                        mu[k] = t[k];                                  // theta_vec.clone()
                        m2[k] = 0.0f;                                  // zeros_like
                    } else {
                        const float d = __fsub_rn(t[k], mu[k]);        // delta
                        mu[k] = __fadd_rn(mu[k], div_scalar<kDiv>(d, nf, inv_nf));
                        const float d2 = __fsub_rn(t[k], mu[k]);       // delta2
                        m2[k] = __fadd_rn(m2[k], __fmul_rn(d, d2));
                    }
                }
                st_stream(mean + i, make_float4(mu[0], mu[1], mu[2], mu[3]));
                st_stream(M2 + i, make_float4(m2[0], m2[1], m2[2], m2[3]));
            }
        }
    }
}

// One tile per CTA, dispatched in address order (see the launch-shape note in bdl_step.cu).
static uint32_t ew_grid(uint32_t n4, int /*per_sm*/) {
    const uint32_t tile_groups = kCapThreads * kCapU;
    return (n4 + tile_groups - 1) / tile_groups;
}

// -------------------------------------------------------------------------------------------
// Sample ring: pure TMA copy.  One warp per CTA; lane 0 drives a kStages-deep pipeline of
//   cp.async.bulk global -> shared (mbarrier complete_tx)  and  cp.async.bulk shared -> global
// so no data ever passes through registers.  SASS: UBLKCP (both directions) + SYNCS.
// -------------------------------------------------------------------------------------------
constexpr uint32_t kRingChunk = 16384;   // bytes per stage
constexpr int kRingStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(32, 3)
ring_copy_kernel(const char* __restrict__ src, char* __restrict__ dst, uint64_t bytes, uint32_t per_cta) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[kRingStages];
    if (threadIdx.x != 0) return;

    const uint64_t nchunks = (bytes + kRingChunk - 1) / kRingChunk;
    // this CTA copies the contiguous chunks [blockIdx.x * per_cta, ...): CTAs are dispatched in address order
    const uint64_t c_begin = static_cast<uint64_t>(blockIdx.x) * per_cta;
    const uint64_t mine = c_begin < nchunks ? (nchunks - c_begin < per_cta ? nchunks - c_begin : per_cta) : 0;
    for (int s = 0; s < kRingStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

    auto chunk_off = [&](uint64_t k) { return (c_begin + k) * static_cast<uint64_t>(kRingChunk); };
    auto chunk_len = [&](uint64_t k) {
        const uint64_t off = chunk_off(k);
        return static_cast<uint32_t>(bytes - off < kRingChunk ? bytes - off : kRingChunk);
    };
    auto issue_load = [&](uint64_t k) {
        const int s = static_cast<int>(k % kRingStages);
        const uint32_t len = chunk_len(k);
        mbar_expect_tx(smem_u32(&bars[s]), len);
        tma_load_1d(smem_u32(smem + s * kRingChunk), src + chunk_off(k), len, smem_u32(&bars[s]));
    };

    const uint64_t pro = mine < kRingStages ? mine : kRingStages;
    for (uint64_t k = 0; k < pro; ++k) issue_load(k);
    for (uint64_t k = 0; k < mine; ++k) {
        const int s = static_cast<int>(k % kRingStages);
        mbar_wait(smem_u32(&bars[s]), static_cast<uint32_t>((k / kRingStages) & 1));
        tma_store_1d(dst + chunk_off(k), smem_u32(smem + s * kRingChunk), chunk_len(k));
        tma_commit();
        // refill the stage used by the *previous* chunk once its store has finished reading shared memory
        if (k >= 1 && k - 1 + kRingStages < mine) {
            tma_wait_read<1>();
            issue_load(k - 1 + kRingStages);
        }
    }
    tma_wait_all<0>();
}

}  // namespace bdl

extern "C" int bdl_moments_avg(const float* theta, float* mom1, float* mom2, uint64_t n, float cnt, float cntp1,
                               int init, int div_mode, void* stream) {
    using namespace bdl;
    if (n == 0) return BDL_OK;                     // empty state: a no-op, pointers may be null
    BDL_REQUIRE(theta && mom1, BDL_ERR_INVALID, "bdl_moments_avg: null theta/mom1");
    BDL_REQUIRE(n % 4 == 0 && (n >> 2) < 0xFFFFFFFFull, BDL_ERR_INVALID, "bdl_moments_avg: bad n");
    BDL_REQUIRE(aligned16(theta) && aligned16(mom1) && aligned16(mom2), BDL_ERR_ALIGN, "bdl_moments_avg: unaligned pointer");
    BDL_REQUIRE(div_mode == BDL_DIV_IEEE || div_mode == BDL_DIV_RECIP, BDL_ERR_INVALID, "bdl_moments_avg: bad div_mode");
    const uint32_t n4 = static_cast<uint32_t>(n >> 2);
    const uint32_t grid = ew_grid(n4, 4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float inv = scalar_reciprocal(cntp1, div_mode);
#define BDL_LAUNCH_MA(D, I, H) moments_avg_kernel<D, I, H><<<grid, kCapThreads, 0, st>>>(theta, mom1, mom2, n4, cnt, cntp1, inv)
    const bool h = mom2 != nullptr;
    if (init) {
        if (h) BDL_LAUNCH_MA(BDL_DIV_IEEE, true, true); else BDL_LAUNCH_MA(BDL_DIV_IEEE, true, false);
    } else if (div_mode == BDL_DIV_IEEE) {
        if (h) BDL_LAUNCH_MA(BDL_DIV_IEEE, false, true); else BDL_LAUNCH_MA(BDL_DIV_IEEE, false, false);
    } else {
        if (h) BDL_LAUNCH_MA(BDL_DIV_RECIP, false, true); else BDL_LAUNCH_MA(BDL_DIV_RECIP, false, false);
    }
#undef BDL_LAUNCH_MA
    return check_cuda(cudaGetLastError(), "moments_avg_kernel launch");
}

extern "C" int bdl_moments_welford(const float* theta, float* mean, float* M2, uint64_t n, float nf, int init,
                                   int div_mode, void* stream) {
    using namespace bdl;
    if (n == 0) return BDL_OK;                     // empty state: a no-op, pointers may be null
    BDL_REQUIRE(theta && mean && M2, BDL_ERR_INVALID, "bdl_moments_welford: null pointer");
    BDL_REQUIRE(n % 4 == 0 && (n >> 2) < 0xFFFFFFFFull, BDL_ERR_INVALID, "bdl_moments_welford: bad n");
    BDL_REQUIRE(aligned16(theta) && aligned16(mean) && aligned16(M2), BDL_ERR_ALIGN, "bdl_moments_welford: unaligned pointer");
    BDL_REQUIRE(div_mode == BDL_DIV_IEEE || div_mode == BDL_DIV_RECIP, BDL_ERR_INVALID, "bdl_moments_welford: bad div_mode");
    const uint32_t n4 = static_cast<uint32_t>(n >> 2);
    const uint32_t grid = ew_grid(n4, 4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float inv = scalar_reciprocal(nf, div_mode);
    if (init) moments_welford_kernel<BDL_DIV_IEEE, true><<<grid, kCapThreads, 0, st>>>(theta, mean, M2, n4, nf, inv);
    else if (div_mode == BDL_DIV_IEEE) moments_welford_kernel<BDL_DIV_IEEE, false><<<grid, kCapThreads, 0, st>>>(theta, mean, M2, n4, nf, inv);
    else moments_welford_kernel<BDL_DIV_RECIP, false><<<grid, kCapThreads, 0, st>>>(theta, mean, M2, n4, nf, inv);
    return check_cuda(cudaGetLastError(), "moments_welford_kernel launch");
}

static thread_local uint32_t g_ring_chunks_per_cta = 4;   // per calling thread (bdl_set_ring_config)
extern "C" int bdl_set_ring_config(int chunks_per_cta) {
    using namespace bdl;
    BDL_REQUIRE(chunks_per_cta >= 1 && chunks_per_cta <= (1 << 20), BDL_ERR_INVALID, "chunks_per_cta out of range");
    g_ring_chunks_per_cta = static_cast<uint32_t>(chunks_per_cta);
    return BDL_OK;
}

extern "C" int bdl_capture_ring(const float* theta, float* ring, uint64_t slot, uint64_t n, void* stream) {
    using namespace bdl;
    if (n == 0) return BDL_OK;                     // empty state: a no-op, pointers may be null
    BDL_REQUIRE(theta && ring, BDL_ERR_INVALID, "bdl_capture_ring: null pointer");
    BDL_REQUIRE(n % 4 == 0, BDL_ERR_INVALID, "bdl_capture_ring: n must be a multiple of 4");
    BDL_REQUIRE(aligned16(theta) && aligned16(ring), BDL_ERR_ALIGN, "bdl_capture_ring: unaligned pointer");
    const int smem_bytes = kRingStages * kRingChunk;
    BDL_CUDA(cudaFuncSetAttribute(ring_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const uint64_t bytes = n * sizeof(float);
    const uint64_t nchunks = (bytes + kRingChunk - 1) / kRingChunk;
    const uint32_t per_cta = g_ring_chunks_per_cta;
    const uint64_t grid = (nchunks + per_cta - 1) / per_cta;
    ring_copy_kernel<<<static_cast<uint32_t>(grid), 32, smem_bytes, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const char*>(theta), reinterpret_cast<char*>(ring + slot * n), bytes, per_cta);
    return check_cuda(cudaGetLastError(), "ring_copy_kernel launch");
}
