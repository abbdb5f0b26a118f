// bdl_selftest.cu -- device self-test of the arithmetic helpers in bdl_common.cuh (test / diagnostics entry, not on the
// hot path).  The optimistic sqrt / reciprocal / quotient helpers claim "same bits as the library intrinsic whenever
// the range flag is clear": this kernel checks the claim for EVERY fp32 bit pattern (sqrt, rcp: 2^32 inputs each) and for
// 2^32 Philox-hashed (x, d) pairs (quotient), counting mismatches and how often the fast path applied.
#include "bdl_common.cuh"

namespace bdl {

__global__ void __launch_bounds__(256) selftest_math_kernel(unsigned long long* __restrict__ out, uint32_t span) {
    // out[0..2]: mismatches sqrt / rcp / div ; out[3..5]: inputs on the fast path
    unsigned long long bad[3] = {0, 0, 0}, fast[3] = {0, 0, 0};
    const uint64_t first = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * span;
    for (uint32_t k = 0; k < span; ++k) {
        const uint32_t bits = static_cast<uint32_t>(first + k);
        const float x = __uint_as_float(bits);
        bool slow = false;
        const float a = sqrt_rn_opt(x, slow);
        if (!slow) {
            ++fast[0];
            bad[0] += __float_as_uint(a) != __float_as_uint(__fsqrt_rn(x));
        }
        slow = false;
        const float b = rcp_rn_opt(x, slow);
        if (!slow) {
            ++fast[1];
            bad[1] += __float_as_uint(b) != __float_as_uint(__frcp_rn(x));
        }
        // quotient: numerator = this bit pattern, divisor hashed from it (both windows are exercised)
        uint32_t h[4];
        philox4x32_10(bits, 0x51u, 0u, 0u, 0xA5A5A5A5u, 0x5A5A5A5Au, h);
        const float d = __uint_as_float(h[0]);
        const float r = __frcp_rn(d);
        slow = false;
        const float q = div_by_rcp_opt(x, d, r, slow);
        if (!slow) {
            ++fast[2];
            bad[2] += __float_as_uint(q) != __float_as_uint(__fdiv_rn(x, d));
        }
        if (__float_as_uint(div_by_rcp(x, d, r)) != __float_as_uint(__fdiv_rn(x, d)) && !(x != x) && !(d != d)) ++bad[2];
    }
    for (int j = 0; j < 3; ++j) {
        if (bad[j]) atomicAdd(out + j, bad[j]);
        atomicAdd(out + 3 + j, fast[j]);
    }
}

// Bare-traffic yardstick (bdl_probe_stream): the memory traffic of a sampler kernel with next to no arithmetic, in the
// product kernels' launch shape (one tile per CTA, CTAs dispatched in address order, one 128-bit group per thread and
// stream, evict-first stores).  kReads / kWrites streams: 4/2 = SGHMC step, 2/1 = posterior draw, 1/1 = copy.
template <int kReads, int kWrites>
__global__ void __launch_bounds__(256) probe_stream_kernel(float* __restrict__ a, float* __restrict__ b, const float* __restrict__ c,
                                                           const float* __restrict__ d, uint32_t n4) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n4) return;
    const uint64_t i = static_cast<uint64_t>(q) << 2;
    float4 x = *reinterpret_cast<const float4*>(c + i);
    if constexpr (kReads >= 2) { const float4 y = *reinterpret_cast<const float4*>(d + i); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
    if constexpr (kReads >= 3) { const float4 y = *reinterpret_cast<const float4*>(a + i); x.x = 0.5f * (x.x + y.x); x.y = 0.5f * (x.y + y.y); x.z = 0.5f * (x.z + y.z); x.w = 0.5f * (x.w + y.w); }
    float4 z = x;
    if constexpr (kReads >= 4) { const float4 y = *reinterpret_cast<const float4*>(b + i); z.x = 0.5f * (x.x - y.x); z.y = 0.5f * (x.y - y.y); z.z = 0.5f * (x.z - y.z); z.w = 0.5f * (x.w - y.w); }
    st_stream(a + i, x);
    if constexpr (kWrites >= 2) st_stream(b + i, z);
}

}  // namespace bdl

extern "C" int bdl_probe_stream(float* a, float* b, const float* c, const float* d, uint64_t n, int reads, int writes,
                                int threads, void* stream) {
    using namespace bdl;
    if (n == 0) return BDL_OK;
    BDL_REQUIRE(a && c && (reads < 2 || d) && ((reads < 4 && writes < 2) || b), BDL_ERR_INVALID, "bdl_probe_stream: null pointer");
    BDL_REQUIRE(n % 4 == 0 && (n >> 2) < 0xFFFFFFFFull, BDL_ERR_INVALID, "bdl_probe_stream: bad n");
    BDL_REQUIRE(aligned16(a) && aligned16(b) && aligned16(c) && aligned16(d), BDL_ERR_ALIGN, "bdl_probe_stream: unaligned pointer");
    BDL_REQUIRE(threads == 64 || threads == 128 || threads == 256, BDL_ERR_INVALID, "bdl_probe_stream: threads must be 64, 128 or 256");
    const uint32_t n4 = static_cast<uint32_t>(n >> 2);
    const uint32_t grid = (n4 + threads - 1) / threads;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (reads == 4 && writes == 2) probe_stream_kernel<4, 2><<<grid, threads, 0, st>>>(a, b, c, d, n4);
    else if (reads == 2 && writes == 1) probe_stream_kernel<2, 1><<<grid, threads, 0, st>>>(a, b, c, d, n4);
    else if (reads == 1 && writes == 1) probe_stream_kernel<1, 1><<<grid, threads, 0, st>>>(a, b, c, d, n4);
    else BDL_REQUIRE(false, BDL_ERR_INVALID, "bdl_probe_stream: (reads, writes) must be (4,2), (2,1) or (1,1)");
    return check_cuda(cudaGetLastError(), "probe_stream_kernel launch");
}

extern "C" int bdl_selftest_math(unsigned long long* out6_dev, void* stream) {
    using namespace bdl;
    BDL_REQUIRE(out6_dev != nullptr, BDL_ERR_INVALID, "bdl_selftest_math: output required");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BDL_CUDA(cudaMemsetAsync(out6_dev, 0, 6 * sizeof(unsigned long long), st));
    constexpr uint32_t span = 1024;                           // 2^32 / (16384 CTAs * 256 threads)
    selftest_math_kernel<<<16384, 256, 0, st>>>(out6_dev, span);
    return check_cuda(cudaGetLastError(), "selftest_math_kernel launch");
}
