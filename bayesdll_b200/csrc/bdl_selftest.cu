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

}  // namespace bdl

extern "C" int bdl_selftest_math(unsigned long long* out6_dev, void* stream) {
    using namespace bdl;
    BDL_REQUIRE(out6_dev != nullptr, BDL_ERR_INVALID, "bdl_selftest_math: output required");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BDL_CUDA(cudaMemsetAsync(out6_dev, 0, 6 * sizeof(unsigned long long), st));
    constexpr uint32_t span = 1024;                           // 2^32 / (16384 CTAs * 256 threads)
    selftest_math_kernel<<<16384, 256, 0, st>>>(out6_dev, span);
    return check_cuda(cudaGetLastError(), "selftest_math_kernel launch");
}
