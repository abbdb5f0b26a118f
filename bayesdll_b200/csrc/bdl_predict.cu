// bdl_predict.cu -- posterior-predictive ensemble and calibration reductions
// (north star (c)).
//
//   bdl_ensemble          log-mean-softmax over S samples     methods/sgld.py:299-300, csgld.py:416-431
//   bdl_ce_err            CE sum + error count                methods/sgld.py:302-306,314-315
//   bdl_lse_*             sample-sharded variant (running logsumexp; all-reduce MAX + SUM, section 8e)
//   bdl_calibrate         reliability bins + NLL              calibration.py:43-65, 246-249
//
// These tensors are tiny ([B,K,S] <= 64*37*40, [N,K] = 3669*37): the kernels are latency-bound, one
// warp per row with shuffle reductions; shared-memory bin counters for the histogram.
#include <math_constants.h>

#include "bdl_common.cuh"

namespace bdl {

__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

constexpr int kPredThreads = 256;
constexpr int kPredWarps = kPredThreads / 32;

__device__ __forceinline__ float mix(float comp, float w, float prev, int mode) {
    if (mode == 0) return comp;
    const float t = __fmul_rn(w, comp);                  // weight * component_out
    return mode == 1 ? t : __fadd_rn(prev, t);           // batch_logits += ...
}

// One warp per batch row b.  lse_s = logsumexp_K(L[b,:,s]) is kept in shared memory (S floats per warp).
__global__ void __launch_bounds__(kPredThreads)
ensemble_kernel(const float* __restrict__ L, uint32_t B, uint32_t K, uint32_t S, float log_S, float weight, int mode,
                float* __restrict__ out) {
    extern __shared__ float s_lse[];                     // [kPredWarps][2*S]: max_s, log-sum_s
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* mx_s = s_lse + static_cast<size_t>(warp) * 2 * S;
    float* ls_s = mx_s + S;
    for (uint32_t b = blockIdx.x * kPredWarps + warp; b < B; b += gridDim.x * kPredWarps) {
        const float* row = L + static_cast<size_t>(b) * K * S;
        // log_softmax over K for every sample s:  (x - max) - log(sum exp(x - max))
        for (uint32_t s = 0; s < S; ++s) {
            float mx = -CUDART_INF_F;
            for (uint32_t k = lane; k < K; k += 32) mx = fmaxf(mx, row[k * S + s]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (uint32_t k = lane; k < K; k += 32) sum += expf(row[k * S + s] - mx);
            sum = warp_sum(sum);
            if (lane == 0) {
                mx_s[s] = mx;
                ls_s[s] = logf(sum);
            }
        }
        __syncwarp();
        // logsumexp over S of the log-probabilities, minus log S
        for (uint32_t k = lane; k < K; k += 32) {
            float m2 = -CUDART_INF_F;
            for (uint32_t s = 0; s < S; ++s) m2 = fmaxf(m2, (row[k * S + s] - mx_s[s]) - ls_s[s]);
            float sum = 0.f;
            for (uint32_t s = 0; s < S; ++s) sum += expf(((row[k * S + s] - mx_s[s]) - ls_s[s]) - m2);
            const float comp = (logf(sum) + m2) - log_S;
            float* o = out + static_cast<size_t>(b) * K + k;
            *o = mix(comp, weight, mode == 2 ? *o : 0.f, mode);
        }
        __syncwarp();
    }
}

// Sample-sharded ensemble: running logsumexp over samples of log_softmax_K(logits), kept as (max, scaled sum) so that
// probabilities that underflow in linear space (p < 1e-38) still combine exactly like the reference's
// logsumexp(log_softmax) (methods/sgld.py:300).  m starts at -inf, s at 0.
__global__ void __launch_bounds__(kPredThreads)
lse_accum_kernel(const float* __restrict__ logits, uint32_t B, uint32_t K, float* __restrict__ m, float* __restrict__ s) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t b = blockIdx.x * kPredWarps + warp; b < B; b += gridDim.x * kPredWarps) {
        const float* row = logits + static_cast<size_t>(b) * K;
        float mx = -CUDART_INF_F;
        for (uint32_t k = lane; k < K; k += 32) mx = fmaxf(mx, row[k]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (uint32_t k = lane; k < K; k += 32) sum += expf(row[k] - mx);
        const float lsum = logf(warp_sum(sum));
        for (uint32_t k = lane; k < K; k += 32) {
            const size_t i = static_cast<size_t>(b) * K + k;
            const float ls = (row[k] - mx) - lsum;                  // log_softmax
            const float mo = m[i], mn = fmaxf(mo, ls);
            const float keep = mo == -CUDART_INF_F ? 0.f : s[i] * expf(mo - mn);
            s[i] = keep + expf(ls - mn);
            m[i] = mn;
        }
    }
}

// after all-reduce(MAX) of m: bring the local sums onto the global maximum so they can be all-reduce(SUM)ed
__global__ void __launch_bounds__(kPredThreads)
lse_rescale_kernel(const float* __restrict__ m_local, const float* __restrict__ m_global, float* __restrict__ s, uint32_t total) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const float ml = m_local[i];
        s[i] = ml == -CUDART_INF_F ? 0.f : s[i] * expf(ml - m_global[i]);
    }
}

__global__ void __launch_bounds__(kPredThreads)
lse_finalize_kernel(const float* __restrict__ m, const float* __restrict__ s, uint32_t total, float log_S, float weight,
                    int mode, float* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const float comp = (logf(s[i]) + m[i]) - log_S;
        out[i] = mix(comp, weight, mode == 2 ? out[i] : 0.f, mode);
    }
}

// Single CTA (B is a batch: tens to hundreds of rows) so that the accumulation order is fixed.
__global__ void __launch_bounds__(kPredThreads)
ce_err_kernel(const float* __restrict__ logits, const int64_t* __restrict__ y, uint32_t B, uint32_t K,
              double* __restrict__ loss_sum, int32_t* __restrict__ err_count) {
    __shared__ float s_loss[kPredWarps];
    __shared__ int s_err[kPredWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float loss = 0.f;
    int err = 0;
    for (uint32_t b = warp; b < B; b += kPredWarps) {
        const float* row = logits + static_cast<size_t>(b) * K;
        float mx = -CUDART_INF_F;
        int arg = 0x7fffffff;
        for (uint32_t k = lane; k < K; k += 32) {
            const float x = row[k];
            if (x > mx) { mx = x; arg = static_cast<int>(k); }       // first max within the lane's stride
        }
        // warp arg-max with lowest-index tie-break (torch.max returns the first maximal index on CUDA ties unspecified;
        // ties have probability ~0 for real logits)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
            if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
        }
        float sum = 0.f;
        for (uint32_t k = lane; k < K; k += 32) sum += expf(row[k] - mx);
        sum = warp_sum(sum);
        if (lane == 0) {
            const int64_t t = y[b];
            loss += -((row[t] - mx) - logf(sum));                    // -log_softmax[y]
            err += (arg != static_cast<int>(t));
        }
    }
    if (lane == 0) { s_loss[warp] = loss; s_err[warp] = err; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tl = 0.f;
        int te = 0;
        for (int w = 0; w < kPredWarps; ++w) { tl += s_loss[w]; te += s_err[w]; }
        *loss_sum += static_cast<double>(tl);
        *err_count += te;
    }
}

// -------------------------------------------------------------------------------------------
// calibration: one warp per row; per-CTA shared-memory bin counters; fp64 compare against the
// host-computed numpy edges (np.linspace(0, 1+1e-8, M+1)[1:]); np.digitize == #edges <= p.
// -------------------------------------------------------------------------------------------
constexpr int kMaxBins = 128;

template <typename T>
__device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <typename T>
__device__ __forceinline__ T t_log(T x);
template <> __device__ __forceinline__ float t_log<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double t_log<double>(double x) { return log(x); }

template <typename T>
__global__ void __launch_bounds__(kPredThreads)
calibrate_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, uint32_t N, uint32_t K, T temp,
                 const double* __restrict__ edges, uint32_t M, double* __restrict__ bin_size,
                 double* __restrict__ acc_sum, double* __restrict__ conf_sum, double* __restrict__ nll_sum,
                 unsigned long long* __restrict__ near_edge, int32_t* __restrict__ binned) {
    __shared__ double s_edges[kMaxBins];
    __shared__ unsigned int s_size[kMaxBins];
    __shared__ unsigned int s_acc[kMaxBins];
    __shared__ double s_conf[kMaxBins];
    __shared__ double s_nll;
    __shared__ unsigned int s_near;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        s_edges[i] = edges[i];
        s_size[i] = 0; s_acc[i] = 0; s_conf[i] = 0.0;
    }
    if (threadIdx.x == 0) { s_nll = 0.0; s_near = 0; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double guess_scale = static_cast<double>(M) / (1.0 + 1e-8);
    const bool scaled = !(temp == T(1));
    for (uint32_t r = blockIdx.x * kPredWarps + warp; r < N; r += gridDim.x * kPredWarps) {
        const float* row = logits + static_cast<size_t>(r) * K;
        const int64_t lab = labels[r];
        T mx = -CUDART_INF;
        for (uint32_t k = lane; k < K; k += 32) {
            const T x = scaled ? static_cast<T>(row[k]) / temp : static_cast<T>(row[k]);
            mx = x > mx ? x : mx;
        }
        mx = warp_max(mx);
        T sum = 0;
        for (uint32_t k = lane; k < K; k += 32) {
            const T x = scaled ? static_cast<T>(row[k]) / temp : static_cast<T>(row[k]);
            sum += t_exp<T>(x - mx);
        }
        sum = warp_sum(sum);
        for (uint32_t k = lane; k < K; k += 32) {
            const T x = scaled ? static_cast<T>(row[k]) / temp : static_cast<T>(row[k]);
            const T pt = t_exp<T>(x - mx) / sum;
            const double p = static_cast<double>(pt);
            int b = static_cast<int>(p * guess_scale);
            b = b < 0 ? 0 : (b > static_cast<int>(M) ? static_cast<int>(M) : b);
            while (b < static_cast<int>(M) && s_edges[b] <= p) ++b;          // digitize: #edges <= p
            while (b > 0 && s_edges[b - 1] > p) --b;
            // certification: is p within 16 ulp(T) of the nearest edge?  (a different-but-valid softmax
            // rounding could then land in the neighbouring bin)
            const double tol = p * (sizeof(T) == 4 ? 16.0 * 5.9604644775390625e-08 : 16.0 * 1.1102230246251565e-16);
            const double dl = b > 0 ? p - s_edges[b - 1] : 1.0;
            const double dr = b + 1 < static_cast<int>(M) ? s_edges[b] - p : 1.0;   // p <= 1 < last edge for any softmax
            if (dl <= tol || dr <= tol) atomicAdd(&s_near, 1u);
            if (binned) binned[static_cast<size_t>(r) * K + k] = b;
            if (b < static_cast<int>(M)) {
                atomicAdd(&s_size[b], 1u);
                if (static_cast<int64_t>(k) == lab) atomicAdd(&s_acc[b], 1u);
                atomicAdd(&s_conf[b], p);
            }
        }
        if (lane == 0) {
            const T xl = scaled ? static_cast<T>(row[lab]) / temp : static_cast<T>(row[lab]);
            const T nll = (t_log<T>(sum) + mx) - xl;                          // logsumexp - logit[y]
            atomicAdd(&s_nll, static_cast<double>(nll));
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        if (s_size[i]) {
            atomicAdd(&bin_size[i], static_cast<double>(s_size[i]));
            atomicAdd(&acc_sum[i], static_cast<double>(s_acc[i]));
            atomicAdd(&conf_sum[i], s_conf[i]);
        }
    }
    if (threadIdx.x == 0) {
        atomicAdd(nll_sum, s_nll);
        if (near_edge && s_near) atomicAdd(near_edge, static_cast<unsigned long long>(s_near));
    }
}

// -------------------------------------------------------------------------------------------
// Bayesian model average over stored raw samples (methods/csghmc_fs.py:349-377): logits are summed model by
// model in fp32 (numpy `all_logits_sum += model_logits`) and divided once by the model count (IEEE).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPredThreads)
bma_mean_kernel(const float* __restrict__ L, uint32_t total, uint32_t S, float S_f, float* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const float* x = L + static_cast<size_t>(i) * S;
        float acc = x[0];
        for (uint32_t s = 1; s < S; ++s) acc = __fadd_rn(acc, x[s]);
        out[i] = __fdiv_rn(acc, S_f);
    }
}

// -------------------------------------------------------------------------------------------
// Temperature-scaling objective (calibration.py:178-184): nll(T) = mean_i( logsumexp_k(l_ik / T) - l_iy / T ), all in
// fp64 (T is an fp64 ndarray in the reference, so `logits / T` promotes).  Two launches with a fixed reduction order,
// so the value is reproducible run to run -- the scipy BFGS driver differentiates it numerically with a 1.5e-8 step.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPredThreads)
nll_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, uint32_t N, uint32_t K, double temp,
                double* __restrict__ row_nll) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t r = blockIdx.x * kPredWarps + warp; r < N; r += gridDim.x * kPredWarps) {
        const float* row = logits + static_cast<size_t>(r) * K;
        double mx = -CUDART_INF;
        for (uint32_t k = lane; k < K; k += 32) mx = fmax(mx, static_cast<double>(row[k]) / temp);
        mx = warp_max(mx);
        double sum = 0.0;
        for (uint32_t k = lane; k < K; k += 32) sum += exp(static_cast<double>(row[k]) / temp - mx);
        sum = warp_sum(sum);                               // xor-butterfly: same value in every lane, fixed order
        if (lane == 0) row_nll[r] = (log(sum) + mx) - static_cast<double>(row[labels[r]]) / temp;
    }
}

constexpr int kSumThreads = 1024;
__global__ void __launch_bounds__(kSumThreads)
mean_f64_kernel(const double* __restrict__ x, uint32_t N, double* __restrict__ out) {
    __shared__ double s[kSumThreads];
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < N; i += kSumThreads) acc += x[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kSumThreads / 2; o > 0; o >>= 1) {
        if (static_cast<int>(threadIdx.x) < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0] / static_cast<double>(N);
}

static uint32_t rows_grid(uint32_t rows) {
    uint32_t need = (rows + kPredWarps - 1) / kPredWarps;
    uint32_t cap = static_cast<uint32_t>(num_sms() * 4);
    return need < cap ? (need ? need : 1) : cap;
}

}  // namespace bdl

extern "C" int bdl_ensemble(const float* logits_all, uint32_t B, uint32_t K, uint32_t S, float log_S, float weight,
                            int mode, float* out, void* stream) {
    using namespace bdl;
    if (B == 0) return BDL_OK;                     // empty batch: a no-op, pointers may be null
    BDL_REQUIRE(logits_all && out, BDL_ERR_INVALID, "bdl_ensemble: null pointer");
    BDL_REQUIRE(K >= 1 && S >= 1, BDL_ERR_INVALID, "bdl_ensemble: K and S must be >= 1");
    BDL_REQUIRE(mode >= 0 && mode <= 2, BDL_ERR_INVALID, "bdl_ensemble: mode must be 0,1,2");
    const size_t smem = static_cast<size_t>(kPredWarps) * 2 * S * sizeof(float);
    BDL_REQUIRE(smem <= 48 * 1024, BDL_ERR_UNSUPPORTED, "bdl_ensemble: S=%u too large", S);
    ensemble_kernel<<<rows_grid(B), kPredThreads, smem, static_cast<cudaStream_t>(stream)>>>(logits_all, B, K, S, log_S,
                                                                                           weight, mode, out);
    return check_cuda(cudaGetLastError(), "ensemble_kernel launch");
}

extern "C" int bdl_ce_err(const float* logits, const int64_t* y, uint32_t B, uint32_t K, double* loss_sum,
                          int32_t* err_count, void* stream) {
    using namespace bdl;
    if (B == 0) return BDL_OK;                     // empty batch: a no-op, pointers may be null
    BDL_REQUIRE(logits && y && loss_sum && err_count, BDL_ERR_INVALID, "bdl_ce_err: null pointer");
    BDL_REQUIRE(K >= 1, BDL_ERR_INVALID, "bdl_ce_err: K must be >= 1");
    ce_err_kernel<<<1, kPredThreads, 0, static_cast<cudaStream_t>(stream)>>>(logits, y, B, K, loss_sum, err_count);
    return check_cuda(cudaGetLastError(), "ce_err_kernel launch");
}

static uint32_t flat_grid(uint64_t total) {
    uint32_t grid = static_cast<uint32_t>((total + bdl::kPredThreads - 1) / bdl::kPredThreads);
    const uint32_t cap = static_cast<uint32_t>(bdl::num_sms() * 4);
    return grid > cap ? cap : grid;
}

extern "C" int bdl_lse_accum(const float* logits, uint32_t B, uint32_t K, float* m, float* s, void* stream) {
    using namespace bdl;
    if (B == 0) return BDL_OK;                     // empty batch: a no-op, pointers may be null
    BDL_REQUIRE(logits && m && s, BDL_ERR_INVALID, "bdl_lse_accum: null pointer");
    BDL_REQUIRE(K >= 1, BDL_ERR_INVALID, "bdl_lse_accum: K must be >= 1");
    lse_accum_kernel<<<rows_grid(B), kPredThreads, 0, static_cast<cudaStream_t>(stream)>>>(logits, B, K, m, s);
    return check_cuda(cudaGetLastError(), "lse_accum_kernel launch");
}

extern "C" int bdl_lse_rescale(const float* m_local, const float* m_global, float* s, uint64_t total, void* stream) {
    using namespace bdl;
    BDL_REQUIRE(m_local && m_global && s, BDL_ERR_INVALID, "bdl_lse_rescale: null pointer");
    BDL_REQUIRE(total < 0xFFFFFFFFull, BDL_ERR_INVALID, "bdl_lse_rescale: too many elements");
    if (total == 0) return BDL_OK;
    lse_rescale_kernel<<<flat_grid(total), kPredThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        m_local, m_global, s, static_cast<uint32_t>(total));
    return check_cuda(cudaGetLastError(), "lse_rescale_kernel launch");
}

extern "C" int bdl_lse_finalize(const float* m, const float* s, uint32_t B, uint32_t K, float log_S, float weight, int mode,
                                float* out, void* stream) {
    using namespace bdl;
    BDL_REQUIRE(m && s && out, BDL_ERR_INVALID, "bdl_lse_finalize: null pointer");
    BDL_REQUIRE(mode >= 0 && mode <= 2, BDL_ERR_INVALID, "bdl_lse_finalize: mode must be 0,1,2");
    const uint64_t total = static_cast<uint64_t>(B) * K;
    BDL_REQUIRE(total < 0xFFFFFFFFull, BDL_ERR_INVALID, "bdl_lse_finalize: B*K too large");
    if (total == 0) return BDL_OK;
    lse_finalize_kernel<<<flat_grid(total), kPredThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        m, s, static_cast<uint32_t>(total), log_S, weight, mode, out);
    return check_cuda(cudaGetLastError(), "lse_finalize_kernel launch");
}

extern "C" int bdl_calibrate(const float* logits, const int64_t* labels, uint64_t N, uint32_t K, double temperature,
                             int use_f64, const double* edges, uint32_t M, double* bin_size, double* acc_sum,
                             double* conf_sum, double* nll_sum, unsigned long long* near_edge, int32_t* binned,
                             void* stream) {
    using namespace bdl;
    if (N == 0) return BDL_OK;                     // empty batch: a no-op, pointers may be null
    BDL_REQUIRE(logits && labels && edges && bin_size && acc_sum && conf_sum && nll_sum, BDL_ERR_INVALID,
                "bdl_calibrate: null pointer");
    BDL_REQUIRE(M >= 1 && M <= static_cast<uint32_t>(kMaxBins), BDL_ERR_UNSUPPORTED, "bdl_calibrate: num_bins=%u not in [1,%d]", M, kMaxBins);
    BDL_REQUIRE(K >= 1 && N < 0xFFFFFFFFull, BDL_ERR_INVALID, "bdl_calibrate: bad N/K");
    BDL_REQUIRE(temperature > 0.0, BDL_ERR_INVALID, "bdl_calibrate: temperature must be > 0");
    const uint32_t grid = rows_grid(static_cast<uint32_t>(N));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (use_f64)
        calibrate_kernel<double><<<grid, kPredThreads, 0, st>>>(logits, labels, static_cast<uint32_t>(N), K, temperature,
                                                                edges, M, bin_size, acc_sum, conf_sum, nll_sum, near_edge, binned);
    else
        calibrate_kernel<float><<<grid, kPredThreads, 0, st>>>(logits, labels, static_cast<uint32_t>(N), K,
                                                               static_cast<float>(temperature), edges, M, bin_size,
                                                               acc_sum, conf_sum, nll_sum, near_edge, binned);
    return check_cuda(cudaGetLastError(), "calibrate_kernel launch");
}

extern "C" int bdl_bma_mean(const float* logits_all, uint32_t B, uint32_t K, uint32_t S, float* out, void* stream) {
    using namespace bdl;
    BDL_REQUIRE(K >= 1 && S >= 1, BDL_ERR_INVALID, "bdl_bma_mean: K and S must be >= 1");
    if (B == 0) return BDL_OK;                     // empty batch: a no-op, pointers may be null
    BDL_REQUIRE(logits_all && out, BDL_ERR_INVALID, "bdl_bma_mean: null pointer");
    const uint64_t total = static_cast<uint64_t>(B) * K;
    BDL_REQUIRE(total < 0xFFFFFFFFull, BDL_ERR_INVALID, "bdl_bma_mean: B*K too large");
    if (total == 0) return BDL_OK;
    bma_mean_kernel<<<flat_grid(total), kPredThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        logits_all, static_cast<uint32_t>(total), S, static_cast<float>(S), out);
    return check_cuda(cudaGetLastError(), "bma_mean_kernel launch");
}

extern "C" int bdl_nll_temperature(const float* logits, const int64_t* labels, uint64_t N, uint32_t K, double temperature,
                                   double* row_nll, double* out_mean, void* stream) {
    using namespace bdl;
    BDL_REQUIRE(logits && labels && row_nll && out_mean, BDL_ERR_INVALID, "bdl_nll_temperature: null pointer");
    BDL_REQUIRE(K >= 1 && N >= 1 && N < 0xFFFFFFFFull, BDL_ERR_INVALID, "bdl_nll_temperature: bad N/K");
    BDL_REQUIRE(temperature == temperature && temperature != 0.0, BDL_ERR_INVALID, "bdl_nll_temperature: temperature is 0 or NaN");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    nll_rows_kernel<<<rows_grid(static_cast<uint32_t>(N)), kPredThreads, 0, st>>>(logits, labels, static_cast<uint32_t>(N), K,
                                                                                temperature, row_nll);
    mean_f64_kernel<<<1, kSumThreads, 0, st>>>(row_nll, static_cast<uint32_t>(N), out_mean);
    return check_cuda(cudaGetLastError(), "nll_temperature launch");
}
