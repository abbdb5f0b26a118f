/*
 * bdl.h -- C ABI of libbdl: B200 (sm_100a) kernels for the SG-MCMC sampler hot path of
 * BayesDLL (omarezz46/BayesDLL).  This is the drop-in boundary.
 *
 * The reference has no FFI today: its hot path is per-tensor Python loops over torch eager ops.
 * Each entry point below names the reference call site (file:line, relative to the reference
 * checkout) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross the boundary.
 *   - every pointer named *_dev is DEVICE memory owned by the caller (torch CUDA tensors in the
 *     Python host).  The library never allocates, frees or retains caller memory.
 *   - `stream` is a cudaStream_t passed as void*; all launches are asynchronous on it; no
 *     hidden synchronisation.
 *   - return value: 0 (BDL_OK) or a negative bdl_status; bdl_last_error() returns a
 *     thread-local human-readable message.  No C++ exception crosses the boundary.
 *   - flat buffers: fp32, base 16-byte aligned, length `n` a multiple of 4 elements.  Tensors
 *     are laid out in named_parameters() order, every tensor start rounded up to a multiple of
 *     4 elements ("padded flat layout"); padding elements are ordinary elements with g = 0.
 *   - Python doubles are rounded to fp32 by the host exactly where the reference's eager ops
 *     round them; bdl_scalars carries the already-rounded values.
 */
#ifndef BDL_H_
#define BDL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BDL_ABI_VERSION 7

typedef enum {
    BDL_OK = 0,
    BDL_ERR_INVALID = -1,   /* bad argument (null pointer, n not multiple of 4, unknown variant ...) */
    BDL_ERR_ALIGN = -2,     /* a pointer is not 16-byte aligned */
    BDL_ERR_CUDA = -3,      /* a CUDA runtime call failed; see bdl_last_error() */
    BDL_ERR_UNSUPPORTED = -4
} bdl_status;

/* Sampler update rules. */
typedef enum {
    BDL_SGLD = 0,          /* methods/sgld.py:469-484 + SGD.step :226 ; methods/csgld.py:665-680 + :253 */
    BDL_SGHMC = 1,         /* methods/sghmc.py:482-510 + SGD(momentum=0).step :229                      */
    BDL_CSGHMC = 2,        /* methods/csghmc.py:747-778 (writes p.data, no optimizer step :304)         */
    BDL_ADAM_SGHMC = 3,    /* methods/adam_sghmc.py:507-553 + SGD(momentum).step :233                   */
    BDL_ADAM_CSGHMC = 4    /* methods/adam_csghmc.py:814-861 + SGD(momentum=0).step :322                */
} bdl_variant;

/* Element class bits (per run). */
#define BDL_CLS_HEAD 1u    /* readout_name in pname  -> lr_head / head noise scale (methods/sghmc.py:485-488) */
#define BDL_CLS_PRIOR 2u   /* prior pull enabled; cleared for 'bias' in pname && bias=='uninformative' (:494) */
#define BDL_CLS_SKIP 4u    /* p.grad is None: the reference leaves such a tensor untouched (:484)            */
#define BDL_CLS_NODROP 8u  /* bdl_dropout_mix only: z = 1 for this run (bias tensors, mc_dropout.py:383-389)        */

/* A run = a contiguous range of the padded flat layout whose elements share one class
 * (and, optionally, one gradient tensor).  Runs are sorted, contiguous and cover [0, n). */
typedef struct {
    uint64_t begin;        /* first element, multiple of 4                                         */
    uint64_t end;          /* one past the last element incl. padding, multiple of 4               */
    uint64_t valid_end;    /* begin + number of real elements; [valid_end, end) is padding         */
    const float* g_dev;    /* optional per-run gradient tensor (element 0 <-> flat index `begin`,   */
                           /* 16-byte aligned, READABLE up to the next 16-byte boundary past its    */
                           /* end: the tail group is fetched as one 128-bit load and the lanes past */
                           /* valid_end are zeroed in registers.  Any allocator with >= 16-byte      */
                           /* granularity (cudaMalloc: 256 B, torch's caching allocator: 512 B)     */
                           /* satisfies this; a sub-allocated view ending at the very end of a       */
                           /* mapping with numel % 4 != 0 does not); NULL -> flat gradient buffer   */
    uint32_t cls;          /* BDL_CLS_* bits                                                       */
    uint32_t reserved;
} bdl_run;

#define BDL_MAX_RUNS 2048

/* Division semantics for `tensor / python_scalar` (last row of the table):
 * the reference on CPU divides (IEEE); the reference on CUDA multiplies by fp32(1.0 / s) where the reciprocal is taken
 * in DOUBLE precision from the Python double and rounded once (probed on this torch build: tools/probe_torch_div.py,
 * profiles/r01_torch_div_probe.log; e.g. s = 1 - 0.99**2 gives a different fp32 factor than 1.0f / fp32(s)). */
#define BDL_DIV_IEEE 0
#define BDL_DIV_RECIP 1

typedef struct {
    float lr[2];             /* [body, head]   fp32(lr_e)                                              */
    float noise_scale[2];    /* [body, head]   SGLD: nd*sqrt(2/(N lr)); SGHMC: nd*sqrt(2 a/(N lr));     */
                             /*                cSGHMC: nd*sqrt(2 a lr)/N   (host fp64 -> fp32)          */
    float one_minus_alpha;   /* fp32(1 - momentum_decay)                                               */
    float sig2;              /* fp32(prior_sig**2);  cSGHMC: fp32(prior_sig) used as an L2 coefficient  */
    float N;                 /* fp32(ND * Ninflate)                                                    */
    float mu;                /* torch SGD momentum (args.momentum); 0 -> no momentum buffer            */
    float beta1, one_minus_beta1, beta2, one_minus_beta2;
    float bias_corr1;        /* fp32(1 - beta1**t), t after the increment (adam_sghmc.py:494,533)      */
    float bias_corr2;        /* fp32(1 - beta2**t)                                                     */
    float eps;               /* Adam epsilon                                                           */
    float two_alpha;         /* fp32(2 * momentum_decay)   (adam_sghmc.py:541)                         */
    float nd;                /* noise discount (Adam variants multiply it in-kernel, :541)             */
    float temperature;       /* Adam-cSGHMC: g / T (adam_csghmc.py:829-831)                            */
    int32_t first_step;      /* 1: SGD momentum buffer is initialised to g' (torch sgd.py)             */
    int32_t add_noise;       /* cSGHMC: should_sample (csghmc.py:769); other variants ignore it        */
    int32_t div_mode;        /* BDL_DIV_IEEE | BDL_DIV_RECIP                                           */
    int32_t reserved;
    /* BDL_DIV_RECIP factors: fp32(1.0 / s_double), computed by the host in double from the ORIGINAL Python doubles
     * (prior_sig**2, ND*Ninflate, 1-beta1**t, 1-beta2**t, temperature).  A field left at 0 makes the library use
     * 1.0f / fp32(s) instead (differs from torch CUDA by one ulp of the factor for a minority of divisors). */
    float inv_sig2, inv_N, inv_bias_corr1, inv_bias_corr2, inv_temperature;
    int32_t reserved2;
} bdl_scalars;

/* Gaussian noise source.  xi_dev != NULL: externally injected N(0,1) draws in the padded flat
 * layout (parity mode: replaces torch.randn_like, methods/sghmc.py:501).  xi_dev == NULL:
 * in-kernel counter-based Philox4x32-10 + Box-Muller keyed by (seed; element/4, stream, subseq):
 * results do not depend on grid shape or on how work is sharded. */
typedef struct {
    const float* xi_dev;
    uint64_t seed;
    uint64_t subseq;         /* step counter (sampler) or packed (batch, cycle, sample) id (draw) */
    uint32_t stream_id;      /* BDL_STREAM_* : separates uses of the same seed                    */
    uint32_t reserved;
} bdl_noise;

#define BDL_STREAM_STEP 0u
#define BDL_STREAM_DRAW 1u
#define BDL_STREAM_USER 2u

int bdl_abi_version(void);
const char* bdl_last_error(void);

/* Optional launch tuning for bdl_step (0 = library default: one tile per CTA, 64-256 threads depending on the
 * variant, 1 float4 group per thread).  ctas_per_sm > 0 caps the grid at #SM * ctas_per_sm persistent CTAs.  Used by bench sweeps and by the
 * launch-shape-independence tests; not needed for correctness.  The setting is THREAD-LOCAL: it affects the launches of the
 * calling thread only, so the library stays re-entrant across threads / devices.  An explicit shape always runs the
 * generic build of the kernel (never the lean kFast / run-table builds). */
int bdl_set_launch_config(int ctas_per_sm, int unroll, int threads);

/* (a1..a5) One fused sampler update over the whole flat state: prior pull, friction/momentum,
 * Adam moments, noise, SGD momentum and the parameter update in a single pass.
 *   theta, v, m, s, buf: updated in place.  g / theta0: read only.
 *   Unused state for a variant may be NULL (v: SGLD; m,s: non-Adam; buf: mu == 0; theta0: cSGHMC).
 *   g_dev may be NULL iff every run carries its own g_dev.
 *   runs_dev: DEVICE array of nruns bdl_run (<= BDL_MAX_RUNS).
 *   runs_host: optional HOST copy of the same table (may be NULL).  Tables of <= 8 runs without per-run gradient
 *   pointers (the usual body | head split) are then passed inside the kernel arguments, which removes every table
 *   load from the kernel; runs_dev may be NULL in that case. */
int bdl_step(int variant, float* theta_dev, const float* g_dev, const float* theta0_dev,
             float* v_dev, float* m_dev, float* s_dev, float* buf_dev, uint64_t n,
             const bdl_run* runs_dev, uint32_t nruns, const bdl_run* runs_host, const bdl_scalars* scalars,
             const bdl_noise* noise, void* stream);

/* Fused step + sample capture.  After burn-in (or in the sampling phase of a cycle) the reference folds the NEW theta
 * into its running moments right after the optimizer step (methods/sghmc.py:242-249, methods/csgld.py:276-293,
 * methods/csghmc.py:327-348) -- with thin = 1 that is every step.  bdl_step_capture does both in one pass: the new
 * theta never leaves the registers, which saves the capture kernel's re-read (SGHMC + moments: 40 instead of
 * 24 + 20 = 44 B/param).  Arithmetic is exactly bdl_step followed by bdl_moments_avg / bdl_moments_welford on the
 * same stream (bit-identical; division semantics = scalars->div_mode).  Tensors skipped by BDL_CLS_SKIP are still
 * captured (their unchanged theta is part of the sample), as parameters_to_vector does. */
#define BDL_CAPTURE_NONE 0
#define BDL_CAPTURE_AVG 1       /* bdl_moments_avg:     first = mom1, second = mom2 (may be NULL: nst == 0)      */
#define BDL_CAPTURE_WELFORD 2   /* bdl_moments_welford: first = mean, second = M2                                */
typedef struct {
    int32_t kind;          /* BDL_CAPTURE_*                                                                     */
    int32_t init;          /* != 0: first sample (avg: mom1 = theta*1.0, mom2 = theta**2; Welford: mean = theta, M2 = 0) */
    float* first_dev;
    float* second_dev;
    float cnt;             /* avg: fp32(cnt) ; Welford: fp32(n)                                                 */
    float cnt_plus_1;      /* avg: fp32(cnt + 1) ; Welford: unused                                              */
} bdl_capture;
/* capture == NULL or kind == BDL_CAPTURE_NONE: identical to bdl_step. */
int bdl_step_capture(int variant, float* theta_dev, const float* g_dev, const float* theta0_dev,
                     float* v_dev, float* m_dev, float* s_dev, float* buf_dev, uint64_t n,
                     const bdl_run* runs_dev, uint32_t nruns, const bdl_run* runs_host, const bdl_scalars* scalars,
                     const bdl_noise* noise, const bdl_capture* capture, void* stream);

/* Gradient-norm clipping between Model.forward and optimizer.step() -- ``torch.nn.utils.clip_grad_norm_(net.parameters(),
 * args.clip_grad)`` at methods/csgld.py:250-251 (p.grad = g + prior + noise) and methods/adam_csghmc.py:319-320 (p.grad = v):
 * every p.grad is scaled by min(1, max_norm / (||all p.grad||_2 + 1e-6)) before the SGD step.  Fused as two passes with
 * no host synchronisation:
 *   bdl_step_gradnorm   recomputes what the reference holds in p.grad (same arithmetic as bdl_step, nothing stored; the
 *                       counter-based noise makes both passes see the same draw) and ADDS the sum of its squares over the
 *                       real elements of every tensor with a gradient to *sumsq_dev (caller zeroes it);
 *   bdl_clip_coef       total_norm = fp32(sqrt(sumsq)); *coef_dev = min(1, max_norm / (total_norm + 1e-6)) (fp32, torch's
 *                       statements); total_norm_dev is optional;
 *   bdl_step_clipped    bdl_step with p.grad scaled by *coef_dev.
 * Variants: BDL_SGLD (sgld / csgld) and BDL_ADAM_CSGHMC.  (csghmc / csghmc_fs also consult args.clip_grad, :301 / :509, but
 * write p.data inside Model.forward, so their clipping never reaches theta: nothing to do.)  runs_host: HOST array with
 * one row per tensor (<= 512 rows, with or without per-run gradient pointers); it travels in the kernel arguments. */
int bdl_step_gradnorm(int variant, const float* theta_dev, const float* g_dev, const float* theta0_dev, const float* v_dev,
                      const float* m_dev, const float* s_dev, const float* buf_dev, uint64_t n, const bdl_run* runs_host,
                      uint32_t nruns, const bdl_scalars* scalars, const bdl_noise* noise, double* sumsq_dev, void* stream);
int bdl_clip_coef(const double* sumsq_dev, float max_norm, float* coef_dev, float* total_norm_dev, void* stream);
int bdl_step_clipped(int variant, float* theta_dev, const float* g_dev, const float* theta0_dev, float* v_dev, float* m_dev,
                     float* s_dev, float* buf_dev, uint64_t n, const bdl_run* runs_host, uint32_t nruns,
                     const bdl_scalars* scalars, const bdl_noise* noise, const float* coef_dev, void* stream);

/* Fill out[0..n) with exactly the N(0,1) stream the step / draw kernels use for (seed, stream_id,
 * subseq).  Test and diagnostics entry (KS / moment tests; external-vs-in-kernel equivalence). */
int bdl_philox_normal(float* out_dev, uint64_t n, uint64_t seed, uint32_t stream_id, uint64_t subseq,
                      void* stream);

/* (a7, a8) Running first/second moments of theta.
 *   init != 0 : mom1 = theta*1.0 ; mom2 = theta**2            (methods/sgld.py:95-102, csgld.py:282-284)
 *   else      : mom <- (theta^k + cnt*mom) / (cnt+1)          (methods/sgld.py:243-245, csgld.py:286-290)
 *   mom2_dev may be NULL (nst == 0). cnt is passed as fp32(cnt) and fp32(cnt+1). */
int bdl_moments_avg(const float* theta_dev, float* mom1_dev, float* mom2_dev, uint64_t n, float cnt,
                    float cnt_plus_1, int init, int div_mode, void* stream);

/* (a8) cSGHMC Welford update (methods/csghmc.py:333-345).
 *   init != 0 : mean = theta ; M2 = 0
 *   else      : d = theta-mean ; mean += d/n_f ; M2 += d*(theta-mean)          (n_f = fp32(n)) */
int bdl_moments_welford(const float* theta_dev, float* mean_dev, float* m2_dev, uint64_t n, float n_f,
                        int init, int div_mode, void* stream);

/* (north star b) Copy theta into slot `slot` of a preallocated sample ring [slots][n] using TMA
 * bulk copies (cp.async.bulk global->shared->global).  Replaces theta_vec.clone() into a dict
 * (methods/csgld.py:278-279). */
int bdl_capture_ring(const float* theta_dev, float* ring_dev, uint64_t slot, uint64_t n, void* stream);
/* Tuning: 16 KiB chunks copied by one CTA of the ring kernel (default 4); thread-local like bdl_set_launch_config. */
int bdl_set_ring_config(int chunks_per_cta);

/* (a9) Posterior draw theta_s = mean + sqrt(var) * eps (methods/sgld.py:292-297, csgld.py:404-413).
 *   var_mode 0: var = max(scale * (second - mean^2), 1e-12)   scale = fp32(ratio)  (sgld.py:338-348)
 *   var_mode 1: var = max(second / scale, 1e-12)              scale = fp32(n-1)    (csghmc.py:451-459)
 *   var_mode 2: var = 1e-12                                   (csghmc.py:458)
 *   var_mode 3: second already holds the variance
 *   var_mode 4: second holds a standard deviation s_: theta = mean + max(s_, 1e-8) * eps -- the VI reparameterisation
 *               draw (methods/vi.py:402-406), no square root
 *   center_dev (optional): draw around this vector instead of `mean` while the variance still comes from
 *   (mean, second) -- cSGLD's cycle likelihoods perturb the *current* theta with the cycle's variance
 *   (methods/csgld.py:518-541).  NULL -> centre = mean. */
int bdl_draw(const float* mean_dev, const float* second_dev, const float* center_dev, float* theta_out_dev, uint64_t n,
             int var_mode, float scale, int div_mode, const bdl_noise* noise, void* stream);

/* (section 8f row 4) MC-Dropout reparameterisation draw (methods/mc_dropout.py:378-394):
 *   z = (u > p_drop) ? 1 : 0 with u ~ U[0,1) ;  theta_out = z*m + (1-z)*theta0 ; runs carrying BDL_CLS_NODROP keep z = 1.
 *   noise->xi_dev == NULL: in-kernel Philox uniforms (same counter layout as the Gaussian stream, 24 bits per value);
 *   noise->xi_dev != NULL: injected uniforms in the padded flat layout (parity mode: replaces torch.rand_like).
 *   z_out_dev (optional): the mask as fp32 0/1, which the reference keeps for its gradient (mc_dropout.py:393-394).
 *   runs_dev / nruns: DEVICE run table (only the BDL_CLS_NODROP bit is read); NULL / 0 = dropout everywhere. */
int bdl_dropout_mix(const float* m_dev, const float* theta0_dev, float* theta_out_dev, float* z_out_dev, uint64_t n,
                    const bdl_run* runs_dev, uint32_t nruns, float p_drop, const bdl_noise* noise, void* stream);

/* (a10) Ensemble average for one test batch (methods/sgld.py:283-305):
 *   logits_all [B,K,S] fp32 (contiguous, S fastest, i.e. torch.stack(outs, 2)) ->
 *   comp[B,K] = logsumexp_S(log_softmax_K(logits_all)) - log_S     (log_S = 0 when nst == 0)
 *   mode 0: out = comp                       (non-cyclical runners)
 *   mode 1: out = weight * comp              (first GMM component, methods/csgld.py:428-429)
 *   mode 2: out += weight * comp             (further components, :430-431; log-space mixture) */
int bdl_ensemble(const float* logits_all_dev, uint32_t B, uint32_t K, uint32_t S, float log_S, float weight,
                 int mode, float* out_logits_dev, void* stream);

/* CE loss sum and error count of [B,K] logits against int64 labels (methods/sgld.py:302-306,314-315):
 *   *loss_sum += sum_b -log_softmax(logits[b])[y[b]] ;  *err_count += sum_b (argmax_k logits[b,k] != y[b]).
 * Accumulating on the device removes the reference's two host syncs per batch. */
int bdl_ce_err(const float* logits_dev, const int64_t* y_dev, uint32_t B, uint32_t K, double* loss_sum_dev,
               int32_t* err_count_dev, void* stream);

/* Sample-sharded ensembles (section 8e): running logsumexp over samples of log_softmax_K(logits[B,K]), kept per
 * element as (m = running max, s = sum of exp(. - m)) so values whose probability underflows in linear space
 * combine exactly like the reference's logsumexp(log_softmax) (methods/sgld.py:300).  Initialise m = -inf, s = 0.
 * Across ranks: all-reduce(MAX) m -> m_global; bdl_lse_rescale; all-reduce(SUM) s; bdl_lse_finalize forms
 * comp = log(s) + m - log_S and applies the same mode / weight rule as bdl_ensemble. */
int bdl_lse_accum(const float* logits_dev, uint32_t B, uint32_t K, float* m_dev, float* s_dev, void* stream);
int bdl_lse_rescale(const float* m_local_dev, const float* m_global_dev, float* s_dev, uint64_t total, void* stream);
int bdl_lse_finalize(const float* m_dev, const float* s_dev, uint32_t B, uint32_t K, float log_S, float weight, int mode,
                     float* out_logits_dev, void* stream);

/* (a11) Calibration bins (calibration.py:24-67) and NLL (calibration.py:246-249).
 *   logits [N,K] fp32, labels [N] int64, edges [M] fp64 right bin boundaries (host: np.linspace(0,1+1e-8,M+1)[1:]).
 *   use_f64 == 0: logits/T and softmax in fp32 (temperature is the int 1 or a Python float in the reference)
 *   use_f64 != 0: logits/T and softmax in fp64 (temperature is an fp64 ndarray in the reference: Topt)
 *   Outputs (device, accumulated into -- caller zeroes): bin_size[M], acc_sum[M], conf_sum[M], nll_sum[1] (fp64)
 *   and near_edge[1] (optional, may be NULL): number of probabilities within 16 ulp of a bin edge; 0 certifies
 *   that any correctly-implemented softmax yields the same bin counts.
 *   binned[N*K] (optional int32, may be NULL): per-probability bin index, np.digitize's return value.
 *   Class-wise over all N*K probabilities; bin index = #edges <= p (np.digitize). */
int bdl_calibrate(const float* logits_dev, const int64_t* labels_dev, uint64_t N, uint32_t K, double temperature,
                  int use_f64, const double* edges_dev, uint32_t M, double* bin_size_dev, double* acc_sum_dev,
                  double* conf_sum_dev, double* nll_sum_dev, unsigned long long* near_edge_dev,
                  int32_t* binned_dev, void* stream);

/* (section 8f row 2) Bayesian model average over stored raw samples, one test batch
 * (methods/csghmc_fs.py:349-377): logits_all [B,K,S] fp32 (S fastest; S = models in sorted file order) ->
 *   out[B,K] = (((l_0 + l_1) + l_2) + ...) / fp32(S)      fp32 running sum in model order, one IEEE division. */
int bdl_bma_mean(const float* logits_all_dev, uint32_t B, uint32_t K, uint32_t S, float* out_logits_dev, void* stream);

/* (section 8f row 4) Temperature-scaling objective of find_optimal_temperature (calibration.py:178-184):
 *   row_nll[i] = logsumexp_k(logits[i,k] / T) - logits[i,y_i] / T   (fp64; `logits / T` promotes in the reference)
 *   out_mean[0] = mean_i row_nll[i]                                  (fixed reduction order: reproducible)
 * row_nll_dev: caller scratch of N doubles (also an output).  The scalar optimiser (scipy BFGS) stays on the host. */
int bdl_nll_temperature(const float* logits_dev, const int64_t* labels_dev, uint64_t N, uint32_t K, double temperature,
                        double* row_nll_dev, double* out_mean_dev, void* stream);

/* Test / diagnostics entry: exhaustive device self-test of the library's correctly rounded fp32 helpers (the
 * branch-free sqrt / reciprocal / quotient fast paths the Adam update rules use, methods/adam_sghmc.py:536-541) against
 * the CUDA intrinsics over all 2^32 bit patterns.  out6_dev: 6 x uint64 = mismatches {sqrt, rcp, div}, then the number
 * of inputs that took the fast path {sqrt, rcp, div}.  ~0.1 s on B200. */
int bdl_selftest_math(unsigned long long* out6_dev, void* stream);

/* Diagnostics entry: the bare-traffic yardstick of the streaming kernels.  Moves the bytes of a sampler kernel with next
 * to no arithmetic, in the product kernels' launch shape (one tile per CTA in address order, one 128-bit group per thread
 * and stream, evict-first stores):  (reads, writes) = (4, 2): a, b read and written, c, d read -- the SGHMC step's
 * 24 B/element (methods/sghmc.py:482-510 + :229);  (2, 1): c, d read, a written -- the posterior draw's 12 B/element
 * (methods/sgld.py:292-297);  (1, 1): c read, a written -- a copy.  bench.py times it next to the real kernels
 * (`roofline.bare_traffic_kernel`); no reference counterpart.  a and b are overwritten with finite averages. */
int bdl_probe_stream(float* a_dev, float* b_dev, const float* c_dev, const float* d_dev, uint64_t n, int reads, int writes,
                     int threads, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer form: a chain whose state is resident in HBM, stepped from HOST memory.
 * (The one place the library owns device memory.)  This is what a CPU-resident caller of the reference's
 * update loop binds: per step only the gradient (in) and theta (out) cross PCIe; theta0, momentum, Adam
 * moments and the SGD buffer never leave the device.  Pinned host buffers are required for the copies to be
 * asynchronous (cudaHostAlloc / torch pin_memory()).
 * ---------------------------------------------------------------------------------------------- */
typedef struct bdl_chain bdl_chain;
enum { BDL_BUF_THETA = 0, BDL_BUF_THETA0 = 1, BDL_BUF_V = 2, BDL_BUF_M = 3, BDL_BUF_S = 4, BDL_BUF_SGD = 5 };

/* n: padded-flat length; chunk_elems: pipeline chunk (0 = default 8 Mi elements). */
int bdl_chain_create(uint64_t n, int variant, int with_sgd_momentum, uint64_t chunk_elems, bdl_chain** out);
int bdl_chain_destroy(bdl_chain* chain);
int bdl_chain_upload(bdl_chain* chain, int which, const float* host);     /* synchronous */
int bdl_chain_download(bdl_chain* chain, int which, float* host);         /* synchronous */
int bdl_chain_device_ptr(bdl_chain* chain, int which, float** out);       /* which == 6: gradient staging buffer */
/* One sampler update: g_host (n floats, in) -> theta_out_host (n floats, out).  runs_host is a HOST array without
 * per-run gradient pointers; in-kernel Philox noise only.  Returns when theta_out_host is complete. */
int bdl_chain_step_host(bdl_chain* chain, const float* g_host, float* theta_out_host, const bdl_run* runs_host,
                        uint32_t nruns, const bdl_scalars* scalars, const bdl_noise* noise);

#ifdef __cplusplus
}
#endif
#endif /* BDL_H_ */
