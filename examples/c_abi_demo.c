/*
 * c_abi_demo.c -- libbdl from plain C: no Python, no torch.  One fused SGHMC update (injected noise, IEEE division) on a
 * small flat state, checked on the host against the reference's arithmetic (methods/sghmc.py:482-510 + SGD.step :229)
 * written out in C with one rounding per operation.
 *
 *   gcc examples/c_abi_demo.c -Iinclude -I/usr/local/cuda/include -Lbayesdll_b200 -lbdl -L/usr/local/cuda/lib64 -lcudart \
 *       -lm -ffp-contract=off -Wl,-rpath,$PWD/bayesdll_b200 -o /tmp/c_abi_demo && /tmp/c_abi_demo
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bdl.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

static float frand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) / 16777216.0f) - 0.5f; }

int main(void) {
    const uint64_t n = 4096 + 64;              /* body 4096 elements, head 64 elements */
    const size_t bytes = n * sizeof(float);
    float *theta = malloc(bytes), *g = malloc(bytes), *theta0 = malloc(bytes), *v = malloc(bytes), *xi = malloc(bytes);
    float *want_t = malloc(bytes), *want_v = malloc(bytes), *got_t = malloc(bytes), *got_v = malloc(bytes);
    unsigned seed = 7;
    for (uint64_t i = 0; i < n; ++i) {
        theta[i] = frand(&seed); g[i] = 0.1f * frand(&seed); theta0[i] = frand(&seed); v[i] = 0.01f * frand(&seed); xi[i] = 4.0f * frand(&seed);
    }
    if (bdl_abi_version() != BDL_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 2; }

    /* hyper-parameters as the reference's Python doubles; rounded to fp32 where they meet a tensor */
    const double ND = 1840, Ninflate = 3, prior_sig = 0.9, nd = 0.7, alpha = 0.18, lr[2] = {1e-3, 2e-2};
    const double N = ND * Ninflate;
    bdl_scalars sc;
    memset(&sc, 0, sizeof sc);
    for (int h = 0; h < 2; ++h) { sc.lr[h] = (float)lr[h]; sc.noise_scale[h] = (float)(nd * sqrt(2 * alpha / (N * lr[h]))); }
    sc.one_minus_alpha = (float)(1 - alpha);
    sc.sig2 = (float)(prior_sig * prior_sig);
    sc.N = (float)N;
    sc.div_mode = BDL_DIV_IEEE;

    bdl_run runs[2];
    memset(runs, 0, sizeof runs);
    runs[0].begin = 0;    runs[0].end = 4096; runs[0].valid_end = 4096; runs[0].cls = BDL_CLS_PRIOR;
    runs[1].begin = 4096; runs[1].end = n;    runs[1].valid_end = n;    runs[1].cls = BDL_CLS_PRIOR | BDL_CLS_HEAD;

    /* host restatement, one rounding per op (compile with -ffp-contract=off); the SGD step is the one fused op */
    for (uint64_t i = 0; i < n; ++i) {
        const int h = i >= 4096;
        float d = theta[i] - theta0[i];
        d = d / sc.sig2;
        d = d / sc.N;
        const float gU = g[i] + d;
        float t1 = v[i] * sc.one_minus_alpha, t2 = sc.lr[h] * gU, t3 = sc.noise_scale[h] * xi[i];
        float vn = t1 + t2;
        vn = vn + t3;
        const float gp = g[i] + vn;
        want_v[i] = vn;
        want_t[i] = fmaf(gp, -sc.lr[h], theta[i]);
    }

    float *d_theta, *d_g, *d_theta0, *d_v, *d_xi;
    CHECK_CUDA(cudaMalloc((void**)&d_theta, bytes)); CHECK_CUDA(cudaMalloc((void**)&d_g, bytes));
    CHECK_CUDA(cudaMalloc((void**)&d_theta0, bytes)); CHECK_CUDA(cudaMalloc((void**)&d_v, bytes));
    CHECK_CUDA(cudaMalloc((void**)&d_xi, bytes));
    CHECK_CUDA(cudaMemcpy(d_theta, theta, bytes, cudaMemcpyHostToDevice)); CHECK_CUDA(cudaMemcpy(d_g, g, bytes, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(d_theta0, theta0, bytes, cudaMemcpyHostToDevice)); CHECK_CUDA(cudaMemcpy(d_v, v, bytes, cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(d_xi, xi, bytes, cudaMemcpyHostToDevice));

    bdl_noise nz;
    memset(&nz, 0, sizeof nz);
    nz.xi_dev = d_xi;                          /* injected noise; xi_dev = NULL would use in-kernel Philox */
    /* two runs without gradient pointers: the table rides in the kernel arguments, no device copy needed */
    int rc = bdl_step(BDL_SGHMC, d_theta, d_g, d_theta0, d_v, NULL, NULL, NULL, n, NULL, 2, runs, &sc, &nz, NULL);
    if (rc != BDL_OK) { fprintf(stderr, "bdl_step: %d %s\n", rc, bdl_last_error()); return 1; }
    CHECK_CUDA(cudaDeviceSynchronize());
    CHECK_CUDA(cudaMemcpy(got_t, d_theta, bytes, cudaMemcpyDeviceToHost));
    CHECK_CUDA(cudaMemcpy(got_v, d_v, bytes, cudaMemcpyDeviceToHost));
    uint64_t bad = 0;
    for (uint64_t i = 0; i < n; ++i) bad += memcmp(&got_t[i], &want_t[i], 4) != 0 || memcmp(&got_v[i], &want_v[i], 4) != 0;
    /* error convention: status code + message, no exception */
    rc = bdl_step(BDL_SGHMC, d_theta, d_g, d_theta0, NULL, NULL, NULL, NULL, n, NULL, 2, runs, &sc, &nz, NULL);
    if (rc != BDL_ERR_INVALID) { fprintf(stderr, "expected BDL_ERR_INVALID, got %d\n", rc); return 1; }
    printf("c_abi_demo: %llu of %llu elements differ from the host restatement; missing-momentum call -> %d (%s)\n",
           (unsigned long long)bad, (unsigned long long)n, rc, bdl_last_error());
    cudaFree(d_theta); cudaFree(d_g); cudaFree(d_theta0); cudaFree(d_v); cudaFree(d_xi);
    puts(bad == 0 ? "OK" : "MISMATCH");
    return bad == 0 ? 0 : 1;
}
