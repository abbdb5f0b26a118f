// Probe: can SM-issued loads/stores on mapped pinned host memory beat the copy engines for the duplex host-buffer step?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/zero_copy_probe.cu -o gpurun_out/zero_copy_probe && gpurun_out/zero_copy_probe
// Reads g (1.22 GB, host) and writes theta (1.22 GB, host) in one kernel, like a fused update with host-resident g / theta.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void duplex(const float4* __restrict__ g_host, float4* __restrict__ theta_host, const float4* __restrict__ v_dev,
                       float4* __restrict__ theta_dev, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 g = __ldcs(g_host + i), v = v_dev[i], t = theta_dev[i];
        t.x += g.x * v.x; t.y += g.y * v.y; t.z += g.z * v.z; t.w += g.w * v.w;
        theta_dev[i] = t;
        __stcs(theta_host + i, t);
    }
}
__global__ void read_only(const float4* __restrict__ g_host, float4* __restrict__ dev, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) dev[i] = __ldcs(g_host + i);
}
__global__ void write_only(const float4* __restrict__ dev, float4* __restrict__ t_host, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(t_host + i, dev[i]);
}

int main() {
    const size_t n = 305548328, n4 = n / 4, bytes = n * 4;
    float *g_h, *t_h, *v_d, *t_d;
    CK(cudaHostAlloc(&g_h, bytes, cudaHostAllocMapped)); CK(cudaHostAlloc(&t_h, bytes, cudaHostAllocMapped));
    CK(cudaMalloc(&v_d, bytes)); CK(cudaMalloc(&t_d, bytes));
    for (size_t i = 0; i < n; i += 1024) g_h[i] = 1.0f;
    CK(cudaMemset(v_d, 0, bytes)); CK(cudaMemset(t_d, 0, bytes));
    float4 *g_m, *t_m;
    CK(cudaHostGetDevicePointer(&g_m, g_h, 0)); CK(cudaHostGetDevicePointer(&t_m, t_h, 0));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int grid : {148, 148 * 4, 148 * 16, 148 * 64}) for (int threads : {128, 512}) {
        float ms[3];
        for (int k = 0; k < 3; ++k) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (k == 0) duplex<<<grid, threads>>>(g_m, t_m, (float4*)v_d, (float4*)t_d, n4);
                if (k == 1) read_only<<<grid, threads>>>(g_m, (float4*)t_d, n4);
                if (k == 2) write_only<<<grid, threads>>>((float4*)t_d, t_m, n4);
                cudaEventRecord(b); CK(cudaEventSynchronize(b));
                cudaEventElapsedTime(&ms[k], a, b);
            }
        }
        printf("grid %5d x %3d: duplex %.2f ms (%.1f GB/s each way)  read-only %.2f ms (%.1f GB/s)  write-only %.2f ms (%.1f GB/s)\n",
               grid, threads, ms[0], bytes / ms[0] / 1e6, ms[1], bytes / ms[1] / 1e6, ms[2], bytes / ms[2] / 1e6);
    }
    return 0;
}
