// Micro-benchmarks that decide the layout of the order-1 node-per-thread kernel:
//  * cost of LDS.128 / LDS.64 when the lanes of a warp read distinct, quarter-warp-uniform or warp-uniform addresses
//  * DFMA dependent-chain latency and the throughput one warp reaches with 1..8 independent chains
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench scripts/ubench/smem_fp64.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE, int WIDTH>
__global__ void lds_kernel(int iters, long long* cycles, double* sink) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1.0 + i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // byte offset of this lane's first access
    int off;
    if (MODE == 0) off = lane * (WIDTH / 8);                 // distinct, contiguous
    else if (MODE == 1) off = (lane >> 3) * 2 * 37;          // quarter-warp uniform: four addresses at odd 16-byte strides
    else if (MODE == 2) off = 0;                             // warp uniform
    else off = (lane >> 3) * 16 * 2 + 0;                     // quarter-warp uniform, the four addresses 256 B apart (same banks)
    off += warp * 8;                                         // warps use different rows
    double acc0 = 0, acc1 = 0;
    const double* base = sm + off;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const double* p = base + ((u * 34 + it) & 255) * 2 * ((MODE == 0) ? 1 : 1);
            if (WIDTH == 128) {
                double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
                acc0 += v.x; acc1 += v.y;
            } else {
                double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
                acc0 += v;
            }
        }
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
}

// pure load issue without the dependent adds: loads into distinct registers, one xor-ish combine at the end of the unroll
template <int MODE>
__global__ void lds128_only_kernel(int iters, long long* cycles, double* sink) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1.0 + i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int off;
    if (MODE == 0) off = lane * 2;
    else if (MODE == 1) off = (lane >> 3) * 2 * 37;
    else if (MODE == 2) off = 0;
    else off = (lane >> 3) * 32;
    off += warp * 8;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm + off);
    unsigned long long x = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        unsigned long long r[32];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(r[2 * u]), "=l"(r[2 * u + 1]) : "r"(base + (((u * 34 + it) & 255) << 4)));
        }
#pragma unroll
        for (int u = 0; u < 32; ++u) x ^= r[u];
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = (double)x;
}

template <int ILP>
__global__ void dfma_kernel(int iters, long long* cycles, double* sink, double m, double c) {
    double a[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) a[j] = threadIdx.x * 1e-3 + j;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int j = 0; j < ILP; ++j) a[j] = fma(a[j], m, c);
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += a[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void drcp_kernel(int iters, long long* cycles, double* sink, double z0) {
    double z = z0 + threadIdx.x * 1e-3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) z = __drcp_rn(z) + 0.5;
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = z;
}

static double report(const char* name, long long* d_cyc, int nwarps, int n_ops) {
    long long h[64];
    cudaMemcpy(h, d_cyc, sizeof(long long) * nwarps, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < nwarps; ++i) if (h[i] > mx) mx = h[i];
    double per = (double)mx / n_ops;
    printf("%-58s warps/SM %2d  cycles per warp-instruction %.2f   (SM-wide cycles per instruction %.2f)\n", name, nwarps, per, per / nwarps);
    return per;
}

int main() {
    long long* d_cyc; double* d_sink;
    CK(cudaMalloc(&d_cyc, 64 * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 2048 * sizeof(double)));
    const int iters = 2000;
    const int smem = 8192 * 8;
    int warps_list[] = {1, 4, 8, 16};
#define RUN_LDS(MODE, WIDTH, label) \
    for (int w : warps_list) { \
        CK(cudaFuncSetAttribute(lds_kernel<MODE, WIDTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        lds_kernel<MODE, WIDTH><<<1, 32 * w, smem>>>(iters, d_cyc, d_sink); CK(cudaDeviceSynchronize()); \
        report(label, d_cyc, w, iters * 16); }
    RUN_LDS(0, 128, "LDS.128 distinct contiguous (+2 DADD)")
    RUN_LDS(1, 128, "LDS.128 quarter-warp uniform, 4 addresses (+2 DADD)")
    RUN_LDS(3, 128, "LDS.128 quarter-warp uniform, 4 addresses same banks (+2 DADD)")
    RUN_LDS(2, 128, "LDS.128 warp uniform (+2 DADD)")
    RUN_LDS(0, 64, "LDS.64 distinct contiguous (+1 DADD)")
    RUN_LDS(1, 64, "LDS.64 quarter-warp uniform (+1 DADD)")
    RUN_LDS(2, 64, "LDS.64 warp uniform (+1 DADD)")
#define RUN_LDSO(MODE, label) \
    for (int w : warps_list) { \
        CK(cudaFuncSetAttribute(lds128_only_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        lds128_only_kernel<MODE><<<1, 32 * w, smem>>>(iters, d_cyc, d_sink); CK(cudaDeviceSynchronize()); \
        report(label, d_cyc, w, iters * 16); }
    RUN_LDSO(0, "LDS.128 only, distinct contiguous")
    RUN_LDSO(1, "LDS.128 only, quarter-warp uniform (4 addresses)")
    RUN_LDSO(3, "LDS.128 only, quarter-warp uniform (4 addr, same banks)")
    RUN_LDSO(2, "LDS.128 only, warp uniform")
#define RUN_DFMA(ILP) \
    for (int w : {1, 4, 8, 16}) { \
        dfma_kernel<ILP><<<1, 32 * w>>>(iters, d_cyc, d_sink, 0.999, 1e-3); CK(cudaDeviceSynchronize()); \
        char lab[64]; snprintf(lab, 64, "DFMA, %d independent chains per thread", ILP); \
        report(lab, d_cyc, w, iters * 16 * ILP); }
    RUN_DFMA(1) RUN_DFMA(2) RUN_DFMA(4) RUN_DFMA(8)
    for (int w : {1, 4, 8}) {
        drcp_kernel<<<1, 32 * w>>>(iters, d_cyc, d_sink, 1.3); CK(cudaDeviceSynchronize());
        report("__drcp_rn + DADD dependent chain", d_cyc, w, iters * 16);
    }
    return 0;
}
