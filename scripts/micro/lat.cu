// FP64 / shared-memory latency and throughput probes on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int iters) {
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lds_chase(int* out, long long* cyc, int iters) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 33) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) p = idx[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// LDS.64 -> DMUL -> DFMA -> STS.64 dependent through memory (like one recurrence node)
__global__ void node_chain(double* out, long long* cyc, int iters) {
    __shared__ double w[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) w[i] = 1.0 + i * 1e-6;
    __syncthreads();
    int lane = threadIdx.x & 31;
    double r = 1.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double nb = w[((it & 31) * 32 + lane + 1) & 2047];
        r = fma(0.5, nb, 1.0000001 * r);
        w[(((it + 1) & 31) * 32 + lane) & 2047] = r;
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = r;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* d; long long* c; int* di;
    cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 8); cudaMalloc(&di, 4096);
    long long h;
    const int iters = 2000;
#define RUN(K, grid, block, ops, label) K<<<grid, block>>>(d, c, iters); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-40s %8.2f cyc per op-group\n", label, (double)h / (ops));
    RUN(dfma_chain<1>, 1, 32, iters * 16.0, "DFMA dependent latency (1 warp, ILP1)");
    RUN(dfma_chain<2>, 1, 32, iters * 16.0, "DFMA 2 chains: cyc per 2 DFMA");
    RUN(dfma_chain<4>, 1, 32, iters * 16.0, "DFMA 4 chains: cyc per 4 DFMA");
    RUN(dfma_chain<8>, 1, 32, iters * 16.0, "DFMA 8 chains: cyc per 8 DFMA");
    RUN(dfma_chain<1>, 1, 128, iters * 16.0, "4 warps ILP1 (1/SMSP): cyc per DFMA/warp");
    RUN(dfma_chain<1>, 1, 512, iters * 16.0, "16 warps ILP1 (4/SMSP): cyc per DFMA/warp");
    RUN(dfma_chain<4>, 1, 512, iters * 16.0, "16 warps ILP4: cyc per 4 DFMA/warp");
    lds_chase<<<1, 32>>>(di, c, iters); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-40s %8.2f cyc\n", "LDS dependent latency", (double)h / iters);
    RUN(node_chain, 1, 32, (double)iters, "LDS->DMUL->DFMA->STS->syncwarp (1 warp)");
    RUN(node_chain, 1, 512, (double)iters, "same, 16 warps per SM");
    return 0;
}
