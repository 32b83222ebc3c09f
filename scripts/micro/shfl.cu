// Throughput probes on sm_100a: SHFL vs LDS/STS vs DFMA, alone and mixed (one SM, 16 warps).
// Question answered: is a register + shuffle exchange cheaper than a shared-memory round trip, and do the
// two contend for the same pipe?  Output: SM cycles per warp-instruction group.
#include <cstdio>
#include <cuda_runtime.h>

#define UNROLL 8

template <int NF, int NS, int NL, int NT>   // per iteration: NF DFMA, NS SHFL.32, NL LDS.64, NT STS.64 (per chain)
__global__ void __launch_bounds__(512) mix(double* out, long long* cyc, int iters) {
    __shared__ double w[16 * 32 * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* mine = w + warp * 256;
    for (int i = lane; i < 256; i += 32) mine[i] = 1.0 + i * 1e-6;
    __syncthreads();
    double a[UNROLL];
    int s[UNROLL];
    for (int i = 0; i < UNROLL; ++i) { a[i] = lane * 1e-9 + i; s[i] = lane + i; }
    const double m = 1.0000001, c = 1e-9;
    const int src = (lane + 1) & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int k = 0; k < NF; ++k) a[u] = fma(a[u], m, c);
#pragma unroll
            for (int k = 0; k < NS; ++k) s[u] = __shfl_sync(0xffffffffu, s[u], src) + 1;
#pragma unroll
            for (int k = 0; k < NL; ++k) a[u] += mine[((it + k) & 7) * 32 + ((lane + u) & 31)];
#pragma unroll
            for (int k = 0; k < NT; ++k) mine[((it + k + 3) & 7) * 32 + ((lane + u) & 31)] = a[u];
        }
    }
    long long t1 = clock64();
    double r = 0;
    for (int i = 0; i < UNROLL; ++i) r += a[i] + s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// double exchanged with the neighbour lane by two SHFL.32, then one DFMA: the z-column recurrence node
template <int NF>
__global__ void __launch_bounds__(512) node_shfl(double* out, long long* cyc, int iters) {
    const int lane = threadIdx.x & 31;
    double a[UNROLL];
    for (int i = 0; i < UNROLL; ++i) a[i] = lane * 1e-9 + i;
    const int src = (lane + 31) & 31;
    const double pc = 0.999;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            double nb = __shfl_sync(0xffffffffu, a[u], src);
            a[u] = fma(pc, a[u], nb * 1e-3);
#pragma unroll
            for (int k = 1; k < NF; ++k) a[u] = fma(a[u], pc, 1e-9);
        }
    }
    long long t1 = clock64();
    double r = 0;
    for (int i = 0; i < UNROLL; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* d; long long* c;
    cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 8);
    long long h;
    const int iters = 2000;
    const int nw = 16;
#define RUN(K, label, groups) K<<<1, nw * 32>>>(d, c, iters); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("%-58s %7.3f SM-cyc per warp-group (%s)\n", label, (double)h / ((double)iters * UNROLL * nw), groups);
    RUN((mix<1, 0, 0, 0>), "1 DFMA", "expect 0.5");
    RUN((mix<0, 1, 0, 0>), "1 SHFL.32 (+IADD)", "?");
    RUN((mix<0, 2, 0, 0>), "2 SHFL.32", "?");
    RUN((mix<0, 0, 1, 0>), "1 LDS.64 (+DADD)", "expect 2");
    RUN((mix<0, 0, 0, 1>), "1 STS.64", "expect 2");
    RUN((mix<0, 0, 1, 1>), "LDS.64 + STS.64", "expect 4");
    RUN((mix<2, 2, 0, 0>), "2 DFMA + 2 SHFL", "overlap?");
    RUN((mix<4, 2, 0, 0>), "4 DFMA + 2 SHFL", "overlap?");
    RUN((mix<0, 2, 1, 0>), "2 SHFL + 1 LDS.64", "same pipe -> additive");
    RUN((mix<0, 2, 1, 1>), "2 SHFL + LDS.64 + STS.64", "");
    RUN((mix<2, 0, 1, 1>), "2 DFMA + LDS.64 + STS.64", "current D-tree node");
    RUN((mix<4, 0, 1, 0>), "4 DFMA + LDS.64", "");
    RUN((node_shfl<1>), "node: shfl(double) + DMUL + DFMA", "z-column x/y step");
    RUN((node_shfl<3>), "node: shfl(double) + DMUL + 3 DFMA", "");
    return 0;
}
