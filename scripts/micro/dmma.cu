// Is the FP64 tensor-core path (mma.sync m8n8k4 / m16n8k8 .f64) a second FP64 pipe on B200, or the same
// units as DFMA?  Full-chip throughput of DMMA alone, DFMA alone, and both interleaved.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int MODE>   // 0: DFMA, 1: DMMA m8n8k4, 2: both, 3: DMMA m16n8k8, 4: m16n8k8 + DFMA
__global__ void __launch_bounds__(256) k(double* out, int iters) {
    double c[8][4], f[8];
    for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-9 + i; for (int j = 0; j < 4; ++j) c[i][j] = i + j; }
    double a[4] = {1.0000001, 0.5, 0.25, 0.125}, b[2] = {1e-9 + threadIdx.x * 1e-12, 2e-9};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 2 || MODE == 4) {
                f[u] = fma(f[u], a[0], b[0]);
                f[u] = fma(f[u], a[0], b[0]);
            }
            if (MODE == 1 || MODE == 2) dmma884(c[u][0], c[u][1], a[0], b[0]);
            if (MODE == 3 || MODE == 4) dmma1688(c[u], a, b);
        }
    }
    double r = 0;
    for (int i = 0; i < 8; ++i) { r += f[i]; for (int j = 0; j < 4; ++j) r += c[i][j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
    double* d;
    cudaMalloc(&d, 1 << 24);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000, grid = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[5] = {"DFMA x2 per slot", "DMMA m8n8k4", "DMMA m8n8k4 + 2 DFMA", "DMMA m16n8k8", "DMMA m16n8k8 + 2 DFMA"};
    // flops per thread-iteration-slot: DFMA 2*2; m8n8k4: 8*8*4*2/32 = 16; m16n8k8: 16*8*8*2/32 = 64
    const double fl[5] = {4, 16, 20, 64, 68};
    for (int mode = 0; mode < 5; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            switch (mode) {
                case 0: k<0><<<grid, 256>>>(d, iters); break;
                case 1: k<1><<<grid, 256>>>(d, iters); break;
                case 2: k<2><<<grid, 256>>>(d, iters); break;
                case 3: k<3><<<grid, 256>>>(d, iters); break;
                case 4: k<4><<<grid, 256>>>(d, iters); break;
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = fl[mode] * 8.0 * iters * 256.0 * grid;
        printf("%-26s %8.3f ms  %7.2f TFLOP/s  (%s)\n", names[mode], ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
