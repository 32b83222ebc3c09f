// Support header for the GENERATED 3c2e kernels (codegen/gen3c2e.py).
//
// The same generated source compiles two ways:
//   * nvcc (default): real kernels, __syncwarp / __shfl_sync / atomicAdd;
//   * -DSI_G3_HOST_EMU with g++: TEST-ONLY warp emulation -- 32 host threads per warp, a
//     std::barrier for __syncwarp -- used by tests/emu to check the generated code without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#include "si_machine.h"

#if defined(SI_G3_HOST_EMU)
struct int2 {   // (built in under nvcc)
    int x, y;
};
#endif

struct BasisDev {
    const int* L;
    const int* nprim;
    const int* poff;
    const int* foff;  // function offsets (sph or cart, chosen by the host)
    const double* cen;
    const double* exps;
    const double* coef;
};

// doubles per primitive pair in the host-built pair table: p, 1/p, P[3], PA[3], exp(-mu AB^2) da db, exponent a,
// exponent b, (pad) -- 96 bytes, so that one coalesced load brings everything a kernel needs about the pair
#define SI_PP_STRIDE 12

struct SiPairItem {   // 64 bytes: everything the kernels need about a shell pair in ONE load
    int sa, sb;       // shell indices; L(sa) >= L(sb)
    int swapped;      // packed layout: 1 if sa is the column shell of the (i>=j) block
    int ppoff;        // first primitive pair of this shell pair in the precomputed pair-data array
    long long prow;   // first row of the pair block in the packed 3-index layout
    int Ka, Kb;       // contraction lengths
    int fa0, fb0;     // spherical function offsets of the two shells
    double AB[3];     // A - B
};

struct SiAuxItem {    // 48 bytes: one auxiliary shell of the class being processed
    int shell, Kc, pc0, foff;
    double C[3];
    double pad;
};

enum StoreMode { STORE_BLOCK = 0, STORE_MATRIX = 1, STORE_3C_FULL = 2, STORE_3C_PACKED = 3 };

struct OutDesc {
    double* out;
    long long ld;           // matrix: nbf ; 3c: naux_local
    long long comp_stride;  // matrix: nbf*nbf
    long long nbf;
    int mode;
};

struct G3Args {
    BasisDev orb, aux;
    const SiPairItem* pairs;
    const double* ppair;  // 9 doubles per primitive pair: p, 1/p, P[3], PA[3], exp(-mu AB^2) da db
    int npairs;
    const double* aux_oexp;  // 1/exponent of every aux primitive
    const int* auxlist;   // aux shells of this class (same Lc)
    const SiAuxItem* auxitems;  // the same shells as packed records (generated kernels)
    int naux_c;
    int col_begin;
    int groups_per_pair;  // ceil(naux_c / tuples-per-warp)
    long long ngroups;    // npairs * groups_per_pair, or the number of groups listed in `tuples`
    const int2* tuples;   // optional explicit tuple table: ngroups x tuples-per-warp entries (pair, aux), -1 = padding
    int chunk;            // unused (the kernels request one group per atomic); kept for ABI stability of the struct
    // opt-in primitive pre-screening (0 = off, the parity / benchmark mode): a primitive triple is skipped when the
    // square of its (ss|s)-type base integral bound is <= screen_eps2 for every tuple of the warp -- the pPRE2
    // estimate of sympleints/graphs/templates/int_3c2e_func.tpl:46-54 (a no-op `continue` in the reference)
    double screen_eps2;
    OutDesc od;
    int* counter;
    SiBoys boys;
};

// generated nuclear-attraction kernels (codegen/gencoul.py): one tuple per shell pair, all charges inside
struct GCArgs {
    const SiPairItem* pairs;
    const double* ppair;   // as in G3Args
    int npairs;
    const double* c4;      // ncharge packed records (x, y, z, q)
    int ncharge;
    int nsplit;            // charge ranges per pair: split 0 writes od.out, split sp >= 1 slice sp-1 of `part`
    double* part;          // (nsplit_max - 1) zero-initialised nbf x nbf slices, summed into od.out afterwards
    long long ngroups;     // ceil(npairs / pairs-per-warp) * nsplit
    OutDesc od;            // matrix: out[(fa0+ia) * ld + fb0+jb]
    int* counter;
    SiBoys boys;
};

#if defined(SI_G3_HOST_EMU)
// ------------------------------------------------------------------ host warp emulation
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <barrier>
using std::max;
using std::min;
struct SiG3EmuWarp {
    std::barrier<>* bar;
    volatile long long xchg[32];
};
extern thread_local SiG3EmuWarp* si_g3_emu_warp;
extern thread_local int si_g3_emu_lane;
#define __device__
#define SI_G3_DEV static inline
// bounds-checked view of the tuple's shared-memory region (host emulation only)
struct SiG3CheckedW {
    double* p;
    int n;
    double& operator[](long long i) const {
        if (i < 0 || i >= n) {
            fprintf(stderr, "g3 emulation: shared-memory index %lld outside [0,%d)\n", i, n);
            abort();
        }
        return p[i];
    }
};
#define SI_G3_WTYPE const SiG3CheckedW
#define SI_G3_WINIT(ptr, n) SiG3CheckedW{ptr, n}
#define SI_G3_GLOBAL(mb) static inline
#define SI_G3_KERNEL_PROLOGUE                                   \
    extern thread_local double* si_g3_emu_smem;                 \
    double* g3_smem = si_g3_emu_smem;                           \
    const int g3_lane = si_g3_emu_lane, g3_warp = 0;           \
    const long long g3_first = 0, g3_stride = 1;
static inline void si_g3_syncwarp() { si_g3_emu_warp->bar->arrive_and_wait(); }
static inline double si_g3_rsqrt(double x) { return 1.0 / sqrt(x); }
static inline long long si_g3_next_group(int* counter, int lane) {
    if (lane == 0) si_g3_emu_warp->xchg[0] = __atomic_fetch_add(counter, 1, __ATOMIC_SEQ_CST);
    si_g3_syncwarp();
    long long v = si_g3_emu_warp->xchg[0];
    si_g3_syncwarp();
    return v;
}
static inline int si_g3_fetch_group(int* counter, int lane) {
    return lane == 0 ? __atomic_fetch_add(counter, 1, __ATOMIC_SEQ_CST) : 0;
}
static inline int si_g3_tri_row(int ai) {
    int r = 0;
    while ((r + 1) * (r + 2) / 2 <= ai) ++r;
    return r;
}
static inline long long si_g3_bcast_group(int v) {
    if (si_g3_emu_lane == 0) si_g3_emu_warp->xchg[0] = v;
    si_g3_syncwarp();
    long long r = si_g3_emu_warp->xchg[0];
    si_g3_syncwarp();
    return r;
}
static inline bool si_g3_warp_all(bool p) {
    si_g3_emu_warp->xchg[si_g3_emu_lane] = p ? 1 : 0;
    si_g3_syncwarp();
    bool a = true;
    for (int i = 0; i < 32; ++i) a = a && si_g3_emu_warp->xchg[i] != 0;
    si_g3_syncwarp();
    return a;
}
static inline int si_g3_warp_max(int v) {
    si_g3_emu_warp->xchg[si_g3_emu_lane] = v;
    si_g3_syncwarp();
    int m = v;
    for (int i = 0; i < 32; ++i) m = (int)si_g3_emu_warp->xchg[i] > m ? (int)si_g3_emu_warp->xchg[i] : m;
    si_g3_syncwarp();
    return m;
}
#else
// ------------------------------------------------------------------ device
#define SI_G3_DEV __device__ __forceinline__
#define SI_G3_WTYPE double* const
#define SI_G3_WINIT(ptr, n) (ptr)
#define SI_G3_GLOBAL(mb) __global__ __launch_bounds__(128, mb)
#define SI_G3_KERNEL_PROLOGUE              \
    asm volatile("griddepcontrol.launch_dependents;");   /* see g3_launch_part_*: dependents may start at once */ \
    extern __shared__ double g3_smem[];    \
    const int g3_lane = threadIdx.x & 31, g3_warp = threadIdx.x >> 5;                         \
    const long long g3_first = (long long)blockIdx.x * (blockDim.x >> 5) + g3_warp,            \
                    g3_stride = (long long)gridDim.x * (blockDim.x >> 5);
__device__ __forceinline__ void si_g3_syncwarp() { __syncwarp(); }
// one-wave grid: at most `occ` resident blocks per SM (the grid passed in is already capped at 4 per SM)
static inline int si_g3_grid(int grid, int occ) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return grid < sms * occ ? grid : sms * occ;
}
__device__ __forceinline__ double si_g3_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ long long si_g3_next_group(int* counter, int lane) {
    int v = 0;
    if (lane == 0) v = atomicAdd(counter, 1);
    return (long long)__shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ int si_g3_fetch_group(int* counter, int lane) {
    // Inline PTX on purpose: for atomicAdd() nvcc emits its warp-aggregation pattern (leader election,
    // ATOMG, then a SHFL that broadcasts the result at once), i.e. the warp waits for the round trip here
    // and the prefetch is lost.  The raw instruction leaves the result untouched until the caller needs it.
    int v = 0;
    if (lane == 0) asm volatile("atom.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(v) : "l"(counter) : "memory");
    return v;
}
__device__ __forceinline__ long long si_g3_bcast_group(int v) {
    return (long long)__shfl_sync(0xffffffffu, v, 0);
}
// row r of the triangular index ai = r (r + 1) / 2 + c, 0 <= c <= r (ai < 2^20: exact in float)
__device__ __forceinline__ int si_g3_tri_row(int ai) {
    return (int)((__fsqrt_rn(8.0f * (float)ai + 1.0f) - 1.0f) * 0.5f);
}
__device__ __forceinline__ int si_g3_warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ bool si_g3_warp_all(bool p) { return __all_sync(0xffffffffu, p) != 0; }
#endif

// addressing of the elements (ia = 0..nsa-1, jb, m) of the (a_sph, b_sph, c_sph) block of pair item `it`:
// element ia at p0[ia * s0]; in the full layout the mirror image (b, a) at p1[ia * s1] (p1 = null otherwise)
SI_HD void si_g3_store_setup(const OutDesc& od, const SiPairItem& it, int fa0, int fb0, int nsa, int nsb, int nm,
                             long long col, int jb, int m, double** p0, long long* s0, double** p1, long long* s1) {
    if (od.mode == STORE_BLOCK) {
        *p0 = od.out + (long long)jb * nm + m;
        *s0 = (long long)nsb * nm;
    } else if (od.mode == STORE_3C_FULL) {
        *p0 = od.out + ((long long)fa0 * od.nbf + (fb0 + jb)) * od.ld + col + m;
        *s0 = od.nbf * od.ld;
        if (it.sa != it.sb) {
            *p1 = od.out + ((long long)(fb0 + jb) * od.nbf + fa0) * od.ld + col + m;
            *s1 = od.ld;
        }
    } else if (it.swapped) {
        *p0 = od.out + (it.prow + (long long)jb * nsa) * od.ld + col + m;
        *s0 = od.ld;
    } else {
        *p0 = od.out + (it.prow + jb) * od.ld + col + m;
        *s0 = (long long)nsb * od.ld;
    }
}

// element (ia, jb, m) of the (a_sph, b_sph, c_sph) block of pair item `it`
SI_HD void si_g3_store(const OutDesc& od, const SiPairItem& it, int fa0, int fb0, int nsa, int nsb, int nm,
                       long long col, int ia, int jb, int m, double val) {
    if (od.mode == STORE_BLOCK) {
        od.out[((long long)ia * nsb + jb) * nm + m] = val;
    } else if (od.mode == STORE_3C_FULL) {
        od.out[((long long)(fa0 + ia) * od.nbf + (fb0 + jb)) * od.ld + col + m] = val;
        if (it.sa != it.sb) od.out[((long long)(fb0 + jb) * od.nbf + (fa0 + ia)) * od.ld + col + m] = val;
    } else {
        const long long row = it.swapped ? it.prow + (long long)jb * nsa + ia : it.prow + (long long)ia * nsb + jb;
        od.out[row * od.ld + col + m] = val;
    }
}
