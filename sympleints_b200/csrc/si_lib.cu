// sympleints_b200 -- CUDA kernels and C ABI (see include/sympleints_b200.h).
//
// Kernels: one warp per shell tuple ("slot").  A slot owns a work array W and a
// scalar table S in shared memory and executes the stage program of its class
// (si_machine.h): per primitive tuple the Boys values and geometry scalars are
// staged into shared memory, the vertical recurrences run stage by stage with
// __syncwarp() in between, the primitive contraction accumulates into X, then
// cart->sph on c, the horizontal recurrence and the remaining cart->sph
// transforms run once per contracted tuple.  Work is handed out through a
// global atomic counter (cost-sorted item lists) so that unevenly contracted
// shells do not leave SMs idle.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/sympleints_b200.h"
#include "si_g3.h"
#include "si_machine.h"
#include "si_program.h"
#include "generated/g3_table.inc"
#include "generated/gc_table.inc"
#include "si_onee.cuh"
#include "si_gto.cuh"
#include "si_schwarz.cuh"

// ---------------------------------------------------------------------------
// device-side data
// ---------------------------------------------------------------------------
__device__ __forceinline__ long long next_item(int* counter, int lane) {
    int v = 0;
    if (lane == 0) v = atomicAdd(counter, 1);
    return (long long)__shfl_sync(0xffffffffu, v, 0);
}

__device__ __forceinline__ void load_consts(const SiProg& P, double* S, int lane) {
    for (int c = lane; c < P.nconst; c += 32) S[SI_NDYN + c] = P.consts[c];
}

__device__ __forceinline__ void contracted_phase(const SiProg& P, const SiStage* stg, double* W, double* S,
                                                 const double* A, const double* B, int lane) {
    auto sync = [] { __syncwarp(); };
    __syncwarp();
    {
        double v[12], kb[5];
#pragma unroll
        for (int q = 0; q < 12; ++q) v[q] = 0.0;
#pragma unroll
        for (int q = 0; q < 5; ++q) kb[q] = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) v[6 + d] = A[d] - B[d];
        si_fill_scalars(S, v, kb, 0.0, lane, 32);
    }
    si_run_stages(P, stg, P.n_prim_stages, P.chunk_begin, W, S, 0, 1, lane, 32, sync);
    for (int m0 = 0; m0 < P.nb_total; m0 += P.chunk) {
        const int nb = min(P.chunk, P.nb_total - m0);
        si_run_stages(P, stg, P.chunk_begin, P.nstages, W, S, m0, nb, lane, 32, sync);
    }
    __syncwarp();
}

// Store the final block W[out_off ...] of a 2-index key.
__device__ __forceinline__ void store_matrix(const SiProg& P, const double* W, const OutDesc& od, int fa,
                                             int fb, bool diag, int lane) {
    const int n0 = P.out_n0, n1 = P.out_n1, nb = P.nb_total;
    const int n01 = n0 * n1;
    const int total = n01 * nb;
    for (int idx = lane; idx < total; idx += 32) {
        const double val = W[P.out_off + idx];
        if (od.mode == STORE_BLOCK) {
            od.out[idx] = val;
            continue;
        }
        int b, i, j;
        if (P.out_batch_major) {
            b = idx / n01;
            int r = idx - b * n01;
            i = r / n1;
            j = r - i * n1;
        } else {
            b = idx % nb;
            int r = idx / nb;
            i = r / n1;
            j = r - i * n1;
        }
        double* o = od.out + (long long)b * od.comp_stride;
        o[(long long)(fa + i) * od.ld + (fb + j)] = val;
        if (!diag) o[(long long)(fb + j) * od.ld + (fa + i)] = val;
    }
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// 1e (ovlp, kin, dpm, qpm, dqpm) and 2c2e: one warp per shell pair.
__global__ void __launch_bounds__(256) k_two_index(SiProg P, SiBoys Bt, BasisDev bs, const SiPairItem* items,
                                                   int nitems, int is_2c2e, double Rx, double Ry, double Rz,
                                                   OutDesc od, int* counter) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (P.w_size + SI_NDYN + P.nconst + 1) & ~1;
    SiStage* stg = reinterpret_cast<SiStage*>(smem);
    const int stg_doubles = (P.nstages * (int)sizeof(SiStage) + 7) / 8;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P.stages);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
        for (int i = threadIdx.x; i < P.nstages * (int)(sizeof(SiStage) / 4); i += blockDim.x) dst[i] = src[i];
    }
    double* W = smem + stg_doubles + (size_t)warp * per;
    double* S = W + P.w_size;
    load_consts(P, S, lane);
    __syncthreads();
    auto sync = [] { __syncwarp(); };
    const double R[3] = {Rx, Ry, Rz};
    for (long long item = next_item(counter, lane); item < nitems; item = next_item(counter, lane)) {
        const SiPairItem it = items[item];
        const int Ka = bs.nprim[it.sa], Kb = bs.nprim[it.sb];
        const int pa = bs.poff[it.sa], pb = bs.poff[it.sb];
        double A[3], B[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            A[d] = bs.cen[3 * it.sa + d];
            B[d] = bs.cen[3 * it.sb + d];
        }
        __syncwarp();
        for (int i = lane; i < P.x_size; i += 32) W[P.x_off + i] = 0.0;
        for (int i = 0; i < Ka; ++i) {
            const double ax = bs.exps[pa + i], da = bs.coef[pa + i];
            for (int j = 0; j < Kb; ++j) {
                const double bx = bs.exps[pb + j], db = bs.coef[pb + j];
                __syncwarp();
                if (is_2c2e)
                    si_setup_2c2e(P, Bt, ax, da, A, bx, db, B, W, S, lane, 32);
                else
                    si_setup_1e(P, ax, da, A, bx, db, B, R, W, S, lane, 32);
                si_run_stages(P, stg, 0, P.n_prim_stages, W, S, 0, 1, lane, 32, sync);
            }
        }
        contracted_phase(P, stg, W, S, A, B, lane);
        store_matrix(P, W, od, bs.foff ? bs.foff[it.sa] : 0, bs.foff ? bs.foff[it.sb] : 0, it.sa == it.sb, lane);
    }
}

// coul: one CTA per shell pair; the warps split the point charges, their partial
// [a0] accumulators are summed in shared memory, warp 0 finishes the pair.
__global__ void __launch_bounds__(256) k_coul(SiProg P, SiBoys Bt, BasisDev bs, const SiPairItem* items,
                                              int nitems, const double* cxyz, const double* cq, int ncharge,
                                              OutDesc od) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int per = (P.w_size + SI_NDYN + P.nconst + 1) & ~1;
    SiStage* stg = reinterpret_cast<SiStage*>(smem);
    const int stg_doubles = (P.nstages * (int)sizeof(SiStage) + 7) / 8;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P.stages);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
        for (int i = threadIdx.x; i < P.nstages * (int)(sizeof(SiStage) / 4); i += blockDim.x) dst[i] = src[i];
    }
    double* W = smem + stg_doubles + (size_t)warp * per;
    double* S = W + P.w_size;
    load_consts(P, S, lane);
    __syncthreads();
    auto sync = [] { __syncwarp(); };
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const SiPairItem it = items[item];
        const int Ka = bs.nprim[it.sa], Kb = bs.nprim[it.sb];
        const int pa = bs.poff[it.sa], pb = bs.poff[it.sb];
        double A[3], B[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            A[d] = bs.cen[3 * it.sa + d];
            B[d] = bs.cen[3 * it.sb + d];
        }
        __syncthreads();
        for (int i = lane; i < P.x_size; i += 32) W[P.x_off + i] = 0.0;
        for (int i = 0; i < Ka; ++i) {
            const double ax = bs.exps[pa + i], da = bs.coef[pa + i];
            for (int j = 0; j < Kb; ++j) {
                const double bx = bs.exps[pb + j], db = bs.coef[pb + j];
                for (int c = warp; c < ncharge; c += nw) {
                    const double Rc[3] = {cxyz[3 * c], cxyz[3 * c + 1], cxyz[3 * c + 2]};
                    __syncwarp();
                    si_setup_coul(P, Bt, ax, da, A, bx, db, B, Rc, cq[c], W, S, lane, 32);
                    si_run_stages(P, stg, 0, P.n_prim_stages, W, S, 0, 1, lane, 32, sync);
                }
            }
        }
        __syncthreads();
        // reduce the per-warp accumulators into warp 0's X (fixed order: deterministic)
        for (int i = threadIdx.x; i < P.x_size; i += blockDim.x) {
            double* base = smem + stg_doubles;
            double s = base[P.x_off + i];
            for (int w = 1; w < nw; ++w) s += base[(size_t)w * per + P.x_off + i];
            base[P.x_off + i] = s;
        }
        __syncthreads();
        if (warp == 0) {
            contracted_phase(P, stg, W, S, A, B, lane);
            store_matrix(P, W, od, bs.foff ? bs.foff[it.sa] : 0, bs.foff ? bs.foff[it.sb] : 0, it.sa == it.sb,
                         lane);
        }
    }
}

// 3c2e: one warp per (shell pair, auxiliary shell).
__global__ void __launch_bounds__(256) k_3c2e(SiProg P, SiBoys Bt, BasisDev orb, BasisDev aux,
                                              const SiPairItem* pairs, int npairs, const int* auxlist,
                                              int naux_c, int col_begin, OutDesc od, int* counter) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (P.w_size + SI_NDYN + P.nconst + 1) & ~1;
    SiStage* stg = reinterpret_cast<SiStage*>(smem);
    const int stg_doubles = (P.nstages * (int)sizeof(SiStage) + 7) / 8;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P.stages);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
        for (int i = threadIdx.x; i < P.nstages * (int)(sizeof(SiStage) / 4); i += blockDim.x) dst[i] = src[i];
    }
    double* W = smem + stg_doubles + (size_t)warp * per;
    double* S = W + P.w_size;
    load_consts(P, S, lane);
    __syncthreads();
    auto sync = [] { __syncwarp(); };
    const long long nitems = (long long)npairs * naux_c;
    for (long long item = next_item(counter, lane); item < nitems; item = next_item(counter, lane)) {
        const int ip = (int)(item / naux_c);
        const int ic = (int)(item - (long long)ip * naux_c);
        const SiPairItem it = pairs[ip];
        const int sc = auxlist ? auxlist[ic] : ic;
        const int Ka = orb.nprim[it.sa], Kb = orb.nprim[it.sb], Kc = aux.nprim[sc];
        const int pa = orb.poff[it.sa], pb = orb.poff[it.sb], pc = aux.poff[sc];
        double A[3], B[3], C[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            A[d] = orb.cen[3 * it.sa + d];
            B[d] = orb.cen[3 * it.sb + d];
            C[d] = aux.cen[3 * sc + d];
        }
        __syncwarp();
        for (int i = lane; i < P.x_size; i += 32) W[P.x_off + i] = 0.0;
        for (int i = 0; i < Ka; ++i) {
            const double ax = orb.exps[pa + i], da = orb.coef[pa + i];
            for (int j = 0; j < Kb; ++j) {
                const double bx = orb.exps[pb + j], db = orb.coef[pb + j];
                for (int k = 0; k < Kc; ++k) {
                    const double cx = aux.exps[pc + k], dc = aux.coef[pc + k];
                    __syncwarp();
                    si_setup_3c2e(P, Bt, ax, da, A, bx, db, B, cx, dc, C, W, S, lane, 32);
                    si_run_stages(P, stg, 0, P.n_prim_stages, W, S, 0, 1, lane, 32, sync);
                }
            }
        }
        contracted_phase(P, stg, W, S, A, B, lane);
        // store: W[out_off + (i*n1 + j)*nm + m]
        const int n0 = P.out_n0, n1 = P.out_n1, nm = P.nb_total;
        const int total = n0 * n1 * nm;
        if (od.mode == STORE_BLOCK) {
            for (int idx = lane; idx < total; idx += 32) od.out[idx] = W[P.out_off + idx];
        } else {
            const long long col = (long long)(aux.foff[sc] - col_begin);
            const int fa = orb.foff[it.sa], fb = orb.foff[it.sb];
            for (int idx = lane; idx < total; idx += 32) {
                const int m = idx % nm;
                const int r = idx / nm;
                const int i = r / n1;
                const int j = r - i * n1;
                const double val = W[P.out_off + idx];
                if (od.mode == STORE_3C_FULL) {
                    od.out[((long long)(fa + i) * od.nbf + (fb + j)) * od.ld + col + m] = val;
                    if (it.sa != it.sb) od.out[((long long)(fb + j) * od.nbf + (fa + i)) * od.ld + col + m] = val;
                } else {
                    const long long row = it.swapped ? it.prow + (long long)j * n0 + i : it.prow + (long long)i * n1 + j;
                    od.out[row * od.ld + col + m] = val;
                }
            }
        }
    }
}

// copy the lower triangle of an n x n matrix into the upper one (exact symmetry, as intor_2c mirrors)
// upper triangle := transpose of the lower triangle, in 32 x 32 tiles through shared memory (coalesced reads and
// writes; tiles above the diagonal are written, diagonal tiles only above the diagonal)
__global__ void __launch_bounds__(256) k_mirror_lower(double* m, long long n) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;   // tile (bi, bj) of the UPPER part is written: bj >= bi
    if (bj < bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const long long row = (long long)bj * 32 + r, col = (long long)bi * 32 + tx;   // source tile (bj, bi), lower part
        if (row < n && col < n) tile[r][tx] = m[row * n + col];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long row = (long long)bi * 32 + r, col = (long long)bj * 32 + tx;
        if (row < n && col < n && col > row) m[row * n + col] = tile[tx][r];
    }
}

// The same for a matrix of which only the shell blocks with L(row shell) >= L(column shell) were computed (2c2e through
// the (a s0|b) classes with La >= Lb): element (r, c) := element (c, r) where L(r) < L(c), or L(r) == L(c) and c > r.
__global__ void __launch_bounds__(256) k_mirror_by_l(double* m, long long n, const signed char* fnL) {
    __shared__ double tile[32][33];
    __shared__ int need;
    const int bi = blockIdx.y, bj = blockIdx.x;   // tile (bi, bj) is (partly) written from tile (bj, bi)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x == 0) need = 0;
    __syncthreads();
    const long long col = (long long)bj * 32 + tx;
    const int lc = col < n ? fnL[col] : 127;
    bool take[4];
    for (int k = 0; k < 4; ++k) {
        const long long row = (long long)bi * 32 + ty + 8 * k;
        take[k] = false;
        if (row < n && col < n) {
            const int lr = fnL[row];
            take[k] = lr < lc || (lr == lc && col > row);
        }
        if (take[k]) need = 1;
    }
    __syncthreads();
    if (!need) return;
    for (int r = ty; r < 32; r += 8) {
        const long long row = (long long)bj * 32 + r, c2 = (long long)bi * 32 + tx;   // source tile (bj, bi)
        if (row < n && c2 < n) tile[r][tx] = m[row * n + c2];
    }
    __syncthreads();
    for (int k = 0; k < 4; ++k) {
        const int r = ty + 8 * k;
        if (take[k]) m[((long long)bi * 32 + r) * n + col] = tile[tx][r];
    }
}

// mode 0: si_boys (the reference's summation order); 1: si_boys_fast, the variant every generated 3c2e kernel
// calls (Horner/FMA Taylor sum on the transposed table, binary-power asymptote); 2: the variant of the generated
// nuclear-attraction kernels (one shared rsqrt per charge, prefactor kept in a register)
__global__ void k_boys_eval(SiBoys Bt, int mode, int n, const double* x, size_t nx, double* out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= nx) return;
    const double T = x[i];
    double v;
    if (mode == 0) {
        v = si_boys(Bt, n, T);
    } else if (mode == 1) {
        v = si_boys_fast(Bt, n, T);
    } else {
        const double xp = Bt.xlarge_pref[n];
        v = (T < 30.0) ? si_boys_taylor(Bt, n, T) : xp * si_boys_xpow_rs(n, si_rsqrt(T));
    }
    out[i] = v;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static thread_local std::string g_last_error;

#define SI_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            g_last_error = std::string(#call) + ": " + cudaGetErrorString(_e);                 \
            return SI_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

struct DevProgram {
    SiProgramHost host;
    SiProg dev;
    void* d_hdr = nullptr;
    void* d_terms = nullptr;
    void* d_stages = nullptr;
    void* d_consts = nullptr;
};

static const int SI_NSTREAMS = 8;
static const int SI_NCOUNTERS = 1024;
#define SI_GC_MAX_SPLIT 8

// Tuning / debugging switches from the environment, read ONCE (C++11 thread-safe static initialisation) --
// no lazily initialised function-level statics in the entry points.
struct SiEnv {
    int nstreams = SI_NSTREAMS;
    int gc_warps_per_sm = 8, gc_max_split = SI_GC_MAX_SPLIT;
    bool gc_profile = false, g3_profile = false, force_lut = false, force_fan = false, tuple_tables = true;
    int g3_warps_per_block = 4, g3_grid_per_sm = 4;
    double g3_small_total = 2.0e9, g3_big_flops = 2.5e8;
    int g3_fan = 2;
    bool graphs = true;
    int pdl = -1;                    // -1: automatic (calls below g3_pdl_total model flops), 0 / 1: forced
    double g3_pdl_total = 1.5e11;
    SiEnv() {
        auto geti = [](const char* n, int dflt) { const char* e = getenv(n); return e ? atoi(e) : dflt; };
        auto getd = [](const char* n, double dflt) { const char* e = getenv(n); return e ? atof(e) : dflt; };
        auto getb = [](const char* n, bool dflt) { const char* e = getenv(n); return e ? (e[0] == '1') : dflt; };
        nstreams = std::max(1, std::min(SI_NSTREAMS, geti("SI_NSTREAMS", SI_NSTREAMS)));
        gc_warps_per_sm = std::max(1, geti("SI_GC_WARPS_PER_SM", 8));
        gc_max_split = std::max(1, std::min(SI_GC_MAX_SPLIT, geti("SI_GC_MAX_SPLIT", SI_GC_MAX_SPLIT)));
        gc_profile = getb("SI_GC_PROFILE", false);
        g3_profile = getb("SI_G3_PROFILE", false);
        force_lut = getb("SI_FORCE_LUT", false);
        force_fan = getb("SI_FORCE_FAN", false);
        tuple_tables = getb("SI_G3_TUPLE_TABLES", true);
        g3_warps_per_block = std::max(1, std::min(4, geti("SI_G3_WARPS_PER_BLOCK", 4)));
        g3_grid_per_sm = std::max(1, geti("SI_G3_GRID_PER_SM", 4));
        g3_small_total = getd("SI_G3_SMALL_TOTAL", 2.0e9);
        g3_big_flops = getd("SI_G3_BIG_FLOPS", 2.5e8);
        pdl = geti("SI_G3_PDL", -1);
        g3_pdl_total = getd("SI_G3_PDL_TOTAL", 1.5e11);
        g3_fan = std::max(1, std::min(SI_NSTREAMS, geti("SI_G3_FAN", 2)));
        graphs = getb("SI_GRAPHS", true);
    }
};
static const SiEnv& si_env() {
    static const SiEnv e;
    return e;
}

struct si_context {
    int device = 0;
    int lmax = 0, lauxmax = 0;
    int sm_count = 148;
    int max_smem = 227 * 1024;
    std::vector<double> boys_host;
    std::vector<double> pref_host;
    int nrows = 0, ncols = 0;
    double* d_boys = nullptr;
    double* d_boysT = nullptr;
    double* d_pref = nullptr;
    SiBoys boys_dev;
    std::map<std::tuple<int, int, int, int, int, int>, DevProgram*> progs;
    long long launches = 0;
    cudaStream_t streams[SI_NSTREAMS];
    cudaEvent_t ev_fork;
    cudaEvent_t ev_join[SI_NSTREAMS];
    // Calls on one context are serialised: on the host by `mu` (entry points may be called from several
    // threads), on the device by `ev_last` (recorded when a call's work has joined the caller's stream; the
    // next call's stream waits for it before the shared work counters are reset) -- callers may therefore
    // pass different streams to consecutive calls.  Independent work belongs on separate contexts.
    std::recursive_mutex mu;
    cudaEvent_t ev_last;
    bool ev_last_valid = false;
    int* d_counters = nullptr;
    int next_counter = 0;
    // grow-only device scratch of the coul path (charges + charge-split partial matrices): no per-call
    // allocation, whose cost through the stream-ordered pool varies by orders of magnitude
    double* d_scratch = nullptr;
    size_t scratch_doubles = 0;
    // host-buffer 3c2e path: grow-only staging buffers, streams and events owned by the context
    double* stage[2] = {nullptr, nullptr};
    size_t stage_elems = 0;
    cudaStream_t s_comp = nullptr, s_copy = nullptr;
    cudaEvent_t ev_done[2], ev_free[2];
    std::vector<cudaEvent_t> ev_time;   // timing events (two per chunk), grow-only
    double breakdown[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // of the last si_int3c2e_h call, see si_host_breakdown()
    // shell sets by content: a set created again with identical content re-attaches to the cached object
    // (device arrays are uploaded again, the shell-batcher's pair sets / tuple tables / aux lists are kept)
    std::vector<struct si_basis*> basis_cache;   // least recently used first
    std::map<size_t, int> onee_occ;              // resident blocks per SM of k_onee by dynamic shared-memory size
    bool small_call = false;                     // int3c2e_impl: the current call is launch-latency bound (see g3_config)
    bool twoc_lower_only = false;                // si_int2c2e: evaluate the classes with La >= Lb only
    bool capturing = false;                      // a call is being captured into a CUDA graph (no ev_last traffic)
    double screen_eps = 0.0;                     // opt-in primitive pre-screening threshold of the 3c2e kernels (0 = off)
};
static const size_t SI_BASIS_CACHE = 8;

static int ctx_scratch(si_context* c, size_t ndoubles, cudaStream_t s, double** out) {
    if (ndoubles > c->scratch_doubles) {
        SI_CUDA(cudaStreamSynchronize(s));
        SI_CUDA(cudaDeviceSynchronize());
        if (c->d_scratch) SI_CUDA(cudaFree(c->d_scratch));
        c->d_scratch = nullptr;
        c->scratch_doubles = 0;
        SI_CUDA(cudaMalloc((void**)&c->d_scratch, ndoubles * sizeof(double)));
        c->scratch_doubles = ndoubles;
    }
    *out = c->d_scratch;
    return SI_OK;
}

static std::atomic<long long> g_basis_uid{0};
// live shell-set objects: a handle that outlived its context (released by si_finalize) is recognised, not dereferenced
static std::mutex g_live_mu;
static std::set<const void*> g_live_bases;
static bool basis_is_live(const void* b) {
    std::lock_guard<std::mutex> lk(g_live_mu);
    return g_live_bases.count(b) != 0;
}
struct si_basis {
    si_context* ctx = nullptr;
    int users = 1;            // live handles returned by si_basis_create for this content
    long long uid = g_basis_uid.fetch_add(1) + 1;   // never reused (keys of caches that outlive a basis)
    int nshell = 0;
    std::vector<int> L, nprim, poff, foff_sph, foff_cart;
    std::vector<double> cen, exps, coef, coef_pgto, oexp;
    int nbf_sph = 0, nbf_cart = 0, lmax = 0;
    long long packed_rows = 0;
    int *d_L = nullptr, *d_nprim = nullptr, *d_poff = nullptr, *d_foff_sph = nullptr, *d_foff_cart = nullptr;
    double *d_cen = nullptr, *d_exps = nullptr, *d_coef = nullptr, *d_coef_pgto = nullptr, *d_oexp = nullptr;
    // shell-pair classes (La >= Lb), cost-sorted, on the device
    struct TupleTable {
        int2* d = nullptr;
        long long ngroups = 0;
    };
    struct PairClass {
        int La, Lb;
        int n = 0;
        SiPairItem* d_items = nullptr;
        std::vector<std::pair<int, int>> kab;   // (Ka, Kb) of every item, in item order (host copy)
        // explicit tuple tables for launches with few aux shells per class (key: aux uid, sb, se, Lc, tpw)
        std::map<std::tuple<long long, int, int, int, int>, TupleTable> tables;
    };
    // A set of shell pairs (i, j), i0 <= i < i1, j <= i, sorted into (La >= Lb) classes, with the
    // precomputed primitive-pair data (9 doubles each) for coef / coef_pgto.
    struct PairSet {
        std::vector<PairClass> classes;
        double* d_ppair[2] = {nullptr, nullptr};
        long long rows = 0;   // packed rows covered by the set
        int i0 = 0, i1 = 0;
    };
    PairSet full;
    std::vector<PairSet*> row_chunks;   // cached chunking of the host-buffer 3c2e path
    long long row_chunk_bytes = 0;
    int row_chunk_ncol = 0;
    bool pairs_built = false;
    bool selected = false;    // si_basis_select_rows restricted `full` to a subset of the pairs
    // cached device lists of the aux shells of a range, grouped by angular momentum (3c2e)
    struct AuxRange {
        int sb, se;
        int* d_list = nullptr;
        SiAuxItem* d_items = nullptr;
        std::map<int, std::pair<int, int>> span;
    };
    mutable std::vector<AuxRange> aux_ranges;
    // 2c2e through the generated 3c2e kernels: (a|b) = (a s0|b) with a unit s shell of exponent 0
    mutable si_basis* twoc_orb = nullptr;     // copy of this basis + the s0 shell (last)
    mutable PairSet* twoc_pairs = nullptr;    // pairs (shell j, s0)
    mutable signed char* d_fn_L = nullptr;    // angular momentum of every spherical function (2c2e mirror)
    // The metric is ~36 short class launches + the mirror kernel: launch-rate bound when issued one by one.  After
    // a first ordinary call (which fills every cache) the call sequence is captured into a CUDA graph per
    // (output pointer, normalisation) and replayed with one cudaGraphLaunch.
    struct TwocGraph {
        double* out = nullptr;
        int normalize = 0;
        cudaGraphExec_t exec = nullptr;
        int kernels = 0;
    };
    mutable std::vector<TwocGraph> twoc_graphs;
    mutable bool twoc_warm = false;
};

const char* si_strerror(int status) {
    switch (status) {
        case SI_OK: return "ok";
        case SI_ERR_ARG: return "invalid argument";
        case SI_ERR_LMAX: return "angular momentum exceeds lmax/lauxmax of the context";
        case SI_ERR_CUDA: return "CUDA error";
        case SI_ERR_KEY: return "unknown integral key or unsupported option";
        case SI_ERR_BOYS: return "Boys order exceeds the table";
        default: return "internal error";
    }
}
const char* si_last_error(void) { return g_last_error.c_str(); }

// Boys table with the reference algorithm (testing/boys.py:18-29,81-91); the top row is
// evaluated with the convergent series F_n(x) = exp(-x) sum_k (2x)^k / ((2n+1)(2n+3)...(2n+2k+1))
// in extended precision instead of adaptive quadrature.
static void build_boys_table(int nmax, std::vector<double>& table, int& nrows, int& ncols) {
    const int order = 6;
    nrows = nmax + order + 1;
    ncols = 301;
    table.assign((size_t)nrows * ncols, 0.0);
    std::vector<double> xs(ncols);
    for (int i = 0; i < ncols; ++i) {
        // numpy.linspace(0, 30, 301): start + i*step with step = 30/300
        const double step = 30.0 / 300.0;
        xs[i] = (i == ncols - 1) ? 30.0 : 0.0 + i * step;
    }
    const int ntop = nmax + order;
    for (int i = 0; i < ncols; ++i) {
        long double x = xs[i];
        long double term = 1.0L / (2 * ntop + 1);
        long double sum = term;
        for (int k = 1; k < 2000; ++k) {
            term *= 2.0L * x / (2 * ntop + 2 * k + 1);
            sum += term;
            if (term < 1e-22L * sum) break;
        }
        table[(size_t)ntop * ncols + i] = (double)(expl(-x) * sum);
    }
    for (int n = ntop - 1; n >= 0; --n)
        for (int i = 0; i < ncols; ++i) {
            double emx = std::exp(-xs[i]);
            table[(size_t)n * ncols + i] = (2.0 * xs[i] * table[(size_t)(n + 1) * ncols + i] + emx) / (2 * n + 1);
        }
}

int si_init(int device, int lmax, int lauxmax, const double* boys_table, int nrows, int ncols, si_context** out) {
    if (!out || lmax < 0 || lauxmax < 0) return SI_ERR_ARG;
    if (lmax > 5 || lauxmax > 2 * 5) {
        g_last_error = "lmax <= 5 and lauxmax <= 10 supported";
        return SI_ERR_LMAX;
    }
    int ndev = 0;
    SI_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        g_last_error = "no such CUDA device";
        return SI_ERR_CUDA;
    }
    SI_CUDA(cudaSetDevice(device));
    si_context* c = new si_context;
    c->device = device;
    c->lmax = lmax;
    c->lauxmax = lauxmax;
    cudaDeviceProp prop;
    SI_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->max_smem = (int)prop.sharedMemPerBlockOptin;
    if (boys_table) {
        if (ncols != 301 || nrows < 8) {
            delete c;
            return SI_ERR_ARG;
        }
        c->boys_host.assign(boys_table, boys_table + (size_t)nrows * ncols);
        c->nrows = nrows;
        c->ncols = ncols;
    } else {
        int need = std::max(std::max(2 * lmax + lauxmax, 2 * lauxmax), 4 * lmax);  // highest Boys order (4 lmax: Schwarz integrals)
        build_boys_table(std::max(need, 8), c->boys_host, c->nrows, c->ncols);
    }
    const int nmax = c->nrows - 7;
    c->pref_host.resize(nmax + 1);
    for (int n = 0; n <= nmax; ++n) {
        double f2 = 1.0;
        for (int k = 2 * n - 1; k > 1; k -= 2) f2 *= k;
        c->pref_host[n] = f2 / std::pow(2.0, n + 1) * std::sqrt(SI_PI);
    }
    SI_CUDA(cudaMalloc(&c->d_boys, c->boys_host.size() * sizeof(double)));
    SI_CUDA(cudaMemcpy(c->d_boys, c->boys_host.data(), c->boys_host.size() * sizeof(double), cudaMemcpyHostToDevice));
    SI_CUDA(cudaMalloc(&c->d_pref, c->pref_host.size() * sizeof(double)));
    SI_CUDA(cudaMemcpy(c->d_pref, c->pref_host.data(), c->pref_host.size() * sizeof(double), cudaMemcpyHostToDevice));
    {
        std::vector<double> tT((size_t)c->nrows * c->ncols);
        for (int r = 0; r < c->nrows; ++r)
            for (int i = 0; i < c->ncols; ++i) tT[(size_t)i * c->nrows + r] = c->boys_host[(size_t)r * c->ncols + i];
        SI_CUDA(cudaMalloc(&c->d_boysT, tT.size() * sizeof(double)));
        SI_CUDA(cudaMemcpy(c->d_boysT, tT.data(), tT.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    c->boys_dev.tableT = c->d_boysT;
    c->boys_dev.table = c->d_boys;
    c->boys_dev.xlarge_pref = c->d_pref;
    c->boys_dev.nrows = c->nrows;
    c->boys_dev.ncols = c->ncols;
    for (int i = 0; i < SI_NSTREAMS; ++i) {
        SI_CUDA(cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking));
        SI_CUDA(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    SI_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    SI_CUDA(cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming));
    SI_CUDA(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
    SI_CUDA(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        SI_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
        SI_CUDA(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
    }
    SI_CUDA(cudaMalloc(&c->d_counters, SI_NCOUNTERS * sizeof(int)));
    SI_CUDA(cudaFuncSetAttribute(k_two_index, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    SI_CUDA(cudaFuncSetAttribute(k_coul, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    SI_CUDA(cudaFuncSetAttribute(k_3c2e, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    SI_CUDA(cudaFuncSetAttribute(k_schwarz, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    SI_CUDA(cudaFuncSetAttribute(k_onee<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    SI_CUDA(cudaFuncSetAttribute(k_onee<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    SI_CUDA(cudaFuncSetAttribute(k_onee<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    {
        // tables of the gto / prefactor kernels (per device)
        static SiGtoConst hg;
        memset(&hg, 0, sizeof(hg));
        for (int l = 0; l <= SI_GTO_LMAX; ++l) {
            std::vector<int> lmn;
            si_cart_order(l, lmn);
            const std::vector<double> M = si_cart2sph(l), f = si_lmn_factors(l);
            const int nc = si_ncart(l), ns = si_nsph(l);
            for (int k = 0; k < nc; ++k) {
                hg.cart[l][k] = lmn[3 * k] | (lmn[3 * k + 1] << 4) | (lmn[3 * k + 2] << 8);
                hg.lmn[l][k] = f[k];
                for (int m = 0; m < ns; ++m) hg.c2s[l][m][k] = M[(size_t)m * nc + k];
            }
        }
        SI_CUDA(cudaMemcpyToSymbol(g_si_gto, &hg, sizeof(hg)));
    }
    // per-device attribute of every generated kernel (one context per device may live in the same process)
    if (g3_set_attributes(c->max_smem) || gc_set_attributes(c->max_smem)) {
        g_last_error = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for a generated kernel";
        return SI_ERR_CUDA;
    }
    *out = c;
    return SI_OK;
}

static void free_program(DevProgram* p) {
    cudaFree(p->d_hdr);
    cudaFree(p->d_terms);
    cudaFree(p->d_stages);
    cudaFree(p->d_consts);
    delete p;
}

static void basis_free(si_basis* b);
int si_finalize(si_context* c) {
    if (!c) return SI_ERR_ARG;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto* b : c->basis_cache) b->twoc_orb = nullptr;   // (internal sets are cache entries themselves)
    for (auto* b : c->basis_cache) basis_free(b);
    c->basis_cache.clear();
    cudaStreamDestroy(c->s_comp);
    cudaStreamDestroy(c->s_copy);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(c->ev_done[i]);
        cudaEventDestroy(c->ev_free[i]);
        cudaFree(c->stage[i]);
    }
    for (auto e : c->ev_time) cudaEventDestroy(e);
    for (auto& kv : c->progs) free_program(kv.second);
    for (int i = 0; i < SI_NSTREAMS; ++i) {
        cudaStreamDestroy(c->streams[i]);
        cudaEventDestroy(c->ev_join[i]);
    }
    cudaEventDestroy(c->ev_fork);
    cudaEventDestroy(c->ev_last);
    cudaFree(c->d_boys);
    cudaFree(c->d_pref);
    cudaFree(c->d_counters);
    cudaFree(c->d_boysT);
    cudaFree(c->d_scratch);
    delete c;
    return SI_OK;
}

int si_boys_table(const si_context* c, double* table, int* nrows, int* ncols) {
    if (!c) return SI_ERR_ARG;
    if (nrows) *nrows = c->nrows;
    if (ncols) *ncols = c->ncols;
    if (table) memcpy(table, c->boys_host.data(), c->boys_host.size() * sizeof(double));
    return SI_OK;
}

int si_boys_eval(si_context* c, int n, const double* x, size_t nx, double* out) {
    return si_boys_eval_mode(c, 0, n, x, nx, out);
}

int si_boys_eval_mode(si_context* c, int mode, int n, const double* x, size_t nx, double* out) {
    if (!c || !x || !out || mode < 0 || mode > 2) return SI_ERR_ARG;
    if (n < 0 || n > c->nrows - 7) return SI_ERR_BOYS;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    SI_CUDA(cudaSetDevice(c->device));
    double *dx = nullptr, *dout = nullptr;
    SI_CUDA(cudaMalloc(&dx, nx * sizeof(double)));
    SI_CUDA(cudaMalloc(&dout, nx * sizeof(double)));
    SI_CUDA(cudaMemcpy(dx, x, nx * sizeof(double), cudaMemcpyHostToDevice));
    k_boys_eval<<<(unsigned)((nx + 255) / 256), 256>>>(c->boys_dev, mode, n, dx, nx, dout);
    c->launches++;
    SI_CUDA(cudaGetLastError());
    SI_CUDA(cudaMemcpy(out, dout, nx * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dx);
    cudaFree(dout);
    return SI_OK;
}

static int get_program(si_context* c, int key, int La, int Lb, int Lc, int sph, int normalize, DevProgram** out) {
    auto k = std::make_tuple(key, La, Lb, Lc, sph, normalize ? 1 : 0);
    auto it = c->progs.find(k);
    if (it != c->progs.end()) {
        *out = it->second;
        return SI_OK;
    }
    DevProgram* p = new DevProgram;
    try {
        p->host = si_build_program(key, La, Lb, Lc, sph, normalize ? 1 : 0);
    } catch (std::exception& e) {
        g_last_error = e.what();
        delete p;
        return SI_ERR_INTERNAL;
    }
    const SiProgramHost& h = p->host;
    // Boys orders needed vs table
    if (key == SI_KEY_COUL || key == SI_KEY_2C2E || key == SI_KEY_3C2E) {
        if (h.desc.boys_nlo + h.desc.nbase - 1 > c->nrows - 7) {
            g_last_error = "Boys table too small for this class";
            delete p;
            return SI_ERR_BOYS;
        }
    }
    auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, std::max<size_t>(bytes, 8));
        if (e != cudaSuccess) return e;
        if (bytes) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
        return e;
    };
    SI_CUDA(up(h.hdr.data(), h.hdr.size() * 4, &p->d_hdr));
    SI_CUDA(up(h.terms.data(), h.terms.size() * 4, &p->d_terms));
    SI_CUDA(up(h.stages.data(), h.stages.size() * sizeof(SiStage), &p->d_stages));
    SI_CUDA(up(h.consts.data(), h.consts.size() * 8, &p->d_consts));
    p->dev = h.desc;
    p->dev.hdr = (const uint32_t*)p->d_hdr;
    p->dev.terms = (const uint32_t*)p->d_terms;
    p->dev.stages = (const SiStage*)p->d_stages;
    p->dev.consts = (const double*)p->d_consts;
    // The uploads above come from pageable memory (cudaMemcpy may return before the DMA has landed) and the
    // kernels that read them run on non-blocking streams, possibly forked before this call: wait explicitly.
    SI_CUDA(cudaDeviceSynchronize());
    c->progs[k] = p;
    *out = p;
    return SI_OK;
}

int si_class_stats(si_context* c, int key, int La, int Lb, int Lc, int sph, int normalize, long long* out) {
    if (!c || !out) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (La < Lb) std::swap(La, Lb);
    DevProgram* p;
    int rc = get_program(c, key, La, Lb, Lc, sph, normalize, &p);
    if (rc) return rc;
    out[0] = p->host.n_fma_prim;
    out[1] = p->host.n_fma_contr;
    out[2] = p->host.desc.w_size + SI_NDYN + p->host.desc.nconst;
    out[3] = (long long)(p->host.hdr.size() + p->host.terms.size());
    return SI_OK;
}

long long si_launch_count(const si_context* c) { return c ? c->launches : -1; }

// ---------------------------------------------------------------------------
// basis
// ---------------------------------------------------------------------------
template <typename T>
static cudaError_t upload(const std::vector<T>& v, T** d) {
    cudaError_t e = cudaMalloc((void**)d, std::max<size_t>(v.size() * sizeof(T), 8));
    if (e != cudaSuccess) return e;
    if (!v.empty()) e = cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

static void free_pair_set(si_basis::PairSet* ps);
int si_basis_create(si_context* c, int nshell, const int* L, const int* nprim, const double* centers,
                    const double* exps, const double* coefs, si_basis** out) {
    if (!c || !out || nshell <= 0 || !L || !nprim || !centers || !exps || !coefs) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    SI_CUDA(cudaSetDevice(c->device));
    {
        // content cache: identical shell data -> the cached object (its pair sets, tuple tables and aux lists are
        // kept); the device copies are uploaded again, so the call still moves this set's bytes host -> device
        size_t np_ = 0;
        for (int i = 0; i < nshell; ++i) np_ += (size_t)std::max(nprim[i], 0);
        for (size_t ci = 0; ci < c->basis_cache.size(); ++ci) {
            si_basis* o = c->basis_cache[ci];
            if (o->nshell != nshell || o->exps.size() != np_) continue;
            if (memcmp(o->L.data(), L, nshell * sizeof(int)) || memcmp(o->nprim.data(), nprim, nshell * sizeof(int)) ||
                memcmp(o->cen.data(), centers, 3 * (size_t)nshell * sizeof(double)) ||
                memcmp(o->exps.data(), exps, np_ * sizeof(double)) || memcmp(o->coef.data(), coefs, np_ * sizeof(double)))
                continue;
            if (c->ev_last_valid) SI_CUDA(cudaEventSynchronize(c->ev_last));   // no kernel of an earlier call reads them now
            if (o->selected && o->users <= 0) {   // a pair selection of an earlier owner does not outlive its handle
                free_pair_set(&o->full);
                for (auto& g : o->twoc_graphs) cudaGraphExecDestroy(g.exec);
                o->twoc_graphs.clear();
                o->pairs_built = false;
                o->selected = false;
            }
            SI_CUDA(cudaMemcpy(o->d_L, L, nshell * sizeof(int), cudaMemcpyHostToDevice));
            SI_CUDA(cudaMemcpy(o->d_nprim, nprim, nshell * sizeof(int), cudaMemcpyHostToDevice));
            SI_CUDA(cudaMemcpy(o->d_cen, centers, 3 * (size_t)nshell * sizeof(double), cudaMemcpyHostToDevice));
            SI_CUDA(cudaMemcpy(o->d_exps, exps, np_ * sizeof(double), cudaMemcpyHostToDevice));
            SI_CUDA(cudaMemcpy(o->d_coef, coefs, np_ * sizeof(double), cudaMemcpyHostToDevice));
            SI_CUDA(cudaDeviceSynchronize());
            o->users++;
            c->basis_cache.erase(c->basis_cache.begin() + ci);
            c->basis_cache.push_back(o);
            *out = o;
            return SI_OK;
        }
    }
    si_basis* b = new si_basis;
    b->ctx = c;
    b->nshell = nshell;
    b->L.assign(L, L + nshell);
    b->nprim.assign(nprim, nprim + nshell);
    b->cen.assign(centers, centers + 3 * (size_t)nshell);
    int np = 0, fs = 0, fc = 0;
    long long prow = 0;
    for (int i = 0; i < nshell; ++i) {
        if (L[i] < 0 || L[i] > std::max(c->lmax, c->lauxmax) || nprim[i] <= 0) {
            delete b;
            g_last_error = "shell angular momentum exceeds lmax/lauxmax (or nprim <= 0)";
            return SI_ERR_LMAX;
        }
        b->poff.push_back(np);
        np += nprim[i];
        b->foff_sph.push_back(fs);
        b->foff_cart.push_back(fc);
        fs += si_nsph(L[i]);
        fc += si_ncart(L[i]);
        b->lmax = std::max(b->lmax, L[i]);
    }
    b->nbf_sph = fs;
    b->nbf_cart = fc;
    for (int i = 0; i < nshell; ++i)
        for (int j = 0; j <= i; ++j) prow += (long long)si_nsph(L[i]) * si_nsph(L[j]);
    b->packed_rows = prow;
    b->exps.assign(exps, exps + np);
    b->coef.assign(coefs, coefs + np);
    // --normalize pgto: radial part of norm_pgto (main.py:152-159) folded into the coefficients
    b->coef_pgto.resize(np);
    for (int i = 0; i < nshell; ++i)
        for (int k = 0; k < nprim[i]; ++k) {
            int idx = b->poff[i] + k;
            double a = b->exps[idx];
            int l = L[i];
            b->coef_pgto[idx] = b->coef[idx] * std::sqrt(std::pow(2.0, 2 * l + 1.5) * std::pow(a, l + 1.5) /
                                                         std::pow(SI_PI, 1.5));
        }
    SI_CUDA(upload(b->L, &b->d_L));
    SI_CUDA(upload(b->nprim, &b->d_nprim));
    SI_CUDA(upload(b->poff, &b->d_poff));
    SI_CUDA(upload(b->foff_sph, &b->d_foff_sph));
    SI_CUDA(upload(b->foff_cart, &b->d_foff_cart));
    SI_CUDA(upload(b->cen, &b->d_cen));
    SI_CUDA(upload(b->exps, &b->d_exps));
    SI_CUDA(upload(b->coef, &b->d_coef));
    SI_CUDA(upload(b->coef_pgto, &b->d_coef_pgto));
    b->oexp.resize(np);
    for (int i = 0; i < np; ++i) b->oexp[i] = 1.0 / b->exps[i];
    SI_CUDA(upload(b->oexp, &b->d_oexp));
    SI_CUDA(cudaDeviceSynchronize());   // pageable uploads complete before any kernel on a non-blocking stream reads them
    {
        std::lock_guard<std::mutex> lk2(g_live_mu);
        g_live_bases.insert(b);
    }
    // into the content cache; unreferenced entries beyond the limit are released, oldest first
    c->basis_cache.push_back(b);
    for (size_t ci = 0; c->basis_cache.size() > SI_BASIS_CACHE && ci < c->basis_cache.size();) {
        if (c->basis_cache[ci]->users <= 0) {
            si_basis* o = c->basis_cache[ci];
            c->basis_cache.erase(c->basis_cache.begin() + ci);
            basis_free(o);
        } else {
            ++ci;
        }
    }
    *out = b;
    return SI_OK;
}

static void free_pair_set(si_basis::PairSet* ps);
// A handle is released; the object stays in the context's content cache until it is evicted or the context ends.
int si_basis_destroy(si_basis* b) {
    if (!b || !basis_is_live(b)) return SI_ERR_ARG;   // (also: a handle whose context is gone)
    std::lock_guard<std::recursive_mutex> lk(b->ctx->mu);
    if (b->users > 0) b->users--;
    return SI_OK;
}

static void basis_free(si_basis* b) {
    {
        std::lock_guard<std::mutex> lk2(g_live_mu);
        g_live_bases.erase(b);
    }
    cudaSetDevice(b->ctx->device);
    cudaDeviceSynchronize();
    cudaFree(b->d_L);
    cudaFree(b->d_nprim);
    cudaFree(b->d_poff);
    cudaFree(b->d_foff_sph);
    cudaFree(b->d_foff_cart);
    cudaFree(b->d_cen);
    cudaFree(b->d_exps);
    cudaFree(b->d_coef);
    cudaFree(b->d_coef_pgto);
    cudaFree(b->d_oexp);
    if (b->d_fn_L) cudaFree(b->d_fn_L);
    if (b->twoc_pairs) {
        free_pair_set(b->twoc_pairs);
        delete b->twoc_pairs;
    }
    for (auto& g : b->twoc_graphs) cudaGraphExecDestroy(g.exec);
    if (b->twoc_orb) si_basis_destroy(b->twoc_orb);   // (drops the reference; the object lives in the cache)
    for (auto& r : b->aux_ranges) {
        cudaFree(r.d_list);
        cudaFree(r.d_items);
    }
    free_pair_set(&b->full);
    for (auto* ps : b->row_chunks) {
        free_pair_set(ps);
        delete ps;
    }
    delete b;
}

int si_basis_nbf(const si_basis* b, int sph) { return b ? (sph ? b->nbf_sph : b->nbf_cart) : -1; }
int si_basis_nshell(const si_basis* b) { return b ? b->nshell : -1; }
long long si_basis_packed_rows(const si_basis* b) { return b ? b->packed_rows : -1; }

static void free_pair_set(si_basis::PairSet* ps) {
    for (auto& pc : ps->classes) {
        cudaFree(pc.d_items);
        for (auto& kv : pc.tables) cudaFree(kv.second.d);
    }
    ps->classes.clear();
    cudaFree(ps->d_ppair[0]);
    cudaFree(ps->d_ppair[1]);
    ps->d_ppair[0] = ps->d_ppair[1] = nullptr;
}

// Host shell-batcher: sort the shell pairs (i, j), i0 <= i < i1, j <= i, into (La >= Lb) classes, most
// expensive (most primitive pairs) first, and precompute the primitive-pair quantities.
static int build_pair_set(si_basis* b, int i0, int i1, si_basis::PairSet* ps, bool skip_diagonal = false) {
    std::map<std::pair<int, int>, std::vector<SiPairItem>> cls;
    long long prow = 0;
    ps->i0 = i0;
    ps->i1 = i1;
    for (int i = i0; i < i1; ++i)
        for (int j = 0; j <= i; ++j) {
            if (skip_diagonal && j == i) continue;
            SiPairItem it;
            memset(&it, 0, sizeof(it));
            if (b->L[i] >= b->L[j]) {
                it.sa = i;
                it.sb = j;
                it.swapped = 0;
            } else {
                it.sa = j;
                it.sb = i;
                it.swapped = 1;
            }
            it.prow = prow;
            it.Ka = b->nprim[it.sa];
            it.Kb = b->nprim[it.sb];
            it.fa0 = b->foff_sph[it.sa];
            it.fb0 = b->foff_sph[it.sb];
            for (int d = 0; d < 3; ++d) it.AB[d] = b->cen[3 * it.sa + d] - b->cen[3 * it.sb + d];
            prow += (long long)si_nsph(b->L[i]) * si_nsph(b->L[j]);
            cls[{b->L[it.sa], b->L[it.sb]}].push_back(it);
        }
    ps->rows = prow;
    {
        size_t npp = 0;
        for (auto& kv : cls)
            for (auto& it : kv.second) {
                it.ppoff = (int)npp;
                npp += (size_t)b->nprim[it.sa] * b->nprim[it.sb];
            }
        for (int variant = 0; variant < 2; ++variant) {
            const std::vector<double>& co = variant ? b->coef_pgto : b->coef;
            std::vector<double> pp(npp * SI_PP_STRIDE, 0.0);
            for (auto& kv : cls)
                for (auto& it : kv.second) {
                    const int Ka = b->nprim[it.sa], Kb = b->nprim[it.sb];
                    const double* A = &b->cen[3 * it.sa];
                    const double* B = &b->cen[3 * it.sb];
                    for (int i = 0; i < Ka; ++i)
                        for (int j = 0; j < Kb; ++j) {
                            double* q = &pp[((size_t)it.ppoff + (size_t)i * Kb + j) * SI_PP_STRIDE];
                            q[9] = b->exps[b->poff[it.sa] + i];
                            q[10] = b->exps[b->poff[it.sb] + j];
                            const double ax = b->exps[b->poff[it.sa] + i], bx = b->exps[b->poff[it.sb] + j];
                            const double p = ax + bx, oop = 1.0 / p, mu = ax * bx * oop;
                            double r2 = 0.0;
                            q[0] = p;
                            q[1] = oop;
                            for (int d = 0; d < 3; ++d) {
                                q[2 + d] = (ax * A[d] + bx * B[d]) * oop;
                                q[5 + d] = q[2 + d] - A[d];
                                r2 += (A[d] - B[d]) * (A[d] - B[d]);
                            }
                            q[8] = std::exp(-mu * r2) * co[b->poff[it.sa] + i] * co[b->poff[it.sb] + j];
                        }
                }
            SI_CUDA(cudaMalloc((void**)&ps->d_ppair[variant], std::max<size_t>(pp.size() * sizeof(double), 8)));
            SI_CUDA(cudaMemcpy(ps->d_ppair[variant], pp.data(), pp.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    for (auto& kv : cls) {
        auto& v = kv.second;
        std::stable_sort(v.begin(), v.end(), [&](const SiPairItem& x, const SiPairItem& y) {
            return b->nprim[x.sa] * b->nprim[x.sb] > b->nprim[y.sa] * b->nprim[y.sb];
        });
        si_basis::PairClass pc;
        pc.La = kv.first.first;
        pc.Lb = kv.first.second;
        pc.n = (int)v.size();
        for (auto& x : v) pc.kab.push_back({x.Ka, x.Kb});
        SI_CUDA(cudaMalloc((void**)&pc.d_items, v.size() * sizeof(SiPairItem)));
        SI_CUDA(cudaMemcpy(pc.d_items, v.data(), v.size() * sizeof(SiPairItem), cudaMemcpyHostToDevice));
        ps->classes.push_back(pc);
    }
    // pageable uploads may still be in flight when cudaMemcpy returns; the consumers run on non-blocking streams
    SI_CUDA(cudaDeviceSynchronize());
    return SI_OK;
}

static int build_pair_classes(si_basis* b) {
    if (b->pairs_built) return SI_OK;
    int rc = build_pair_set(b, 0, b->nshell, &b->full);
    if (rc) return rc;
    b->pairs_built = true;
    return SI_OK;
}

// Restrict the shell pairs the batched entry points work on to (i, j <= i) with shell_begin <= i < shell_end
// (optionally without the diagonal pairs i == j).  Used by the per-class micro-benchmark (pure same-class
// batches: one shell of La after n shells of Lb) and for row-wise evaluation; (0, nshell, 0) restores the full set.
int si_basis_select_rows(si_basis* b, int shell_begin, int shell_end, int skip_diagonal) {
    if (!b || !basis_is_live(b)) return SI_ERR_ARG;
    if (shell_begin < 0 || shell_end > b->nshell || shell_begin >= shell_end) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(b->ctx->mu);
    SI_CUDA(cudaSetDevice(b->ctx->device));
    SI_CUDA(cudaDeviceSynchronize());
    free_pair_set(&b->full);
    for (auto& g : b->twoc_graphs) cudaGraphExecDestroy(g.exec);
    b->twoc_graphs.clear();
    b->pairs_built = false;
    int rc = build_pair_set(b, shell_begin, shell_end, &b->full, skip_diagonal != 0);
    if (rc) return rc;
    b->pairs_built = true;
    b->selected = !(shell_begin == 0 && shell_end == b->nshell && !skip_diagonal);
    return SI_OK;
}
long long si_basis_selected_rows(const si_basis* b) {
    if (!b) return -1;
    return b->pairs_built ? b->full.rows : b->packed_rows;
}

static BasisDev basis_dev(const si_basis* b, int sph, int normalize) {
    BasisDev d;
    d.L = b->d_L;
    d.nprim = b->d_nprim;
    d.poff = b->d_poff;
    d.foff = sph ? b->d_foff_sph : b->d_foff_cart;
    d.cen = b->d_cen;
    d.exps = b->d_exps;
    d.coef = (normalize == SI_NORM_PGTO) ? b->d_coef_pgto : b->d_coef;
    return d;
}

// fork/join helpers: class launches fan out over the context's streams
struct Fan {
    si_context* c;
    cudaStream_t user;
    int next = 0;
    int limit = SI_NSTREAMS;   // streams used by this fan
    bool used[SI_NSTREAMS];
    bool open = false;
    // The destructor joins the forked streams on every exit path (error returns inside the launch loops).
    ~Fan() {
        if (open) end();
    }
    int begin() {
        memset(used, 0, sizeof(used));
        // order this call after the previous call on this context, whatever stream that one was given
        if (c->ev_last_valid && !c->capturing) SI_CUDA(cudaStreamWaitEvent(user, c->ev_last, 0));
        SI_CUDA(cudaMemsetAsync(c->d_counters, 0, SI_NCOUNTERS * sizeof(int), user));
        c->next_counter = 0;
        SI_CUDA(cudaEventRecord(c->ev_fork, user));
        open = true;
        return SI_OK;
    }
    int stream(cudaStream_t* s) {
        const int nstreams = si_env().nstreams;
        int i = next++ % std::min(nstreams, limit);
        if (!used[i]) {
            SI_CUDA(cudaStreamWaitEvent(c->streams[i], c->ev_fork, 0));
            used[i] = true;
        }
        *s = c->streams[i];
        return SI_OK;
    }
    int end() {
        open = false;
        for (int i = 0; i < SI_NSTREAMS; ++i)
            if (used[i]) {
                used[i] = false;
                SI_CUDA(cudaEventRecord(c->ev_join[i], c->streams[i]));
                SI_CUDA(cudaStreamWaitEvent(user, c->ev_join[i], 0));
            }
        if (!c->capturing) {
            SI_CUDA(cudaEventRecord(c->ev_last, user));
            c->ev_last_valid = true;
        }
        return SI_OK;
    }
};

// marks the end of a call's device work on the caller's stream (see si_context::ev_last)
static int mark_done(si_context* c, cudaStream_t user, int rc) {
    if (rc == SI_OK) {
        SI_CUDA(cudaEventRecord(c->ev_last, user));
        c->ev_last_valid = true;
    }
    return rc;
}

static int pick_warps(const si_context* c, const SiProg& P, int max_warps, size_t* smem_bytes) {
    const size_t per = (size_t)((P.w_size + SI_NDYN + P.nconst + 1) & ~1) * sizeof(double);
    const size_t stg = (size_t)((P.nstages * sizeof(SiStage) + 7) / 8) * sizeof(double);
    int nw = (int)std::min<size_t>((size_t)max_warps, ((size_t)c->max_smem - stg) / per);
    if (nw < 1) return 0;
    *smem_bytes = stg + per * nw;
    return nw;
}

static int check_key_2idx(int key) {
    return (key == SI_OVLP || key == SI_KIN || key == SI_DPM || key == SI_QPM || key == SI_DQPM) ? 1 : 0;
}

static int run_two_index(si_context* c, const si_basis* cb, int key, int sph, int normalize, const double* R,
                         int ncharge, const double* d_cxyz, const double* d_q, double* out_dev, cudaStream_t user) {
    si_basis* b = const_cast<si_basis*>(cb);
    SI_CUDA(cudaSetDevice(c->device));
    int rc = build_pair_classes(b);
    if (rc) return rc;
    const int lim = (key == SI_2C2E) ? c->lauxmax : c->lmax;
    if (b->lmax > lim) {
        g_last_error = "basis angular momentum exceeds the context limit for this key";
        return SI_ERR_LMAX;
    }
    const long long nbf = sph ? b->nbf_sph : b->nbf_cart;
    OutDesc od;
    od.out = out_dev;
    od.ld = nbf;
    od.comp_stride = nbf * nbf;
    od.nbf = nbf;
    od.mode = STORE_MATRIX;
    BasisDev bd = basis_dev(b, sph, normalize);
    Fan fan{c, user};
    rc = fan.begin();
    if (rc) return rc;
    for (auto& pc : b->full.classes) {
        DevProgram* p;
        rc = get_program(c, key, pc.La, pc.Lb, 0, sph, normalize, &p);
        if (rc) return rc;
        size_t smem = 0;
        cudaStream_t s;
        rc = fan.stream(&s);
        if (rc) return rc;
        if (key == SI_COUL) {
            int nw = pick_warps(c, p->dev, 8, &smem);
            if (nw < 1) return SI_ERR_INTERNAL;
            int grid = std::min(pc.n, c->sm_count * 16);
            k_coul<<<grid, nw * 32, smem, s>>>(p->dev, c->boys_dev, bd, pc.d_items, pc.n, d_cxyz, d_q, ncharge, od);
        } else {
            int nw = pick_warps(c, p->dev, 4, &smem);
            if (nw < 1) return SI_ERR_INTERNAL;
            int grid = std::min((pc.n + nw - 1) / nw, c->sm_count * 16);
            int* counter = c->d_counters + (c->next_counter++ % SI_NCOUNTERS);
            k_two_index<<<grid, nw * 32, smem, s>>>(p->dev, c->boys_dev, bd, pc.d_items, pc.n, key == SI_2C2E ? 1 : 0,
                                                    R ? R[0] : 0.0, R ? R[1] : 0.0, R ? R[2] : 0.0, od, counter);
        }
        c->launches++;
        SI_CUDA(cudaGetLastError());
    }
    return fan.end();
}

// ovlp / kin / dpm / qpm (dqpm) through the fused kernel k_onee: ONE launch for all shell-pair classes and all
// requested operators (mask bit c = component c of si_onee.cuh; out[c] its (nbf, nbf) matrix)
static int run_onee(si_context* c, const si_basis* cb, unsigned mask, double* const* outc, int normalize, const double* R,
                    cudaStream_t user, int origin_p = 0) {
    si_basis* b = const_cast<si_basis*>(cb);
    SI_CUDA(cudaSetDevice(c->device));
    int rc = build_pair_classes(b);
    if (rc) return rc;
    if (b->lmax > c->lmax || b->lmax > ONEE_MAXL) {
        g_last_error = "basis angular momentum exceeds the context limit for this key";
        return SI_ERR_LMAX;
    }
    std::vector<const si_basis::PairClass*> cls;
    for (auto& pc : b->full.classes) cls.push_back(&pc);
    if (cls.size() > 15) return SI_ERR_INTERNAL;
    // heaviest blocks first
    std::sort(cls.begin(), cls.end(), [](const si_basis::PairClass* x, const si_basis::PairClass* y) {
        return si_ncart(x->La) * si_ncart(x->Lb) > si_ncart(y->La) * si_ncart(y->Lb);
    });
    // three launches (classes by block size, see k_onee), concurrent on three streams
    Fan fan{c, user};
    fan.limit = 3;
    rc = fan.begin();
    if (rc) return rc;
    for (int big = 2; big >= 0; --big) {
        OneEArgs A;
        memset(&A, 0, sizeof(A));
        int wsz_max = 0, g = 0, ns = 0;
        for (size_t i = 0; i < cls.size(); ++i) {
            const int la = cls[i]->La, lb = cls[i]->Lb, tpw = 32 / onee_ept(la, lb);
            if (onee_size_class(la, lb) != big) continue;
            A.seg[ns].items = cls[i]->d_items;
            A.seg[ns].n = cls[i]->n;
            A.seg[ns].la = la;
            A.seg[ns].lb = lb;
            g += (cls[i]->n + tpw - 1) / tpw;
            A.seg[ns].group_end = g;
            wsz_max = std::max(wsz_max, onee_wsz(la, lb) * tpw);
            ++ns;
        }
        if (!ns) continue;
        A.nseg = ns;
        A.ngroups = g;
        A.ppair = b->full.d_ppair[normalize == SI_NORM_PGTO ? 1 : 0];
        A.exps = b->d_exps;
        A.poff = b->d_poff;
        for (int d = 0; d < 3; ++d) A.R[d] = R ? R[d] : 0.0;
        A.origin_p = origin_p;
        for (int k = 0; k < ONEE_NCOMP; ++k) A.out[k] = ((mask >> k) & 1u) ? outc[k] : nullptr;
        A.nbf = b->nbf_sph;
        A.counter = c->d_counters + (c->next_counter++ % SI_NCOUNTERS);
        const int warps = 4;
        const size_t smem = (size_t)warps * wsz_max * sizeof(double);
        // resident blocks per SM for this shared-memory size (cached per context: the occupancy query costs ~10 us of host time)
        const size_t okey = smem * 4 + big;
        auto occ_it = c->onee_occ.find(okey);
        if (occ_it == c->onee_occ.end()) {
            int occ = 0;
            cudaError_t e = big == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_onee<4, 2>, warps * 32, smem)
                            : big == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_onee<4, 1>, warps * 32, smem)
                                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_onee<4, 0>, warps * 32, smem);
            if (e != cudaSuccess || occ < 1) occ = 1;
            occ_it = c->onee_occ.emplace(okey, occ).first;
        }
        const int grid = std::max(1, std::min((g + warps - 1) / warps, c->sm_count * occ_it->second));
        cudaStream_t s;
        rc = fan.stream(&s);
        if (rc) return rc;
        if (big == 2)
            k_onee<4, 2><<<grid, warps * 32, smem, s>>>(A, mask, wsz_max);
        else if (big == 1)
            k_onee<4, 1><<<grid, warps * 32, smem, s>>>(A, mask, wsz_max);
        else
            k_onee<4, 0><<<grid, warps * 32, smem, s>>>(A, mask, wsz_max);
        c->launches++;
        SI_CUDA(cudaGetLastError());
    }
    return fan.end();
}

static bool onee_usable(const si_context* c, const si_basis* b, int sph, int normalize) {
    return sph && normalize != SI_NORM_NONE && !si_env().force_lut && b->lmax <= ONEE_MAXL && b->lmax <= c->lmax;
}

int si_int1e_set(si_context* c, const si_basis* b, int sph, int normalize, const double* R, double* out_dev, void* stream) {
    if (!c || !b || !out_dev) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    const long long nbf = sph ? b->nbf_sph : b->nbf_cart;
    if (onee_usable(c, b, sph, normalize)) {
        double* outc[ONEE_NCOMP];
        for (int k = 0; k < ONEE_NCOMP; ++k) outc[k] = out_dev + (long long)k * nbf * nbf;
        return mark_done(c, (cudaStream_t)stream, run_onee(c, b, 0x7FFu, outc, normalize, R, (cudaStream_t)stream));
    }
    // Cartesian / un-normalised output: the four operators one after the other on the stage machine
    const int keys[4] = {SI_OVLP, SI_KIN, SI_DPM, SI_QPM}, offs[4] = {0, 1, 2, 5};
    for (int k = 0; k < 4; ++k) {
        int rc = run_two_index(c, b, keys[k], sph, normalize, R, 0, nullptr, nullptr, out_dev + (long long)offs[k] * nbf * nbf,
                               (cudaStream_t)stream);
        if (rc) return rc;
    }
    return mark_done(c, (cudaStream_t)stream, SI_OK);
}

// multipole3d_sph: the nine real regular solid harmonic operators R_lm, l <= 2, m = -l..l (coefficients of
// sympleints/sym_solid_harmonics.py, natural m order) from the ten Cartesian multipoles 1, x..z, xx..zz about P
__global__ void k_multi_sph_combine(const double* cart11, double* out9, long long n2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const double s3 = 1.7320508075688772;
    const double S = cart11[i], x = cart11[2 * n2 + i], y = cart11[3 * n2 + i], z = cart11[4 * n2 + i];
    const double xx = cart11[5 * n2 + i], xy = cart11[6 * n2 + i], xz = cart11[7 * n2 + i], yy = cart11[8 * n2 + i],
                 yz = cart11[9 * n2 + i], zz = cart11[10 * n2 + i];
    out9[i] = S;
    out9[n2 + i] = y;
    out9[2 * n2 + i] = z;
    out9[3 * n2 + i] = x;
    out9[4 * n2 + i] = s3 * xy;
    out9[5 * n2 + i] = s3 * yz;
    out9[6 * n2 + i] = zz - 0.5 * (xx + yy);
    out9[7 * n2 + i] = s3 * xz;
    out9[8 * n2 + i] = 0.5 * s3 * (xx - yy);
}

// per shell block: operators of order l > La + Lb are set to zero, as the reference does (defs/multipole.py:118-121)
__global__ void k_multi_sph_truncate(double* out9, long long nbf, const int* shell_of_fn, const int* L) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbf * nbf) return;
    const int lsum = L[shell_of_fn[i / nbf]] + L[shell_of_fn[i % nbf]];
    if (lsum < 1)
        for (int c = 1; c < 4; ++c) out9[c * nbf * nbf + i] = 0.0;
    if (lsum < 2)
        for (int c = 4; c < 9; ++c) out9[c * nbf * nbf + i] = 0.0;
}

int si_int1e(si_context* c, const si_basis* b, int key, int sph, int normalize, const double* R, double* out_dev,
             void* stream) {
    if (!c || !b || !out_dev) return SI_ERR_ARG;
    if (key == SI_MULTI_SPH) {
        // (sph + cgto|pgto only: served by the fused kernel with the origin at P of every primitive pair)
        if (!onee_usable(c, b, sph, normalize)) {
            g_last_error = "multi_sph is available for spherical, normalised output only";
            return SI_ERR_KEY;
        }
        std::lock_guard<std::recursive_mutex> lk1(c->mu);
        cudaStream_t s = (cudaStream_t)stream;
        const long long nbf = b->nbf_sph, n2 = nbf * nbf;
        double* scr = nullptr;
        int rc = ctx_scratch(c, (size_t)11 * n2 + (size_t)nbf, s, &scr);
        if (rc) return rc;
        double* outc[ONEE_NCOMP];
        for (int k = 0; k < ONEE_NCOMP; ++k) outc[k] = scr + (long long)k * n2;
        rc = run_onee(c, b, 0x7FDu, outc, normalize, nullptr, s, 1);
        if (rc) return rc;
        k_multi_sph_combine<<<(unsigned)((n2 + 255) / 256), 256, 0, s>>>(scr, out_dev, n2);
        // shell of every function (uploaded next to the Cartesian scratch)
        std::vector<int> sof((size_t)nbf);
        for (int sh = 0; sh < b->nshell; ++sh)
            for (int m = 0; m < si_nsph(b->L[sh]); ++m) sof[(size_t)b->foff_sph[sh] + m] = sh;
        int* d_sof = reinterpret_cast<int*>(scr + (size_t)11 * n2);
        SI_CUDA(cudaMemcpyAsync(d_sof, sof.data(), (size_t)nbf * sizeof(int), cudaMemcpyHostToDevice, s));
        k_multi_sph_truncate<<<(unsigned)((n2 + 255) / 256), 256, 0, s>>>(out_dev, nbf, d_sof, b->d_L);
        c->launches += 2;
        SI_CUDA(cudaGetLastError());
        SI_CUDA(cudaStreamSynchronize(s));   // `sof` is a host temporary
        return mark_done(c, s, SI_OK);
    }
    if (!check_key_2idx(key)) return SI_ERR_KEY;
    if (onee_usable(c, b, sph, normalize)) {
        std::lock_guard<std::recursive_mutex> lk1(c->mu);
        const long long n2 = (long long)b->nbf_sph * b->nbf_sph;
        double* outc[ONEE_NCOMP] = {nullptr};
        unsigned mask = 0;
        switch (key) {
            case SI_OVLP: mask = 1u; outc[0] = out_dev; break;
            case SI_KIN: mask = 2u; outc[1] = out_dev; break;
            case SI_DPM: mask = 0x1Cu; for (int k = 0; k < 3; ++k) outc[2 + k] = out_dev + k * n2; break;
            case SI_QPM: mask = 0x7E0u; for (int k = 0; k < 6; ++k) outc[5 + k] = out_dev + k * n2; break;
            default: mask = (1u << 5) | (1u << 8) | (1u << 10); outc[5] = out_dev; outc[8] = out_dev + n2; outc[10] = out_dev + 2 * n2; break;
        }
        return mark_done(c, (cudaStream_t)stream, run_onee(c, b, mask, outc, normalize, R, (cudaStream_t)stream));
    }
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    return mark_done(c, (cudaStream_t)stream,
                     run_two_index(c, b, key, sph, normalize, R, 0, nullptr, nullptr, out_dev, (cudaStream_t)stream));
}

static int int3c2e_impl(si_context* c, const si_basis* corb, const si_basis* caux, int sb, int se, int normalize,
                        int layout, const si_basis::PairSet* pset, double* out_dev, size_t out_len, void* stream);
static bool g3_disabled();
static bool g3_config(const si_context* c, const int* g3, long long ngroups, int* grid, int* block, size_t* smem);
static const int* gc_find(int La, int Lb) {
    const int id = La * 10 + Lb;
    const int n = (int)(sizeof(gc_table) / sizeof(gc_table[0]));
    for (int i = 0; i < n; ++i)
        if (gc_table[i][0] == id) return gc_table[i];
    return nullptr;
}

int si_int2c2e(si_context* c, const si_basis* b, int sph, int normalize, double* out_dev, void* stream) {
    if (!c || !b || !out_dev) return SI_ERR_ARG;
    if (b->lmax > c->lauxmax) return SI_ERR_LMAX;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (!sph || normalize == SI_NORM_NONE || g3_disabled())
        return mark_done(c, (cudaStream_t)stream, run_two_index(c, b, SI_2C2E, sph, normalize, nullptr, 0, nullptr, nullptr,
                                                                out_dev, (cudaStream_t)stream));
    // Spherical 2c2e through the generated 3c2e kernels: (a|b) = (a s0|b), s0 = unit s function with exponent 0
    // (then p = a, P = A, K_ab = 1 and the 3c2e base integral reduces to the 2c2e one, defs/coulomb.py:157-160
    // vs :317-323).  Pairs (shell j, s0) in shell order give the rows of the (naux, naux) matrix.
    SI_CUDA(cudaSetDevice(c->device));
    if (!b->twoc_orb) {
        std::vector<int> L = b->L, np = b->nprim;
        std::vector<double> cen = b->cen, ex = b->exps, co = b->coef;
        L.push_back(0);
        np.push_back(1);
        cen.insert(cen.end(), {0.0, 0.0, 0.0});
        ex.push_back(0.0);
        co.push_back(1.0);
        si_basis* tb = nullptr;
        int rc = si_basis_create(c, (int)L.size(), L.data(), np.data(), cen.data(), ex.data(), co.data(), &tb);
        if (rc) return rc;
        // the pgto radial factor of an exponent-0 function is 0: the dummy keeps coefficient 1 in both variants
        tb->coef_pgto.back() = 1.0;
        SI_CUDA(cudaMemcpy(tb->d_coef_pgto + (tb->coef_pgto.size() - 1), &tb->coef_pgto.back(), sizeof(double),
                           cudaMemcpyHostToDevice));
        auto* ps = new si_basis::PairSet;
        rc = build_pair_set(tb, tb->nshell - 1, tb->nshell, ps, true);
        if (rc) return rc;
        b->twoc_orb = tb;
        b->twoc_pairs = ps;
    }
    if (!b->d_fn_L) {
        std::vector<signed char> fl((size_t)b->nbf_sph);
        for (int sh = 0; sh < b->nshell; ++sh)
            for (int m = 0; m < si_nsph(b->L[sh]); ++m) fl[(size_t)b->foff_sph[sh] + m] = (signed char)b->L[sh];
        SI_CUDA(cudaMalloc(&b->d_fn_L, fl.size()));
        SI_CUDA(cudaMemcpy(b->d_fn_L, fl.data(), fl.size(), cudaMemcpyHostToDevice));
        SI_CUDA(cudaDeviceSynchronize());
    }
    const size_t n = (size_t)b->nbf_sph;
    cudaStream_t user = (cudaStream_t)stream;
    auto enqueue = [&]() -> int {
        // only the classes (La s0|Lb) with La >= Lb are evaluated (21 of the 36 for lauxmax = 5); the other blocks are
        // their transposes.  Source and destination elements of the mirror are disjoint, so it runs in place.
        c->twoc_lower_only = true;
        int rc = int3c2e_impl(c, b->twoc_orb, b, 0, b->nshell, normalize, SI_LAYOUT_PACKED, b->twoc_pairs, out_dev, n * n, stream);
        c->twoc_lower_only = false;
        if (rc) return rc;
        const unsigned nt = (unsigned)((n + 31) / 32);
        k_mirror_by_l<<<dim3(nt, nt), 256, 0, user>>>(out_dev, (long long)n, b->d_fn_L);
        c->launches++;
        SI_CUDA(cudaGetLastError());
        return SI_OK;
    };
    if (!si_env().graphs || !b->twoc_warm || si_env().g3_profile || user == nullptr) {
        // (the legacy default stream cannot be captured)
        int rc = enqueue();
        if (rc == SI_OK) b->twoc_warm = true;
        return mark_done(c, user, rc);
    }
    const si_basis::TwocGraph* tg = nullptr;
    for (auto& g : b->twoc_graphs)
        if (g.out == out_dev && g.normalize == normalize) tg = &g;
    if (!tg) {
        if (c->ev_last_valid) SI_CUDA(cudaStreamWaitEvent(user, c->ev_last, 0));
        const long long l0 = c->launches;
        SI_CUDA(cudaStreamBeginCapture(user, cudaStreamCaptureModeThreadLocal));
        c->capturing = true;
        int rc = enqueue();
        c->capturing = false;
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(user, &graph);
        const int nk = (int)(c->launches - l0);
        c->launches = l0;
        if (rc != SI_OK || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            // capture not possible (e.g. a cache had to be refilled): issue the launches directly
            rc = enqueue();
            return mark_done(c, user, rc);
        }
        si_basis::TwocGraph g;
        g.out = out_dev;
        g.normalize = normalize;
        g.kernels = nk;
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            g_last_error = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e);
            return SI_ERR_CUDA;
        }
        if (b->twoc_graphs.size() >= 4) {
            cudaGraphExecDestroy(b->twoc_graphs.front().exec);
            b->twoc_graphs.erase(b->twoc_graphs.begin());
        }
        b->twoc_graphs.push_back(g);
        tg = &b->twoc_graphs.back();
    } else if (c->ev_last_valid) {
        SI_CUDA(cudaStreamWaitEvent(user, c->ev_last, 0));
    }
    SI_CUDA(cudaGraphLaunch(tg->exec, user));
    c->launches += tg->kernels;
    return mark_done(c, user, SI_OK);
}

// out[i] += sum_{sp >= 1} part[sp - 1][i] in a fixed order (charge-split partial matrices of the coul kernels)
__global__ void k_sum_slices(double* out, const double* part, long long n, int nslices) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = out[i];
    for (int sp = 0; sp < nslices; ++sp) v += part[(long long)sp * n + i];
    out[i] = v;
}

// Nuclear attraction through the generated McMurchie-Davidson kernels (codegen/gencoul.py): one tuple per
// shell pair, all charges inside; every class on its own stream so that the tails overlap.
static int run_coul_generated(si_context* c, const si_basis* cb, int normalize, int ncharge, const double* d_c4,
                              double* out_dev, cudaStream_t user) {
    si_basis* b = const_cast<si_basis*>(cb);
    int rc = build_pair_classes(b);
    if (rc) return rc;
    if (b->lmax > c->lmax) {
        g_last_error = "basis angular momentum exceeds the context limit for this key";
        return SI_ERR_LMAX;
    }
    // highest Boys order of the basis (La + Lb <= 2 lmax) against the table (IndexError in the reference)
    if (2 * b->lmax > c->nrows - 7) {
        g_last_error = "Boys table too small for this basis";
        return SI_ERR_BOYS;
    }
    const long long nbf = b->nbf_sph;
    OutDesc od;
    od.out = out_dev;
    od.ld = nbf;
    od.comp_stride = nbf * nbf;
    od.nbf = nbf;
    od.mode = STORE_MATRIX;
    // Charge split: a class with few shell pairs would leave most SMs idle, so its charges are divided over
    // nsplit tuples per pair; split 0 writes `out`, split sp >= 1 the slice sp-1 of a zeroed scratch buffer.
    const int target_warps = c->sm_count * si_env().gc_warps_per_sm, max_split = si_env().gc_max_split;
    std::vector<int> nsplits;
    int top_split = 1;
    for (auto& pc : b->full.classes) {
        const int* gc = gc_find(pc.La, pc.Lb);
        if (!gc) return SI_ERR_INTERNAL;
        const long long pg = (pc.n + gc[3] - 1) / gc[3];
        int ns = (int)std::min<long long>(std::min(max_split, std::max(1, ncharge / 16)), (target_warps + pg - 1) / pg);
        ns = std::max(1, ns);
        nsplits.push_back(ns);
        top_split = std::max(top_split, ns);
    }
    double* d_part = nullptr;
    if (top_split > 1) {
        // the charges occupy the first 4 * ncharge doubles of the context scratch (si_coul)
        double* base = nullptr;
        rc = ctx_scratch(c, (size_t)4 * ncharge + (size_t)(top_split - 1) * nbf * nbf, user, &base);
        if (rc) return rc;
        if (base + 0 != d_c4) {
            g_last_error = "coul scratch moved";
            return SI_ERR_INTERNAL;
        }
        d_part = base + (size_t)4 * ncharge;
        SI_CUDA(cudaMemsetAsync(d_part, 0, (size_t)(top_split - 1) * nbf * nbf * sizeof(double), user));
    }
    const bool profile = si_env().gc_profile;   // SI_GC_PROFILE=1: classes one after the other, each timed (tuning only)
    Fan fan{c, user};
    rc = fan.begin();
    if (rc) return rc;
    size_t ci = 0;
    for (auto& pc : b->full.classes) {
        const int* gc = gc_find(pc.La, pc.Lb);
        cudaStream_t s;
        if (profile) {
            s = user;
        } else {
            rc = fan.stream(&s);
            if (rc) return rc;
        }
        GCArgs ga;
        ga.pairs = pc.d_items;
        ga.ppair = b->full.d_ppair[normalize == SI_NORM_PGTO ? 1 : 0];
        ga.npairs = pc.n;
        ga.c4 = d_c4;
        ga.ncharge = ncharge;
        ga.nsplit = nsplits[ci++];
        ga.ngroups = (long long)((pc.n + gc[3] - 1) / gc[3]) * ga.nsplit;
        ga.od = od;
        ga.part = d_part;
        ga.counter = c->d_counters + (c->next_counter++ % SI_NCOUNTERS);
        ga.boys = c->boys_dev;
        int grid, block;
        size_t smem;
        c->small_call = false;
        if (!g3_config(c, gc, ga.ngroups, &grid, &block, &smem)) return SI_ERR_INTERNAL;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (profile) {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, s);
        }
        if (gc_launch(gc[1], gc[0], ga, grid, block, smem, c->max_smem, s)) return SI_ERR_INTERNAL;
        c->launches++;
        SI_CUDA(cudaGetLastError());
        if (profile) {
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            fprintf(stderr, "GCPROF %d%d %.4f nsplit=%d grid=%d pairs=%d\n", pc.La, pc.Lb, ms, ga.nsplit, grid, pc.n);
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
    }
    rc = fan.end();
    if (rc) return rc;
    if (d_part) {
        const long long n = nbf * nbf;
        k_sum_slices<<<(unsigned)((n + 255) / 256), 256, 0, user>>>(out_dev, d_part, n, top_split - 1);
        c->launches++;
        SI_CUDA(cudaGetLastError());
    }
    return SI_OK;
}

int si_coul(si_context* c, const si_basis* b, int sph, int normalize, int ncharge, const double* xyz, const double* q,
            double* out_dev, void* stream) {
    if (!c || !b || !out_dev || ncharge <= 0 || !xyz || !q) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    SI_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const bool generated = sph && normalize != SI_NORM_NONE && !g3_disabled() && b->lmax <= 4;
    // scratch: [xyz 3n][q n][partial matrices of the charge split]; sized for the largest split up front so
    // that it does not move between the copy of the charges and the launches
    const size_t nbf = (size_t)b->nbf_sph;
    double* dx = nullptr;
    int rc = ctx_scratch(c, (size_t)4 * ncharge + (generated ? (size_t)(SI_GC_MAX_SPLIT - 1) * nbf * nbf : 0), s, &dx);
    if (rc) return rc;
    if (generated) {
        // packed records (x, y, z, q): one 32-byte sector per charge
        std::vector<double> c4((size_t)4 * ncharge);
        for (int i = 0; i < ncharge; ++i) {
            c4[4 * (size_t)i] = xyz[3 * (size_t)i];
            c4[4 * (size_t)i + 1] = xyz[3 * (size_t)i + 1];
            c4[4 * (size_t)i + 2] = xyz[3 * (size_t)i + 2];
            c4[4 * (size_t)i + 3] = q[i];
        }
        SI_CUDA(cudaMemcpyAsync(dx, c4.data(), c4.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        SI_CUDA(cudaStreamSynchronize(s));   // c4 is a temporary (pageable copies are staged, but be explicit)
        return mark_done(c, s, run_coul_generated(c, b, normalize, ncharge, dx, out_dev, s));
    }
    double* dq = dx + (size_t)3 * ncharge;
    SI_CUDA(cudaMemcpyAsync(dx, xyz, (size_t)ncharge * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    SI_CUDA(cudaMemcpyAsync(dq, q, (size_t)ncharge * sizeof(double), cudaMemcpyHostToDevice, s));
    return mark_done(c, s, run_two_index(c, b, SI_COUL, sph, normalize, nullptr, ncharge, dx, dq, out_dev, s));
}

// ---- generated class-specialised 3c2e kernels (codegen/gen3c2e.py) ----------------------------
static const int* g3_find(int La, int Lb, int Lc) {
    const int id = La * 100 + Lb * 10 + Lc;
    const int n = (int)(sizeof(g3_table) / sizeof(g3_table[0]));
    for (int i = 0; i < n; ++i)
        if (g3_table[i][0] == id) return g3_table[i];
    return nullptr;
}
static bool g3_disabled() { return si_env().force_lut; }
static bool g3_profile() { return si_env().g3_profile; }
static void g3_fill_args(G3Args& ga, si_context* c, const BasisDev& orb, const BasisDev& aux, const double* aux_oexp,
                         const SiPairItem* pairs,
                         const double* ppair, int npairs, const int* auxlist, const SiAuxItem* auxitems, int naux_c, int col_begin,
                         const OutDesc& od, int* counter, int tpw) {
    ga.orb = orb;
    ga.aux = aux;
    ga.aux_oexp = aux_oexp;
    ga.pairs = pairs;
    ga.ppair = ppair;
    ga.npairs = npairs;
    ga.auxlist = auxlist;
    ga.auxitems = auxitems;
    ga.naux_c = naux_c;
    ga.col_begin = col_begin;
    ga.groups_per_pair = (naux_c + tpw - 1) / tpw;
    ga.ngroups = (long long)npairs * ga.groups_per_pair;
    ga.tuples = nullptr;
    ga.chunk = 1;
    ga.screen_eps2 = c->screen_eps * c->screen_eps;
    ga.od = od;
    ga.counter = counter;
    ga.boys = c->boys_dev;
}
static bool g3_config(const si_context* c, const int* g3, long long ngroups, int* grid, int* block, size_t* smem) {
    const size_t per_warp = (size_t)g3[4] * g3[3] * sizeof(double);  // WSZ * TPW
    const int wpb = si_env().g3_warps_per_block;   // warps per block (the kernels are compiled for up to 4)
    int nw = (int)std::min<size_t>((size_t)wpb, (size_t)c->max_smem / per_warp);
    if (nw < 1) return false;
    *block = nw * 32;
    *smem = per_warp * nw;
    // persistent blocks: at most 4 x 4 warps are resident per SM.  A call that is small as a whole (the 2c2e metric) fans
    // its launches out over all streams; one block per SM each lets them run side by side (330 -> 290 us for the metric)
    const int per_sm = c->small_call ? 1 : si_env().g3_grid_per_sm;
    *grid = (int)std::max<long long>(1, std::min<long long>((ngroups + nw - 1) / nw, (long long)c->sm_count * per_sm * (4 / nw)));
    return true;
}

static int int3c2e_impl(si_context* c, const si_basis* corb, const si_basis* caux, int sb, int se, int normalize,
                        int layout, const si_basis::PairSet* pset, double* out_dev, size_t out_len, void* stream);

int si_int3c2e(si_context* c, const si_basis* corb, const si_basis* caux, int sb, int se, int normalize, int layout,
               double* out_dev, size_t out_len, void* stream) {
    if (!c) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    return int3c2e_impl(c, corb, caux, sb, se, normalize, layout, nullptr, out_dev, out_len, stream);
}

// Explicit tuple table of one class launch: runs of pairs with equal contraction lengths, their (pair, aux)
// tuples pair-major, tpw per group, the last group of a run padded with -1.  Built once per (aux range, class).
static int get_tuple_table(const si_basis::PairClass* cpc, long long aux_uid, int sb, int se, int Lc, int naux_c,
                           int tpw, const si_basis::TupleTable** out) {
    si_basis::PairClass* pc = const_cast<si_basis::PairClass*>(cpc);
    auto key = std::make_tuple(aux_uid, sb, se, Lc, tpw);
    auto itb = pc->tables.find(key);
    if (itb == pc->tables.end()) {
        std::vector<int2> tab;
        tab.reserve((size_t)pc->n * naux_c + 64);
        int p0 = 0;
        while (p0 < pc->n) {
            int p1 = p0;
            while (p1 < pc->n && pc->kab[p1] == pc->kab[p0]) ++p1;
            for (int ip = p0; ip < p1; ++ip)
                for (int ic = 0; ic < naux_c; ++ic) tab.push_back(make_int2(ip, ic));
            while (tab.size() % tpw) tab.push_back(make_int2(-1, -1));
            p0 = p1;
        }
        si_basis::TupleTable tt;
        tt.ngroups = (long long)(tab.size() / tpw);
        SI_CUDA(cudaMalloc((void**)&tt.d, tab.size() * sizeof(int2)));
        SI_CUDA(cudaMemcpy(tt.d, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice));
        // cudaMemcpy from pageable memory may return before the DMA has landed, and the class kernels run on
        // non-blocking streams that were forked BEFORE this upload: wait for it explicitly (cache miss only)
        SI_CUDA(cudaDeviceSynchronize());
        if (pc->tables.size() >= 64) {   // bound the cache
            cudaFree(pc->tables.begin()->second.d);
            pc->tables.erase(pc->tables.begin());
        }
        itb = pc->tables.emplace(key, tt).first;
    }
    *out = &itb->second;
    return SI_OK;
}

static int int3c2e_impl(si_context* c, const si_basis* corb, const si_basis* caux, int sb, int se, int normalize,
                        int layout, const si_basis::PairSet* pset, double* out_dev, size_t out_len, void* stream) {
    if (!c || !corb || !caux || !out_dev) return SI_ERR_ARG;
    si_basis* orb = const_cast<si_basis*>(corb);
    const si_basis* aux = caux;
    if (sb < 0 || se > aux->nshell || sb >= se) return SI_ERR_ARG;
    if ((orb->lmax > c->lmax && pset == nullptr) || aux->lmax > c->lauxmax) return SI_ERR_LMAX;
    SI_CUDA(cudaSetDevice(c->device));
    int rc = SI_OK;
    if (!pset) {
        rc = build_pair_classes(orb);
        if (rc) return rc;
    }
    const int col_begin = aux->foff_sph[sb];
    const int col_end = (se == aux->nshell) ? aux->nbf_sph : aux->foff_sph[se];
    const long long naux_local = col_end - col_begin;
    if (!pset) pset = &orb->full;
    if (pset != &orb->full && layout != SI_LAYOUT_PACKED) return SI_ERR_ARG;
    const long long rows = (layout == SI_LAYOUT_FULL) ? (long long)orb->nbf_sph * orb->nbf_sph : pset->rows;
    if ((unsigned long long)(rows * naux_local) > out_len) {
        g_last_error = "output buffer too small";
        return SI_ERR_ARG;
    }
    OutDesc od;
    od.out = out_dev;
    od.ld = naux_local;
    od.comp_stride = 0;
    od.nbf = orb->nbf_sph;
    od.mode = (layout == SI_LAYOUT_FULL) ? STORE_3C_FULL : STORE_3C_PACKED;
    cudaStream_t user = (cudaStream_t)stream;
    // aux shells of the range by angular momentum (device list cached per range)
    const si_basis::AuxRange* ar = nullptr;
    for (auto& r : aux->aux_ranges)
        if (r.sb == sb && r.se == se) ar = &r;
    if (!ar) {
        std::map<int, std::vector<int>> by_l;
        for (int s = sb; s < se; ++s) by_l[aux->L[s]].push_back(s);
        std::vector<int> flat;
        si_basis::AuxRange r;
        r.sb = sb;
        r.se = se;
        for (auto& kv : by_l) {
            r.span[kv.first] = {(int)flat.size(), (int)kv.second.size()};
            flat.insert(flat.end(), kv.second.begin(), kv.second.end());
        }
        SI_CUDA(cudaMalloc((void**)&r.d_list, flat.size() * sizeof(int)));
        SI_CUDA(cudaMemcpy(r.d_list, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice));
        std::vector<SiAuxItem> aitems(flat.size());
        for (size_t i = 0; i < flat.size(); ++i) {
            const int sh = flat[i];
            memset(&aitems[i], 0, sizeof(SiAuxItem));
            aitems[i].shell = sh;
            aitems[i].Kc = aux->nprim[sh];
            aitems[i].pc0 = aux->poff[sh];
            aitems[i].foff = aux->foff_sph[sh];
            for (int d = 0; d < 3; ++d) aitems[i].C[d] = aux->cen[3 * sh + d];
        }
        SI_CUDA(cudaMalloc((void**)&r.d_items, aitems.size() * sizeof(SiAuxItem)));
        SI_CUDA(cudaMemcpy(r.d_items, aitems.data(), aitems.size() * sizeof(SiAuxItem), cudaMemcpyHostToDevice));
        SI_CUDA(cudaDeviceSynchronize());   // (see build_pair_set)
        if (aux->aux_ranges.size() >= 64) {
            cudaFree(aux->aux_ranges.front().d_list);
            cudaFree(aux->aux_ranges.front().d_items);
            aux->aux_ranges.erase(aux->aux_ranges.begin());
        }
        aux->aux_ranges.push_back(r);
        ar = &aux->aux_ranges.back();
    }
    const std::map<int, std::pair<int, int>>& span = ar->span;
    int* d_list = ar->d_list;
    const SiAuxItem* d_aitems = ar->d_items;
    BasisDev od_orb = basis_dev(orb, 1, normalize);
    BasisDev od_aux = basis_dev(aux, 1, normalize);
    Fan fan{c, user};
    rc = fan.begin();
    if (rc) return rc;
    // heaviest classes first
    struct Job {
        const si_basis::PairClass* pc;
        int Lc;
        double cost;
    };
    std::vector<Job> jobs;
    for (auto& pc : pset->classes)
        for (auto& kv : span) {
            if (c->twoc_lower_only && kv.first > pc.La) continue;   // 2c2e: (a|b) blocks with La < Lb come from the mirror
            DevProgram* p;
            rc = get_program(c, SI_3C2E, pc.La, pc.Lb, kv.first, 1, normalize, &p);
            if (rc) return rc;
            double cost = (double)(p->host.n_fma_prim + p->host.n_fma_contr) * pc.n * kv.second.second;
            jobs.push_back({&pc, kv.first, cost});
        }
    std::sort(jobs.begin(), jobs.end(), [](const Job& a, const Job& b) { return a.cost > b.cost; });
    bool use_pdl = false;
    {
        // a call that is small as a whole (the 2c2e metric, tiny systems) is launch-latency bound: use all streams
        double total = 0.0;
        for (auto& j : jobs) total += j.cost;
        // Programmatic dependent launch (one stream, every launch's blocks fill the SMs as the previous launch's
        // blocks retire) pays when the launches are short, e.g. one rank's shard of 4 or 8: measured 9.36 -> 9.09 ms
        // for a 1/8 shard, but 65.4 -> 66.1 ms for the whole tensor, where two alternating streams stay the default.
        use_pdl = si_env().pdl < 0 ? (2.0 * total < si_env().g3_pdl_total && 2.0 * total >= si_env().g3_small_total)
                                   : si_env().pdl != 0;
        fan.limit = use_pdl ? 1 : si_env().g3_fan;
        c->small_call = false;
        if (2.0 * total < si_env().g3_small_total && !use_pdl) {
            fan.limit = SI_NSTREAMS;
            c->small_call = true;
        }
    }
    for (auto& job : jobs) {
        DevProgram* p;
        rc = get_program(c, SI_3C2E, job.pc->La, job.pc->Lb, job.Lc, 1, normalize, &p);
        if (rc) return rc;
        auto sp = span.at(job.Lc);
        cudaStream_t s;
        // launches that fill the GPU on their own run back to back on one stream (concurrent residency
        // of kernels with different shared-memory carve-outs costs more than their tails); small ones fan out
        const bool force_fan = si_env().force_fan;
        // "big" = long enough that its tail (one shell-tuple latency, 20-100 us) does not matter; shorter
        // launches (small systems, one rank's shard of many) alternate between two streams so that the
        // next kernel's blocks fill the tail of the previous one
        const bool big = !force_fan && 2.0 * job.cost >= si_env().g3_big_flops;
        if (big) fan.next = 0;
        rc = fan.stream(&s);
        if (rc) return rc;
        int* counter = c->d_counters + (c->next_counter++ % SI_NCOUNTERS);
        const int* g3 = g3_find(job.pc->La, job.pc->Lb, job.Lc);
        if (g3 && normalize != SI_NORM_NONE && !g3_disabled()) {
            G3Args ga;
            g3_fill_args(ga, c, od_orb, od_aux, aux->d_oexp, job.pc->d_items, pset->d_ppair[normalize == SI_NORM_PGTO ? 1 : 0], job.pc->n, d_list + sp.first, d_aitems + sp.first, sp.second, col_begin,
                         od, counter, g3[3]);
            {
                // per-pair packing leaves lanes idle when the range holds few aux shells of this class: switch to an
                // explicit tuple table (cached) when more than ~10 % of the tuple slots would be empty
                const bool use_tables = si_env().tuple_tables;
                const int tpw = g3[3];
                const double fill = (double)sp.second / ((double)ga.groups_per_pair * tpw);
                const long long nt = (long long)job.pc->n * sp.second;
                if (use_tables && tpw >= 4 && fill < 0.9 && pset == &orb->full && nt * 8 <= (256ll << 20)) {
                    const si_basis::TupleTable* tt = nullptr;
                    rc = get_tuple_table(job.pc, aux->uid, sb, se, job.Lc, sp.second, tpw, &tt);
                    if (rc) return rc;
                    ga.tuples = tt->d;
                    ga.ngroups = tt->ngroups;
                }
            }
            if (ga.ngroups >= (1ll << 31) - 65536) {
                // 32-bit work counter and group arithmetic in the kernels (2^31 groups = ~1e10 shell triples of one class)
                g_last_error = "too many shell tuples in one class for one call: split the auxiliary range";
                return SI_ERR_ARG;
            }
            int grid, block;
            size_t smem;
            if (!g3_config(c, g3, ga.ngroups, &grid, &block, &smem)) {
                g_last_error = "generated class does not fit into shared memory";
                return SI_ERR_INTERNAL;
            }
            if (g3_profile()) {
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                cudaDeviceSynchronize();
                cudaEventRecord(e0, s);
                if (g3_launch(g3[1], g3[0], ga, grid, block, smem, c->max_smem, s)) return SI_ERR_INTERNAL;
                cudaEventRecord(e1, s);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                fprintf(stderr, "G3PROF %d%d%d %.4f\n", job.pc->La, job.pc->Lb, job.Lc, ms);
                cudaEventDestroy(e0);
                cudaEventDestroy(e1);
            } else if (g3_launch(g3[1], g3[0], ga, grid, block, smem, use_pdl ? -1 : c->max_smem, s)) return SI_ERR_INTERNAL;
        } else {
            size_t smem = 0;
            int nw = pick_warps(c, p->dev, 4, &smem);
            if (nw < 1) {
                g_last_error = "class does not fit into shared memory";
                return SI_ERR_INTERNAL;
            }
            long long nitems = (long long)job.pc->n * sp.second;
            int grid = (int)std::min<long long>((nitems + nw - 1) / nw, (long long)c->sm_count * 16);
            k_3c2e<<<grid, nw * 32, smem, s>>>(p->dev, c->boys_dev, od_orb, od_aux, job.pc->d_items, job.pc->n,
                                               d_list + sp.first, sp.second, col_begin, od, counter);
        }
        c->launches++;
        SI_CUDA(cudaGetLastError());
    }
    rc = fan.end();
    return rc;
}

// ---- Schwarz integrals (ab|ab) ---------------------------------------------------
int si_schwarz(si_context* c, const si_basis* cb, int normalize, double* D_dev, double* Q_dev, void* stream) {
    if (!c || !cb || !D_dev) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    si_basis* b = const_cast<si_basis*>(cb);
    if (b->lmax > c->lmax || b->lmax > SI_GTO_LMAX) return SI_ERR_LMAX;
    if (4 * b->lmax > c->nrows - 7) {
        g_last_error = "Boys table too small for the Schwarz integrals of this basis (orders up to 4 lmax)";
        return SI_ERR_BOYS;
    }
    SI_CUDA(cudaSetDevice(c->device));
    int rc = build_pair_classes(b);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (c->ev_last_valid) SI_CUDA(cudaStreamWaitEvent(s, c->ev_last, 0));
    for (auto& pc : b->full.classes) {
        SchwarzArgs A;
        A.items = pc.d_items;
        A.n = pc.n;
        A.la = pc.La;
        A.lb = pc.Lb;
        A.ppair = b->full.d_ppair[normalize == SI_NORM_PGTO ? 1 : 0];
        A.D = D_dev;
        A.nbf = b->nbf_sph;
        A.boys = c->boys_dev;
        A.use_lmn = normalize != SI_NORM_NONE;
        const int L = pc.La + pc.Lb, M1 = 2 * L + 1, NH = (L + 1) * (L + 2) * (L + 3) / 6;
        const int ncomp = si_nsph(pc.La) * si_nsph(pc.Lb), e1sz = (pc.La + 1) * (pc.Lb + 1) * (L + 2);
        const int NW = (NH + 31) / 32;
        const size_t fixed = (size_t)ncomp + M1 + (M1 & 1) + 6 * (size_t)e1sz + 2 * (size_t)M1 * M1 * M1 + (size_t)NH + 4;
        const size_t budget = (size_t)c->max_smem / sizeof(double);
        const size_t per_comp = 2 * (size_t)NH + NW;
        if (fixed + per_comp > budget) return SI_ERR_INTERNAL;
        A.chunk = (int)std::min<size_t>((size_t)ncomp, (budget - fixed) / per_comp);
        const size_t smem = (fixed + (size_t)A.chunk * per_comp) * sizeof(double);
        // threads per shell pair: the work items of the two heavy stages are (component pair, Hermite index)
        const int nthr = (size_t)ncomp * NH >= 4096 ? 512 : ((size_t)ncomp * NH >= 512 ? 256 : 128);
        k_schwarz<<<pc.n, nthr, smem, s>>>(A);
        c->launches++;
        SI_CUDA(cudaGetLastError());
    }
    if (Q_dev) {
        const int n2 = b->nshell * b->nshell;
        k_schwarz_q<<<(n2 + 127) / 128, 128, 0, s>>>(D_dev, b->nbf_sph, b->nshell, b->d_foff_sph, b->d_L, Q_dev);
        c->launches++;
        SI_CUDA(cudaGetLastError());
    }
    return mark_done(c, s, SI_OK);
}

// ---- gto on a grid ------------------------------------------------------------
int si_gto(si_context* c, const si_basis* b, int sph, int normalize, long long npoints, const double* points_xyz,
           double* out_dev, void* stream) {
    if (!c || !b || !points_xyz || !out_dev || npoints <= 0) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (b->lmax > std::max(c->lmax, c->lauxmax) || b->lmax > SI_GTO_LMAX) return SI_ERR_LMAX;
    SI_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    double* dpts = nullptr;
    int rc = ctx_scratch(c, (size_t)3 * npoints, s, &dpts);
    if (rc) return rc;
    if (c->ev_last_valid) SI_CUDA(cudaStreamWaitEvent(s, c->ev_last, 0));   // the scratch may be in use by an earlier call
    SI_CUDA(cudaMemcpyAsync(dpts, points_xyz, (size_t)3 * npoints * sizeof(double), cudaMemcpyHostToDevice, s));
    BasisDev bd = basis_dev(b, sph, normalize);
    dim3 grid((unsigned)((npoints + 127) / 128), (unsigned)b->nshell);
    k_gto<<<grid, 128, 0, s>>>(bd, b->nshell, npoints, dpts, sph, normalize != SI_NORM_NONE, out_dev, npoints);
    c->launches++;
    SI_CUDA(cudaGetLastError());
    SI_CUDA(cudaStreamSynchronize(s));   // points_xyz is the caller's host buffer
    return mark_done(c, s, SI_OK);
}

int si_gto_h(si_context* c, const si_basis* b, int sph, int normalize, long long npoints, const double* points_xyz,
             double* out_host) {
    if (!c || !b || !out_host || npoints <= 0) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    SI_CUDA(cudaSetDevice(c->device));
    const size_t n = (size_t)(sph ? b->nbf_sph : b->nbf_cart) * (size_t)npoints;
    double* d = nullptr;
    SI_CUDA(cudaMalloc((void**)&d, n * sizeof(double)));
    int rc = si_gto(c, b, sph, normalize, npoints, points_xyz, d, nullptr);
    if (rc == SI_OK && cudaMemcpy(out_host, d, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = SI_ERR_CUDA;
    cudaFree(d);
    return rc;
}

// ---- host-buffer variants ---------------------------------------------------
static int run_host(si_context* c, size_t n, double* out_host, const std::function<int(double*)>& f) {
    SI_CUDA(cudaSetDevice(c->device));
    double* d = nullptr;
    SI_CUDA(cudaMalloc((void**)&d, n * sizeof(double)));
    SI_CUDA(cudaMemset(d, 0, n * sizeof(double)));
    int rc = f(d);
    if (rc == SI_OK) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(out_host, d, n * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            g_last_error = cudaGetErrorString(e);
            rc = SI_ERR_CUDA;
        }
    }
    cudaFree(d);
    return rc;
}

int si_int1e_h(si_context* c, const si_basis* b, int key, int sph, int normalize, const double* R, double* out) {
    if (!c || !b || !out) return SI_ERR_ARG;
    if (!check_key_2idx(key) && key != SI_MULTI_SPH) return SI_ERR_KEY;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    size_t nbf = sph ? b->nbf_sph : b->nbf_cart;
    return run_host(c, nbf * nbf * (key == SI_MULTI_SPH ? 9 : si_ncomp(key)), out,
                    [&](double* d) { return si_int1e(c, b, key, sph, normalize, R, d, nullptr); });
}
int si_coul_h(si_context* c, const si_basis* b, int sph, int normalize, int ncharge, const double* xyz,
              const double* q, double* out) {
    if (!c || !b || !out) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    size_t nbf = sph ? b->nbf_sph : b->nbf_cart;
    return run_host(c, nbf * nbf, out, [&](double* d) { return si_coul(c, b, sph, normalize, ncharge, xyz, q, d, nullptr); });
}
int si_int2c2e_h(si_context* c, const si_basis* b, int sph, int normalize, double* out) {
    if (!c || !b || !out) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    size_t nbf = sph ? b->nbf_sph : b->nbf_cart;
    return run_host(c, nbf * nbf, out, [&](double* d) { return si_int2c2e(c, b, sph, normalize, d, nullptr); });
}
// Host-buffer 3c2e.  The tensor is produced chunk by chunk into two device staging buffers; while chunk
// k+1 computes, chunk k is copied to the caller's host buffer on a copy stream (pin the host buffer for
// full overlap).  PACKED layout: chunks are ranges of the first shell index, i.e. contiguous row blocks
// of the host tensor -> one contiguous cudaMemcpyAsync each.  FULL layout: chunks are aux-shell ranges
// (column blocks, cudaMemcpy2DAsync).  Staging buffers, streams and events belong to the context (grow-only);
// the row chunking and its pair sets are cached on the (content-cached) orbital shell set.
static double wall_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int si_int3c2e_h(si_context* c, const si_basis* corb, const si_basis* aux, int sb, int se, int normalize, int layout,
                 double* out, size_t out_len) {
    if (!c || !corb || !aux || !out) return SI_ERR_ARG;
    if (sb < 0 || se > aux->nshell || sb >= se) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    si_basis* orb = const_cast<si_basis*>(corb);
    SI_CUDA(cudaSetDevice(c->device));
    const double t_begin = wall_now();
    double t_batcher = 0.0, t_alloc = 0.0;
    int rc = build_pair_classes(orb);
    if (rc) return rc;
    const long long rows = (layout == SI_LAYOUT_FULL) ? (long long)orb->nbf_sph * orb->nbf_sph : orb->packed_rows;
    const int col_begin = aux->foff_sph[sb];
    const int col_end = (se == aux->nshell) ? aux->nbf_sph : aux->foff_sph[se];
    const long long ncol = col_end - col_begin;
    if ((unsigned long long)(rows * ncol) > out_len) {
        g_last_error = "output buffer too small";
        return SI_ERR_ARG;
    }
    // staging size: at least ~12 chunks per call (compute/copy overlap needs several), 64 MiB ... 2 GiB each
    // (SI_STAGE_BYTES overrides; used by the tests)
    long long stage_bytes = std::min<long long>(1ll << 31, std::max<long long>(64ll << 20, rows * ncol * 8 / 12));
    if (const char* e = getenv("SI_STAGE_BYTES")) stage_bytes = std::max(1024ll, atoll(e));
    cudaStream_t s_comp = c->s_comp, s_copy = c->s_copy;
    // order this call after earlier work of the context (the staging buffers and events are shared)
    if (c->ev_last_valid) SI_CUDA(cudaStreamWaitEvent(s_comp, c->ev_last, 0));
    auto ensure_stage = [&](size_t elems) -> int {
        if (elems <= c->stage_elems) return SI_OK;
        const double t0 = wall_now();
        SI_CUDA(cudaDeviceSynchronize());
        for (int i = 0; i < 2; ++i) {
            if (c->stage[i]) SI_CUDA(cudaFree(c->stage[i]));
            c->stage[i] = nullptr;
        }
        c->stage_elems = 0;
        for (int i = 0; i < 2; ++i) SI_CUDA(cudaMalloc((void**)&c->stage[i], elems * sizeof(double)));
        c->stage_elems = elems;
        t_alloc += wall_now() - t0;
        return SI_OK;
    };
    auto timing_events = [&](size_t n) -> int {
        while (c->ev_time.size() < n) {
            cudaEvent_t e;
            SI_CUDA(cudaEventCreate(&e));
            c->ev_time.push_back(e);
        }
        return SI_OK;
    };
    size_t nchunks = 0;
    double t_enqueue0 = 0.0;
    if (layout == SI_LAYOUT_PACKED) {
        // (re)build the cached row chunking: consecutive first-shell ranges of <= stage_bytes each
        if (orb->row_chunks.empty() || orb->row_chunk_ncol != (int)ncol || orb->row_chunk_bytes != stage_bytes) {
            const double t0 = wall_now();
            for (auto* ps : orb->row_chunks) {
                free_pair_set(ps);
                delete ps;
            }
            orb->row_chunks.clear();
            const long long max_rows = std::max<long long>(1, stage_bytes / (ncol * 8));
            int i0 = 0;
            long long acc = 0;
            for (int i = 0; i < orb->nshell; ++i) {
                long long r = 0;
                for (int j = 0; j <= i; ++j) r += (long long)si_nsph(orb->L[i]) * si_nsph(orb->L[j]);
                if (acc > 0 && acc + r > max_rows) {
                    auto* ps = new si_basis::PairSet;
                    rc = build_pair_set(orb, i0, i, ps);
                    orb->row_chunks.push_back(ps);
                    if (rc) return rc;
                    i0 = i;
                    acc = 0;
                }
                acc += r;
            }
            auto* ps = new si_basis::PairSet;
            rc = build_pair_set(orb, i0, orb->nshell, ps);
            orb->row_chunks.push_back(ps);
            if (rc) return rc;
            orb->row_chunk_ncol = (int)ncol;
            orb->row_chunk_bytes = stage_bytes;
            t_batcher += wall_now() - t0;
        }
        size_t need = 0;
        for (auto* ps : orb->row_chunks) need = std::max(need, (size_t)(ps->rows * ncol));
        rc = ensure_stage(need);
        if (rc) return rc;
        nchunks = orb->row_chunks.size();
        rc = timing_events(2 * nchunks);
        if (rc) return rc;
        t_enqueue0 = wall_now();
        long long row0 = 0;
        for (size_t k = 0; k < nchunks && rc == SI_OK; ++k) {
            const int b = (int)(k & 1);
            auto* ps = orb->row_chunks[k];
            if (k >= 2) SI_CUDA(cudaStreamWaitEvent(s_comp, c->ev_free[b], 0));
            SI_CUDA(cudaEventRecord(c->ev_time[2 * k], s_comp));
            rc = int3c2e_impl(c, orb, aux, sb, se, normalize, layout, ps, c->stage[b], (size_t)(ps->rows * ncol), (void*)s_comp);
            if (rc != SI_OK) break;
            SI_CUDA(cudaEventRecord(c->ev_time[2 * k + 1], s_comp));
            SI_CUDA(cudaEventRecord(c->ev_done[b], s_comp));
            SI_CUDA(cudaStreamWaitEvent(s_copy, c->ev_done[b], 0));
            SI_CUDA(cudaMemcpyAsync(out + row0 * ncol, c->stage[b], (size_t)(ps->rows * ncol) * sizeof(double),
                                    cudaMemcpyDeviceToHost, s_copy));
            SI_CUDA(cudaEventRecord(c->ev_free[b], s_copy));
            row0 += ps->rows;
        }
    } else {
        const long long max_cols = std::max<long long>(1, stage_bytes / (rows * 8));
        const long long want = std::max<long long>(1, std::min<long long>(max_cols, (ncol + 3) / 4));
        std::vector<int> cuts;
        cuts.push_back(sb);
        long long acc = 0;
        for (int s = sb; s < se; ++s) {
            acc += si_nsph(aux->L[s]);
            if (acc >= want && s + 1 < se) {
                cuts.push_back(s + 1);
                acc = 0;
            }
        }
        cuts.push_back(se);
        long long maxc = 0;
        for (size_t k = 0; k + 1 < cuts.size(); ++k) {
            int e = cuts[k + 1];
            maxc = std::max<long long>(maxc, ((e == aux->nshell) ? aux->nbf_sph : aux->foff_sph[e]) - aux->foff_sph[cuts[k]]);
        }
        rc = ensure_stage((size_t)rows * maxc);
        if (rc) return rc;
        nchunks = cuts.size() - 1;
        rc = timing_events(2 * nchunks);
        if (rc) return rc;
        t_enqueue0 = wall_now();
        for (size_t k = 0; k + 1 < cuts.size() && rc == SI_OK; ++k) {
            const int b = (int)(k & 1);
            const int cs = cuts[k], ce = cuts[k + 1];
            const long long c0 = aux->foff_sph[cs] - col_begin;
            const long long cc = ((ce == aux->nshell) ? aux->nbf_sph : aux->foff_sph[ce]) - aux->foff_sph[cs];
            if (k >= 2) SI_CUDA(cudaStreamWaitEvent(s_comp, c->ev_free[b], 0));
            SI_CUDA(cudaEventRecord(c->ev_time[2 * k], s_comp));
            rc = int3c2e_impl(c, orb, aux, cs, ce, normalize, layout, nullptr, c->stage[b], (size_t)(rows * cc), (void*)s_comp);
            if (rc != SI_OK) break;
            SI_CUDA(cudaEventRecord(c->ev_time[2 * k + 1], s_comp));
            SI_CUDA(cudaEventRecord(c->ev_done[b], s_comp));
            SI_CUDA(cudaStreamWaitEvent(s_copy, c->ev_done[b], 0));
            SI_CUDA(cudaMemcpy2DAsync(out + c0, (size_t)ncol * sizeof(double), c->stage[b], (size_t)cc * sizeof(double),
                                      (size_t)cc * sizeof(double), (size_t)rows, cudaMemcpyDeviceToHost, s_copy));
            SI_CUDA(cudaEventRecord(c->ev_free[b], s_copy));
        }
    }
    const double t_enqueued = wall_now();
    cudaError_t e1 = cudaStreamSynchronize(s_comp);
    const double t_comp_done = wall_now();
    cudaError_t e2 = cudaStreamSynchronize(s_copy);
    const double t_end = wall_now();
    // later calls on the context (any stream) are ordered after this one
    if (cudaEventRecord(c->ev_last, s_copy) == cudaSuccess) c->ev_last_valid = true;
    if (rc == SI_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        g_last_error = cudaGetErrorString(e1 != cudaSuccess ? e1 : e2);
        rc = SI_ERR_CUDA;
    }
    double t_kernels = 0.0;
    if (rc == SI_OK)
        for (size_t k = 0; k < nchunks; ++k) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, c->ev_time[2 * k], c->ev_time[2 * k + 1]) == cudaSuccess) t_kernels += ms * 1e-3;
        }
    c->breakdown[0] = t_end - t_begin;           // total
    c->breakdown[1] = t_batcher;                 // host shell-batcher (row-chunk pair sets; 0 when cached)
    c->breakdown[2] = t_alloc;                   // staging (re)allocation (0 when large enough)
    c->breakdown[3] = t_enqueued - t_enqueue0;   // host time to enqueue all chunks (launches + copies)
    c->breakdown[4] = t_kernels;                 // device time of the compute chunks (sum of per-chunk event times)
    c->breakdown[5] = t_end - t_comp_done;       // wait for the copies after the last kernel finished
    c->breakdown[6] = (double)nchunks;
    c->breakdown[7] = (double)stage_bytes;
    return rc;
}

int si_set_screening(si_context* c, double eps) {
    if (!c || !(eps >= 0.0)) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    c->screen_eps = eps;
    return SI_OK;
}

int si_host_breakdown(const si_context* c, double* out8) {
    if (!c || !out8) return SI_ERR_ARG;
    memcpy(out8, c->breakdown, sizeof(c->breakdown));
    return SI_OK;
}

// ---- FP64 peak microbenchmark (roofline denominator; not in MEASURED_PEAKS.json) ------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c);
            a1 = fma(a1, m, c);
            a2 = fma(a2, m, c);
            a3 = fma(a3, m, c);
            a4 = fma(a4, m, c);
            a5 = fma(a5, m, c);
            a6 = fma(a6, m, c);
            a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int si_fp64_peak(si_context* c, double seconds, double* tflops) {
    if (!c || !tflops) return SI_ERR_ARG;
    SI_CUDA(cudaSetDevice(c->device));
    const int blocks = c->sm_count * 8, threads = 256;
    double* d = nullptr;
    SI_CUDA(cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    SI_CUDA(cudaEventCreate(&e0));
    SI_CUDA(cudaEventCreate(&e1));
    int iters = 2000;
    double best = 0.0, spent = 0.0;
    for (int rep = 0; rep < 50 && spent < seconds; ++rep) {
        SI_CUDA(cudaEventRecord(e0));
        k_dfma_peak<<<blocks, threads>>>(d, iters);
        SI_CUDA(cudaEventRecord(e1));
        SI_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        SI_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        spent += ms * 1e-3;
        if (ms < 20.0f) iters *= 2;
    }
    c->launches += 1;
    cudaFree(d);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return SI_OK;
}

// ---- single block with the generated-function signature ---------------------
int si_eval_block(si_context* c, int key, int La, int Lb, int Lc, int sph, int normalize, int Ka, const double* ax,
                  const double* da, const double* A, int Kb, const double* bx, const double* db, const double* B,
                  int Kc, const double* cx, const double* dc, const double* C, const double* R, double* result) {
    if (!c || !ax || !da || !A || !result || Ka <= 0) return SI_ERR_ARG;
    if (key < 0 || key > SI_MULTI_SPH) return SI_ERR_KEY;
    if (key != SI_GTO && (!bx || !db || !B || Kb <= 0)) return SI_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (key == SI_GTO) {
        // one-index function f(ax, da, A, R, result): a one-shell set evaluated at the single point R
        if (!R) return SI_ERR_ARG;
        if (La < 0 || La > c->lmax) return SI_ERR_LMAX;
        si_basis* ob1 = nullptr;
        int rc1 = si_basis_create(c, 1, &La, &Ka, A, ax, da, &ob1);
        if (rc1) return rc1;
        rc1 = si_gto_h(c, ob1, sph, normalize, 1, R, result);
        si_basis_destroy(ob1);
        return rc1;
    }
    if (key == SI_MULTI_SPH) {
        if (La < 0 || Lb < 0 || La > c->lmax || Lb > c->lmax) return SI_ERR_LMAX;
        SI_CUDA(cudaSetDevice(c->device));
        int Ls2[2] = {La, Lb}, nps2[2] = {Ka, Kb};
        std::vector<double> cen2(6), ex2, co2;
        for (int d = 0; d < 3; ++d) {
            cen2[d] = A[d];
            cen2[3 + d] = B[d];
        }
        ex2.insert(ex2.end(), ax, ax + Ka);
        ex2.insert(ex2.end(), bx, bx + Kb);
        co2.insert(co2.end(), da, da + Ka);
        co2.insert(co2.end(), db, db + Kb);
        si_basis* ob2 = nullptr;
        int rc2 = si_basis_create(c, 2, Ls2, nps2, cen2.data(), ex2.data(), co2.data(), &ob2);
        if (rc2) return rc2;
        const int na = si_nsph(La), nb = si_nsph(Lb), nbf = na + nb;
        double* d9 = nullptr;
        rc2 = cudaMalloc((void**)&d9, (size_t)9 * nbf * nbf * sizeof(double)) == cudaSuccess ? SI_OK : SI_ERR_CUDA;
        if (rc2 == SI_OK) rc2 = si_int1e(c, ob2, SI_MULTI_SPH, sph, normalize, nullptr, d9, nullptr);
        std::vector<double> h((size_t)9 * nbf * nbf);
        if (rc2 == SI_OK && cudaMemcpy(h.data(), d9, h.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc2 = SI_ERR_CUDA;
        cudaFree(d9);
        si_basis_destroy(ob2);
        if (rc2) return rc2;
        for (int q = 0; q < 9; ++q)
            for (int ia = 0; ia < na; ++ia)
                for (int jb = 0; jb < nb; ++jb) result[((size_t)q * na + ia) * nb + jb] = h[((size_t)q * nbf + ia) * nbf + na + jb];
        return SI_OK;
    }
    if (key == SI_PREFACTOR) {
        if (!R) return SI_ERR_ARG;
        if (La < 0 || Lb < 0 || La > c->lmax || Lb > c->lmax) return SI_ERR_LMAX;
        SI_CUDA(cudaSetDevice(c->device));
        // the reference's functions are primitive ones (scalar da, db); several primitives are summed
        double sa = 0.0, sb = 0.0;
        for (int i = 0; i < Ka; ++i)
            sa += da[i] * (normalize == SI_NORM_PGTO ? std::sqrt(std::pow(2.0, 2 * La + 1.5) * std::pow(ax[i], La + 1.5) / std::pow(SI_PI, 1.5)) : 1.0);
        for (int i = 0; i < Kb; ++i)
            sb += db[i] * (normalize == SI_NORM_PGTO ? std::sqrt(std::pow(2.0, 2 * Lb + 1.5) * std::pow(bx[i], Lb + 1.5) / std::pow(SI_PI, 1.5)) : 1.0);
        const int na = sph ? si_nsph(La) : si_ncart(La), nb = sph ? si_nsph(Lb) : si_ncart(Lb);
        double* d = nullptr;
        int rc1 = ctx_scratch(c, (size_t)na * nb, nullptr, &d);
        if (rc1) return rc1;
        if (c->ev_last_valid) SI_CUDA(cudaEventSynchronize(c->ev_last));
        k_prefactor<<<1, 512>>>(La, Lb, sph, normalize != SI_NORM_NONE, sa * sb, R[0], R[1], R[2], d);
        c->launches++;
        SI_CUDA(cudaGetLastError());
        SI_CUDA(cudaMemcpy(result, d, (size_t)na * nb * sizeof(double), cudaMemcpyDeviceToHost));
        return SI_OK;
    }
    const bool three = (key == SI_3C2E);
    if (three && (!cx || !dc || !C || Kc <= 0)) return SI_ERR_ARG;
    if (three) sph = 1;
    const int lim = (key == SI_2C2E) ? c->lauxmax : c->lmax;
    if (La < 0 || Lb < 0 || La > lim || Lb > lim || (three && (Lc < 0 || Lc > c->lauxmax))) return SI_ERR_LMAX;
    SI_CUDA(cudaSetDevice(c->device));
    const bool swap = La < Lb;
    // two-shell orbital set (a = higher L first) and one-shell auxiliary set, packed into ONE host image that is
    // uploaded with a single copy into a context-owned device scratch: one upload, one launch, one read-back per block
    const int Ls[2] = {swap ? Lb : La, swap ? La : Lb};
    const int nps[2] = {swap ? Kb : Ka, swap ? Ka : Kb};
    const double* A_ = swap ? B : A;
    const double* B_ = swap ? A : B;
    const double* ex_[2] = {swap ? bx : ax, swap ? ax : bx};
    const double* co_[2] = {swap ? db : da, swap ? da : db};
    DevProgram* p;
    int rc = get_program(c, key, Ls[0], Ls[1], three ? Lc : 0, sph, normalize, &p);
    if (rc) return rc;
    const SiProg& P = p->dev;
    const int n = P.out_n0 * P.out_n1 * P.nb_total;
    size_t smem = 0;
    const int nw = pick_warps(c, P, 1, &smem);
    if (nw < 1) return SI_ERR_INTERNAL;
    const int npo = nps[0] + nps[1], npc = three ? Kc : 0, npair = nps[0] * nps[1];
    auto pgto = [&](double a, int l) {
        return normalize == SI_NORM_PGTO ? std::sqrt(std::pow(2.0, 2 * l + 1.5) * std::pow(a, l + 1.5) / std::pow(SI_PI, 1.5)) : 1.0;
    };
    // image layout (doubles): [ints: 16 slots][cen_o 6][ex_o][co_o][cen_a 3][ex_a][co_a][oex_a][pp 9*npair][R 3][q 1]
    //                         [pair item 8][aux item 6][list + counter 1][out n]
    size_t o = 0;
    auto take = [&](size_t nd) { size_t r = o; o += nd; return r; };
    const size_t o_int = take(8), o_ceno = take(6), o_exo = take(npo), o_coo = take(npo), o_cena = take(3), o_exa = take(npc),
                 o_coa = take(npc), o_oexa = take(npc), o_pp = take((size_t)SI_PP_STRIDE * npair), o_R = take(3), o_q = take(1),
                 o_item = take(sizeof(SiPairItem) / 8), o_aitem = take(sizeof(SiAuxItem) / 8), o_cnt = take(1), o_out = take(n);
    std::vector<double> img(o, 0.0);
    int* hi = reinterpret_cast<int*>(&img[o_int]);   // L[2] nprim[2] poff[2] | aux: L nprim poff
    hi[0] = Ls[0]; hi[1] = Ls[1]; hi[2] = nps[0]; hi[3] = nps[1]; hi[4] = 0; hi[5] = nps[0];
    hi[8] = Lc; hi[9] = npc; hi[10] = 0;
    for (int d = 0; d < 3; ++d) {
        img[o_ceno + d] = A_[d];
        img[o_ceno + 3 + d] = B_[d];
        if (three) img[o_cena + d] = C[d];
        img[o_R + d] = R ? R[d] : 0.0;
    }
    img[o_q] = 1.0;
    for (int sh = 0, k = 0; sh < 2; ++sh)
        for (int i = 0; i < nps[sh]; ++i, ++k) {
            img[o_exo + k] = ex_[sh][i];
            img[o_coo + k] = co_[sh][i] * pgto(ex_[sh][i], Ls[sh]);
        }
    for (int i = 0; i < npc; ++i) {
        img[o_exa + i] = cx[i];
        img[o_coa + i] = dc[i] * pgto(cx[i], Lc);
        img[o_oexa + i] = 1.0 / cx[i];
    }
    for (int i = 0; i < nps[0]; ++i)   // primitive-pair data of the single (a = shell 0, b = shell 1) item
        for (int j = 0; j < nps[1]; ++j) {
            double* q = &img[o_pp + ((size_t)i * nps[1] + j) * SI_PP_STRIDE];
            q[9] = ex_[0][i];
            q[10] = ex_[1][j];
            const double ax_ = ex_[0][i], bx_ = ex_[1][j];
            const double p_ = ax_ + bx_, oop = 1.0 / p_, mu = ax_ * bx_ * oop;
            double r2 = 0.0;
            q[0] = p_;
            q[1] = oop;
            for (int d = 0; d < 3; ++d) {
                q[2 + d] = (ax_ * A_[d] + bx_ * B_[d]) * oop;
                q[5 + d] = q[2 + d] - A_[d];
                r2 += (A_[d] - B_[d]) * (A_[d] - B_[d]);
            }
            q[8] = std::exp(-mu * r2) * img[o_coo + i] * img[o_coo + nps[0] + j];
        }
    SiPairItem item;
    memset(&item, 0, sizeof(item));
    item.sa = 0;
    item.sb = 1;
    item.Ka = nps[0];
    item.Kb = nps[1];
    for (int d = 0; d < 3; ++d) item.AB[d] = A_[d] - B_[d];
    memcpy(&img[o_item], &item, sizeof(item));
    SiAuxItem aitem;
    memset(&aitem, 0, sizeof(aitem));
    aitem.Kc = npc;
    if (three)
        for (int d = 0; d < 3; ++d) aitem.C[d] = C[d];
    memcpy(&img[o_aitem], &aitem, sizeof(aitem));
    if (c->ev_last_valid) SI_CUDA(cudaEventSynchronize(c->ev_last));   // the scratch may still be read by an earlier call
    double* dimg = nullptr;
    rc = ctx_scratch(c, img.size(), nullptr, &dimg);
    if (rc) return rc;
    SI_CUDA(cudaMemcpy(dimg, img.data(), (img.size() - n) * sizeof(double), cudaMemcpyHostToDevice));
    SI_CUDA(cudaDeviceSynchronize());   // pageable upload landed before a kernel on any stream reads it
    const int* di = reinterpret_cast<const int*>(dimg + o_int);
    BasisDev bd, ad;
    bd.L = di;
    bd.nprim = di + 2;
    bd.poff = di + 4;
    bd.foff = nullptr;
    bd.cen = dimg + o_ceno;
    bd.exps = dimg + o_exo;
    bd.coef = dimg + o_coo;
    ad.L = di + 8;
    ad.nprim = di + 9;
    ad.poff = di + 10;
    ad.foff = nullptr;
    ad.cen = dimg + o_cena;
    ad.exps = dimg + o_exa;
    ad.coef = dimg + o_coa;
    const SiPairItem* d_item = reinterpret_cast<const SiPairItem*>(dimg + o_item);
    const SiAuxItem* d_aitem = reinterpret_cast<const SiAuxItem*>(dimg + o_aitem);
    int* d_counter = reinterpret_cast<int*>(dimg + o_cnt);
    OutDesc od;
    od.out = dimg + o_out;
    od.ld = 0;
    od.comp_stride = 0;
    od.nbf = 0;
    od.mode = STORE_BLOCK;
    cudaError_t e = cudaSuccess;
    if (three) {
        const int* g3 = g3_find(Ls[0], Ls[1], Lc);
        if (g3 && normalize != SI_NORM_NONE && !g3_disabled()) {
            G3Args ga;
            g3_fill_args(ga, c, bd, ad, dimg + o_oexa, d_item, dimg + o_pp, 1, reinterpret_cast<const int*>(dimg + o_cnt) + 1,
                         d_aitem, 1, 0, od, d_counter, g3[3]);
            if (g3_launch(g3[1], g3[0], ga, 1, 32, (size_t)g3[4] * g3[3] * sizeof(double), c->max_smem, 0))
                e = cudaErrorInvalidDeviceFunction;
        } else {
            k_3c2e<<<1, 32, smem>>>(P, c->boys_dev, bd, ad, d_item, 1, nullptr, 1, 0, od, d_counter);
        }
    } else if (key == SI_COUL) {
        k_coul<<<1, 32, smem>>>(P, c->boys_dev, bd, d_item, 1, dimg + o_R, dimg + o_q, 1, od);
    } else {
        k_two_index<<<1, 32, smem>>>(P, c->boys_dev, bd, d_item, 1, key == SI_2C2E ? 1 : 0, R ? R[0] : 0.0, R ? R[1] : 0.0,
                                     R ? R[2] : 0.0, od, d_counter);
    }
    c->launches++;
    if (e == cudaSuccess) e = cudaGetLastError();
    std::vector<double> tmp(n);
    if (e == cudaSuccess) e = cudaMemcpy(tmp.data(), dimg + o_out, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        g_last_error = cudaGetErrorString(e);
        return SI_ERR_CUDA;
    }
    if (!swap) {
        memcpy(result, tmp.data(), n * sizeof(double));
    } else {
        // computed (comp, b, a[, c]) -> (comp, a, b[, c])   (resort_ba_ab / resort_bac_abc)
        const int nb_ = P.out_n0, na_ = P.out_n1;  // computed block is (n0 = size of Lb-shell..)
        const int ncomp = P.out_batch_major ? P.nb_total : 1;
        const int nc = P.out_batch_major ? 1 : P.nb_total;
        for (int q = 0; q < ncomp; ++q)
            for (int ib = 0; ib < nb_; ++ib)
                for (int ia = 0; ia < na_; ++ia)
                    for (int m = 0; m < nc; ++m)
                        result[((size_t)(q * na_ + ia) * nb_ + ib) * nc + m] =
                            tmp[((size_t)(q * nb_ + ib) * na_ + ia) * nc + m];
    }
    return SI_OK;
}
