// Fused one-electron kernel: ovlp, kin, dpm, qpm (and dqpm = the diagonal of qpm) of a whole shell set in ONE
// launch.  Included by si_lib.cu.
//
// Maths (reference): 1-D Obara-Saika multipole recurrence  defs/multipole.py:17-32
//     I(i+1,j,e) = PA I(i,j,e) + 1/(2p) [ i I(i-1,j,e) + j I(i,j-1,e) + e I(i,j,e-1) ]     (PB / PR for j / e)
// with base sqrt(pi/p) exp(-mu AB_d^2) (defs/twocenter1d.py:14-51), 3-D value Ix Iy Iz (multipole.py:35-39),
// component order canonical_order(Le) outermost (multipole.py:42-47: dpm x,y,z; qpm xx,xy,xz,yy,yz,zz);
// kinetic energy  defs/kinetic.py:15-50:
//     T00 = (a - 2a^2 (PA^2 + 1/(2p))) S00
//     T(i+1,j) = PA T(i,j) + 1/(2p)[i T(i-1,j) + j T(i,j-1)] + (b/p)[2a S(i+1,j) - i S(i-1,j)]   (ket: a<->b, PB)
//     T = Tx Sy Sz + Sx Ty Sz + Sx Sy Tz
// lmn factors and C2S on both indices as main.py:122-189,250-272.
//
// Execution model: EPT lanes of a warp per shell pair.  Per primitive pair 6 (LB + 1) lanes build the 1-D tables
// (lane = (multipole | kinetic, direction, ket index j), rows in registers, neighbours by warp shuffle) and publish
// them in shared memory; then the lanes take Cartesian component pairs (a, b), form
// the products for the requested operators and accumulate the contraction; per shell pair the two half
// transforms C2S(a), C2S(b) run through shared memory and the block and its mirror image are stored.
// The three-dimensional exp(-mu AB^2) da db (pi/p)^(3/2) is put into the x table; the y and z tables start at 1.
#pragma once
#include <type_traits>

#define ONEE_NCOMP 11   // 0 ovlp, 1 kin, 2-4 dpm x y z, 5-10 qpm xx xy xz yy yz zz
#define ONEE_MAXL 4
#define ONEE_KCH 12     // primitive pairs staged in shared memory at a time (SI_PP_STRIDE = 12 doubles each)

#include "generated/onee_tables.inc"

struct OneEClassSeg {
    const SiPairItem* items;
    int n;           // shell pairs of the class
    int la, lb;
    int group_end;   // prefix sum of warp-groups (exclusive end) over the classes of this launch
};

struct OneEArgs {
    OneEClassSeg seg[15];
    int nseg;
    int ngroups;
    const double* ppair;      // 9 doubles per primitive pair (see build_pair_set)
    const double* exps;       // primitive exponents of the shell set
    const int* poff;          // first primitive of every shell
    double R[3];              // multipole origin
    int origin_p;             // 1: the origin is the centre P of every primitive pair (multi_sph), R unused
    double* out[ONEE_NCOMP];  // (nbf, nbf) matrices, null = operator not requested
    long long nbf;
    int* counter;
};

template <int L>
struct OneECart {
    static constexpr int N = (L + 1) * (L + 2) / 2;
};

// lanes per pair / smem doubles per pair of a class
__host__ __device__ constexpr int onee_ept(int la, int lb) {
    const int nab = ((la + 1) * (la + 2) / 2) * ((lb + 1) * (lb + 2) / 2);
    return nab <= 8 ? 8 : (nab <= 16 ? 16 : 32);
}
__host__ __device__ constexpr int onee_cp(int la, int lb) {   // components per pass
    const int nab = ((la + 1) * (la + 2) / 2) * ((lb + 1) * (lb + 2) / 2);
    return nab <= 60 ? ONEE_NCOMP : (nab <= 100 ? 6 : 3);
}
__host__ __device__ constexpr int onee_wsz(int la, int lb) {
    const int na = (la + 1) * (la + 2) / 2, nb = (lb + 1) * (lb + 2) / 2;
    const int ts = (la + 1) * (lb + 1);
    int w = 15 * ts + onee_cp(la, lb) * (na * nb + (2 * la + 1) * nb) + ONEE_KCH * 12;
    return (w + 1) & ~1;
}

// Lane-parallel table build: lane (type, d, j) of a pair's lane group owns column j of direction d -- type 0 the
// multipole tables I(i, j, e), e = 0..2, type 1 the overlap/kinetic pair S(i, j), T(i, j) -- and walks the rows
// i = 0..LA with everything in registers; the only inter-lane dependence, the (i-1, j-1) neighbour of the bra
// reduction (and (0, j-1), (0, j-2) in the first row), arrives by warp shuffle.  The dependent chain per primitive
// pair is ~(LA + LB) steps of a shuffle and a few FMAs instead of (LA+1)(LB+1) x 3 entries read back through shared
// memory by a single lane.  Needs 6 (LB + 1) <= EPT lanes per pair (30 for LB = 4).
template <int LA, int LB>
__device__ __forceinline__ void onee_tables_parallel(double* TB, const int q, const double* pp, const double* AB,
                                                     const double* Rorg, const int origin_p, const bool live) {
    constexpr int NJ = LB + 1, TS = (LA + 1) * NJ, DS = 5 * TS;
    const int type = q / (3 * NJ), r = q - type * 3 * NJ, d = min(r / NJ, 2), j = r - (r / NJ) * NJ;
    const bool act = q < 6 * NJ;
    const double oop = pp[1], oo2p = 0.5 * oop;
    const double PA = pp[5 + d], PB = PA + AB[d], PR = origin_p ? 0.0 : pp[2 + d] - Rorg[d];
    const double a = pp[9], b = pp[10], bop = b * oop, aop = a * oop;
    const double base = d == 0 ? (live ? pp[8] : 0.0) * (SI_PI * oop) * sqrt(SI_PI * oop) : 1.0;
    const double fj = (double)j;
    // channels: multipole lanes e = 0, 1, 2; kinetic lanes ch0 = S, ch1 = T
    double c0, c1, c2;          // row i - 1 ... becomes row i
    double p0 = 0.0, p1 = 0.0, p2 = 0.0;   // row i - 2
    // ---- row i = 0: column 0, then the ket recursion along j (lane j takes its turn at step jj == j)
    if (type == 0) {
        c0 = base;
        c1 = PR * c0;
        c2 = fma(oo2p, c0, PR * c1);
    } else {
        c0 = base;
        c1 = (a - 2.0 * a * a * fma(PA, PA, oo2p)) * base;
        c2 = 0.0;
    }
#pragma unroll
    for (int jj = 1; jj <= LB; ++jj) {
        const double m0 = __shfl_up_sync(0xffffffffu, c0, 1), m1 = __shfl_up_sync(0xffffffffu, c1, 1),
                     m2 = __shfl_up_sync(0xffffffffu, c2, 1);
        const double n0 = __shfl_up_sync(0xffffffffu, c0, 2), n1 = __shfl_up_sync(0xffffffffu, c1, 2),
                     n2 = __shfl_up_sync(0xffffffffu, c2, 2);
        if (j == jj) {
            const double f = oo2p * (double)(jj - 1);
            if (type == 0) {   // I(0, j, e) = PB I(0, j-1, e) + 1/(2p) [(j-1) I(0, j-2, e) + e I(0, j-1, e-1)]
                c0 = fma(f, jj > 1 ? n0 : 0.0, PB * m0);
                c1 = fma(oo2p, m0, fma(f, jj > 1 ? n1 : 0.0, PB * m1));
                c2 = fma(2.0 * oo2p, m1, fma(f, jj > 1 ? n2 : 0.0, PB * m2));
            } else {           // S(0, j), T(0, j)
                const double sm2 = jj > 1 ? n0 : 0.0, tm2 = jj > 1 ? n1 : 0.0;
                c0 = fma(f, sm2, PB * m0);
                c1 = fma(aop, fma(-(double)(jj - 1), sm2, 2.0 * b * c0), fma(f, tm2, PB * m1));
            }
        }
    }
    double* dstM = TB + d * DS + j;            // multipole: + e * TS + i * NJ
    double* dstT = TB + d * DS + 3 * TS + j;   // kinetic T: + i * NJ
    if (act) {
        if (type == 0) {
            dstM[0] = c0;
            dstM[TS] = c1;
            dstM[2 * TS] = c2;
        } else {
            dstT[0] = c1;
        }
    }
    // ---- rows i = 1..LA: bra reduction; the (i-1, j-1) neighbour comes from the lane below
#pragma unroll
    for (int i = 1; i <= LA; ++i) {
        const double m0 = __shfl_up_sync(0xffffffffu, c0, 1), m1 = __shfl_up_sync(0xffffffffu, c1, 1),
                     m2 = __shfl_up_sync(0xffffffffu, c2, 1);
        const double fi = (double)(i - 1);
        double n0, n1, n2;
        if (type == 0) {   // I(i,j,e) = PA I(i-1,j,e) + 1/(2p) [(i-1) I(i-2,j,e) + j I(i-1,j-1,e) + e I(i-1,j,e-1)]
            n0 = fma(oo2p, fma(fi, p0, fj * m0), PA * c0);
            n1 = fma(oo2p, fma(fi, p1, fma(fj, m1, c0)), PA * c1);
            n2 = fma(oo2p, fma(fi, p2, fma(fj, m2, 2.0 * c1)), PA * c2);
        } else {           // S(i,j), T(i,j) = PA T(i-1,j) + 1/(2p)[..] + (b/p) [2a S(i,j) - (i-1) S(i-2,j)]
            n0 = fma(oo2p, fma(fi, p0, fj * m0), PA * c0);
            n1 = fma(bop, fma(-fi, p0, 2.0 * a * n0), fma(oo2p, fma(fi, p1, fj * m1), PA * c1));
            n2 = 0.0;
        }
        p0 = c0; p1 = c1; p2 = c2;
        c0 = n0; c1 = n1; c2 = n2;
        if (act) {
            if (type == 0) {
                dstM[i * NJ] = c0;
                dstM[TS + i * NJ] = c1;
                dstM[2 * TS + i * NJ] = c2;
            } else {
                dstT[i * NJ] = c1;
            }
        }
    }
}

template <int LA, int LB>
__device__ void onee_body(const OneEArgs& A, const OneEClassSeg& sg, int grp_in_class, int lane, double* wbase,
                          unsigned mask) {
    constexpr int NA = OneECart<LA>::N, NB = OneECart<LB>::N, NAB = NA * NB;
    constexpr int SA = 2 * LA + 1, SB = 2 * LB + 1;
    constexpr int TS = (LA + 1) * (LB + 1), DS = 5 * TS;
    constexpr int EPT = onee_ept(LA, LB), TPW = 32 / EPT, WSZ = onee_wsz(LA, LB), CP = onee_cp(LA, LB);
    constexpr int NIT = (NAB + EPT - 1) / EPT;
    const int q = lane % EPT, tl = lane / EPT;
    double* W = wbase + tl * WSZ;
    double* TB = W;                       // [d][kind][i][j]: kind 0..2 multipole order e, 3 kinetic, 4 kinetic lane's overlap
    double* CART = W + 15 * TS;           // [c][a][b] of the current pass; later the spherical block [c][ia][jb]
    double* HALF = CART + CP * NAB;       // [c][ia][b]
    // table offsets (x | y << 8 | z << 16) of this lane's Cartesian component pairs
    int toff[NIT];
#pragma unroll
    for (int t = 0; t < NIT; ++t) {
        const int ab = min(q + EPT * t, NAB - 1);
        const int ai = ab / NB, bi = ab - ai * NB;
        const int pa = onee_cart[LA][ai], pb = onee_cart[LB][bi];
        toff[t] = ((pa & 15) * (LB + 1) + (pb & 15)) | ((((pa >> 4) & 15) * (LB + 1) + ((pb >> 4) & 15)) << 8) |
                  (((pa >> 8) * (LB + 1) + (pb >> 8)) << 16);
    }
    double* PP = HALF + CP * SA * NB;     // staged primitive-pair data: p.., exponent a, exponent b (12 doubles each)
    int ip_ = grp_in_class * TPW + tl;
    const bool valid = ip_ < sg.n;
    if (!valid) ip_ = sg.n - 1;
    const SiPairItem it = sg.items[ip_];
    const int K = it.Ka * it.Kb;
    int Kmax = K;
#pragma unroll
    for (int o = 16; o >= EPT; o >>= 1) Kmax = max(Kmax, __shfl_xor_sync(0xffffffffu, Kmax, o));
    const int emax = (mask & 0x7E0u) ? 2 : ((mask & 0x1Cu) ? 1 : 0);
    const bool want_kin = (mask & 2u) != 0;
    // Stage the data of primitive pairs [k0, k0 + ONEE_KCH) with coalesced loads: one global round trip per chunk
    // instead of one (dependent) per primitive pair.
    auto stage = [&](int k0) {
        const int nk = max(0, min(ONEE_KCH, K - k0));   // (0 for pairs with fewer primitives: they keep their last chunk)
        const double* src = A.ppair + (size_t)(it.ppoff + k0) * SI_PP_STRIDE;   // (the exponents a, b ride along)
        for (int i = q; i < nk * SI_PP_STRIDE; i += EPT) PP[i] = src[i];
        __syncwarp();
    };
    auto build_tables = [&](int kk, bool live, int e_hi, bool kin) {
        // (pi/p)^(3/2) exp(-mu AB^2) da db goes into the x tables, 1 into y and z; pairs with fewer primitives add zeros
        (void)e_hi;
        (void)kin;
        onee_tables_parallel<LA, LB>(TB, q, PP + kk * 12, it.AB, A.R, A.origin_p, live);
        __syncwarp();
    };
    // with one primitive pair the tables are built once and shared by all passes
    const bool once = (Kmax == 1) && CP < ONEE_NCOMP;
    if (once) {
        stage(0);
        build_tables(0, true, emax, want_kin);
    }
    auto run_pass = [&](auto c0_tag) {
        constexpr int c0 = decltype(c0_tag)::value;
        const unsigned pmask = (mask >> c0) & ((1u << CP) - 1u);
        if (!pmask) return;
        const bool pass_kin = want_kin && c0 <= 1 && 1 < c0 + CP;
        const int pass_emax = (c0 + CP > 5) ? emax : min(emax, (c0 + CP > 2) ? 1 : 0);
        double acc[NIT][CP];   // contraction accumulators of this lane's Cartesian pairs (registers)
#pragma unroll
        for (int t = 0; t < NIT; ++t)
#pragma unroll
            for (int c = 0; c < CP; ++c) acc[t][c] = 0.0;
        for (int ip = 0; ip < Kmax; ++ip) {
            if (!once) {
                const int ipk = ip < K ? ip : K - 1;
                if (ip % ONEE_KCH == 0) stage(ip);
                build_tables(ipk % ONEE_KCH, ip < K, pass_emax, pass_kin);
            }
#pragma unroll
            for (int t = 0; t < NIT; ++t) {
                const double* X = TB + (toff[t] & 255);
                const double* Y = TB + DS + ((toff[t] >> 8) & 255);
                const double* Z = TB + 2 * DS + (toff[t] >> 16);
                const double sx = X[0], sy = Y[0], sz = Z[0];
                const double syz = sy * sz, sxz = sx * sz, sxy = sx * sy;
                // component `comp` lives in slot comp - c0 of this pass (compile-time slot, run-time test)
#define ONEE_PUT(comp, expr)                                                                     \
    if constexpr ((comp) >= c0 && (comp) < c0 + CP) {                                            \
        if ((pmask >> ((comp) - c0)) & 1u) acc[t][((comp) - c0) < CP ? ((comp) - c0) : 0] += (expr); \
    }
                ONEE_PUT(0, sx * syz)
                if (pass_kin) { ONEE_PUT(1, fma(X[3 * TS], syz, fma(Y[3 * TS], sxz, Z[3 * TS] * sxy))) }
                if (pass_emax >= 1) {
                    const double dx = X[TS], dy = Y[TS], dz = Z[TS];
                    ONEE_PUT(2, dx * syz)
                    ONEE_PUT(3, dy * sxz)
                    ONEE_PUT(4, dz * sxy)
                    if (pass_emax >= 2) {
                        ONEE_PUT(5, X[2 * TS] * syz)
                        ONEE_PUT(6, dx * dy * sz)
                        ONEE_PUT(7, dx * dz * sy)
                        ONEE_PUT(8, Y[2 * TS] * sxz)
                        ONEE_PUT(9, dy * dz * sx)
                        ONEE_PUT(10, Z[2 * TS] * sxy)
                    }
                }
#undef ONEE_PUT
            }
            __syncwarp();
        }
#pragma unroll
        for (int t = 0; t < NIT; ++t) {
            const int ab = q + EPT * t;
            if (ab < NAB) {
#pragma unroll
                for (int c = 0; c < CP; ++c)
                    if ((pmask >> c) & 1u) CART[c * NAB + ab] = acc[t][c];
            }
        }
        __syncwarp();
        // C2S(a): lanes <-> (component, b); HALF[c][ia][b] = sum_a coef[ia][a] CART[c][a][b], rows unrolled
        for (int item = q; item < CP * NB; item += EPT) {
            const int c = item / NB, bi = item - c * NB;
            if (!((pmask >> c) & 1u)) continue;
            const double* src = CART + c * NAB + bi;
            double* dst = HALF + c * SA * NB + bi;
#pragma unroll
            for (int ia = 0; ia < SA; ++ia) dst[ia * NB] = onee_c2s_row(LA, ia, src, NB);
        }
        __syncwarp();
        // C2S(b): lanes <-> (component, ia); OUTB[c][ia][jb] over the (dead) Cartesian buffer
        for (int item = q; item < CP * SA; item += EPT) {
            if (!((pmask >> (item / SA)) & 1u)) continue;
            const double* src = HALF + item * NB;
            double* dst = CART + item * SB;
#pragma unroll
            for (int jb = 0; jb < SB; ++jb) dst[jb] = onee_c2s_row(LB, jb, src, 1);
        }
        __syncwarp();
        // stores: block rows (c, ia) with lanes along jb, then mirror rows (c, jb) with lanes along ia
        if (valid) {
            {
                constexpr int ROWS = EPT / SB;
                const int r0 = q / SB, j0 = q - r0 * SB;
                if (r0 < ROWS)
                    for (int row = r0; row < CP * SA; row += ROWS) {
                        const int c = row / SA, ia = row - c * SA;
                        if ((pmask >> c) & 1u)
                            A.out[c0 + c][(long long)(it.fa0 + ia) * A.nbf + it.fb0 + j0] = CART[row * SB + j0];
                    }
            }
            if (it.sa != it.sb) {
                constexpr int ROWS = EPT / SA;
                const int r0 = q / SA, i0 = q - r0 * SA;
                if (r0 < ROWS)
                    for (int row = r0; row < CP * SB; row += ROWS) {
                        const int c = row / SB, jb = row - c * SB;
                        if ((pmask >> c) & 1u)
                            A.out[c0 + c][(long long)(it.fb0 + jb) * A.nbf + it.fa0 + i0] = CART[(c * SA + i0) * SB + jb];
                    }
            }
        }
        __syncwarp();
    };
    run_pass(std::integral_constant<int, 0>{});
    if constexpr (CP < ONEE_NCOMP) run_pass(std::integral_constant<int, CP>{});
    if constexpr (2 * CP < ONEE_NCOMP) run_pass(std::integral_constant<int, 2 * CP>{});
    if constexpr (3 * CP < ONEE_NCOMP) run_pass(std::integral_constant<int, 3 * CP>{});
}

// Three instantiations by block size (Cartesian component pairs NAB): SIZE 0 = NAB <= 18 ((ss) ... (dp), (fs), (gs): most
// of the shell pairs; small shared-memory and register footprint -> many resident warps, which is what hides the
// dependent global-load chains of these short groups), SIZE 1 = NAB <= 60, SIZE 2 the rest ((ff), (gd), (gf), (gg)).
// The launches of a call run concurrently on the context's streams.
__host__ __device__ constexpr int onee_size_class(int la, int lb) {
    const int nab = ((la + 1) * (la + 2) / 2) * ((lb + 1) * (lb + 2) / 2);
    return nab <= 18 ? 0 : (nab <= 60 ? 1 : 2);
}
#ifndef ONEE_MB0
#define ONEE_MB0 6
#endif

template <int WARPS, int SIZE>
__global__ void __launch_bounds__(WARPS * 32, SIZE == 0 ? ONEE_MB0 : (SIZE == 1 ? 4 : 3)) k_onee(OneEArgs A, unsigned mask, int wsz_max) {
    extern __shared__ double onee_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* wbase = onee_smem + (size_t)warp * wsz_max;
    // dynamic scheduling over the (heaviest-first) group list; static round-robin was 14 % slower (the long
    // high-contraction s/p groups sit at the end of the list)
    int g = 0;
    if (lane == 0) g = atomicAdd(A.counter, 1);
    g = __shfl_sync(0xffffffffu, g, 0);
    while (g < A.ngroups) {
        // next group: raw atomic (atomicAdd() compiles to warp aggregation + an immediate SHFL, i.e. a wait);
        // its round trip overlaps this group's work
        int gn = 0;
        if (lane == 0) asm volatile("atom.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(gn) : "l"(A.counter) : "memory");
        int s = 0;
        while (g >= A.seg[s].group_end) ++s;
        const OneEClassSeg& sg = A.seg[s];
        const int gi = g - (s ? A.seg[s - 1].group_end : 0);
        switch (sg.la * 5 + sg.lb) {
#define ONEE_CASE(LA, LB)                                                                      \
    case LA * 5 + LB:                                                                          \
        if constexpr (onee_size_class(LA, LB) == SIZE) onee_body<LA, LB>(A, sg, gi, lane, wbase, mask); \
        break;
            ONEE_CASE(0, 0) ONEE_CASE(1, 0) ONEE_CASE(1, 1) ONEE_CASE(2, 0) ONEE_CASE(2, 1) ONEE_CASE(2, 2)
            ONEE_CASE(3, 0) ONEE_CASE(3, 1) ONEE_CASE(3, 2) ONEE_CASE(3, 3)
            ONEE_CASE(4, 0) ONEE_CASE(4, 1) ONEE_CASE(4, 2) ONEE_CASE(4, 3) ONEE_CASE(4, 4)
#undef ONEE_CASE
        }
        g = __shfl_sync(0xffffffffu, gn, 0);
    }
}
