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
// Execution model: EPT lanes of a warp per shell pair.  Per primitive pair six lanes build the 1-D tables
// (one lane per direction for the multipole tables e = 0..2, one per direction for the kinetic table) in
// shared memory; then the lanes take Cartesian component pairs (a, b), form
// the products for the requested operators and accumulate the contraction; per shell pair the two half
// transforms C2S(a), C2S(b) run through shared memory and the block and its mirror image are stored.
// The three-dimensional exp(-mu AB^2) da db (pi/p)^(3/2) is put into the x table; the y and z tables start at 1.
#pragma once
#include <type_traits>

#define ONEE_NCOMP 11   // 0 ovlp, 1 kin, 2-4 dpm x y z, 5-10 qpm xx xy xz yy yz zz
#define ONEE_MAXL 4
#define ONEE_KCH 12     // primitive pairs staged in shared memory at a time (12 doubles each)

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

// one lane: all multipole tables I(i, j, e), e <= emax, of one direction, built in place in shared memory (the
// lane reads back its own stores; no register arrays, so the kernel keeps a small register footprint)
template <int LA, int LB>
__device__ __forceinline__ void onee_multipole_tables(double* TB, int emax, double base, double PA, double PB, double PR,
                                                      double oo2p) {
    constexpr int TS = (LA + 1) * (LB + 1), JS = LB + 1;
    for (int e = 0; e <= emax; ++e) {
        double* I = TB + e * TS;
        const double* Im = I - TS;   // order e - 1
#pragma unroll
        for (int j = 0; j <= LB; ++j)
#pragma unroll
            for (int i = 0; i <= LA; ++i) {
                double v;
                if (i > 0) {   // reduce the bra index
                    double t = 0.0;
                    if (i > 1) t = (double)(i - 1) * I[(i - 2) * JS + j];
                    if (j > 0) t = fma((double)j, I[(i - 1) * JS + j - 1], t);
                    if (e > 0) t = fma((double)e, Im[(i - 1) * JS + j], t);
                    v = fma(oo2p, t, PA * I[(i - 1) * JS + j]);
                } else if (j > 0) {   // i == 0: reduce the ket index
                    double t = 0.0;
                    if (j > 1) t = (double)(j - 1) * I[j - 2];
                    if (e > 0) t = fma((double)e, Im[j - 1], t);
                    v = fma(oo2p, t, PB * I[j - 1]);
                } else if (e > 0) {   // i == j == 0: reduce the multipole index
                    v = PR * Im[0];
                    if (e > 1) v = fma(oo2p * (double)(e - 1), Im[-TS], v);
                } else {
                    v = base;
                }
                I[i * JS + j] = v;
            }
    }
}

// one lane: kinetic table of one direction (needs the overlap table e = 0 of the same direction; the lane
// rebuilds it in `S`, a scratch table of its own, so that it does not wait for the multipole lane)
template <int LA, int LB>
__device__ __forceinline__ void onee_kinetic_table(double* T, double* S, double base, double PA, double PB, double oo2p,
                                                   double a, double b, double oop) {
    constexpr int JS = LB + 1;
    const double bop = b * oop, aop = a * oop;
#pragma unroll
    for (int j = 0; j <= LB; ++j)
#pragma unroll
        for (int i = 0; i <= LA; ++i) {
            double s, t;
            if (i > 0) {
                double us = 0.0, ut = 0.0;
                if (i > 1) {
                    us = (double)(i - 1) * S[(i - 2) * JS + j];
                    ut = (double)(i - 1) * T[(i - 2) * JS + j];
                }
                if (j > 0) {
                    us = fma((double)j, S[(i - 1) * JS + j - 1], us);
                    ut = fma((double)j, T[(i - 1) * JS + j - 1], ut);
                }
                s = fma(oo2p, us, PA * S[(i - 1) * JS + j]);
                t = fma(oo2p, ut, PA * T[(i - 1) * JS + j]);
                double w = 2.0 * a * s;
                if (i > 1) w = fma(-(double)(i - 1), S[(i - 2) * JS + j], w);
                t = fma(bop, w, t);
            } else if (j > 0) {
                s = PB * S[j - 1];
                t = PB * T[j - 1];
                if (j > 1) {
                    s = fma(oo2p * (double)(j - 1), S[j - 2], s);
                    t = fma(oo2p * (double)(j - 1), T[j - 2], t);
                }
                double w = 2.0 * b * s;
                if (j > 1) w = fma(-(double)(j - 1), S[j - 2], w);
                t = fma(aop, w, t);
            } else {
                s = base;
                t = (a - 2.0 * a * a * fma(PA, PA, oo2p)) * s;
            }
            S[i * JS + j] = s;
            T[i * JS + j] = t;
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
    const int pa0 = A.poff[it.sa], pb0 = A.poff[it.sb];
    // Stage the data of primitive pairs [k0, k0 + ONEE_KCH) with coalesced loads: one global round trip per chunk
    // instead of one (dependent) per primitive pair.
    auto stage = [&](int k0) {
        const int nk = max(0, min(ONEE_KCH, K - k0));   // (0 for pairs with fewer primitives: they keep their last chunk)
        for (int i = q; i < nk * 12; i += EPT) {
            const int k = i / 12, j = i - k * 12;
            double v = 0.0;
            if (j < 9) {
                v = A.ppair[(size_t)(it.ppoff + k0 + k) * 9 + j];
            } else if (j == 9) {
                v = A.exps[pa0 + (k0 + k) / it.Kb];
            } else if (j == 10) {
                v = A.exps[pb0 + (k0 + k) % it.Kb];
            }
            PP[i] = v;
        }
        __syncwarp();
    };
    auto build_tables = [&](int kk, bool live, int e_hi, bool kin) {
        const double* pp = PP + kk * 12;
        if (q < 6) {
            const int d = q % 3;
            const double oop = pp[1];
            const double PA = pp[5 + d], PB = PA + it.AB[d], oo2p = 0.5 * oop;
            // (pi/p)^(3/2) exp(-mu AB^2) da db in the x table, 1 in y and z; pairs with fewer primitives add zeros
            const double base = d == 0 ? (live ? pp[8] : 0.0) * (SI_PI * oop) * sqrt(SI_PI * oop) : 1.0;
            if (q < 3) {
                onee_multipole_tables<LA, LB>(TB + d * DS, e_hi, base, PA, PB, A.origin_p ? 0.0 : pp[2 + d] - A.R[d], oo2p);
            } else if (kin) {
                onee_kinetic_table<LA, LB>(TB + d * DS + 3 * TS, TB + d * DS + 4 * TS, base, PA, PB, oo2p, pp[9], pp[10], oop);
            }
        }
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

// Two instantiations: BIG = false holds the classes with at most 60 Cartesian component pairs (small shared-memory
// and register footprint -> many resident warps, which is what hides the long dependent chains of these groups),
// BIG = true the rest ((ff), (gd), (gf), (gg)).  The two launches of a call run concurrently on two streams.
__host__ __device__ constexpr bool onee_is_big(int la, int lb) {
    return ((la + 1) * (la + 2) / 2) * ((lb + 1) * (lb + 2) / 2) > 60;
}

template <int WARPS, bool BIG>
__global__ void __launch_bounds__(WARPS * 32, BIG ? 3 : 4) k_onee(OneEArgs A, unsigned mask, int wsz_max) {
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
        if constexpr (onee_is_big(LA, LB) == BIG) onee_body<LA, LB>(A, sg, gi, lane, wbase, mask); \
        break;
            ONEE_CASE(0, 0) ONEE_CASE(1, 0) ONEE_CASE(1, 1) ONEE_CASE(2, 0) ONEE_CASE(2, 1) ONEE_CASE(2, 2)
            ONEE_CASE(3, 0) ONEE_CASE(3, 1) ONEE_CASE(3, 2) ONEE_CASE(3, 3)
            ONEE_CASE(4, 0) ONEE_CASE(4, 1) ONEE_CASE(4, 2) ONEE_CASE(4, 3) ONEE_CASE(4, 4)
#undef ONEE_CASE
        }
        g = __shfl_sync(0xffffffffu, gn, 0);
    }
}
