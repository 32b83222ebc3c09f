// sympleints_b200 -- recurrence "stage machine" shared by the CUDA kernels and
// the host-side test emulator.
//
// A shell class (key, La, Lb[, Lc]) is compiled on the host (si_program.cpp)
// into a *program*: an ordered list of stages.  Each stage is a set of
// independent elements
//     W[dst + b*dbs] (+)= sum_k  S[sid_k] * W[src_k + b*sbs]          (linear)
//     W[dst + b*dbs] (+)= sum_k  W[s3k] * W[s3k+1] * W[s3k+2]         (PROD3)
// over a flat work array W (shared memory on the device) and a scalar table S
// (per primitive tuple: PA, PC, k/(2p), ...; followed by per-class constants,
// e.g. cart2sph coefficients).  This is the flat-work-array formulation of
// the reference's graph generator (sympleints/graphs/generate.py:20-67) turned
// into data that one kernel executes for every class.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SI_HD __host__ __device__ __forceinline__
#else
#define SI_HD inline
#endif

// ---- scalar table layout (dynamic part) ------------------------------------
#define SI_KMAX 10
#define SI_SID_ONE 0
#define SI_SID_V0 1   // 3 entries each
#define SI_SID_V1 4
#define SI_SID_V2 7
#define SI_SID_V3 10
#define SI_SID_K0 13  // SI_KMAX entries each: (k+1)*kbase
#define SI_SID_K1 23
#define SI_SID_K2 33
#define SI_SID_K3 43
#define SI_SID_K4 53
#define SI_SID_S0 63
#define SI_NDYN 64

// ---- stage flags ------------------------------------------------------------
#define SI_ST_ACCUM 1   // W[dst] += value
#define SI_ST_PROD3 2   // sum of triple products (1e integrals)
#define SI_ST_SYNC 4    // a barrier is required before this stage

struct SiStage {
    int32_t hdr_off;   // into hdr[]: dst address per element
    int32_t term_off;  // into terms[]: terms[term_off + k*nelem + el]
    int32_t nelem;
    int32_t nterms;
    int32_t batched;   // 1: element is repeated over the batch index b
    int32_t sbs, dbs;  // batch strides for src / dst
    int32_t scs, dcs;  // chunk-offset strides (multiplied by first batch index of the chunk)
    int32_t flags;
    uint32_t magic;    // floor(2^32 / nelem) + 1 for flat-index division
    int32_t nb_fixed;  // >0: stage has its own batch count (C2S on c, batched over a')
};

// Program descriptor as seen by the executor (pointers are host or device).
struct SiProg {
    const uint32_t* hdr;
    const uint32_t* terms;
    const SiStage* stages;
    const double* consts;
    int32_t nconst;
    int32_t nstages;
    int32_t n_prim_stages;  // stages [0, n_prim_stages) run once per primitive tuple
    int32_t chunk_begin;    // stages [chunk_begin, nstages) run once per batch chunk
    int32_t nb_total;       // batch extent (2Lc+1 for 3c2e, ncomp for 1e, else 1)
    int32_t chunk;          // batch indices processed per chunk pass
    int32_t w_size;         // doubles of work array per slot
    int32_t x_off, x_size;  // contraction accumulator (zeroed per shell tuple)
    int32_t base_off;       // W address of the first base value
    int32_t nbase;          // number of base values (Boys orders / 1-D bases)
    int32_t boys_nlo;       // lowest Boys order
    int32_t out_off;        // final block, layout [n0][n1][nb_total] or [nb][n0][n1]
    int32_t out_n0, out_n1;
    int32_t out_batch_major;  // 1: out[(b*n0 + i)*n1 + j] (1e components), 0: out[(i*n1+j)*nb + b]
    int32_t La, Lb, Lc;
};

// Execute stages [s0, s1) for batch indices [m0, m0+nb).
// Each lane works on SI_UNROLL independent elements at a time so that the look-up-table loads,
// the shared-memory operand loads and the FMAs of different elements overlap.
#ifndef SI_UNROLL
#define SI_UNROLL 1
#endif
template <typename SyncF>
SI_HD void si_run_stages(const SiProg& P, const SiStage* stages, int s0, int s1, double* W, const double* S,
                         int m0, int nb, int lane, int nlanes, SyncF sync) {
    for (int s = s0; s < s1; ++s) {
        const SiStage st = stages[s];
        if (st.flags & SI_ST_SYNC) sync();
        const int nbb = st.nb_fixed > 0 ? st.nb_fixed : (st.batched ? nb : 1);
        const int total = st.nelem * nbb;
        const uint32_t* hdr = P.hdr + st.hdr_off;
        const uint32_t* trm = P.terms + st.term_off;
        const int soff = st.scs * m0, doff = st.dcs * m0;
        const int nelem = st.nelem, nterms = st.nterms;
        for (int e0 = lane; e0 < total; e0 += nlanes * SI_UNROLL) {
            int el[SI_UNROLL], sb[SI_UNROLL], db[SI_UNROLL];
            bool ok[SI_UNROLL];
            double acc[SI_UNROLL];
#pragma unroll
            for (int u = 0; u < SI_UNROLL; ++u) {
                int e = e0 + u * nlanes;
                ok[u] = e < total;
                if (!ok[u]) e = e0;
                int b = 0, x = e;
                if (nbb > 1 && nelem == 1) {
                    b = e;
                    x = 0;
                } else if (nbb > 1) {
#if defined(__CUDA_ARCH__)
                    b = (int)__umulhi((uint32_t)e, st.magic);
#else
                    b = (int)(((uint64_t)(uint32_t)e * st.magic) >> 32);
#endif
                    x = e - b * nelem;
                }
                el[u] = x;
                sb[u] = soff + b * st.sbs;
                db[u] = doff + b * st.dbs;
                acc[u] = 0.0;
            }
            if (st.flags & SI_ST_PROD3) {
                for (int k = 0; k < nterms; k += 3) {
#pragma unroll
                    for (int u = 0; u < SI_UNROLL; ++u) {
                        uint32_t t0 = trm[(k + 0) * nelem + el[u]];
                        uint32_t t1 = trm[(k + 1) * nelem + el[u]];
                        uint32_t t2 = trm[(k + 2) * nelem + el[u]];
                        acc[u] += W[(t0 & 0xffffu)] * W[(t1 & 0xffffu)] * W[(t2 & 0xffffu)];
                    }
                }
            } else {
                for (int k = 0; k < nterms; ++k) {
                    uint32_t t[SI_UNROLL];
#pragma unroll
                    for (int u = 0; u < SI_UNROLL; ++u) t[u] = trm[k * nelem + el[u]];
#pragma unroll
                    for (int u = 0; u < SI_UNROLL; ++u) acc[u] += S[t[u] >> 16] * W[(int)(t[u] & 0xffffu) + sb[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < SI_UNROLL; ++u) {
                if (ok[u]) {
                    const int d = (int)hdr[el[u]] + db[u];
                    if (st.flags & SI_ST_ACCUM)
                        W[d] += acc[u];
                    else
                        W[d] = acc[u];
                }
            }
        }
    }
}

// ---- Boys function: algorithm of sympleints/testing/boys.py:96-140 ----------
// table: row-major [nrows][ncols] with ncols = 301, dx = 0.1, xlarge = 30.
struct SiBoys {
    const double* table;   // (nmax + 7) x 301
    const double* xlarge_pref;  // (2n-1)!!/2^(n+1) sqrt(pi), n = 0..nmax
    int32_t nrows, ncols;
    // the same table transposed, 301 x (nmax + 7): the six Taylor coefficients F_n..F_{n+5} of one grid
    // point are contiguous (one or two cache lines instead of six), and all orders of one tuple share them
    const double* tableT;
};

SI_HD double si_boys(const SiBoys& B, int n, double x) {
    if (x == 0.0) return B.table[n * B.ncols];
    if (x < 30.0) {
        const int ind0 = (int)(x * 10.0);
        const double xg = ind0 * 0.1;
        const double mdx = -(x - xg);
        const double* t = B.table + n * B.ncols + ind0;
        // result += table[n+k, ind0] * (-dx)^k / k!   (same summation order)
        double r = t[0];
        double pw = mdx;
        r += t[B.ncols] * pw * 1.0;
        pw *= mdx;
        r += t[2 * B.ncols] * pw * 0.5;
        pw *= mdx;
        r += t[3 * B.ncols] * pw * (1.0 / 6.0);
        pw *= mdx;
        r += t[4 * B.ncols] * pw * (1.0 / 24.0);
        pw *= mdx;
        r += t[5 * B.ncols] * pw * (1.0 / 120.0);
        return r;
    }
    // x >= 30: xlarge_pref[n] * x^(-n - 1/2)
#if defined(__CUDA_ARCH__)
    return B.xlarge_pref[n] * pow(x, -(double)n - 0.5);
#else
    return B.xlarge_pref[n] * __builtin_pow(x, -(double)n - 0.5);
#endif
}

// Same algorithm, cheaper evaluation for the generated kernels: Horner form of the 6-term Taylor sum and
// x^(-n-1/2) = sqrt(1/x) (1/x)^n by binary exponentiation instead of pow() (differences ~1e-16 relative).
SI_HD double si_boys_taylor(const SiBoys& B, int n, double x) {
    // (x == 0 gives ind0 = 0, mdx = 0 and the sum reduces to the tabulated F_n(0) exactly)
    const int ind0 = (int)(x * 10.0);
    const double mdx = ind0 * 0.1 - x;
    const double* t = B.tableT + ind0 * B.nrows + n;
    const double t0 = t[0], t1 = t[1], t2 = t[2], t3 = t[3], t4 = t[4], t5 = t[5];
    double r = t5 * (1.0 / 120.0);
    r = fma(r, mdx, t4 * (1.0 / 24.0));
    r = fma(r, mdx, t3 * (1.0 / 6.0));
    r = fma(r, mdx, t2 * 0.5);
    r = fma(r, mdx, t1);
    r = fma(r, mdx, t0);
    return r;
}
// x^(-n-1/2) from ONE reciprocal square root (a division and a square root cost ~3x as much; point charges
// far from the pair make the asymptotic branch the common one in the nuclear-attraction kernels)
SI_HD double si_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
// rs^(2n+1) for rs = x^(-1/2), by binary powering (n < 32)
SI_HD double si_boys_xpow_rs(int n, double rs) {
    double acc = rs, b = rs * rs;
    int e = n;
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        if (e & 1) acc *= b;
        b *= b;
        e >>= 1;
    }
    return acc;
}
SI_HD double si_boys_xpow(int n, double x) {
    const double rs = si_rsqrt(x);
    double acc = rs, b = rs * rs;
    int e = n;
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        if (e & 1) acc *= b;
        b *= b;
        e >>= 1;
    }
    return acc;
}
SI_HD double si_boys_fast(const SiBoys& B, int n, double x) {
    if (x < 30.0) return si_boys_taylor(B, n, x);
    return B.xlarge_pref[n] * si_boys_xpow(n, x);
}
// pref_n = B.xlarge_pref[n], for callers with a fixed order per lane (kept in a register)
SI_HD double si_boys_fast_pref(const SiBoys& B, int n, double x, double pref_n) {
    if (x < 30.0) return si_boys_taylor(B, n, x);
    return pref_n * si_boys_xpow(n, x);
}

// ---- dynamic scalar table fill ----------------------------------------------
// v[12]: four 3-vectors (V0..V3); kb[5]: bases of the k-multiples; s0: single.
SI_HD void si_fill_scalars(double* S, const double* v, const double* kb, double s0, int lane,
                           int nlanes) {
    for (int idx = lane; idx < SI_NDYN; idx += nlanes) {
        double val = 1.0;
        if (idx >= 1 && idx < 13) {
            const int j = idx - 1;
#pragma unroll
            for (int q = 0; q < 12; ++q)
                if (j == q) val = v[q];
        } else if (idx >= 13 && idx < 63) {
            const int j = (idx - 13) / SI_KMAX;
            const int k = (idx - 13) - j * SI_KMAX;
            double base = 0.0;
#pragma unroll
            for (int q = 0; q < 5; ++q)
                if (j == q) base = kb[q];
            val = (double)(k + 1) * base;
        } else if (idx == 63) {
            val = s0;
        }
        S[idx] = val;
    }
}

#define SI_PI 3.14159265358979323846
#define SI_PI_25 17.493418327624862846  // pi^2.5

// Per-primitive set-up.  All lanes compute the (uniform) geometry redundantly,
// then fill S cooperatively and write the base values cooperatively.

// 3c2e (ab|c): defs/coulomb.py:238-323.
SI_HD void si_setup_3c2e(const SiProg& P, const SiBoys& B, double ax, double da, const double* A,
                         double bx, double db, const double* Bc, double cx, double dc,
                         const double* C, double* W, double* S, int lane, int nlanes) {
    const double p = ax + bx;
    const double oop = 1.0 / p;
    const double mu = ax * bx * oop;
    double v[12], kb[5];
    double r2ab = 0.0, r2pc = 0.0;
    const double rho = p * cx / (p + cx);
    const double rho_p = rho * oop;
    const double rho_c = rho / cx;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double Pd = (ax * A[d] + bx * Bc[d]) * oop;
        const double ab = A[d] - Bc[d];
        const double pc = Pd - C[d];
        r2ab += ab * ab;
        r2pc += pc * pc;
        v[0 + d] = Pd - A[d];     // V0 = PA
        v[3 + d] = -rho_p * pc;   // V1 = -(rho/p) PC
        v[6 + d] = ab;            // V2 = AB
        v[9 + d] = rho_c * pc;    // V3 = (rho/c) PC
    }
    kb[0] = 0.5 * oop;             // K0 = k/(2p)
    kb[1] = -0.5 * oop * rho_p;    // K1 = -k rho/(2 p^2)
    kb[2] = 0.5 / (p + cx);        // K2 = k/(2(p+c))  == (rho/c) k/(2p)
    kb[3] = 0.0;
    kb[4] = 0.0;
    si_fill_scalars(S, v, kb, 0.0, lane, nlanes);
    const double T = rho * r2pc;
    const double pref = 2.0 * SI_PI_25 / (sqrt(p + cx) * p * cx) * exp(-mu * r2ab) * (da * db * dc);
    for (int j = lane; j < P.nbase; j += nlanes) W[P.base_off + j] = pref * si_boys(B, P.boys_nlo + j, T);
}

// coul (a|1/r_R|b): defs/coulomb.py:35-80 (VRR on a) + graphs/defs/int_nucattr.py:14-31 (HRR).
SI_HD void si_setup_coul(const SiProg& P, const SiBoys& B, double ax, double da, const double* A,
                         double bx, double db, const double* Bc, const double* R, double q,
                         double* W, double* S, int lane, int nlanes) {
    const double p = ax + bx;
    const double oop = 1.0 / p;
    const double mu = ax * bx * oop;
    double v[12], kb[5];
    double r2ab = 0.0, r2pr = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double Pd = (ax * A[d] + bx * Bc[d]) * oop;
        const double ab = A[d] - Bc[d];
        const double pr = Pd - R[d];
        r2ab += ab * ab;
        r2pr += pr * pr;
        v[0 + d] = Pd - A[d];  // PA
        v[3 + d] = -pr;        // -PR
        v[6 + d] = ab;         // AB
        v[9 + d] = Pd - Bc[d]; // PB
    }
    kb[0] = 0.5 * oop;
    kb[1] = -0.5 * oop;
    kb[2] = kb[3] = kb[4] = 0.0;
    si_fill_scalars(S, v, kb, 0.0, lane, nlanes);
    const double T = p * r2pr;
    const double pref = 2.0 * SI_PI * oop * exp(-mu * r2ab) * (da * db * q);
    for (int j = lane; j < P.nbase; j += nlanes) W[P.base_off + j] = pref * si_boys(B, P.boys_nlo + j, T);
}

// 2c2e (a|b): defs/coulomb.py:120-160.
SI_HD void si_setup_2c2e(const SiProg& P, const SiBoys& B, double ax, double da, const double* A,
                         double bx, double db, const double* Bc, double* W, double* S, int lane,
                         int nlanes) {
    const double p = ax + bx;
    const double oop = 1.0 / p;
    const double mu = ax * bx * oop;
    double v[12], kb[5];
    double r2ab = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double Pd = (ax * A[d] + bx * Bc[d]) * oop;
        const double ab = A[d] - Bc[d];
        r2ab += ab * ab;
        v[0 + d] = Pd - A[d];   // PA
        v[3 + d] = Pd - Bc[d];  // PB
        v[6 + d] = ab;
        v[9 + d] = 0.0;
    }
    kb[0] = 0.5 / ax;               // k/(2a)
    kb[1] = -0.5 / ax * (bx * oop); // -k (b/p)/(2a)
    kb[2] = 0.5 * oop;              // k/(2p)
    kb[3] = 0.5 / bx;               // k/(2b)
    kb[4] = -0.5 / bx * (ax * oop); // -k (a/p)/(2b)
    si_fill_scalars(S, v, kb, 0.0, lane, nlanes);
    const double T = mu * r2ab;
    const double pref = 2.0 * SI_PI_25 / (sqrt(p) * ax * bx) * (da * db);
    for (int j = lane; j < P.nbase; j += nlanes) W[P.base_off + j] = pref * si_boys(B, P.boys_nlo + j, T);
}

// 1e multipole/kinetic: defs/multipole.py:17-32, defs/kinetic.py:15-44.
// Base values: S_d(0,0,0) for d = x,y,z (coefficient product folded into x) and
// T_d(0,0) for the kinetic program (nbase == 6).
SI_HD void si_setup_1e(const SiProg& P, double ax, double da, const double* A, double bx,
                       double db, const double* Bc, const double* R, double* W, double* S,
                       int lane, int nlanes) {
    const double p = ax + bx;
    const double oop = 1.0 / p;
    const double mu = ax * bx * oop;
    double v[12], kb[5];
    double s00[3], t00[3];
    const double sq = sqrt(SI_PI * oop);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double Pd = (ax * A[d] + bx * Bc[d]) * oop;
        const double ab = A[d] - Bc[d];
        const double pa = Pd - A[d];
        v[0 + d] = pa;           // PA
        v[3 + d] = Pd - Bc[d];   // PB
        v[6 + d] = Pd - R[d];    // PR
        v[9 + d] = 0.0;
        s00[d] = sq * exp(-mu * ab * ab);
        if (d == 0) s00[d] *= da * db;
        t00[d] = (ax - 2.0 * ax * ax * (pa * pa + 0.5 * oop)) * s00[d];
    }
    kb[0] = 0.5 * oop;    // k/(2p)
    kb[1] = -bx * oop;    // -k b/p
    kb[2] = -ax * oop;    // -k a/p
    kb[3] = kb[4] = 0.0;
    // S0 = 2ab/p
    si_fill_scalars(S, v, kb, 2.0 * ax * bx * oop, lane, nlanes);
    for (int j = lane; j < P.nbase; j += nlanes) W[P.base_off + j] = (j < 3) ? s00[j] : t00[j - 3];
}
