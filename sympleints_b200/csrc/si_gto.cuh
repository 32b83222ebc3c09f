// Section 8(f)-1 keys of the reference that need no recurrence:
//   * gto        {cart,sph}_gto3d_<La>(ax, da, A, R, result): values of a shell's functions at the point R,
//                sum_k d_k (R-A)^lmn exp(-a_k |R-A|^2), normalised and C2S-transformed like every other key
//                (sympleints/defs/gto.py:9-28, main.py:520-549) -- here also batched: all functions of a shell set
//                on a grid of points (the density-evaluation use the reference names it for);
//   * prefactor  prefactors_<La><Lb>(ax, da, A, bx, db, B, result) = da db rP^(a+b) per Cartesian component pair,
//                normalised and C2S-transformed on both indices (defs/overlap.py:16-34, main.py:481-518).  The
//                reference's generated code reads rP / rP2 as module-level names; here rP arrives as the R argument.
// Included by si_lib.cu.
#pragma once

#define SI_GTO_LMAX 5
struct SiGtoConst {
    double c2s[SI_GTO_LMAX + 1][11][21];   // real cart->sph matrices, rows m = -l..l (cart2sph.py:163-202)
    double lmn[SI_GTO_LMAX + 1][21];       // 1/sqrt((2l-1)!!(2m-1)!!(2n-1)!!)  (main.py:175-189)
    int cart[SI_GTO_LMAX + 1][21];         // ax | ay << 4 | az << 8, canonical order
};
__device__ SiGtoConst g_si_gto;

__device__ __forceinline__ double si_ipow(double x, int n) {
    double r = 1.0;
    for (int i = 0; i < n; ++i) r *= x;
    return r;
}

// one thread per (point, shell): out[(foff + m) * ld + point]
__global__ void __launch_bounds__(128) k_gto(BasisDev bs, int nshell, long long npts, const double* pts, int sph, int use_lmn,
                                             double* out, long long ld) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int sh = blockIdx.y;
    if (g >= npts || sh >= nshell) return;
    const int l = bs.L[sh], np = bs.nprim[sh], p0 = bs.poff[sh];
    const double x = pts[3 * g] - bs.cen[3 * sh], y = pts[3 * g + 1] - bs.cen[3 * sh + 1], z = pts[3 * g + 2] - bs.cen[3 * sh + 2];
    const double r2 = x * x + y * y + z * z;
    double rad = 0.0;
    for (int k = 0; k < np; ++k) rad = fma(bs.coef[p0 + k], exp(-bs.exps[p0 + k] * r2), rad);
    const int nc = (l + 1) * (l + 2) / 2;
    double c[21];
    for (int k = 0; k < nc; ++k) {
        const int pk = g_si_gto.cart[l][k];
        double v = si_ipow(x, pk & 15) * si_ipow(y, (pk >> 4) & 15) * si_ipow(z, pk >> 8) * rad;
        if (use_lmn) v *= g_si_gto.lmn[l][k];
        c[k] = v;
    }
    double* o = out + (long long)bs.foff[sh] * ld + g;
    if (sph) {
        for (int m = 0; m < 2 * l + 1; ++m) {
            double acc = 0.0;
            for (int k = 0; k < nc; ++k) acc = fma(g_si_gto.c2s[l][m][k], c[k], acc);
            o[(long long)m * ld] = acc;
        }
    } else {
        for (int k = 0; k < nc; ++k) o[(long long)k * ld] = c[k];
    }
}

// one block; thread t = (ia, jb) of the output block
__global__ void k_prefactor(int la, int lb, int sph, int use_lmn, double dsum, double rx, double ry, double rz, double* out) {
    const int nca = (la + 1) * (la + 2) / 2, ncb = (lb + 1) * (lb + 2) / 2;
    const int na = sph ? 2 * la + 1 : nca, nb = sph ? 2 * lb + 1 : ncb;
    const int t = threadIdx.x;
    if (t >= na * nb) return;
    const int ia = t / nb, jb = t - ia * nb;
    double acc = 0.0;
    for (int a = 0; a < nca; ++a) {
        const double ca = sph ? g_si_gto.c2s[la][ia][a] : (a == ia ? 1.0 : 0.0);
        if (ca == 0.0) continue;
        const int pa = g_si_gto.cart[la][a];
        for (int b = 0; b < ncb; ++b) {
            const double cb = sph ? g_si_gto.c2s[lb][jb][b] : (b == jb ? 1.0 : 0.0);
            if (cb == 0.0) continue;
            const int pb = g_si_gto.cart[lb][b];
            double v = si_ipow(rx, (pa & 15) + (pb & 15)) * si_ipow(ry, ((pa >> 4) & 15) + ((pb >> 4) & 15)) *
                       si_ipow(rz, (pa >> 8) + (pb >> 8));
            if (use_lmn) v *= g_si_gto.lmn[la][a] * g_si_gto.lmn[lb][b];
            acc = fma(ca * cb, v, acc);
        }
    }
    out[t] = dsum * acc;
}
