// Schwarz integrals (ab|ab) of a shell set: the diagonal of the four-centre integral the reference
// defines on its graph path (sympleints/graphs/defs/int_4c2e.py:11-75: base 2 pi^2.5 / (p q sqrt(p+q)) K_AB K_CD F_n,
// all four indices spherical; classes (La Lb|La Lb) of l_iters.py:23-28 `schwarz_iter`).  Included by si_lib.cu.
//
// The reference builds them with the Obara-Saika/HRR transformations of that file (Fortran only).  Only the diagonal is
// needed here, which the McMurchie-Davidson form gives directly as a bilinear form per spherical component pair (i, j):
//     (ij|ij) = sum_{P,Q prim. pairs} 2 pi^2.5 / (p q sqrt(p+q)) sum_h E^P_{ij,h} sum_h' (-1)^{|h'|} R_{h+h'}(alpha, P-Q) E^Q_{ij,h'}
// with the Hermite expansion coefficients E (1-D recurrences, contraction coefficients, lmn factors and C2S on both
// indices folded in) and the Hermite integrals R (one-centre recursion from (-2 alpha)^n F_n(alpha |PQ|^2)).
// One CTA of 128 threads per shell pair; component pairs in chunks so that two E buffers + two R cubes fit shared memory.
#pragma once

struct SchwarzArgs {
    const SiPairItem* items;
    int n;
    int la, lb;
    const double* ppair;
    double* D;          // (nbf, nbf): D[mu, nu] = (mu nu|mu nu)
    long long nbf;
    SiBoys boys;
    int use_lmn;
    int chunk;          // spherical component pairs per pass
};

__device__ __forceinline__ void schwarz_e1d(double* e, int la, int lb, int ts, double base, double PA, double PB, double oo2p) {
    // e[(i * (lb+1) + j) * ts + t], ts = la + lb + 2 (one zero pad), one thread
    const int nj = lb + 1;
    for (int k = 0; k < (la + 1) * nj * ts; ++k) e[k] = 0.0;
    e[0] = base;
    for (int i = 0; i < la; ++i)
        for (int t = 0; t <= i + 1; ++t) {
            const double* s = e + (i * nj) * ts;
            e[((i + 1) * nj) * ts + t] = (t > 0 ? s[t - 1] * oo2p : 0.0) + PA * s[t] + (double)(t + 1) * s[t + 1];
        }
    for (int i = 0; i <= la; ++i)
        for (int j = 0; j < lb; ++j)
            for (int t = 0; t <= i + j + 1; ++t) {
                const double* s = e + (i * nj + j) * ts;
                e[(i * nj + j + 1) * ts + t] = (t > 0 ? s[t - 1] * oo2p : 0.0) + PB * s[t] + (double)(t + 1) * s[t + 1];
            }
}

__global__ void __launch_bounds__(512) k_schwarz(SchwarzArgs A) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int la = A.la, lb = A.lb, L = la + lb, M = 2 * L, M1 = M + 1;
    const int nca = (la + 1) * (la + 2) / 2, ncb = (lb + 1) * (lb + 2) / 2, nsa = 2 * la + 1, nsb = 2 * lb + 1, ncomp = nsa * nsb;
    const int NH = (L + 1) * (L + 2) * (L + 3) / 6, ts = L + 2, e1sz = (la + 1) * (lb + 1) * ts;
    const int CH = A.chunk;
    // shared-memory layout
    double* acc = sm;                        // ncomp
    double* fn = acc + ncomp;                // M1 Boys-type base values
    double* e1d = fn + M1 + (M1 & 1);        // [2][3][e1sz]
    double* EP = e1d + 6 * e1sz;             // [CH][NH]
    double* EQ = EP + CH * NH;               // [CH][NH]
    double* Ra = EQ + CH * NH;               // [M1^3]
    double* Rb = Ra + M1 * M1 * M1;
    int* hl = reinterpret_cast<int*>(Rb + M1 * M1 * M1);   // NH packed (t | u << 8 | v << 16)
    int* hb = hl + NH;                                      // NH: offset of (t, u, v) in an R cube | odd parity << 31
    const int NW = (NH + 31) / 32;                          // warps (of 32 Hermite indices) per component pair
    double* part = reinterpret_cast<double*>(hb + NH);               // [CH][NW] partial sums of the bilinear form
    if (tid == 0) {
        int h = 0;
        for (int t = 0; t <= L; ++t)
            for (int u = 0; u <= L - t; ++u)
                for (int v = 0; v <= L - t - u; ++v) {
                    hl[h] = t | (u << 8) | (v << 16);
                    hb[h] = ((t * M1 + u) * M1 + v) | (((t + u + v) & 1) << 31);
                    ++h;
                }
    }
    for (int i = tid; i < ncomp; i += nthr) acc[i] = 0.0;
    const SiPairItem it = A.items[blockIdx.x];
    const int K = it.Ka * it.Kb;
    __syncthreads();

    auto build_e = [&](int ip, double* e1, double* E, int c0, int nc) {
        const double* pp = A.ppair + (size_t)(it.ppoff + ip) * SI_PP_STRIDE;
        if (tid < 3) {
            const int d = tid;
            const double PA = pp[5 + d], PB = PA + it.AB[d];
            schwarz_e1d(e1 + d * e1sz, la, lb, ts, d == 0 ? pp[8] : 1.0, PA, PB, 0.5 * pp[1]);
        }
        __syncthreads();
        const double *ex = e1, *ey = e1 + e1sz, *ez = e1 + 2 * e1sz;
        for (int item = tid; item < nc * NH; item += nthr) {
            const int c = item / NH, h = item - c * NH;
            const int comp = c0 + c, i = comp / nsb, j = comp - i * nsb;
            const int t = hl[h] & 255, u = (hl[h] >> 8) & 255, v = hl[h] >> 16;
            double s = 0.0;
            for (int a = 0; a < nca; ++a) {
                double ca = g_si_gto.c2s[la][i][a];
                if (ca == 0.0) continue;
                if (A.use_lmn) ca *= g_si_gto.lmn[la][a];
                const int pa = g_si_gto.cart[la][a], ax = pa & 15, ay = (pa >> 4) & 15, az = pa >> 8;
                for (int b = 0; b < ncb; ++b) {
                    double cb = g_si_gto.c2s[lb][j][b];
                    if (cb == 0.0) continue;
                    if (A.use_lmn) cb *= g_si_gto.lmn[lb][b];
                    const int pb = g_si_gto.cart[lb][b], bx = pb & 15, by = (pb >> 4) & 15, bz = pb >> 8;
                    if (t > ax + bx || u > ay + by || v > az + bz) continue;
                    s = fma(ca * cb, ex[(ax * (lb + 1) + bx) * ts + t] * ey[(ay * (lb + 1) + by) * ts + u] * ez[(az * (lb + 1) + bz) * ts + v], s);
                }
            }
            E[item] = s;
        }
        __syncthreads();
    };

    for (int c0 = 0; c0 < ncomp; c0 += CH) {
        const int nc = min(CH, ncomp - c0);
        for (int ipP = 0; ipP < K; ++ipP) {
            build_e(ipP, e1d, EP, c0, nc);
            const double* ppP = A.ppair + (size_t)(it.ppoff + ipP) * SI_PP_STRIDE;
            for (int ipQ = ipP; ipQ < K; ++ipQ) {
                const double* EQp = EP;
                if (ipQ != ipP) {
                    build_e(ipQ, e1d + 3 * e1sz, EQ, c0, nc);
                    EQp = EQ;
                }
                const double* ppQ = A.ppair + (size_t)(it.ppoff + ipQ) * SI_PP_STRIDE;
                const double p = ppP[0], q = ppQ[0], alpha = p * q / (p + q);
                const double X = ppP[2] - ppQ[2], Y = ppP[3] - ppQ[3], Z = ppP[4] - ppQ[4];
                const double T = alpha * (X * X + Y * Y + Z * Z);
                // base values (-2 alpha)^n F_n(T)
                for (int n = tid; n <= M; n += nthr) {
                    double f = 1.0;
                    for (int k = 0; k < n; ++k) f *= -2.0 * alpha;
                    fn[n] = f * si_boys(A.boys, n, T);
                }
                __syncthreads();
                // Hermite integrals by downward recursion in n: level n in cur, level n + 1 in prev (cube layouts)
                double *cur = Ra, *prev = Rb;
                for (int n = M; n >= 0; --n) {
                    const int lim = M - n;
                    for (int idx = tid; idx < (lim + 1) * (lim + 1) * (lim + 1); idx += nthr) {
                        const int t = idx / ((lim + 1) * (lim + 1)), r = idx - t * (lim + 1) * (lim + 1), u = r / (lim + 1), v = r - u * (lim + 1);
                        if (t + u + v > lim) continue;
                        double val;
                        if (t + u + v == 0) {
                            val = fn[n];
                        } else if (t > 0) {
                            val = X * prev[((t - 1) * M1 + u) * M1 + v];
                            if (t > 1) val = fma((double)(t - 1), prev[((t - 2) * M1 + u) * M1 + v], val);
                        } else if (u > 0) {
                            val = Y * prev[(t * M1 + u - 1) * M1 + v];
                            if (u > 1) val = fma((double)(u - 1), prev[(t * M1 + u - 2) * M1 + v], val);
                        } else {
                            val = Z * prev[(t * M1 + u) * M1 + v - 1];
                            if (v > 1) val = fma((double)(v - 1), prev[(t * M1 + u) * M1 + v - 2], val);
                        }
                        cur[(t * M1 + u) * M1 + v] = val;
                    }
                    __syncthreads();
                    double* tmp = cur;
                    cur = prev;
                    prev = tmp;
                }
                const double* R0 = prev;   // level 0
                const double pref = 2.0 * SI_PI_25 / (p * q * sqrt(p + q)) * (ipQ == ipP ? 1.0 : 2.0);
                // bilinear form: work items (component pair c, Hermite index h), the 32 lanes of a warp hold consecutive h of
                // one c; the inner sum over h' reads R (gather), E^Q (broadcast) and the packed offsets (broadcast)
                for (int item = tid; item < nc * NW * 32; item += nthr) {
                    const int c = item / (NW * 32), h = item - c * (NW * 32);
                    double sh = 0.0;
                    if (h < NH) {
                        const double a1 = EP[c * NH + h];
                        if (a1 != 0.0) {
                            const double* e2 = EQp + c * NH;
                            const double* Rh = R0 + (hb[h] & 0x7fffffff);
                            double i0 = 0.0, i1 = 0.0;
                            int g = 0;
                            for (; g + 1 < NH; g += 2) {
                                const int w0 = hb[g], w1 = hb[g + 1];
                                const double r0 = Rh[w0 & 0x7fffffff], r1 = Rh[w1 & 0x7fffffff];
                                i0 = fma(w0 < 0 ? -r0 : r0, e2[g], i0);
                                i1 = fma(w1 < 0 ? -r1 : r1, e2[g + 1], i1);
                            }
                            if (g < NH) {
                                const int w0 = hb[g];
                                const double r0 = Rh[w0 & 0x7fffffff];
                                i0 = fma(w0 < 0 ? -r0 : r0, e2[g], i0);
                            }
                            sh = a1 * (i0 + i1);
                        }
                    }
                    for (int o = 16; o > 0; o >>= 1) sh += __shfl_xor_sync(0xffffffffu, sh, o);
                    if ((tid & 31) == 0) part[c * NW + (h >> 5)] = sh;
                }
                __syncthreads();
                for (int c = tid; c < nc; c += nthr) {
                    double sc = 0.0;
                    for (int w = 0; w < NW; ++w) sc += part[c * NW + w];
                    acc[c0 + c] += pref * sc;
                }
                __syncthreads();
            }
        }
    }
    for (int c = tid; c < ncomp; c += nthr) {
        const int i = c / nsb, j = c - i * nsb;
        if (it.sa == it.sb && i > j) continue;   // same shell twice: (i j|i j) == (j i|j i), written once from i <= j
        A.D[(long long)(it.fa0 + i) * A.nbf + it.fb0 + j] = acc[c];
        A.D[(long long)(it.fb0 + j) * A.nbf + it.fa0 + i] = acc[c];
    }
}

// Q[sa, sb] = sqrt(max over the block |D|): one thread per shell pair
__global__ void k_schwarz_q(const double* D, long long nbf, int nshell, const int* foff, const int* L, double* Q) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nshell * nshell) return;
    const int sa = idx / nshell, sb = idx - sa * nshell;
    double m = 0.0;
    for (int i = 0; i < 2 * L[sa] + 1; ++i)
        for (int j = 0; j < 2 * L[sb] + 1; ++j) m = fmax(m, fabs(D[(long long)(foff[sa] + i) * nbf + foff[sb] + j]));
    Q[idx] = sqrt(m);
}
