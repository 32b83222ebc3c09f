"""Code generator: class-specialised CUDA kernels for the nuclear-attraction matrix ("coul").

    V_ab = sum_C q_C (a | 1/r_C | b),        spherical, normalize cgto|pgto

The reference evaluates every (a, b, C) block with the two-index Obara-Saika recurrence
(sympleints/defs/coulomb.py:18-121; python path) or VRR + HRR (graphs/defs/int_nucattr.py).  With ~10^3
point charges per shell pair that is the wrong shape for a GPU: the Cartesian two-index recursion
is repeated for every charge.  These kernels use the McMurchie-Davidson factorisation instead, which
moves everything that depends on (a, b) out of the charge loop:

    (a|1/r_C|b) = (2 pi / p) K_ab  sum_{tuv} E^{ab}_{tuv} R_{tuv}(p, P - C)

  * per (primitive pair, charge): Hermite integrals R^0_{tuv}, t+u+v <= La+Lb, by the one-centre recursion
        R^N_{t+1,u,v} = t R^{N+1}_{t-1,u,v} + (P-C)_x R^{N+1}_{t,u,v},   R^N_000 = (-2p)^N F_N(p |PC|^2)
    (2 flops per element), accumulated over the charges IN REGISTERS: Rbar_{tuv} = sum_C q_C R^0_{tuv};
  * per primitive pair: E^{ab}_{tuv} = Ex[ax][bx][t] Ey[ay][by][u] Ez[az][bz][v] from the 1-D two-index
    recurrences, contracted with Rbar into the Cartesian block;
  * per shell pair: C2S(b), C2S(a) with the lmn factors folded in, store + mirror.

No horizontal recurrence is involved, so the result has none of the (p/a)^Lb amplification that makes
VRR+HRR fail the parity criterion for very unequal exponents.  It agrees with the reference's two-index
recurrence to rounding (tests/test_generated_kernels_emulated.py, tests/test_gpu_batched.py).

Execution model as in gen3c2e.py: EPT lanes of a warp per shell pair (32/EPT pairs per warp); each
tuple loops over its primitive pairs and, inside, over ALL charges.
"""
from __future__ import annotations

import json as _json
import math
import os as _os

from .gen3c2e import cart2sph, cart_order, cidx, lit, lmn_factors, ncart, nsph, off, pow2ceil

_TUNE_FILE = _os.path.join(_os.path.dirname(__file__), "tuning_coul.json")
TUNING = {}
if _os.path.exists(_TUNE_FILE) and not _os.environ.get("SI_GC_NO_TUNING"):
    TUNING = _json.load(open(_TUNE_FILE))
DEFAULT_MB = int(_os.environ.get("SI_GC_MB", "1"))


class CoulGen:
    def __init__(self, La, Lb, ept=None, mb=None):
        assert La >= Lb
        tune = TUNING.get(f"{La}{Lb}", {})
        if ept is None:
            ept = tune.get("ept")
        if mb is None:
            mb = tune.get("mb", DEFAULT_MB)
        self.MB = mb
        self.La, self.Lb = La, Lb
        self.L = L = La + Lb
        self.NA = off(L + 1)
        self.nca, self.ncb = ncart(La), ncart(Lb)
        self.nsa, self.nsb = nsph(La), nsph(Lb)
        if ept is None:
            ept = min(32, pow2ceil(max(1, (ncart(L) + 1) // 2)))
            sh = int(_os.environ.get("SI_GC_EPT_SHIFT", "0"))
            if sh < 0:
                ept = max(1, ept >> (-sh))
            elif sh > 0:
                ept = min(32, ept << sh)
        self.EPT = ept
        self.TPW = 32 // ept
        self.name = f"{La}{Lb}"
        self.lines = []
        self.ind = 0
        self._layout()

    def _layout(self):
        L, La, Lb = self.L, self.La, self.Lb
        o = 0
        # (1) charge loop: Hermite recursion arrays V(n, jn), shell n at auxiliary order jn (jn = 0 is never
        # stored for n >= 1)
        self.V_OFF = {}
        for n in range(L + 1):
            for jn in range(L - n + 1):
                if n >= 1 and jn == 0:
                    continue
                self.V_OFF[(n, jn)] = o
                o += ncart(n)
        self.SX_OFF = o          # (P - C), two buffers by charge parity (lanes whose direction is not uniform)
        o += 6
        end1 = o
        # (2) after the charge loop, over the dead recursion arrays: Rbar (flat over shells 0..L) and the
        # E tables [d][i][j][t]
        o = 0
        self.RB_OFF = o
        o += self.NA
        self.TS = L + 1
        self.E_OFF = o
        self.ESZ = (La + 1) * (Lb + 1) * self.TS
        o += 3 * self.ESZ
        end2 = o
        # (3) after the primitive loop: Cartesian block [ai][bi], then [ai][jb] after C2S(b)
        o = 0
        self.XC_OFF = o
        o += self.nca * self.ncb
        self.Z_OFF = o
        o += self.nca * self.nsb
        end3 = o
        o = max(end1, end2, end3)
        # tuple stride = odd multiple of EPT modulo 16 (see gen3c2e.ClassGen._layout)
        e = self.EPT if self.EPT <= 8 else 1    # (EPT >= 16: the tuples of a warp sit in different half-warps)
        while o % (2 * e) != e:
            o += 1
        self.WSZ = o

    def emit(self, s=""):
        self.lines.append("    " * self.ind + s)

    def open(self, s):
        self.emit(s + " {")
        self.ind += 1

    def close(self, s="}"):
        self.ind -= 1
        self.emit(s)

    def nslots(self, count):
        return (count + self.EPT - 1) // self.EPT

    def sync(self, comment=""):
        if self.EPT > 1:
            self.emit("si_g3_syncwarp();" + (f"   // {comment}" if comment else ""))

    def vrr_meta(self):
        tab = {}
        for n in range(1, self.L + 1):
            row = []
            for a in cart_order(n):
                d = next(i for i in range(3) if a[i] > 0)
                am = list(a)
                am[d] -= 1
                k1 = cidx(am)
                fac = a[d] - 1
                k2 = 0
                if fac > 0:
                    amm = list(am)
                    amm[d] -= 1
                    k2 = cidx(amm)
                row.append((d, k1, k2, fac))
            tab[n] = row
        return tab

    def generate(self):
        La, Lb, L = self.La, self.Lb, self.L
        EPT, TPW = self.EPT, self.TPW
        name = self.name
        E = self.emit
        self.lines = []
        self.ind = 0
        vm = self.vrr_meta()
        for n in range(1, L + 1):
            vals = [d | (fac << 2) | (k1 << 8) | (k2 << 16) for (d, k1, k2, fac) in vm[n]]
            E(f"__device__ const int gc_vrr_{name}_{n}[{len(vals)}] = {{{', '.join(map(str, vals))}}};")
        # Cartesian component pairs: ax | ay<<4 | az<<8 | bx<<12 | by<<16 | bz<<20
        vals = []
        for a in cart_order(La):
            for b in cart_order(Lb):
                vals.append(a[0] | (a[1] << 4) | (a[2] << 8) | (b[0] << 12) | (b[1] << 16) | (b[2] << 20))
        nab = len(vals)
        E(f"__device__ const int gc_ab_{name}[{nab}] = {{{', '.join(map(str, vals))}}};")
        E("")
        E(f"// coul class ({La}{Lb}): EPT={EPT} lanes per shell pair, {TPW} pairs per warp, {self.WSZ} doubles of shared memory per pair")
        E(f"SI_G3_DEV void gc_body_{name}(const GCArgs& A, const int lane, double* wbase, const long long g_first, const long long g_stride) {{")
        self.ind = 1
        E(f"constexpr int EPT = {EPT}, TPW = {TPW}, WSZ = {self.WSZ};")
        E("const int q = lane % EPT, tl = lane / EPT;")
        E("SI_G3_WTYPE W = SI_G3_WINIT(wbase + tl * WSZ, WSZ);")
        E("(void)q;")
        nsl_v = {n: self.nslots(ncart(n)) for n in range(1, L + 1)}
        for n in range(1, L + 1):
            for s in range(nsl_v[n]):
                E(f"const int kc_{n}_{s} = min(q + EPT * {s}, {ncart(n) - 1}); const int vm_{n}_{s} = gc_vrr_{name}_{n}[kc_{n}_{s}];")
        nsl_b = self.nslots(L + 1)
        nsl_ab = self.nslots(nab)
        E("")
        E("// static round-robin over the warps of a one-wave grid (see gen3c2e.py)")
        self.open("for (long long grp = g_first; grp < A.ngroups; grp += g_stride)")
        E("// group -> (pair group, charge split): the charges are divided over A.nsplit tuples per pair, each")
        E("// writing its partial matrix slice; the slices are summed in a fixed order afterwards")
        E("const int pg = (int)(grp / A.nsplit), sp = (int)(grp - (long long)pg * A.nsplit);")
        E("const int c0 = (int)((long long)A.ncharge * sp / A.nsplit), c1 = (int)((long long)A.ncharge * (sp + 1) / A.nsplit);")
        E("double* const outp = sp == 0 ? A.od.out : A.part + (long long)(sp - 1) * A.od.ld * A.od.ld;")
        E("int ip = pg * TPW + tl;")
        E("const bool valid = ip < A.npairs;")
        E("if (!valid) ip = A.npairs - 1;")
        E("const SiPairItem it = A.pairs[ip];")
        E("const int npp = it.Ka * it.Kb;")
        E("const int npp_max = si_g3_warp_max(npp);")
        E("double " + ", ".join(f"xc_{s} = 0.0" for s in range(nsl_ab)) + ";   // Cartesian block (registers)")
        self.open("for (int ipp = 0; ipp < npp_max; ++ipp)")
        E("// pairs with fewer primitive pairs repeat their last one with zero weight")
        E("const int ippc = ipp < npp ? ipp : npp - 1;")
        E("const double* pp = A.ppair + (size_t)(it.ppoff + ippc) * SI_PP_STRIDE;")
        E("const double p = pp[0], oop = pp[1];")
        E("const double P0 = pp[2], P1 = pp[3], P2 = pp[4];")
        E("const double pref = ipp < npp ? 2.0 * SI_PI * oop * pp[8] : 0.0;")
        # (-2p)^j per lane
        for s in range(nsl_b):
            E(f"const int jb_{s} = min(q + EPT * {s}, {L}); double mp_{s} = 1.0; "
              f"{{ double bb = -2.0 * p; int e_ = jb_{s}; for (int i_ = 0; i_ < 4; ++i_) {{ if (e_ & 1) mp_{s} *= bb; bb *= bb; e_ >>= 1; }} }}")
        for s in range(nsl_b):
            E(f"const double xp_{s} = A.boys.xlarge_pref[jb_{s}];   // asymptotic Boys prefactor of this lane's order")
        # accumulators over the charges
        E("double yr_0 = 0.0;")
        for n in range(1, L + 1):
            E("double " + ", ".join(f"yr_{n}_{s} = 0.0" for s in range(nsl_v[n])) + ";")
        # which (shell, slot) have lanes of different recursion directions?
        mixed = {}
        for n in range(1, L + 1):
            for s in range(nsl_v[n]):
                ds = {vm[n][min(q_ + EPT * s, ncart(n) - 1)][0] for q_ in range(EPT)}
                mixed[(n, s)] = None if len(ds) > 1 else ds.pop()
        need_sx = any(v is None for v in mixed.values())
        E("// charges are packed (x, y, z, q); the next record is loaded one iteration ahead")
        E("double cn0 = A.c4[4 * c0], cn1 = A.c4[4 * c0 + 1], cn2 = A.c4[4 * c0 + 2], cn3 = A.c4[4 * c0 + 3];")
        self.open("for (int ic = c0; ic < c1; ++ic)")
        E("const double X0 = P0 - cn0, X1 = P1 - cn1, X2 = P2 - cn2, qc = cn3;")
        E("{ const double* cc_ = A.c4 + 4 * min(ic + 1, c1 - 1); cn0 = cc_[0]; cn1 = cc_[1]; cn2 = cc_[2]; cn3 = cc_[3]; }")
        E("const double T = p * (X0 * X0 + X1 * X1 + X2 * X2);")
        # base values: V(0, j) = qc (-2p)^j F_j(T)
        # (one reciprocal square root per charge for all orders of the lane in the asymptotic branch)
        E("double " + ", ".join(f"bv_{s}" for s in range(nsl_b)) + ";")
        self.open("if (T < 30.0)")
        for s in range(nsl_b):
            E(f"bv_{s} = qc * mp_{s} * si_boys_taylor(A.boys, jb_{s}, T);")
        self.close()
        self.open("else")
        E("const double rs_ = si_rsqrt(T);")
        for s in range(nsl_b):
            E(f"bv_{s} = qc * mp_{s} * xp_{s} * si_boys_xpow_rs(jb_{s}, rs_);")
        self.close()
        E("yr_0 += bv_0;   // lane q == 0 holds j = 0")
        if L >= 1:
            for s in range(nsl_b):
                E(f"W[{self.V_OFF[(0, 0)]} + jb_{s}] = bv_{s};")
            if need_sx:
                E(f"const int sx_ = {self.SX_OFF} + 3 * (ic & 1);")
                E("if (q == 0) { W[sx_] = X0; W[sx_ + 1] = X1; W[sx_ + 2] = X2; }")
            self.sync()
        for n in range(1, L + 1):
            njn = L - n + 1          # orders jn = 0..njn-1 of this shell
            self.open("")
            for s in range(nsl_v[n]):
                E(f"const int m{s}_ = vm_{n}_{s}; const int d{s}_ = m{s}_ & 3; const int k1_{s} = (m{s}_ >> 8) & 0xFF; const int k2_{s} = (m{s}_ >> 16) & 0xFF;")
                if mixed[(n, s)] is None:
                    E(f"const double cp{s}_ = W[sx_ + d{s}_];   // lanes of this slot differ in direction")
                else:
                    E(f"const double cp{s}_ = X{mixed[(n, s)]};   // every lane of this slot recurs along the same direction")
                for jn in range(1, njn + 1):
                    E(f"const double u{s}_{jn} = W[{self.V_OFF[(n - 1, jn)]} + k1_{s}];")
                    if n >= 2:
                        E(f"const double w{s}_{jn} = W[{self.V_OFF[(n - 2, jn)]} + k2_{s}];")
            for s in range(nsl_v[n]):
                if n >= 2:
                    E(f"const double fc{s}_ = (double)((m{s}_ >> 2) & 0xF);")
                for jn in range(njn):
                    if n >= 2:
                        E(f"const double v{s}_{jn} = fma(fc{s}_, w{s}_{jn + 1}, cp{s}_ * u{s}_{jn + 1});")
                    else:
                        E(f"const double v{s}_{jn} = cp{s}_ * u{s}_{jn + 1};")
                E(f"yr_{n}_{s} += v{s}_0;")
            for s in range(nsl_v[n]):
                for jn in range(1, njn):
                    E(f"W[{self.V_OFF[(n, jn)]} + kc_{n}_{s}] = v{s}_{jn};")
            self.close()
            if n < L:
                self.sync()
        if 1 <= L <= 2:
            self.sync("the next charge overwrites V(0, .)")
        self.close()  # charges
        # ---- Rbar to shared memory, E tables (both overlay the recursion arrays)
        if L >= 3:
            self.sync("last charge's readers of the recursion arrays are done")
        E(f"if (q == 0) W[{self.RB_OFF}] = yr_0;")
        for n in range(1, L + 1):
            for s in range(nsl_v[n]):
                partial = (s + 1) * EPT > ncart(n)
                E((f"if (q + EPT * {s} < {ncart(n)}) " if partial else "") + f"W[{self.RB_OFF + off(n)} + kc_{n}_{s}] = yr_{n}_{s};")
        if L >= 1:
            E("// 1-D expansion coefficients E[d][i][j][t] of x_A^i x_B^j in Hermite Gaussians about P (one lane per direction)")
            self.open("for (int d = q; d < 3; d += EPT)")
            E("const double pa = pp[5 + d], pb = pa + it.AB[d], h = 0.5 * oop;")
            E(f"const int eb = {self.E_OFF} + d * {self.ESZ};")
            E(f"W[eb] = 1.0;")
            TS = self.TS
            jst = TS
            ist = (Lb + 1) * TS
            if La >= 1:
                self.open(f"for (int i = 0; i < {La}; ++i)")
                self.open("for (int t = 0; t <= i + 1; ++t)")
                E(f"const int e0 = eb + i * {ist} + t;")
                E("double v = 0.0;")
                E("if (t >= 1) v = h * W[e0 - 1];")
                E("if (t <= i) v = fma(pa, W[e0], v);")
                E("if (t + 1 <= i) v = fma((double)(t + 1), W[e0 + 1], v);")
                E(f"W[e0 + {ist}] = v;")
                self.close()
                self.close()
            if Lb >= 1:
                self.open(f"for (int j = 0; j < {Lb}; ++j)")
                self.open(f"for (int i = 0; i <= {La}; ++i)")
                self.open("for (int t = 0; t <= i + j + 1; ++t)")
                E(f"const int e0 = eb + i * {ist} + j * {jst} + t;")
                E("double v = 0.0;")
                E("if (t >= 1) v = h * W[e0 - 1];")
                E("if (t <= i + j) v = fma(pb, W[e0], v);")
                E("if (t + 1 <= i + j) v = fma((double)(t + 1), W[e0 + 1], v);")
                E(f"W[e0 + {jst}] = v;")
                self.close()
                self.close()
                self.close()
            self.close()
        self.sync()
        # ---- contraction with E
        for s in range(nsl_ab):
            self.open("")
            E(f"const int ab = gc_ab_{name}[min(q + EPT * {s}, {nab - 1})];")
            if L == 0:
                E(f"xc_{s} = fma(pref, W[{self.RB_OFF}], xc_{s});")
            else:
                E("const int ax = ab & 15, ay = (ab >> 4) & 15, az = (ab >> 8) & 15, bx = (ab >> 12) & 15, by = (ab >> 16) & 15, bz = (ab >> 20) & 15;")
                ist = (Lb + 1) * self.TS
                jst = self.TS
                E(f"const int ex = {self.E_OFF} + ax * {ist} + bx * {jst};")
                E(f"const int ey = {self.E_OFF + self.ESZ} + ay * {ist} + by * {jst};")
                E(f"const int ez = {self.E_OFF + 2 * self.ESZ} + az * {ist} + bz * {jst};")
                E("double sx = 0.0;")
                self.open("for (int t = 0; t <= ax + bx; ++t)")
                E("double sy = 0.0;")
                self.open("for (int u = 0; u <= ay + by; ++u)")
                E("double sz = 0.0;")
                self.open("for (int v = 0; v <= az + bz; ++v)")
                E("const int n_ = t + u + v, i_ = u + v;")
                E(f"sz = fma(W[ez + v], W[{self.RB_OFF} + n_ * (n_ + 1) * (n_ + 2) / 6 + i_ * (i_ + 1) / 2 + v], sz);")
                self.close()
                E("sy = fma(W[ey + u], sz, sy);")
                self.close()
                E("sx = fma(W[ex + t], sy, sx);")
                self.close()
                E(f"xc_{s} = fma(pref, sx, xc_{s});")
            self.close()
        self.sync("Rbar and the E tables are rewritten by the next primitive pair")
        self.close()  # primitive pairs
        # ---- C2S
        for s in range(nsl_ab):
            partial = (s + 1) * EPT > nab
            E((f"if (q + EPT * {s} < {nab}) " if partial else "") + f"W[{self.XC_OFF} + q + EPT * {s}] = xc_{s};")
        self.sync()
        Ma, Mb = cart2sph(La), cart2sph(Lb)
        fa, fb = lmn_factors(La), lmn_factors(Lb)
        nca, ncb, nsa, nsb = self.nca, self.ncb, self.nsa, self.nsb
        if Lb > 0:
            self.open(f"for (int ai = q; ai < {nca}; ai += EPT)")
            for b in range(ncb):
                E(f"const double x{b} = W[{self.XC_OFF} + ai * {ncb} + {b}];")
            for jb in range(nsb):
                expr = None
                for b in range(ncb):
                    if Mb[jb][b] == 0.0:
                        continue
                    coef = Mb[jb][b] * fb[b]
                    expr = f"{lit(coef)} * x{b}" if expr is None else f"fma({lit(coef)}, x{b}, {expr})"
                E(f"W[{self.Z_OFF} + ai * {nsb} + {jb}] = {expr};")
            self.close()
            self.sync()
            zoff = self.Z_OFF
        else:
            zoff = self.XC_OFF
        self.open(f"for (int jb = q; jb < {nsb}; jb += EPT)")
        for a in range(nca):
            E(f"const double z{a} = W[{zoff} + {a * nsb} + jb];")
        for ia in range(nsa):
            expr = None
            for a in range(nca):
                if Ma[ia][a] == 0.0:
                    continue
                coef = Ma[ia][a] * fa[a]
                expr = f"{lit(coef)} * z{a}" if expr is None else f"fma({lit(coef)}, z{a}, {expr})"
            E(f"const double o{ia} = {expr};")
        self.open("if (valid)")
        for ia in range(nsa):
            E(f"outp[(long long)(it.fa0 + {ia}) * A.od.ld + it.fb0 + jb] = o{ia};")
        self.open("if (it.sa != it.sb)")
        for ia in range(nsa):
            E(f"outp[(long long)(it.fb0 + jb) * A.od.ld + it.fa0 + {ia}] = o{ia};")
        self.close()
        self.close()
        self.close()
        self.sync()
        self.close()  # groups
        self.close()  # function
        E("")
        E(f"SI_G3_GLOBAL({self.MB}) void gc_kernel_{name}(GCArgs A) {{")
        E("    SI_G3_KERNEL_PROLOGUE")
        E(f"    gc_body_{name}(A, g3_lane, g3_smem + (size_t)g3_warp * {self.WSZ * TPW}, g3_first, g3_stride);")
        E("}")
        E("")
        return "\n".join(self.lines)


HEADER = """// GENERATED by sympleints_b200/codegen/gencoul.py -- DO NOT EDIT.
#include "../si_g3.h"
"""


def all_classes(lmax=4):
    return [(a, b) for a in range(lmax + 1) for b in range(a + 1)]


def write_sources(outdir, lmax=4, nparts=4):
    """Write generated/gc_part_XX.cu and generated/gc_table.inc.  Returns the list of .cu files."""
    _os.makedirs(outdir, exist_ok=True)
    classes = all_classes(lmax)
    gens = {c: CoulGen(*c) for c in classes}
    parts = [[] for _ in range(nparts)]
    load = [0] * nparts
    for c in sorted(classes, key=lambda c: -off(c[0] + c[1] + 1)):
        i = load.index(min(load))
        parts[i].append(c)
        load[i] += off(c[0] + c[1] + 1)
    files, table = [], []
    for pi, cl in enumerate(parts):
        if not cl:
            continue
        src = [HEADER]
        for c in cl:
            src.append(gens[c].generate())
        src.append("#if !defined(SI_G3_HOST_EMU)")
        src.append(f"int gc_launch_part_{pi}(int id, const GCArgs& A, int grid, int block, size_t smem, int max_smem, cudaStream_t s) {{")
        src.append("    switch (id) {")
        for c in cl:
            g = gens[c]
            cid = c[0] * 10 + c[1]
            src.append(f"        case {cid}: {{ int occ = 0; "
                       f"if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gc_kernel_{g.name}, block, smem) != cudaSuccess || occ < 1) occ = 1; "
                       f"gc_kernel_{g.name}<<<si_g3_grid(grid, occ), block, smem, s>>>(A); return 0; }}")
            table.append((cid, pi, g.EPT, g.TPW, g.WSZ))
        src.append("        default: return 1;")
        src.append("    }")
        src.append("}")
        # the dynamic shared-memory limit is a per-device function attribute: set for every kernel of the part
        # once per context (si_init), not behind a process-wide flag
        src.append(f"int gc_attr_part_{pi}(int max_smem) {{")
        for c in cl:
            src.append(f"    if (cudaFuncSetAttribute(gc_kernel_{gens[c].name}, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) return 1;")
        src.append("    return 0;")
        src.append("}")
        src.append("#endif")
        fn = _os.path.join(outdir, f"gc_part_{pi:02d}.cu")
        text = "\n".join(src) + "\n"
        if not _os.path.exists(fn) or open(fn).read() != text:
            open(fn, "w").write(text)
        files.append(fn)
    lines = ["// GENERATED -- class table of the specialised coul kernels: {id, part, EPT, TPW, WSZ}"]
    lines.append("static const int gc_table[][5] = {")
    for t in sorted(table):
        lines.append("    {%d, %d, %d, %d, %d}," % t)
    lines.append("};")
    for pi, cl in enumerate(parts):
        if cl:
            lines.append(f"int gc_launch_part_{pi}(int id, const GCArgs& A, int grid, int block, size_t smem, int max_smem, cudaStream_t s);")
    for pi, cl in enumerate(parts):
        if cl:
            lines.append(f"int gc_attr_part_{pi}(int max_smem);")
    lines.append("static int gc_set_attributes(int max_smem) {")
    for pi, cl in enumerate(parts):
        if cl:
            lines.append(f"    if (gc_attr_part_{pi}(max_smem)) return 1;")
    lines.append("    return 0;")
    lines.append("}")
    lines.append("static int gc_launch(int part, int id, const GCArgs& A, int grid, int block, size_t smem, int max_smem, cudaStream_t s) {")
    lines.append("    switch (part) {")
    for pi, cl in enumerate(parts):
        if cl:
            lines.append(f"        case {pi}: return gc_launch_part_{pi}(id, A, grid, block, smem, max_smem, s);")
    lines.append("        default: return 1;")
    lines.append("    }")
    lines.append("}")
    fn = _os.path.join(outdir, "gc_table.inc")
    text = "\n".join(lines) + "\n"
    if not _os.path.exists(fn) or open(fn).read() != text:
        open(fn, "w").write(text)
    return files


def write_emu_source(path, classes):
    """TEST-ONLY: one translation unit with the bodies of `classes`, for tests/emu/gc_emu.cpp."""
    src = ["#define SI_G3_HOST_EMU 1", HEADER.replace("../si_g3.h", "../../sympleints_b200/csrc/si_g3.h")]
    for c in classes:
        src.append(CoulGen(*c).generate())
    src.append("struct GCEmuClass { int id, ept, tpw, wsz; void (*body)(const GCArgs&, int, double*, long long, long long); };")
    src.append("static const GCEmuClass gc_emu_classes[] = {")
    for c in classes:
        g = CoulGen(*c)
        src.append(f"    {{{c[0] * 10 + c[1]}, {g.EPT}, {g.TPW}, {g.WSZ}, gc_body_{g.name}}},")
    src.append("};")
    src.append(f"static const int gc_emu_nclasses = {len(classes)};")
    text = "\n".join(src) + "\n"
    if not _os.path.exists(path) or open(path).read() != text:
        open(path, "w").write(text)
