"""Code generator: class-specialised CUDA kernels for the 3c2e_sph tensor (ab|c).

This is the CUDA counterpart of the reference's generators (sympleints/main.py:206-323 for the
sympy path, sympleints/graphs/generate.py + graphs/templates/int_3c2e_func.tpl for the graph
path): for every class (La >= Lb, Lc) it emits straight-line device code for

    Boys -> VRR(a, up to La+Lb) -> truncated aux VRR(c) -> contraction
         -> C2S(c) -> HRR(b) + C2S(b) -> C2S(a) -> store

(defs/coulomb.py:218-341; stage order of graphs/defs/int_3c2e.py:7-43).

Execution model: EPT lanes of a warp co-operate on one shell tuple (32/EPT tuples per warp,
EPT = 1 for the smallest classes = one thread per tuple, EPT = 32 for the largest = one warp per
tuple).  Intermediates live in the tuple's shared-memory region; every shared-memory address is
`immediate + per-lane register`, the per-lane registers (neighbour offsets a-1_i, factors a_i) are
loaded once per kernel from small per-class tables.  All recurrence coefficients are per-tuple
uniform registers.  The horizontal recurrence is evaluated per target a-component entirely in
registers ("tile"), followed directly by C2S(b).
"""
from __future__ import annotations

import math
from fractions import Fraction

import json as _json
import os as _os

# per-class tuning (lanes per tuple, min CTAs/SM for __launch_bounds__) found by scripts/tune_g3.py
DEFAULT_MB = int(_os.environ.get("SI_G3_MB", "1"))
YREG_MAX = int(_os.environ.get("SI_G3_YREG", "36"))   # max doubles of register accumulators per lane
SKIP = set(filter(None, _os.environ.get("SI_G3_SKIP", "").split(",")))   # ablation timing only
_TUNE_FILE = _os.path.join(_os.path.dirname(__file__), "tuning.json")
TUNING = {}
if _os.path.exists(_TUNE_FILE) and not _os.environ.get("SI_G3_NO_TUNING"):
    TUNING = _json.load(open(_TUNE_FILE))
if _os.environ.get("SI_G3_EPT_SHIFT"):
    pass



# cache of the layout / HRR-item-order search (ClassGen._layout): deterministic, but a few seconds per class
_HRR_FILE = _os.path.join(_os.path.dirname(__file__), "hrr_orders.json")
_HRR_CACHE = {}
if _os.path.exists(_HRR_FILE) and not _os.environ.get("SI_G3_NO_HRR_CACHE"):
    try:
        _HRR_CACHE = _json.load(open(_HRR_FILE))
    except Exception:
        _HRR_CACHE = {}


def _save_hrr_cache():
    try:
        tmp = _HRR_FILE + ".tmp%d" % _os.getpid()
        _json.dump(_HRR_CACHE, open(tmp, "w"), sort_keys=True)
        _os.replace(tmp, _HRR_FILE)
    except Exception:
        pass


# ----------------------------------------------------------------------------------------
# small helpers (orderings identical to sympleints/helpers.py:42-49)
# ----------------------------------------------------------------------------------------
def ncart(l):
    return (l + 1) * (l + 2) // 2


def nsph(l):
    return 2 * l + 1


def cart_order(L):
    out = []
    for i in range(L + 1):
        l = L - i
        for n in range(i + 1):
            out.append((l, i - n, n))
    return out


def cidx(a):
    """Index of Cartesian component a within its shell."""
    L = sum(a)
    i = L - a[0]
    return i * (i + 1) // 2 + a[2]


def off(n):
    """Flat offset of shell n when shells 0..n-1 are concatenated."""
    return sum(ncart(j) for j in range(n))


def fact2(n):
    r = 1
    while n > 1:
        r *= n
        n -= 2
    return r


def _c2s_complex(lx, ly, lz, m):
    l = lx + ly + lz
    am = abs(m)
    j2 = lx + ly - am
    if j2 < 0 or j2 % 2:
        return 0.0
    j = j2 // 2
    lmm = l - am
    sign = 1 if m >= 0 else -1
    tot = 0
    for i in range(lmm // 2 + 1):
        if j > i:
            continue
        ifact = Fraction(math.comb(l, i) * math.comb(i, j) * (-1) ** i * math.factorial(2 * l - 2 * i),
                         math.factorial(lmm - 2 * i))
        ks = 0
        for k in range(j + 1):
            if lx - 2 * k < 0 or lx - 2 * k > am:
                continue
            e2 = sign * (am - lx + 2 * k)
            ks += (1j ** (e2 % 4)) * math.comb(j, k) * math.comb(am, lx - 2 * k)
        tot += complex(ks) * float(ifact)
    rad = Fraction(math.factorial(2 * lx) * math.factorial(2 * ly) * math.factorial(2 * lz) * math.factorial(l)
                   * math.factorial(lmm),
                   math.factorial(2 * l) * math.factorial(lx) * math.factorial(ly) * math.factorial(lz)
                   * math.factorial(l + am))
    return tot * math.sqrt(float(rad)) / (2**l * math.factorial(l))


_C2S_CACHE = {}


def cart2sph(l):
    """Real C2S matrix rows m=-l..l (sympleints/cart2sph.py:163-202), list of lists."""
    if l in _C2S_CACHE:
        return _C2S_CACHE[l]
    comps = cart_order(l)
    M = [[0.0] * len(comps) for _ in range(2 * l + 1)]
    for ci, (lx, ly, lz) in enumerate(comps):
        for mi, m in enumerate(range(-l, l + 1)):
            c = _c2s_complex(lx, ly, lz, m)
            if m != 0:
                cm = _c2s_complex(lx, ly, lz, -m)
                if m < 0:
                    c = (c + cm) / math.sqrt(2)
                else:
                    c = (c - cm) / (1j * math.sqrt(2))
            v = complex(c).real
            assert abs(complex(c).imag) < 1e-13
            M[mi][ci] = 0.0 if abs(v) <= 1e-14 else v
    _C2S_CACHE[l] = M
    return M


def lmn_factors(l):
    return [1.0 / math.sqrt(fact2(2 * a - 1) * fact2(2 * b - 1) * fact2(2 * c - 1)) for (a, b, c) in cart_order(l)]


def lit(x):
    """C++ double literal with full precision."""
    s = repr(float(x))
    if "e" not in s and "." not in s and "inf" not in s:
        s += ".0"
    return s


def pow2ceil(n):
    p = 1
    while p < n:
        p *= 2
    return p


# ----------------------------------------------------------------------------------------
class ClassGen:
    """Emit the device function + kernel for one class."""

    def __init__(self, La, Lb, Lc, ept=None, mb=None):
        assert La >= Lb
        tune = TUNING.get(f"{La}{Lb}{Lc}", {})
        if ept is None:
            ept = tune.get("ept")
        if mb is None:
            mb = tune.get("mb", DEFAULT_MB)
        self.MB = mb
        self.La, self.Lb, self.Lc = La, Lb, Lc
        self.L = L = La + Lb
        self.NA = off(L + 1)                 # flat a size, shells 0..L
        self.n_ap = self.NA - off(La)        # target a' (shells La..L)
        self.ncc = ncart(Lc)
        self.nm = nsph(Lc)
        self.nca, self.ncb = ncart(La), ncart(Lb)
        self.nsa, self.nsb = nsph(La), nsph(Lb)
        if ept is None:
            ept = min(32, pow2ceil(max(1, (self.n_ap + 1) // 2)))
            sh = int(_os.environ.get("SI_G3_EPT_SHIFT", "0"))
            if sh < 0:
                ept = max(1, ept >> (-sh))
            elif sh > 0:
                ept = min(32, ept << sh)
        self.EPT = ept
        self.TPW = 32 // ept
        self.name = f"{La}{Lb}{Lc}"
        # spherical accumulators Y[m][a'] in registers (per lane: nm x target slots) when they fit
        tslots = [s for s in range((self.NA + ept - 1) // ept) if (s + 1) * ept - 1 >= off(La) and s * ept < self.NA]
        self.tslots = tslots
        ymax = tune.get("yreg", YREG_MAX)
        self.YREG = Lc > 0 and self.nm * len(tslots) <= ymax
        self.lines = []
        self.ind = 1
        # Lane-private start of the VRR ("chain"): shells 1..N0 are built by the lanes that own the components of shell
        # N0 (the largest shell that fits one slot), each along its own path z..z y..y x..x from shell 0, all Boys orders
        # in registers -- the same FP64 instruction count as one staged slot per shell, but no shared-memory round trips
        # and no barriers in between.  Lower shells are stored only where a later stage reads them.
        n0 = 0
        if _os.environ.get("SI_G3_CHAIN", "1") == "1" and _os.environ.get("SI_G3_PASEL", "1") == "1" and "vrr" not in SKIP:
            for n in range(1, L + 1):
                if ncart(n) <= ept:
                    n0 = n
        self.N0 = n0 if n0 >= 2 else 0
        self._layout()

    def chain_stores(self):
        """{shell n < N0: number of orders jn to store} for the lower shells of the chain (empty entries left out)."""
        L, n0 = self.L, self.N0
        need_lo = self.amin(0) if self.Lc > 0 else self.La     # lowest shell whose order-0 column is read later
        out = {}
        for n in range(1, n0):
            cnt = 0
            if n >= need_lo:
                cnt = 1
            if n == n0 - 1 and L > n0:
                cnt = L - n0 + 1          # the (n-2) operand of the first staged shell N0+1, orders 0..L-N0
            if cnt:
                out[n] = cnt
        return out

    def has_table_variant(self):
        """Classes with >= 4 tuples per warp also get a kernel that takes its tuples from a host-built table."""
        return self.TPW >= 4

    # ---- shared-memory layout (doubles, per tuple) ------------------------------------
    def amin(self, level):
        return max(0, self.La - (self.Lc - level))

    def _layout(self, ypad=None, zpad=None):
        """Shared-memory layout.  The strides of the Y rows and Z slabs are padded such that the
        strided accesses of the contracted phase (lanes differ in (component, m)) spread over the
        banks; the paddings are found by simulating the half-warp bank pattern (_conflicts)."""
        if ypad is None:
            # the search result is cached (per class and lanes-per-tuple) in codegen/hrr_orders.json
            key = f"{self.La}{self.Lb}{self.Lc}_e{self.EPT}_y{int(self.YREG)}"
            hit = _HRR_CACHE.get(key)
            if hit is not None:
                self._layout(hit["yp"], hit["zp"])
                self.hrr_order = [tuple(x) for x in hit["order"]] if hit["order"] is not None else None
                self.hrr_cost, self.hrr_cost_natural, self.hrr_cost_ideal = hit.get("cost"), hit.get("natural"), hit.get("ideal")
                return
            best = None
            for yp in range(0, 8):
                for zp in (range(0, 4) if self.Lb > 0 else [0]):
                    self._layout(yp, zp)
                    self.hrr_order = None
                    order, c = self._order_hrr_items(descent=False)   # constructive orders only: cheap
                    c = (c + self._conflicts_c2sa(), self.WSZ)
                    if best is None or c < best[0]:
                        best = (c, yp, zp, order)
            self._layout(best[1], best[2])
            self.hrr_order, self.hrr_cost = self._order_hrr_items(descent=True)
            _HRR_CACHE[key] = {"yp": best[1], "zp": best[2], "order": self.hrr_order, "cost": self.hrr_cost,
                               "natural": getattr(self, "hrr_cost_natural", None), "ideal": getattr(self, "hrr_cost_ideal", None)}
            _save_hrr_cache()
            return
        L, Lc = self.L, self.Lc
        o = 0
        self.S_OFF = o          # scalars: PA[3], CPC[3], BPOW[L+1]
        self.S_PA, self.S_CPC, self.S_BPOW = 0, 3, 6
        o += 6 + (L + 1)
        if o % 2:
            o += 1
        # per-item metadata of the HRR stage (packed ints, two per double), filled once per kernel
        self.HM_OFF = o
        if self.Lb > 0:
            o += (self.nca * self.nm + 1) // 2
        # Y[m][a'] persistent over primitives: the leaves of the aux recurrence are folded straight
        # into the spherical components (C2S on c), so no Cartesian [a0|c] block is ever stored
        self.YS = self.n_ap + ypad                 # stride between m rows of Y
        # Z (after HRR + C2S(b)) is laid out [a][jb][m], m fastest: the C2S(a) stage then reads consecutive
        # words per lane (its items are (jb, m), m fastest, which also keeps the global stores coalesced)
        self.ZS = self.YS if self.Lb == 0 else None
        self.ZA = None
        self.Y_OFF = o
        o += self.nm * self.YS
        prim_start = o
        # primitive temporaries.  With Lc > 0 the VRR runs on the scaled quantities binv^n V(n, jn, k) (the scale the aux
        # recurrence wants: binv folded into PA, CPC and the two (n-2) coefficients), and the lowest-order columns
        # V(n, 0, :) of all shells ARE the level-0 array of the aux recurrence (flat index off(n) + k): no copy stage.
        self.V_OFF = {}
        self.D_OFF = {}
        if Lc > 0:
            for jn in range(L + 1):
                self.V_OFF[(0, jn)] = o
                o += 1
            self.D_OFF[0] = o
            o += self.NA
            cs = self.chain_stores()
            for n in range(1, L + 1):
                self.V_OFF[(n, 0)] = self.D_OFF[0] + off(n)
                for jn in range(1, L - n + 1):
                    if n < self.N0 and jn >= cs.get(n, 0):
                        continue         # lives in the registers of the lane-private chain only
                    self.V_OFF[(n, jn)] = o
                    o += ncart(n)
        else:
            cs = self.chain_stores()
            for n in range(L + 1):
                for jn in range(L - n + 1):
                    if 1 <= n < self.N0 and jn >= cs.get(n, 0):
                        continue
                    self.V_OFF[(n, jn)] = o
                    o += ncart(n)
        # D-tree arrays per level 1..Lc-1 (non-leaf), flat index relative to off(amin(level))
        for lev in range(0, max(Lc, 1)):
            if lev in self.D_OFF:
                continue
            self.D_OFF[lev] = o          # full-size arrays: address of flat index f is D_OFF[lev] + f
            o += self.NA
        prim_end = o
        # contracted phase: Z[m][a][jb] over the (dead) primitive temporaries
        if self.Lb == 0:
            self.Z_OFF = self.Y_OFF     # Z == Y (n_ap == nca, nsb == 1)
            z_end = prim_start
        else:
            self.Z_OFF = prim_start
            # Z[a][jb][m]: the stride between target components a is padded so that the stores of one HRR round
            # (lanes = (a, m) items, one store per jb) spread over the banks
            self.ZA = self.nsb * self.nm + zpad
            z_end = prim_start + self.nca * self.ZA
        w = max(prim_end, z_end)
        # Several tuples share a warp when EPT < 32; a half-warp (one shared-memory wavefront of 8-byte
        # accesses, 16 banks) then holds 16/EPT tuples whose lanes touch ~EPT consecutive words each, so the
        # tuple stride must be an odd multiple of EPT modulo 16 for the tuples to tile the banks
        e = self.EPT if self.EPT <= 8 else 1    # (EPT >= 16: the tuples of a warp sit in different half-warps)
        while w % (2 * e) != e:
            w += 1
        self.WSZ = w

    def _half_warp_cost(self, addrs):
        """Shared-memory wavefronts of one 8-byte access of a half-warp: the largest number of DISTINCT addresses
        that fall into one of the 16 banks (equal addresses are a broadcast)."""
        banks = {}
        for a in addrs:
            if a is not None:
                banks.setdefault(a % 16, set()).add(a)
        return max((len(v) for v in banks.values()), default=0)

    def _hrr_loads(self):
        ks = [k for n in range(self.Lb + 1) for k in cart_order(n)]
        out = []
        for k in ks:
            ik = k[1] + k[2]
            out.append((ik, off(self.La + sum(k)) - off(self.La) + ik * (ik + 1) // 2 + k[2]))
        return out

    def _order_hrr_items(self, descent=True):
        """Order of the HRR work items (target component ai, spherical aux component m) over the lanes.

        Every item gathers T(Lb) inputs Y[m][a + k]; with the items in natural order (a fastest) the lanes of a
        half-warp span several rows of the component triangle and nearly every gather with k_y + k_z > 0 is 2-way
        bank-conflicted (measured: 46 % excess wavefronts in the gather, the largest single consumer of the
        shared-memory pipe for Lb >= 2).  Items of the SAME triangle row differ by a k-independent address offset,
        so a half-warp filled from one row (and m values whose row offsets tile the banks) is conflict-free for all
        k at once.  Found here by a greedy construction + pairwise-swap descent on the exact wavefront model
        (gather loads + Z stores), per class, at generation time."""
        if self.Lb == 0:
            return None, 0
        import random
        EPT, nm, nca, WSZ = self.EPT, self.nm, self.nca, self.WSZ
        loads = self._hrr_loads()
        row = {}
        for ai in range(nca):
            r = 0
            while (r + 1) * (r + 2) // 2 <= ai:
                r += 1
            row[ai] = r
        items = [(m, ai) for m in range(nm) for ai in range(nca)]
        nit = len(items)
        nround = (nit + EPT - 1) // EPT
        lanes_per_half = min(EPT, 16)
        tuples_per_half = 16 // lanes_per_half

        import numpy as np
        nld = len(loads)
        # address of every access of an item: T(Lb) gathers, then nsb stores (distinct items never share an address)
        A = np.empty((nit, nld + self.nsb), dtype=np.int64)
        for idx, (m, ai) in enumerate(items):
            for j, ld in enumerate(loads):
                A[idx, j] = self.Y_OFF + m * self.YS + ai + row[ai] * ld[0] + ld[1]
            for jb in range(self.nsb):
                A[idx, nld + jb] = self.Z_OFF + ai * self.ZA + jb * nm + m
        toff = (np.arange(tuples_per_half, dtype=np.int64) * WSZ)[:, None, None]
        eye = np.arange(16, dtype=np.int64)

        def half_cost(idxs):
            """Wavefronts of all accesses of one half-warp holding the items `idxs` (of each of its tuples)."""
            banks = ((A[idxs][None, :, :] + toff) % 16).reshape(-1, A.shape[1])        # (lanes, accesses)
            cnt = (banks[:, :, None] == eye).sum(axis=0)                                    # (accesses, 16)
            return int(cnt.max(axis=1).sum())

        def round_cost(chunk):
            return sum(half_cost(chunk[h0:h0 + lanes_per_half]) for h0 in range(0, len(chunk), lanes_per_half))

        def total(order):
            return sum(round_cost(order[r * EPT:(r + 1) * EPT]) for r in range(nround))

        pos = {it: i for i, it in enumerate(items)}
        natural = list(range(nit))                                  # a fastest (round-1 order)
        mfast = [pos[(m, ai)] for ai in range(nca) for m in range(nm)]
        rowm = sorted(range(nit), key=lambda i: (-row[items[i][1]], items[i][0], items[i][1]))
        rowa = sorted(range(nit), key=lambda i: (-row[items[i][1]], items[i][1], items[i][0]))
        cands = [natural, mfast, rowm, rowa]
        best = min(cands, key=total)
        bc = total(best)
        self.hrr_cost_natural = total(natural)
        nhalf = sum((min(nit - r * EPT, EPT) + lanes_per_half - 1) // lanes_per_half for r in range(nround))
        self.hrr_cost_ideal = nhalf * A.shape[1]
        if not descent or bc <= self.hrr_cost_ideal:
            return [items[i] for i in best], bc
        # simulated annealing over pairwise swaps between different half-warps (deterministic seed per class)
        rng = random.Random(1234 + 100 * self.La + 10 * self.Lb + self.Lc)
        order = list(best)
        nh = (nit + lanes_per_half - 1) // lanes_per_half
        hc = [half_cost(order[h * lanes_per_half:(h + 1) * lanes_per_half]) for h in range(nh)]
        cur = sum(hc)
        best_order, best_cost = list(order), cur
        iters = int(_os.environ.get("SI_G3_HRR_ITERS", "12000"))
        for t in range(iters):
            if best_cost <= self.hrr_cost_ideal:
                break
            temp = 1.5 * (1.0 - t / iters) + 0.05
            i, j = rng.randrange(nit), rng.randrange(nit)
            hi, hj = i // lanes_per_half, j // lanes_per_half
            if hi == hj:
                continue
            order[i], order[j] = order[j], order[i]
            ci = half_cost(order[hi * lanes_per_half:(hi + 1) * lanes_per_half])
            cj = half_cost(order[hj * lanes_per_half:(hj + 1) * lanes_per_half])
            delta = ci + cj - hc[hi] - hc[hj]
            if delta <= 0 or rng.random() < math.exp(-delta / temp):
                hc[hi], hc[hj] = ci, cj
                cur += delta
                if cur < best_cost:
                    best_order, best_cost = list(order), cur
            else:
                order[i], order[j] = order[j], order[i]
        return [items[i] for i in best_order], best_cost

    def _conflicts_c2sa(self):
        """Wavefronts of the C2S(a) loads for the current layout (see _half_warp_cost)."""
        EPT, nm, nca, nsb = self.EPT, self.nm, self.nca, self.nsb
        lanes_per_half = min(EPT, 16)
        tuples_per_half = 16 // lanes_per_half
        total = 0
        nit = nsb * nm
        for r in range((nit + EPT - 1) // EPT):
            chunk = [q + EPT * r for q in range(EPT) if q + EPT * r < nit]
            for h0 in range(0, len(chunk), lanes_per_half):
                part = chunk[h0:h0 + lanes_per_half]
                for a in range(nca):
                    if self.Lb > 0:
                        ad = [self.Z_OFF + a * self.ZA + item + t * self.WSZ for t in range(tuples_per_half) for item in part]
                    else:
                        ad = [self.Z_OFF + (item % nm) * self.ZS + a * nsb + item // nm + t * self.WSZ
                              for t in range(tuples_per_half) for item in part]
                    total += self._half_warp_cost(ad)
        return total

    # ---- emission helpers ---------------------------------------------------------------
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
        """Warp barrier between stages; not needed when every lane owns its whole tuple."""
        if self.EPT > 1:
            self.emit("si_g3_syncwarp();" + (f"   // {comment}" if comment else ""))

    # ---- per-class metadata tables ----------------------------------------------------
    def vrr_meta(self):
        """Per shell n >= 1, per component k: direction d, index k1 in shell n-1, k2 in shell n-2, fac."""
        tab = {}
        for n in range(1, self.L + 1):
            row = []
            for a in cart_order(n):
                d = next(i for i in range(3) if a[i] > 0)
                am = list(a)
                am[d] -= 1
                k1 = cidx(am)
                fac = a[d] - 1
                if fac > 0:
                    amm = list(am)
                    amm[d] -= 1
                    k2 = cidx(amm)
                else:
                    k2 = 0
                row.append((d, k1, k2, fac))
            tab[n] = row
        return tab

    def dtree_meta(self):
        """Per flat index f (shells 0..L): for d in x,y,z: (f - index read for the a-1_d neighbour, a_d).

        Elements without a neighbour (a_d == 0) multiply whatever they read by 0.  Reading their own index
        would add a second address stream to the wavefront (bank conflicts), so they are pointed at the
        neighbour address of the nearest element of their half-warp group that has one (a broadcast)."""
        flat = []
        for n in range(self.L + 1):
            for a in cart_order(n):
                flat.append((n, a))
        tgt = []
        for f, (n, a) in enumerate(flat):
            row = []
            for d in range(3):
                if a[d] > 0:
                    am = list(a)
                    am[d] -= 1
                    row.append(off(n - 1) + cidx(am))
                else:
                    row.append(None)
            tgt.append(row)
        B = min(self.EPT, 16)
        lo_top = off(self.La)
        out = []
        for f, (n, a) in enumerate(flat):
            ent = []
            for d in range(3):
                if tgt[f][d] is not None:
                    ent.append((f - tgt[f][d], a[d]))
                    continue
                t = f
                blk = range((f // B) * B, min((f // B) * B + B, self.NA))
                up = [g for g in blk if g > f and tgt[g][d] is not None]
                dn = [g for g in blk if g < f and g >= lo_top and tgt[g][d] is not None]
                if up:
                    t = tgt[up[0]][d]
                elif dn:
                    t = tgt[dn[-1]][d]
                ent.append((f - t, 0))
            out.append((n, ent))
        return out

    # ---- the generator ----------------------------------------------------------------------
    def generate(self):
        La, Lb, Lc, L = self.La, self.Lb, self.Lc, self.L
        EPT, TPW = self.EPT, self.TPW
        name = self.name
        E = self.emit
        self.lines = []
        self.ind = 0
        vm = self.vrr_meta()
        dm = self.dtree_meta()
        # ---- tables
        # VRR: packed int per (n,k): d | fac<<2 | k1<<8 | k2<<16
        for n in range(1, L + 1):
            vals = [d | (fac << 2) | (k1 << 8) | (k2 << 16) for (d, k1, k2, fac) in vm[n]]
            E(f"__device__ const int g3_vrr_{name}_{n}[{len(vals)}] = {{{', '.join(map(str, vals))}}};")
        if self.N0:
            comps = cart_order(self.N0)
            vals = [a[0] | (a[1] << 4) | (a[2] << 8) for a in comps]
            E(f"__device__ const int g3_pc_{name}[{len(vals)}] = {{{', '.join(map(str, vals))}}};")
            for n in self.chain_stores():
                k = self.N0 - n
                vals = [cidx((a[0] - k, a[1], a[2])) if a[0] >= k else -1 for a in comps]
                E(f"__device__ const int g3_pci_{name}_{n}[{len(vals)}] = {{{', '.join(map(str, vals))}}};")
        if Lc > 0:
            # D-tree: per flat index: byte... element deltas (<=0) packed 3x10 bits (as positive magnitudes), shell, factors 3x4 bits
            vals = []
            for (n, ent) in dm:
                v = 0
                for d in range(3):
                    assert -16 <= ent[d][0] < 240
                    v |= ((ent[d][0] + 16) & 0xFF) << (8 * d)
                v |= (n & 0xF) << 24
                vals.append(v)
            E(f"__device__ const int g3_dd_{name}[{len(vals)}] = {{{', '.join(map(str, vals))}}};")
            vals = []
            for (n, ent) in dm:
                v = 0
                for d in range(3):
                    v |= (ent[d][1] & 0xF) << (4 * d)
                vals.append(v)
            E(f"__device__ const int g3_df_{name}[{len(vals)}] = {{{', '.join(map(str, vals))}}};")
        if Lb > 0:
            vals = []
            for (m, ai) in self.hrr_order:
                r = 0
                while (r + 1) * (r + 2) // 2 <= ai:
                    r += 1
                vals.append((self.Y_OFF + m * self.YS + ai) | (r << 13) | ((self.Z_OFF + ai * self.ZA + m) << 16))
            E(f"__device__ const int g3_hm_{name}[{len(vals)}] = {{{', '.join(map(str, vals))}}};")
        E("")
        E(f"// class ({La}{Lb}|{Lc}): EPT={EPT} lanes per tuple, {TPW} tuples per warp, {self.WSZ} doubles of shared memory per tuple")
        E("// TAB: the group's tuples come from the host-built table A.tuples.  A template parameter, not a run-time")
        E("// branch: in the default instantiation the pair index stays warp-uniform for the compiler (uniform")
        E("// registers for the pair record), which is worth 5 % of the step.")
        E(f"template <bool TAB> SI_G3_DEV void g3_body_{name}(const G3Args& A, const int lane, double* wbase) {{")
        self.ind = 1
        E(f"constexpr int EPT = {EPT}, TPW = {TPW}, WSZ = {self.WSZ};")
        E("const int q = lane % EPT, tl = lane / EPT;")
        E("SI_G3_WTYPE W = SI_G3_WINIT(wbase + tl * WSZ, WSZ);")
        E("(void)q;")
        # ---- per-lane metadata
        nsl_v = {n: self.nslots(ncart(n)) for n in range(1, L + 1)}
        for n in range(1, L + 1):
            for s in range(nsl_v[n]):
                E(f"const int kc_{n}_{s} = min(q + EPT * {s}, {ncart(n) - 1}); const int vm_{n}_{s} = g3_vrr_{name}_{n}[kc_{n}_{s}];")
                if n >= 2 and n > self.N0:
                    # (n-2) factor of the staged shells as a per-lane constant (the int -> double conversion is slow)
                    E(f"const double vf_{n}_{s} = (double)((vm_{n}_{s} >> 2) & 0xF);")
        if self.N0:
            n0 = self.N0
            E(f"const int pk_ = g3_pc_{name}[kc_{n0}_0]; const int cay_ = (pk_ >> 4) & 15, caz_ = pk_ >> 8; (void)cay_; (void)caz_;")
            for n in self.chain_stores():
                E(f"const int ci_{n} = g3_pci_{name}_{n}[kc_{n0}_0];")
            # direction and (n-2) factor of every chain step: per-lane constants of the kernel
            for k in range(1, n0 + 1):
                E(f"const int cd{k}_ = {k} <= caz_ ? 2 : ({k} <= caz_ + cay_ ? 1 : 0);")
                if k >= 2:
                    E(f"const double cf{k}_ = (double)(cd{k}_ == 2 ? {k - 1} : (cd{k}_ == 1 ? {k - 1} - caz_ : {k - 1} - caz_ - cay_));")
        nsl_d = self.nslots(self.NA)
        if Lc > 0:
            for s in range(nsl_d):
                E(f"int nx_{s}, ny_{s}, nz_{s}, ds_{s}; double dfx_{s}, dfy_{s}, dfz_{s}; const int fc_{s} = min(q + EPT * {s}, {self.NA - 1}); "
                  f"{{ const int f = fc_{s}; int v = g3_dd_{name}[f]; int w = g3_df_{name}[f]; ds_{s} = (v >> 24) & 0xF; "
                  f"nx_{s} = f + 16 - (v & 0xFF); ny_{s} = f + 16 - ((v >> 8) & 0xFF); nz_{s} = f + 16 - ((v >> 16) & 0xFF); "
                  f"dfx_{s} = (double)(w & 0xF); dfy_{s} = (double)((w >> 4) & 0xF); dfz_{s} = (double)((w >> 8) & 0xF); }}")
                # leaf targets: index into a Y row (clamped) and validity of this lane's element
                E(f"const bool yv_{s} = (q + EPT * {s} >= {off(La)}) && (q + EPT * {s} < {self.NA}); "
                  f"const int yc_{s} = min(max(q + EPT * {s}, {off(La)}), {self.NA - 1}) - {off(La)};")
        E("")
        # Lanes whose element has no a-1 neighbour read some other element's neighbour slot and multiply it by a zero
        # factor, and pruned stores leave slots that are read but never written: the tuple's area starts out as zeros
        # so that every such value is finite.
        E("for (int i_ = q; i_ < WSZ; i_ += EPT) W[i_] = 0.0;")
        self.sync()
        if Lb > 0:
            nit_h = self.nca * self.nm
            E("// HRR work items (target a component, m) in the bank-conflict-free order found by the generator:")
            E("// Y base address | row of a << 13 | Z address << 16")
            self.open(f"for (int item = q; item < {nit_h}; item += EPT)")
            E(f"((int*)&W[{self.HM_OFF}])[item] = g3_hm_{name}[item];")
            self.close()
            self.sync()
        # ---- work loop over tuple groups
        E("// Work is handed out through one atomic counter per launch, one group per request; the next index is")
        E("// requested after the primitive phase and consumed (broadcast) at the end of the group.")
        E("long long ch_next = si_g3_next_group(A.counter, lane);")
        self.open("for (;;)")
        E("const long long grp = ch_next;")
        E("if (grp >= A.ngroups) break;")
        self.open("")
        E("// group -> TPW tuples.  Default: one pair, TPW consecutive aux shells of the class; tuples beyond the run")
        E("// end recompute the last valid one.  With a host-built tuple table (used when only a few aux shells of")
        E("// the class are in the range, e.g. one rank's shard of many, so that per-pair packing would leave most")
        E("// lanes idle) the group's tuples are listed explicitly: (pair, aux) of equal contraction lengths, -1 =")
        E("// padding (recomputes the group's first tuple).")
        E("int2 t_ = TAB ? A.tuples[grp * TPW + tl] : int2{0, 0};")
        E("if (TAB && t_.x < 0) t_ = A.tuples[grp * TPW];")
        E("const int ip = TAB ? t_.x : (int)((unsigned)grp / (unsigned)A.groups_per_pair);   // group indices fit 31 bits (32-bit work counter; checked by the host)")
        E("const int gi = (int)(grp - (long long)ip * A.groups_per_pair);")
        E("int ic = TAB ? t_.y : gi * TPW + tl;")
        E("const bool valid = TAB ? (A.tuples[grp * TPW + tl].x >= 0) : (ic < A.naux_c);")
        E("if (!TAB && !valid) ic = A.naux_c - 1;")
        E("const SiPairItem it = A.pairs[ip];")
        E("const SiAuxItem ai_ = A.auxitems[ic];")
        E("const int sc = ai_.shell; (void)sc;")
        E("const int Ka = it.Ka, Kb = it.Kb, Kc = ai_.Kc, pc0 = ai_.pc0;")
        E("const double Cx[3] = {ai_.C[0], ai_.C[1], ai_.C[2]};")
        E("const double AB[3] = {it.AB[0], it.AB[1], it.AB[2]};")
        E("const int Kc_max = si_g3_warp_max(Kc);")
        if self.YREG:
            for m in range(self.nm):
                E("double " + ", ".join(f"ya_{m}_{s} = 0.0" for s in self.tslots) + ";   // contraction accumulators (registers)")
        elif Lc == 0 and Y0REG:
            # s-type aux shell: Y[0][a'] accumulates V(n, 0, k) of the shells La..L, per (shell, slot) in registers
            self.y0 = [(n, s_) for n in range(La, L + 1) for s_ in range(self.nslots(ncart(n)))]
            E("double " + ", ".join(f"y0_{n}_{s_} = 0.0" for (n, s_) in self.y0) + ";   // contraction accumulators (registers)")
        else:
            E(f"for (int i_ = q; i_ < {self.nm * self.YS}; i_ += EPT) W[{self.Y_OFF} + i_] = 0.0;   // contraction accumulator")
        self.open("for (int ia = 0; ia < Ka; ++ia)")
        self.open("for (int ib = 0; ib < Kb; ++ib)")
        E("// primitive-pair data precomputed by the host shell-batcher: p, 1/p, P[3], PA[3], exp(-mu AB^2) da db")
        E("const double* pp = A.ppair + (size_t)(it.ppoff + ia * Kb + ib) * SI_PP_STRIDE;")
        E("const double p = pp[0], oop = pp[1];")
        E("const double P[3] = {pp[2], pp[3], pp[4]};")
        E("const double PA[3] = {pp[5], pp[6], pp[7]};")
        E("const double Kab = pp[8];")
        self.open("for (int ik = 0; ik < Kc_max; ++ik)")
        E("// tuples with fewer aux primitives repeat their last one with zero weight")
        E("const int ikk = ik < Kc ? ik : Kc - 1;")
        E("const double cx = A.aux.exps[pc0 + ikk], dc = ik < Kc ? A.aux.coef[pc0 + ikk] : 0.0;")
        E("const double ocx = A.aux_oexp[pc0 + ikk];   // 1/c, precomputed by the host")
        if "prim" in SKIP:      # ablation: keep only the loads of the primitive data
            E("if (q == 0) W[0] = p + oop + P[0] + PA[0] + Kab + cx + dc + ocx;")
        else:
            self._emit_primitive()
        self.close()  # ik
        self.close()  # ib
        self.close()  # ia
        E("// Fetch the next chunk index now: raw atomic result of lane 0, consumed (broadcast) at the chunk end, so")
        E("// its round trip overlaps the contracted phase.  (Issued earlier, the result register gets spilled, and")
        E("// the spill store waits for the atomic.)")
        E("const int ch_raw = si_g3_fetch_group(A.counter, lane);")
        if Lc == 0 and Y0REG:
            self.sync("last primitive's readers of the overlaid arrays are done")
            for (n, s_) in self.y0:
                kk = f"kc_{n}_{s_}" if n >= 1 else "0"
                E(f"W[{self.Y_OFF + off(n) - off(La)} + {kk}] = y0_{n}_{s_};")
        if self.YREG:
            self.sync("last primitive's readers of the overlaid arrays are done")
            for m in range(self.nm):
                for s in self.tslots:
                    E(f"if (yv_{s}) W[{self.Y_OFF + m * self.YS} + yc_{s}] = ya_{m}_{s};")
        self.sync()
        if Lb > 0 and "hrr" not in SKIP:
            self._emit_hrr_c2sb()
            self.sync()
        if "c2sa" not in SKIP:
            self._emit_c2s_a_store()
        else:
            E("if (valid && q == 0) A.od.out[0] = W[0];")
        self.sync()
        E("ch_next = si_g3_bcast_group(ch_raw);")
        self.close()  # group scope
        self.close()  # for(;;)
        self.close()  # function
        E("")
        # kernel wrapper
        E(f"SI_G3_GLOBAL({self.MB}) void g3_kernel_{name}(G3Args A) {{")
        E("    SI_G3_KERNEL_PROLOGUE")
        E(f"    g3_body_{name}<false>(A, g3_lane, g3_smem + (size_t)g3_warp * {self.WSZ * TPW});")
        E("}")
        if self.has_table_variant():
            E(f"SI_G3_GLOBAL({self.MB}) void g3_tkernel_{name}(G3Args A) {{")
            E("    SI_G3_KERNEL_PROLOGUE")
            E(f"    g3_body_{name}<true>(A, g3_lane, g3_smem + (size_t)g3_warp * {self.WSZ * TPW});")
            E("}")
        E("")
        return "\n".join(self.lines)

    # ---- primitive phase ----------------------------------------------------------------
    def _emit_primitive(self):
        La, Lb, Lc, L = self.La, self.Lb, self.Lc, self.L
        EPT = self.EPT
        E = self.emit
        vm_n = self.vrr_meta()
        E("const double pc_ = p + cx;")
        E("const double rs_ = si_g3_rsqrt(pc_);   // 1/sqrt(p+c)")
        E("const double rpc = rs_ * rs_;")
        E("const double rho = p * cx * rpc;")
        E("const double rho_p = cx * rpc;          // rho / p")
        E("double PC[3]; double r2pc = 0.0;")
        self.open("for (int d = 0; d < 3; ++d)")
        E("PC[d] = P[d] - Cx[d]; r2pc += PC[d] * PC[d];")
        self.close()
        E("const double T = rho * r2pc;")
        E("// 2 pi^2.5 / (sqrt(p+c) p c) * K_ab * d_c ;  1/sqrt(p+c) = sqrt(p+c) / (p+c)")
        E("const double pref = 2.0 * SI_PI_25 * rs_ * oop * ocx * Kab * dc;")
        E("// opt-in pre-screening: bound of the squared (ss|s)-type base integral, F_0(T)^2 <= min(1, pi / (4 T)); warp-uniform")
        E("if (A.screen_eps2 > 0.0 && si_g3_warp_all(pref * pref * fmin(1.0, 0.78539816339744831 / fmax(T, 1e-300)) <= A.screen_eps2)) continue;")
        E("const double binv = 2.0 * pc_;")
        if Lc > 0:
            # VRR on the scaled quantities binv^n V(n): the (n-2) terms pick up binv^2
            E("const double oo2p = 0.5 * oop * binv * binv, c2 = -rho_p * oo2p;")
        else:
            E("const double oo2p = 0.5 * oop, c2 = -0.5 * oop * rho_p;")
        E("const double beta = 0.5 * rpc;      // (rho/c)/(2p)")
        E("const double rho_c = p * rpc;       // rho / c")
        E("const double alx = rho_c * PC[0], aly = rho_c * PC[1], alz = rho_c * PC[2];")
        E("(void)alx; (void)aly; (void)alz; (void)beta; (void)binv;")
        # No barrier is needed before overwriting the scalar area and the Boys values: their last readers
        # (VRR shells, start of the aux recurrence) are separated from here by at least one barrier of the
        # previous primitive whenever several lanes share a tuple (EPT > 1 implies La + Lb >= 1).
        # Exception: with Lc == 1 the aux recurrence has no barrier of its own, so the beta powers read at
        # its start could still be in use.
        if L == 0 or Lc == 1:
            self.sync("previous primitive's readers are done")
        # scalars to shared memory (lane q==0 of the tuple)
        # VRR direction factors: per-lane selects from registers (PASEL) instead of shared-memory scalars
        pasel = self.PASEL = _os.environ.get("SI_G3_PASEL", "1") == "1" and L >= 1
        if pasel:
            sc = "binv * " if Lc > 0 else ""
            for d in range(3):
                E(f"const double pas{d}_ = {sc}PA[{d}], cps{d}_ = -{sc}rho_p * PC[{d}];")
        self.open("if (q == 0)")
        for d in range(3):
            if pasel:
                continue
            if Lc > 0:
                E(f"W[{self.S_PA + d}] = binv * PA[{d}]; W[{self.S_CPC + d}] = -binv * rho_p * PC[{d}];")
            else:
                E(f"W[{self.S_PA + d}] = PA[{d}]; W[{self.S_CPC + d}] = -rho_p * PC[{d}];")
        if Lc > 0 and EPT <= 2:
            E("{ double bp = 1.0;")
            for n in range(L + 1):
                E(f"  W[{self.S_BPOW + n}] = bp; bp *= beta;")
            E("}")
        self.close()
        if Lc > 0 and EPT > 2:
            # beta^j, one power per lane (binary powering) instead of a serial chain on one lane
            self.open(f"for (int j = q; j < {L + 1}; j += EPT)")
            E("double bp = 1.0, bb = beta; int e_ = j;")
            E("for (int it_ = 0; it_ < 4; ++it_) { if (e_ & 1) bp *= bb; bb *= bb; e_ >>= 1; }")
            E(f"W[{self.S_BPOW} + j] = bp;")
            self.close()
        # Boys base values: V(0, jn, 0) = pref * F_{Lc+jn}(T)
        nb = L + 1
        self.open(f"for (int j = q; j < {nb}; j += EPT)")
        if "boys" in SKIP:
            E(f"const double bv_ = pref * (T + j);")
        else:
            E(f"const double bv_ = pref * si_boys_fast(A.boys, {Lc} + j, T);")
        E(f"W[{self.V_OFF[(0, 0)]} + j] = bv_;")
        if Lc > 0:
            E(f"if (j == 0) W[{self.D_OFF[0]}] = bv_;   // element 0 of the aux recurrence's level-0 array")
        self.close()
        # note: V(0,jn,0) are contiguous since ncart(0)=1
        for jn in range(nb):
            assert self.V_OFF[(0, jn)] == self.V_OFF[(0, 0)] + jn
        self.sync()
        # ---- lane-private chain over shells 1..N0
        n0 = self.N0
        if n0:
            assert pasel
            self.open("")
            for jn in range(L + 1):
                E(f"const double c0_{jn} = W[{self.V_OFF[(0, jn)]}];")
            cs = self.chain_stores()
            stores = []
            for k in range(1, n0 + 1):
                E(f"const double pa{k}_ = cd{k}_ == 0 ? pas0_ : (cd{k}_ == 1 ? pas1_ : pas2_), cp{k}_ = cd{k}_ == 0 ? cps0_ : (cd{k}_ == 1 ? cps1_ : cps2_);")
                if k >= 2:
                    E(f"const double f1{k}_ = cf{k}_ * oo2p, f2{k}_ = cf{k}_ * c2;")
                for jn in range(L - k + 1):
                    if k >= 2:
                        E(f"const double c{k}_{jn} = fma(f2{k}_, c{k - 2}_{jn + 1}, fma(f1{k}_, c{k - 2}_{jn}, fma(pa{k}_, c{k - 1}_{jn}, cp{k}_ * c{k - 1}_{jn + 1})));")
                    else:
                        E(f"const double c{k}_{jn} = fma(pa{k}_, c{k - 1}_{jn}, cp{k}_ * c{k - 1}_{jn + 1});")
                if k < n0 and k in cs:
                    for jn in range(cs[k]):
                        stores.append(f"if (ci_{k} >= 0) W[{self.V_OFF[(k, jn)]} + ci_{k}] = c{k}_{jn};")
            y0direct = Lc == 0 and Y0REG
            for jn in range(L - n0 + 1):
                if y0direct and n0 == L:
                    continue          # top shell of an s-type aux class: goes straight into the register accumulator
                stores.append(f"W[{self.V_OFF[(n0, jn)]} + kc_{n0}_0] = c{n0}_{jn};")
            if y0direct and n0 >= La:
                E(f"y0_{n0}_0 += c{n0}_0;")
            for st in stores:
                E(st)
            self.close()
            self.sync()
        # ---- VRR shells
        for n in range(n0 + 1, (0 if "vrr" in SKIP else L) + 1):
            nc = ncart(n)
            njn = L - n + 1
            # Shared-memory loads and stores of one stage touch disjoint arrays, but the compiler cannot
            # know that (one array, per-lane indices) and would serialise LDS behind every STS.  So each
            # stage is emitted as: all loads, all arithmetic, all stores.
            self.open("")
            nsl_n = self.nslots(nc)
            for s in range(nsl_n):
                # branch-free: lanes beyond the shell recompute (and re-store) its last component
                E(f"const int m{s}_ = vm_{n}_{s}; const int d{s}_ = m{s}_ & 3; const int k1_{s} = (m{s}_ >> 8) & 0xFF; const int k2_{s} = (m{s}_ >> 16) & 0xFF;")
                if pasel:
                    E(f"const double pa{s}_ = d{s}_ == 0 ? pas0_ : (d{s}_ == 1 ? pas1_ : pas2_), cp{s}_ = d{s}_ == 0 ? cps0_ : (d{s}_ == 1 ? cps1_ : cps2_);")
                else:
                    E(f"const double pa{s}_ = W[{self.S_PA} + d{s}_], cp{s}_ = W[{self.S_CPC} + d{s}_];")
                for jn in range(njn + 1):
                    E(f"const double u{s}_{jn} = W[{self.V_OFF[(n - 1, jn)]} + k1_{s}];")
                    if n >= 2:
                        E(f"const double w{s}_{jn} = W[{self.V_OFF[(n - 2, jn)]} + k2_{s}];")
            for s in range(nsl_n):
                if n >= 2:
                    E(f"const double f1_{s} = vf_{n}_{s} * oo2p, f2_{s} = vf_{n}_{s} * c2;")
                for jn in range(njn):
                    if n >= 2:
                        E(f"const double v{s}_{jn} = fma(f2_{s}, w{s}_{jn + 1}, fma(f1_{s}, w{s}_{jn}, fma(pa{s}_, u{s}_{jn}, cp{s}_ * u{s}_{jn + 1})));")
                    else:
                        E(f"const double v{s}_{jn} = fma(pa{s}_, u{s}_{jn}, cp{s}_ * u{s}_{jn + 1});")
            y0direct = Lc == 0 and Y0REG
            for s in range(nsl_n):
                if y0direct and n >= La:
                    E(f"y0_{n}_{s} += v{s}_0;")
                for jn in range(njn):
                    if y0direct and n == L:
                        continue      # top shell: nobody reads it from shared memory
                    E(f"W[{self.V_OFF[(n, jn)]} + kc_{n}_{s}] = v{s}_{jn};")
            self.close()
            self.sync()
        # ---- D-tree
        if Lc == 0 and Y0REG:
            for (n, s_) in self.y0:
                if n >= max(self.N0, 1) and "vrr" not in SKIP:
                    continue          # accumulated from the registers of its VRR stage
                kk = f"kc_{n}_{s_}" if n >= 1 else "0"
                E(f"y0_{n}_{s_} += W[{self.V_OFF[(n, 0)]} + {kk}];")
            return
        if Lc == 0:
            # Y[0][a'] (+)= V(n,0,k) for shells La..L   (C2S of an s function is the identity)
            for n in range(La, L + 1):
                nc = ncart(n)
                for s in range(self.nslots(nc)):
                    partial = (s + 1) * EPT > nc
                    self.open("")
                    ya = self.Y_OFF + off(n) - off(La)
                    kk = f"kc_{n}_{s}" if n >= 1 else "0"
                    E(f"const int k_ = {kk}; const double v_ = W[{ya} + k_] + W[{self.V_OFF[(n, 0)]} + k_];")
                    E((f"if (q + EPT * {s} < {nc}) " if partial else "") + f"W[{ya} + k_] = v_;")
                    self.close()
            return
        if "dtree" in SKIP:
            return
        # root: h[f] = V(n, 0, k) * binv^n for shells amin(0)..L is what the scaled VRR left at D_OFF[0] + f
        # DFS over monomials; direction sequences non-decreasing (x.., y.., z..)
        nsl = self.nslots(self.NA)
        # per-slot: own root values + beta powers for the leaves
        for s in range(nsl):
            lo = off(self.amin(0))
            E(f"const double r0_{s} = W[{self.D_OFF[0]} + fc_{s}];")
        for s in range(nsl):
            E(f"const double bp_{s} = W[{self.S_BPOW} + ds_{s}];")
        al = ["alx", "aly", "alz"]
        dfn = ["dfx", "dfy", "dfz"]
        Mc = cart2sph(Lc)
        fc = lmn_factors(Lc)
        seen = set()     # (m, slot) pairs that already received their first contribution

        def rec(c, level, lastd):
            # values of node c are in registers r{level}_{s}; its array is at D_OFF[level]
            for d in range(lastd, 3):
                c2 = list(c)
                c2[d] += 1
                lev2 = level + 1
                leaf = lev2 == Lc
                lo2 = off(self.amin(lev2))
                self.open("")
                E(f"// node {tuple(c2)} <- {tuple(c)} (+1_{'xyz'[d]})")
                leaf_updates = []
                act = []
                for s in range(nsl):
                    smin, smax = s * EPT, s * EPT + EPT - 1
                    if smax < lo2 or smin >= self.NA:
                        if not leaf:
                            E(f"const double r{lev2}_{s} = 0.0;")
                        continue
                    act.append(s)
                for s in act:
                    E(f"const double nb{lev2}_{s} = W[{self.D_OFF[level]} + {'nx ny nz'.split()[d]}_{s}];")
                for s in act:
                    E(f"const double r{lev2}_{s} = fma({dfn[d]}_{s}, nb{lev2}_{s}, {al[d]} * r{level}_{s});")
                for s in act:
                    if leaf:
                        ci = cidx(c2)
                        E(f"const double t{lev2}_{s} = bp_{s} * r{lev2}_{s};")
                        for m in range(self.nm):
                            if Mc[m][ci] == 0.0:
                                continue
                            coef = lit(Mc[m][ci] * fc[ci])
                            ya = self.Y_OFF + m * self.YS
                            if self.YREG:
                                E(f"ya_{m}_{s} = fma({coef}, t{lev2}_{s}, ya_{m}_{s});")
                            else:
                                leaf_updates.append((s, f"W[{ya} + yc_{s}]", coef, f"t{lev2}_{s}"))
                    else:
                        # stored only where the next level reads it as an a-1 neighbour: shells amin(lev2+1)-1 .. L-1
                        # (the top shell has no reader; about a third of the elements)
                        f_lo, f_hi = off(max(0, self.amin(lev2 + 1) - 1)), off(L)
                        if s * EPT < f_hi and s * EPT + EPT - 1 >= f_lo:
                            E(f"W[{self.D_OFF[lev2]} + fc_{s}] = r{lev2}_{s};")
                # read-modify-write of the spherical accumulators: all loads, all FMAs, then predicated stores
                for k_, (s, addr, coef, t) in enumerate(leaf_updates):
                    E(f"const double yo{k_} = {addr};")
                for k_, (s, addr, coef, t) in enumerate(leaf_updates):
                    E(f"const double yn{k_} = fma({coef}, {t}, yo{k_});")
                for k_, (s, addr, coef, t) in enumerate(leaf_updates):
                    E(f"if (yv_{s}) {addr} = yn{k_};")
                if not leaf:
                    self.sync()
                    rec(c2, lev2, d)
                    self.sync("children done before the array is overwritten")
                self.close()

        self.open("")
        rec([0, 0, 0], 0, 0)
        self.close()

    # ---- HRR tile per (target a, m) + C2S(b): Z[m][a][jb] ---------------------------------
    def _emit_hrr_c2sb(self):
        E = self.emit
        La, Lb, EPT = self.La, self.Lb, self.EPT
        nca, nm, nsb, n_ap = self.nca, self.nm, self.nsb, self.n_ap
        nit = nca * nm
        Mb = cart2sph(Lb)
        fb = lmn_factors(Lb)
        # symbolic tile: nodes (k, b) = (a + k, b); b built by first-nonzero direction
        memo = {}
        order = []

        def node(k, b):
            key = (tuple(k), tuple(b))
            if key in memo:
                return memo[key]
            if sum(b) == 0:
                name = "i" + "_".join(map(str, k))
                memo[key] = name
                order.append(("in", key, name))
                return name
            d = next(i for i in range(3) if b[i] > 0)
            bm = list(b)
            bm[d] -= 1
            kp = list(k)
            kp[d] += 1
            n1 = node(kp, bm)
            n2 = node(k, bm)
            name = "t" + "_".join(map(str, k)) + "__" + "_".join(map(str, b))
            memo[key] = name
            order.append(("op", key, name, n1, n2, d))
            return name

        targets = [node([0, 0, 0], list(b)) for b in cart_order(Lb)]
        self.open(f"for (int item = q; item < {nit}; item += EPT)")
        E(f"const int hm_ = ((const int*)&W[{self.HM_OFF}])[item];   // items (m, a), a fastest: near-consecutive addresses")
        E("const int yb = hm_ & 0x1FFF, ia_ = (hm_ >> 13) & 7, zb = hm_ >> 16;   // ia_: row of the target a (l = La - ia_)")
        for ik in range(Lb + 1):
            E(f"const int yb{ik} = yb + ia_ * {ik};")
        # inputs: idx(a+k) - idx-base: off(La+|k|)-off(La) + row(ik) + nz_k + ia*ik   (+ ai)
        for ent in order:
            if ent[0] == "in":
                k = ent[1][0]
                ik = k[1] + k[2]
                nk = sum(k)
                const = off(La + nk) - off(La) + ik * (ik + 1) // 2 + k[2]
                E(f"const double {ent[2]} = W[yb{ik} + {const}];")
        for ent in order:
            if ent[0] != "in":
                _, key, name, n1, n2, d = ent
                E(f"const double {name} = fma(AB[{d}], {n2}, {n1});")
        # C2S(b)
        for jb in range(nsb):
            expr = None
            for j in range(self.ncb):
                coef = Mb[jb][j] * fb[j]
                if Mb[jb][j] == 0.0:
                    continue
                expr = f"{lit(coef)} * {targets[j]}" if expr is None else f"fma({lit(coef)}, {targets[j]}, {expr})"
            E(f"const double zb{jb} = {expr};")
        for jb in range(nsb):
            E(f"W[zb + {jb * nm}] = zb{jb};")
        self.close()

    # ---- C2S(a) + global store --------------------------------------------------------------
    def _emit_c2s_a_store(self):
        E = self.emit
        La, EPT = self.La, self.EPT
        nca, nm, nsb, nsa = self.nca, self.nm, self.nsb, self.nsa
        Ma = cart2sph(La)
        fa = lmn_factors(La)
        nit = nsb * nm
        E("const long long col = (long long)(ai_.foff - A.col_begin);")
        E("const int fa0 = it.fa0, fb0 = it.fb0;")
        self.open(f"for (int item = q; item < {nit}; item += EPT)")
        E(f"const int jb = item / {nm}, m = item - jb * {nm};")
        for a in range(nca):
            if self.Lb > 0:
                E(f"const double z{a} = W[{self.Z_OFF + a * self.ZA} + item];")
            else:
                E(f"const double z{a} = W[{self.Z_OFF} + m * {self.ZS} + {a * nsb} + jb];")
        for ia in range(nsa):
            expr = None
            for a in range(nca):
                if Ma[ia][a] == 0.0:
                    continue
                coef = Ma[ia][a] * fa[a]
                expr = f"{lit(coef)} * z{a}" if expr is None else f"fma({lit(coef)}, z{a}, {expr})"
            E(f"const double o{ia} = {expr};")
        # one address computation per item: element ia lives at p0 + ia * s0 (and its mirror at p1 + ia * s1)
        E(f"double* p0; long long s0; double* p1 = nullptr; long long s1 = 0;")
        E(f"si_g3_store_setup(A.od, it, fa0, fb0, {nsa}, {nsb}, {nm}, col, jb, m, &p0, &s0, &p1, &s1);")
        self.open("if (valid)")
        for ia in range(nsa):
            E(f"p0[{ia} * s0] = o{ia};")
        self.open("if (p1)")
        for ia in range(nsa):
            E(f"p1[{ia} * s1] = o{ia};")
        self.close()
        self.close()
        self.close()


Y0REG = _os.environ.get("SI_G3_Y0REG", "1") == "1"

HEADER = """// GENERATED by sympleints_b200/codegen/gen3c2e.py -- DO NOT EDIT.
#include "../si_g3.h"
"""


def generate_file(classes):
    parts = [HEADER]
    for (La, Lb, Lc) in classes:
        parts.append(ClassGen(La, Lb, Lc).generate())
    return "\n".join(parts)


def all_classes(lmax=4, lauxmax=5):
    """The 3c2e classes (La >= Lb <= lmax, Lc <= lauxmax) plus (La, 0 | Lc) for lmax < La <= lauxmax: the
    2c2e metric (a|b) is evaluated as (a s0 | b) with a unit s "function" of exponent 0 (p = a, P = A)."""
    cl = [(a, b, c) for a in range(lmax + 1) for b in range(a + 1) for c in range(lauxmax + 1)]
    cl += [(a, 0, c) for a in range(lmax + 1, lauxmax + 1) for c in range(lauxmax + 1)]
    return cl


def class_info(lmax=4, lauxmax=5):
    return {(a, b, c): ClassGen(a, b, c) for (a, b, c) in all_classes(lmax, lauxmax)}


# ----------------------------------------------------------------------------------------
# file emission: several translation units (parallel nvcc) + a class table + launcher
# ----------------------------------------------------------------------------------------
def write_sources(outdir, lmax=4, lauxmax=5, nparts=15):
    """Write generated/g3_part_XX.cu, generated/g3_table.inc.  Returns the list of .cu files."""
    import os

    os.makedirs(outdir, exist_ok=True)
    classes = all_classes(lmax, lauxmax)
    gens = {c: ClassGen(*c) for c in classes}
    # balance parts by generated size (roughly WSZ-proportional): group by (La, Lb)
    pairs = sorted({(a, b) for (a, b, _) in classes})
    parts = [[] for _ in range(nparts)]
    load = [0] * nparts
    for pr in sorted(pairs, key=lambda p: -(off(p[0] + p[1] + 1))):
        i = load.index(min(load))
        cl = [c for c in classes if (c[0], c[1]) == pr]
        parts[i].extend(cl)
        load[i] += sum(gens[c].WSZ for c in cl)
    files = []
    table = []
    for pi, cl in enumerate(parts):
        if not cl:
            continue
        src = [HEADER]
        for c in cl:
            src.append(gens[c].generate())
        # launcher for this part
        src.append("#if !defined(SI_G3_HOST_EMU)")
        src.append(f"int g3_launch_part_{pi}(int id, const G3Args& A, int grid, int block, size_t smem, int max_smem, cudaStream_t s) {{")
        src.append("    // max_smem < 0: programmatic dependent launch (the kernels trigger their dependents at once, so the next")
        src.append("    // launch of the stream fills the SMs as this one's blocks retire; the launches are independent of each other)")
        src.append("    cudaLaunchConfig_t cfg = {};")
        src.append("    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;")
        src.append("    cudaLaunchAttribute at[1];")
        src.append("    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;")
        src.append("    cfg.attrs = at; cfg.numAttrs = max_smem < 0 ? 1 : 0;")
        src.append("    switch (id) {")
        for c in cl:
            g = gens[c]
            cid = c[0] * 100 + c[1] * 10 + c[2]
            tk = (f"if (A.tuples) return cudaLaunchKernelEx(&cfg, g3_tkernel_{g.name}, A) != cudaSuccess; "
                  if g.has_table_variant() else "if (A.tuples) return 1; ")
            src.append(f"        case {cid}: {{ {tk}return cudaLaunchKernelEx(&cfg, g3_kernel_{g.name}, A) != cudaSuccess; }}")
            table.append((cid, pi, g.EPT, g.TPW, g.WSZ))
        src.append("        default: return 1;")
        src.append("    }")
        src.append("}")
        # the dynamic shared-memory limit is a per-device function attribute: set for every kernel of the part
        # once per context (si_init), not behind a process-wide flag
        src.append(f"int g3_attr_part_{pi}(int max_smem) {{")
        for c in cl:
            g = gens[c]
            names = [f"g3_kernel_{g.name}"] + ([f"g3_tkernel_{g.name}"] if g.has_table_variant() else [])
            for kn in names:
                src.append(f"    if (cudaFuncSetAttribute({kn}, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) return 1;")
        src.append("    return 0;")
        src.append("}")
        src.append("#endif")
        fn = os.path.join(outdir, f"g3_part_{pi:02d}.cu")
        text = "\n".join(src) + "\n"
        if not os.path.exists(fn) or open(fn).read() != text:
            open(fn, "w").write(text)
        files.append(fn)
    # table include
    lines = ["// GENERATED -- class table of the specialised 3c2e kernels: {id, part, EPT, TPW, WSZ}"]
    lines.append("static const int g3_table[][5] = {")
    for t in sorted(table):
        lines.append("    {%d, %d, %d, %d, %d}," % t)
    lines.append("};")
    for pi, cl in enumerate(parts):
        if cl:
            lines.append(f"int g3_launch_part_{pi}(int id, const G3Args& A, int grid, int block, size_t smem, int max_smem, cudaStream_t s);")
    for pi, cl in enumerate(parts):
        if cl:
            lines.append(f"int g3_attr_part_{pi}(int max_smem);")
    lines.append("static int g3_set_attributes(int max_smem) {")
    for pi, cl in enumerate(parts):
        if cl:
            lines.append(f"    if (g3_attr_part_{pi}(max_smem)) return 1;")
    lines.append("    return 0;")
    lines.append("}")
    lines.append("static int g3_launch(int part, int id, const G3Args& A, int grid, int block, size_t smem, int max_smem, cudaStream_t s) {")
    lines.append("    switch (part) {")
    for pi, cl in enumerate(parts):
        if cl:
            lines.append(f"        case {pi}: return g3_launch_part_{pi}(id, A, grid, block, smem, max_smem, s);")
    lines.append("        default: return 1;")
    lines.append("    }")
    lines.append("}")
    fn = os.path.join(outdir, "g3_table.inc")
    text = "\n".join(lines) + "\n"
    if not os.path.exists(fn) or open(fn).read() != text:
        open(fn, "w").write(text)
    return files


def write_emu_source(path, classes):
    """TEST-ONLY: one translation unit with the bodies of `classes` and a switch, for tests/emu."""
    src = ["#define SI_G3_HOST_EMU 1", HEADER.replace("../si_g3.h", "../../sympleints_b200/csrc/si_g3.h")]
    for c in classes:
        src.append(ClassGen(*c).generate())
    src.append("struct G3EmuClass { int id, ept, tpw, wsz; void (*body)(const G3Args&, int, double*); void (*tbody)(const G3Args&, int, double*); };")
    src.append("static const G3EmuClass g3_emu_classes[] = {")
    for c in classes:
        g = ClassGen(*c)
        src.append(f"    {{{c[0] * 100 + c[1] * 10 + c[2]}, {g.EPT}, {g.TPW}, {g.WSZ}, g3_body_{g.name}<false>, g3_body_{g.name}<true>}},")
    src.append("};")
    src.append(f"static const int g3_emu_nclasses = {len(classes)};")
    text = "\n".join(src) + "\n"
    import os
    if not os.path.exists(path) or open(path).read() != text:
        open(path, "w").write(text)


if __name__ == "__main__":
    import sys
    out = sys.argv[1] if len(sys.argv) > 1 else "sympleints_b200/csrc/generated"
    for f in write_sources(out):
        print(f)
