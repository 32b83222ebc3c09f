#!/usr/bin/env python
"""Benchmark of the hot path: the 3c2e_sph density-fitting tensor (BASELINE.json config 5).

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU path

One "step" = one full evaluation of the (a>=b)-unique 3c2e tensor (ab|c), lmax=4 / lauxmax=5,
for the 300-shell / 1300-function orbital basis x 1200-shell / 5000-function auxiliary basis of
SURVEY.md section 8d (4.246e9 integrals, 33.97 GB), sharded over contiguous auxiliary-shell
ranges across the N ranks (strong scaling: total work fixed; no collective on the data path).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "3c2e_sph integrals per sec (unique a>=b; lmax=4, lauxmax=5; 1300-function / 5000-aux-function synthetic system)"
UNIT = "integrals/s"


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ts, line in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU baseline: the reference's generated numpy code on a stratified sample of shell triples
# --------------------------------------------------------------------------------------
def _eval_triple_cpu(a, b, c):
    """One contracted block on the CPU; returns (seconds, kind)."""
    from oracle import refmod
    if "have" not in _REF:
        _REF["have"] = refmod.available("3c2e")
        if _REF["have"]:
            refmod.funcs("3c2e")   # import the reference's generated module outside the timed call
    t0 = time.perf_counter()
    if _REF["have"]:
        refmod.eval_triple(a, b, c)   # the reference's own generated numpy code (oracle/_ref)
        kind = "reference"
    else:
        from oracle import ints
        ints.int3c2e_sph(a.L, b.L, c.L, a.exps, a.coeffs, a.center, b.exps, b.coeffs, b.center,
                         c.exps, c.coeffs, c.center)
        kind = "port"
    return time.perf_counter() - t0, kind


_REF = {}


def _sample_worker(args):
    shells, aux, items = args
    out = []
    kind = None
    for (i, j, k) in items:
        dt, kind = _eval_triple_cpu(shells[i], shells[j], aux[k])
        out.append(dt)
    return out, kind


def cpu_baseline_3c2e(shells, aux, per_class=2, nproc=1, seed=0, pool=None):
    """Stratified sample: `per_class` random shell triples for every (La>=Lb, Kab, Lc) class, timed with
    the reference's generated numpy code; extrapolated with the class histogram."""
    import collections
    import multiprocessing as mp

    rng = np.random.default_rng(seed)
    cls = collections.defaultdict(list)
    n = len(shells)
    for i in range(n):
        for j in range(i + 1):
            a, b = shells[i], shells[j]
            cls[(max(a.L, b.L), min(a.L, b.L), len(a.exps) * len(b.exps))].append((i, j))
    auxcls = collections.defaultdict(list)
    for k, c in enumerate(aux):
        auxcls[(c.L, len(c.exps))].append(k)
    items, meta = [], []
    for pk, pairs in cls.items():
        for ak, auxs in auxcls.items():
            ntrip = len(pairs) * len(auxs)
            nint = (2 * pk[0] + 1) * (2 * pk[1] + 1) * (2 * ak[0] + 1)
            for _ in range(per_class):
                i, j = pairs[rng.integers(len(pairs))]
                items.append((i, j, auxs[rng.integers(len(auxs))]))
                meta.append((ntrip / per_class, nint))
    order = rng.permutation(len(items))
    chunks = [[items[q] for q in order[w::nproc]] for w in range(nproc)]
    t0 = time.perf_counter()
    if nproc > 1 and pool is not None:
        res = pool.map(_sample_worker, [(shells, aux, ch) for ch in chunks])
    elif nproc > 1:
        with mp.get_context("fork").Pool(nproc) as pool_:
            res = pool_.map(_sample_worker, [(shells, aux, ch) for ch in chunks])
    else:
        res = [_sample_worker((shells, aux, chunks[0]))]
    wall = time.perf_counter() - t0
    times = np.empty(len(items))
    for w, (ts, kind) in enumerate(res):
        times[order[w::nproc]] = ts
    weights = np.array([m[0] for m in meta])
    nints = np.array([m[1] for m in meta])
    t_full_1core = float((weights * times).sum())           # core-seconds for the whole tensor
    total_int = float((weights * nints).sum())
    p_eff = float(times.sum() / wall) if wall > 0 else 1.0   # cores effectively busy
    value = total_int / (t_full_1core / max(p_eff, 1e-9))
    return dict(value=value, unit=UNIT, cores=nproc, kind=kind,
                sample=f"{len(items)} shell triples ({per_class} per (La,Lb,Kab,Lc,Kc) class, {len(cls) * len(auxcls)} classes), "
                       f"{times.sum():.1f} core-s measured, extrapolated with the class histogram to "
                       f"{t_full_1core:.3e} core-s for the full tensor",
                wall_s=wall, core_seconds_full_tensor=t_full_1core)


# --------------------------------------------------------------------------------------
# parity of what the timed steps wrote: sampled shell blocks against the reference's generated numpy code
# --------------------------------------------------------------------------------------
def _ref_pair(key, a, b, R, dtype):
    from oracle import ints, refmod
    if refmod.available(key):
        return refmod.eval_pair(key, a, b, R=R, dtype=dtype), "reference"
    f, ncomp = ints.KEY_FUNCS[key]
    c = lambda x: np.asarray(x, dtype=dtype)
    return f(a.L, b.L, c(a.exps), c(a.coeffs), c(a.center), c(b.exps), c(b.coeffs), c(b.center), R=c(R),
             sph=True, normalize="cgto").reshape(ncomp, 2 * a.L + 1, 2 * b.L + 1), "port"


def _ref_triple(a, b, c, dtype):
    from oracle import ints, refmod
    if refmod.available("3c2e"):
        return refmod.eval_triple(a, b, c, dtype), "reference"
    cv = lambda x: np.asarray(x, dtype=dtype)
    return ints.int3c2e_sph(a.L, b.L, c.L, cv(a.exps), cv(a.coeffs), cv(a.center), cv(b.exps), cv(b.coeffs), cv(b.center),
                            cv(c.exps), cv(c.coeffs), cv(c.center)).reshape(2 * a.L + 1, 2 * b.L + 1, -1), "port"


def sample_triples(shells, aux, sb, se, per_class, seed):
    """(i >= j, k): `per_class` triples for every (L_i, K_i, L_j, K_j) x L_k combination in the aux range --
    all 90 native classes, the swapped shell orders and every contraction length."""
    rng = np.random.default_rng(seed)
    pairs = {}
    for i, a in enumerate(shells):
        for j in range(i + 1):
            lst = pairs.setdefault((a.L, len(a.exps), shells[j].L, len(shells[j].exps)), [])
            if len(lst) < 64:
                lst.append((i, j))
    auxby = {}
    for k in range(sb, se):
        auxby.setdefault(aux[k].L, []).append(k)
    out = []
    for key in sorted(pairs):
        for lc in sorted(auxby):
            for _ in range(per_class):
                i, j = pairs[key][rng.integers(len(pairs[key]))]
                out.append((i, j, int(rng.choice(auxby[lc]))))
    return out


def parity_3c2e(T, shells, aux, sb, se, per_class, seed):
    """T: this rank's (packed_rows, ncol) slab (torch tensor on the device or numpy host buffer)."""
    from oracle import refmod
    sizes = np.array([2 * s.L + 1 for s in shells], dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(sizes)])
    before = np.concatenate([[0], np.cumsum(sizes * csum[1:])])
    offa = np.concatenate([[0], np.cumsum([2 * s.L + 1 for s in aux])])
    worst, failed, kind = 0.0, 0, None
    triples = sample_triples(shells, aux, sb, se, per_class, seed)
    for (i, j, k) in triples:
        r0 = int(before[i] + sizes[i] * csum[j])
        c0 = int(offa[k] - offa[sb])
        g = T[r0:r0 + int(sizes[i] * sizes[j]), c0:c0 + 2 * aux[k].L + 1]
        g = (g.cpu().numpy() if hasattr(g, "cpu") else np.array(g)).reshape(sizes[i], sizes[j], -1)
        r64, kind = _ref_triple(shells[i], shells[j], aux[k], np.float64)
        r80, _ = _ref_triple(shells[i], shells[j], aux[k], np.longdouble)
        ok, ratio = refmod.block_check(g, r64, r80)
        worst = max(worst, ratio)
        failed += 0 if ok else 1
    return {"blocks": len(triples), "worst_tol_ratio": worst, "failed": failed, "against": kind}


def sample_pairs(shells, per_combo, seed):
    import itertools
    rng = np.random.default_rng(seed)
    by = {}
    for i, s in enumerate(shells):
        by.setdefault((s.L, len(s.exps)), []).append(i)
    return [(int(rng.choice(by[ka])), int(rng.choice(by[kb])))
            for ka, kb in itertools.product(sorted(by), repeat=2) for _ in range(per_combo)]


def parity_2idx(M, key, shells, R, per_combo, seed):
    """M: (ncomp, nbf, nbf) matrix (device tensor or numpy)."""
    from oracle import refmod
    off = np.concatenate([[0], np.cumsum([2 * s.L + 1 for s in shells])])
    worst, failed, kind = 0.0, 0, None
    pairs = sample_pairs(shells, per_combo, seed)
    for (i, j) in pairs:
        g = M[:, off[i]:off[i + 1], off[j]:off[j + 1]]
        g = g.cpu().numpy() if hasattr(g, "cpu") else np.array(g)
        r64, kind = _ref_pair(key, shells[i], shells[j], R, np.float64)
        r80, _ = _ref_pair(key, shells[i], shells[j], R, np.longdouble)
        ok, ratio = refmod.block_check(g, r64, r80)
        worst = max(worst, ratio)
        failed += 0 if ok else 1
    return {"blocks": len(pairs), "worst_tol_ratio": worst, "failed": failed, "against": kind}


def _coul_ref_block(a, b, xyz, q, dtype=np.float64):
    """sum_C q_C (a|1/r_C|b) in one call of the reference's generated function (charge axis = leading broadcast axis)."""
    from oracle import ints, refmod
    c = lambda x: np.asarray(x, dtype=dtype)
    if refmod.available("coul"):
        res = np.zeros((2 * a.L + 1) * (2 * b.L + 1), dtype=dtype)
        da = c(q).reshape(-1, 1, 1) * c(a.coeffs).reshape(1, -1, 1)
        refmod.funcs("coul")[(a.L, b.L)](c(a.exps).reshape(1, -1, 1), da, c(a.center), c(b.exps).reshape(1, 1, -1),
                                         c(b.coeffs).reshape(1, 1, -1), c(b.center), c(xyz).T.reshape(3, -1, 1, 1), res)
        return res.reshape(2 * a.L + 1, 2 * b.L + 1), "reference"
    tot = 0
    for rr, qq in zip(xyz, q):
        tot = tot + qq * ints.coul(a.L, b.L, c(a.exps), c(a.coeffs), c(a.center), c(b.exps), c(b.coeffs), c(b.center),
                                   R=c(rr), sph=True, normalize="cgto")
    return np.asarray(tot).reshape(2 * a.L + 1, 2 * b.L + 1), "port"


def parity_coul(V, shells, xyz, q, seed):
    off = np.concatenate([[0], np.cumsum([2 * s.L + 1 for s in shells])])
    worst, failed, kind = 0.0, 0, None
    pairs = sample_pairs(shells, 1, seed)
    for (i, j) in pairs:
        g = V[off[i]:off[i + 1], off[j]:off[j + 1]]
        g = g.cpu().numpy() if hasattr(g, "cpu") else np.array(g)
        r64, kind = _coul_ref_block(shells[i], shells[j], xyz, q)
        # the charges have both signs: the tolerance scale is the sum of magnitudes (estimated on 64 charges)
        mag, _ = _coul_ref_block(shells[i], shells[j], xyz[:64], np.abs(q[:64]))
        scale = max(float(np.abs(mag).max()) * len(q) / 64.0, float(np.abs(r64).max()))
        ratio = float(np.abs(g - r64).max()) / (1e-14 + 1e-12 * scale)
        worst = max(worst, ratio)
        failed += 0 if ratio <= 1.0 else 1
    return {"blocks": len(pairs), "worst_tol_ratio": worst, "failed": failed, "against": kind,
            "tolerance": "1e-14 + 1e-12 * sum_C |q_C| max|block_C| per shell block"}


# --------------------------------------------------------------------------------------
# CPU baselines of the 2-index keys: the reference's generated numpy code (oracle/_ref) driven like
# testing/intor.py:18-76 on a stratified sample of shell pairs, extrapolated with the class histogram
# --------------------------------------------------------------------------------------
def cpu_baseline_2idx(key, shells, ncomp, per_class=3, seed=11, ncharge=0, xyz=None):
    import collections
    rng = np.random.default_rng(seed)
    cls = collections.defaultdict(list)
    for i, a in enumerate(shells):
        for j in range(i + 1):
            lst = cls[(a.L, len(a.exps), shells[j].L, len(shells[j].exps))]
            lst.append((i, j))
    t_full, n_meas, t_meas, kind = 0.0, 0, 0.0, None
    nsample_c = 4
    for ck, pairs in cls.items():
        ts = []
        for _ in range(per_class):
            i, j = pairs[rng.integers(len(pairs))]
            if key == "coul":   # one call per charge, as test_intor.py:37-43 does
                t0 = time.perf_counter()
                for rr in xyz[:nsample_c]:
                    _, kind = _ref_pair("coul", shells[i], shells[j], rr, np.float64)
                ts.append((time.perf_counter() - t0) * ncharge / nsample_c)
            else:
                t0 = time.perf_counter()
                _, kind = _ref_pair(key, shells[i], shells[j], np.zeros(3), np.float64)
                ts.append(time.perf_counter() - t0)
        t_meas += sum(ts) if key != "coul" else sum(ts) * nsample_c / ncharge
        n_meas += per_class
        t_full += float(np.mean(ts)) * len(pairs)
    nbf = sum(2 * s.L + 1 for s in shells)
    uniq = ncomp * nbf * (nbf + 1) // 2
    return {"value": uniq / t_full, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{n_meas} shell pairs ({per_class} per (La,Ka,Lb,Kb) class, {len(cls)} classes)"
                      + (f" x {nsample_c} of {ncharge} charges" if key == "coul" else "")
                      + f", {t_meas:.2f} core-s measured, extrapolated to {t_full:.3e} core-s for all unique pairs"}


def numba_baseline_worker():
    """Subprocess: the reference's NumbaRenderer output (oracle/_ref/ref_sph_cgto_l2_laux2_numba; generated for
    lmax = 2 because import = eager JIT of every class) on sampled shell pairs with L <= 2, primitive functions
    called as the reference's numba driver does (loop over primitive pairs, result accumulated)."""
    import importlib.util
    from oracle import refmod
    from sympleints_b200 import synthetic
    shells, _ = synthetic.orbital_basis(natoms=50)
    shells = [s for s in shells if s.L <= 2]
    d = refmod.REFDIR / "ref_sph_cgto_l2_laux2_numba"
    out = {}
    for key, (modname, ncomp) in (("ovlp", ("ovlp3d", 1)), ("kin", ("kinetic3d", 1)), ("dpm", ("dipole3d", 3)), ("qpm", ("quadrupole3d", 6))):
        t0 = time.perf_counter()
        spec = importlib.util.spec_from_file_location("nb_" + modname, d / f"{modname}.py")
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        t_jit = time.perf_counter() - t0
        rng = np.random.default_rng(5)
        pairs = [(int(rng.integers(len(shells))), int(rng.integers(len(shells)))) for _ in range(400)]
        R = np.zeros(3)
        nint, t0 = 0, time.perf_counter()
        for (i, j) in pairs:
            a, b = shells[i], shells[j]
            f = getattr(m, f"{modname}_{a.L}{b.L}")
            res = np.zeros(ncomp * (2 * a.L + 1) * (2 * b.L + 1))
            for ea, da in zip(a.exps, a.coeffs):
                for eb, db in zip(b.exps, b.coeffs):
                    f(ea, da, a.center, eb, db, b.center, R, res)
            nint += res.size
        dt = time.perf_counter() - t0
        out[key] = {"value": nint / dt, "unit": UNIT, "cores": 1, "kind": "reference (NumbaRenderer output)",
                    "import_and_jit_s": t_jit,
                    "sample": f"400 random shell pairs of the L <= 2 shells of the benchmark basis, {dt:.3f} s; classes with "
                              "L > 2 not generated (eager JIT of every class at import)"}
    print("NUMBA_JSON " + json.dumps(out))


def numba_baseline(timeout_s=240):
    from oracle import refmod
    if not (refmod.REFDIR / "ref_sph_cgto_l2_laux2_numba" / "ovlp3d.py").exists():
        return {"unavailable": "oracle/_ref/ref_sph_cgto_l2_laux2_numba not generated (needs /root/reference in the build container)"}
    try:
        import numba  # noqa: F401
    except Exception as e:   # pragma: no cover
        return {"unavailable": f"numba not importable: {e}"}
    try:
        p = subprocess.run([sys.executable, str(REPO / "bench.py"), "--numba-worker"], capture_output=True, text=True,
                           timeout=timeout_s, env=dict(os.environ, NUMBA_CACHE_DIR="/tmp/si_numba_cache", CUDA_VISIBLE_DEVICES=""))
        for line in p.stdout.splitlines():
            if line.startswith("NUMBA_JSON "):
                return json.loads(line[len("NUMBA_JSON "):])
        return {"unavailable": "numba worker failed: " + (p.stderr.strip().splitlines() or ["?"])[-1][:200]}
    except subprocess.TimeoutExpired:
        return {"unavailable": f"numba import/JIT exceeded {timeout_s} s"}


# --------------------------------------------------------------------------------------
AUX_ORDER = "class"


def build_system(natoms, aux_order=None):
    """Orbital and auxiliary shells of the synthetic system.  ``aux_order`` "class" (default): the auxiliary shells
    sorted by angular momentum (stable; shells.class_major_order) -- a column permutation of the same tensor that
    the CUDA kernels write faster; "atom": the atom-major order synthetic.aux_basis generates."""
    from sympleints_b200 import synthetic
    from sympleints_b200.shells import class_major_order
    shells, sites = synthetic.orbital_basis(natoms=natoms)
    aux = synthetic.aux_basis(sites)
    if (aux_order or AUX_ORDER) == "class":
        aux = [aux[i] for i in class_major_order(aux)]
    return shells, sites, aux


def run_reference(args, rank, world):
    if rank != 0:
        return
    shells, sites, aux = build_system(args.natoms)
    from sympleints_b200 import workmodel
    work = workmodel.tensor_3c2e_work(shells, aux)
    import multiprocessing as mp
    ncores = len(os.sched_getaffinity(0))
    # load the reference's generated module ONCE in the parent; forked workers inherit it
    _eval_triple_cpu(shells[0], shells[0], aux[0])
    per_class = args.ref_per_class * max(1, ncores // 8)   # a bounded sample that keeps every core busy for seconds
    with mp.get_context("fork").Pool(ncores) as pool:
        for _ in range(args.warmup):
            cpu_baseline_3c2e(shells, aux, per_class=max(1, per_class // 8), nproc=ncores, seed=123, pool=pool)
        vals, t0 = [], time.perf_counter()
        last = None
        for s in range(args.steps):
            last = cpu_baseline_3c2e(shells, aux, per_class=per_class, nproc=ncores, seed=1000 + s, pool=pool)
            vals.append(last["value"])
        wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(shells, aux, work, world),
        "note": "reference generated numpy code (oracle/_ref) on host cores",
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": last["kind"], "sample": last["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def config_dict(shells, aux, work, world):
    return {
        "workload": "3c2e_sph DF tensor (ab|c), lmax=4 lauxmax=5, --sph --normalize cgto, "
                    f"{len(shells)} orbital shells / {sum(2 * s.L + 1 for s in shells)} functions x "
                    f"{len(aux)} aux shells / {sum(2 * s.L + 1 for s in aux)} functions (BASELINE configs[4])",
        "integrals_per_step": work["integrals"], "shell_triples": work["triples"],
        "primitive_triples": work["prim_triples"], "model_flops_per_step": work["flops"],
        "aux_order": "class-major (aux shells sorted by L; columns are a permutation of the atom-major tensor)"
                     if AUX_ORDER == "class" else "atom-major",
        "parallelism": f"aux-shard x{world} (contiguous cost-balanced aux-shell ranges, no collective)",
        "l2": "every step writes 34 GB / N of output (>> 126 MB L2) and re-reads only ~0.1 MB of shell data and "
              "look-up tables; no explicit flush",
    }


def bind_to_gpu_numa_node(torch, local_rank):
    """Run this rank on the CPU cores local to its GPU (sysfs local_cpulist of the PCI device), so that the
    pinned host buffer of the end-to-end leg is allocated on the NUMA node the GPU's PCIe link hangs off.
    (No effect on the single-node-slice boxes of this pool, where all 16 cores are local; the end-to-end
    time there varies 0.68-0.97 s between boxes and runs.)  SI_BENCH_NO_AFFINITY=1 disables it."""
    if os.environ.get("SI_BENCH_NO_AFFINITY") == "1":
        return None
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return txt
    except Exception:
        pass
    return None


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import sympleints_b200
    from sympleints_b200 import dist as sdist
    from sympleints_b200 import workmodel
    from sympleints_b200.shells import shells_to_soa

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the single JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug_%h_%p.log")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    shells, sites, aux = build_system(args.natoms)
    from oracle import boys as oboys   # the checker's Boys table (the reference's): parity at 1e-12 needs the same table
    eng = sympleints_b200.get_engine(local_rank, 4, 5, boys_table=oboys.table())
    ob, ab = eng.basis(shells), eng.basis(aux)
    costs = workmodel.aux_shell_costs(shells, aux)
    ranges = sdist.partition_aux_shells(costs, world)
    my = ranges[rank]
    if args.shard_of > 1:   # debugging aid: time rank 0's shard of an N-way partition on one GPU
        my = sdist.partition_aux_shells(costs, args.shard_of)[0]
    work = workmodel.tensor_3c2e_work(shells, aux)
    mywork = workmodel.tensor_3c2e_work(shells, aux, my)
    ncol = eng.aux_cols(ab, *my)
    out = torch.empty((ob.packed_rows, ncol), dtype=torch.float64, device=eng.tdev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.int3c2e(ob, ab, shell_range=my, layout="packed", out=out)

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    tw1 = time.time()
    launches = eng.launch_count() - l0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=eng.tdev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    ms_per_step = ms_max / args.steps
    value = work["integrals"] / (ms_per_step * 1e-3)

    # ---- parity of the tensor the last timed step wrote (every rank samples its own shard) ----
    def reduce_parity(par):
        v = torch.tensor([par["blocks"], par["failed"]], dtype=torch.float64, device=eng.tdev)
        w = torch.tensor([par["worst_tol_ratio"]], dtype=torch.float64, device=eng.tdev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        return {"blocks": int(v[0].item()), "worst_tol_ratio": float(w.item()), "failed": int(v[1].item()),
                "against": par["against"],
                "criterion": "per shell block: max|x-ref64| <= 1e-14 + 1e-12 max|ref64|, or within 10x the reference's own "
                             "float64 noise of its extended-precision evaluation (tests/golden_util.py)"}

    parity = None
    if not args.no_parity:
        parity = reduce_parity(parity_3c2e(out, shells, aux, my[0], my[1], 2 if world == 1 else 1, seed=17 + rank))

    # ---- end to end through the C ABI with host buffers (H2D of the shell data, D2H of the slab) ----
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty((ob.packed_rows, ncol), dtype=torch.float64, pin_memory=True)
            pinned = True
        except Exception:
            host = torch.empty((ob.packed_rows, ncol), dtype=torch.float64)
            pinned = False
        host_np = host.numpy()
        soa_o, soa_a = shells_to_soa(shells), shells_to_soa(aux)
        h2d = int(sum(x.nbytes for x in soa_o) + sum(x.nbytes for x in soa_a))

        def e2e_step():
            o2, a2 = eng.basis(shells), eng.basis(aux)   # H2D of this step's inputs
            eng.int3c2e_host(o2, a2, shell_range=my, layout="packed", out=host_np)
            o2.close()
            a2.close()

        e2e_step()  # warm-up
        barrier()
        n_e2e = max(1, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        t_basis = 0.0
        for _ in range(n_e2e):
            tb = time.perf_counter()
            o2, a2 = eng.basis(shells), eng.basis(aux)   # H2D of this step's inputs
            t_basis += time.perf_counter() - tb
            eng.int3c2e_host(o2, a2, shell_range=my, layout="packed", out=host_np)
            o2.close()
            a2.close()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=eng.tdev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        breakdown = eng.host_breakdown()
        breakdown["basis_create_h2d_s"] = t_basis / n_e2e
        # D2H ceiling: the same bytes from a device buffer into the same pinned host buffer, plain cudaMemcpyAsync in
        # chunks of the staging size, all ranks at the same time (the host side is shared between the GPUs of a box)
        barrier()
        t0 = time.perf_counter()
        flat_h, flat_d = host.view(-1), out.view(-1)
        chunk = 1 << 27   # 1 GiB of doubles
        for o in range(0, flat_h.numel(), chunk):
            flat_h[o:o + chunk].copy_(flat_d[o:o + chunk], non_blocking=True)
        torch.cuda.synchronize()
        t_probe_local = time.perf_counter() - t0
        barrier()
        dp = torch.tensor([t_probe_local], dtype=torch.float64, device=eng.tdev)
        if world > 1:
            dist.all_reduce(dp, op=dist.ReduceOp.MAX)
        d2h_peak = work["bytes"] / float(dp.item()) / 1e9 if world > 1 else mywork["bytes"] / float(dp.item()) / 1e9
        d2h_gbs = work["bytes"] / float(dt.item()) / 1e9
        e2e = {"value": work["integrals"] / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(mywork["bytes"]), "steps": n_e2e, "s_per_step": float(dt.item()),
               "d2h_GBs": d2h_gbs, "d2h_peak_GBs": d2h_peak, "frac_of_d2h_peak": d2h_gbs / d2h_peak,
               "d2h_peak_def": "aggregate over the N ranks: plain pinned cudaMemcpyAsync of the same slab bytes, all ranks "
                               "concurrently, max over ranks",
               "breakdown": breakdown,
               "pinned_host_buffer": pinned, "cpu_affinity": numa,
               "path": "si_basis_create (H2D) + si_int3c2e_h (row-chunked compute overlapped with contiguous cudaMemcpyAsync D2H)"}
        if not args.no_parity:
            eng.int3c2e_host(ob, ab, shell_range=my, layout="packed", out=host_np)
            e2e["parity"] = reduce_parity(parity_3c2e(host_np, shells, aux, my[0], my[1], 1, seed=91 + rank))
        del host

    # ---- on-request all-gather of the sharded tensor (NCCL over NVLink), timed separately: NOT part of `value` ----
    allgather = None
    if world > 1 and not args.no_allgather:
        cols = sdist.shard_columns([2 * s.L + 1 for s in aux], ranges)
        try:
            full = torch.empty((ob.packed_rows, cols[-1][1]), dtype=torch.float64, device=eng.tdev)
            sdist.allgather_slabs(out, cols, out=full)   # warm-up (NCCL channel set-up)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            sdist.allgather_slabs(out, cols, out=full)
            g1.record()
            barrier()
            tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=eng.tdev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            # every rank now holds the whole tensor: compare a foreign column block with what its owner computed
            other = (rank + 1) % world
            c0, c1 = cols[other]
            chk = parity_3c2e(full[:, c0:c1], shells, aux, ranges[other][0], ranges[other][1], 1, seed=55 + rank)
            okv = torch.tensor([chk["failed"]], dtype=torch.float64, device=eng.tdev)
            dist.all_reduce(okv, op=dist.ReduceOp.SUM)
            recv = work["bytes"] - mywork["bytes"]
            allgather = {"ms": float(tg.item()), "bytes_received_per_rank": int(recv),
                         "GBs_per_rank": recv / (float(tg.item()) * 1e-3) / 1e9,
                         "how": "dist.all_gather_into_tensor on slabs padded to the widest shard, row-blocked staging, "
                                "then strided copies into the column ranges (sympleints_b200/dist.py); excluded from value",
                         "parity_of_received_blocks": {"blocks": chk["blocks"] * world, "failed": int(okv.item())}}
            del full
        except Exception as e:   # e.g. out of memory for the full tensor next to the shard
            allgather = {"unavailable": str(e)[:200]}
        torch.cuda.empty_cache()

    if rank == 0:
        fp64_peak = eng.fp64_peak(0.5)
        peaks = {}
        try:
            peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        tfl = work["flops"] / (ms_per_step * 1e-3) / 1e12 / world   # per GPU
        # DRAM traffic of the kernel family: NOT measured in this run -- read from the committed ncu launch list of the same
        # command (profiles/r02_end_launches_full.csv for the class-major aux order, r02_final_launches_full.csv for the
        # atom-major one; summed over the 90 class launches of one N=1 step)
        traffic, traffic_note = None, None
        tname = "r02_end" if AUX_ORDER == "class" else "r02_final"
        try:
            tj = json.loads((REPO / "profiles" / f"{tname}_traffic.json").read_text())
            if world == 1 and args.shard_of == 1 and args.natoms == 50:
                traffic = (tj["dram_read_GB_per_step"] + tj["dram_write_GB_per_step"]) * 1e9
                traffic_note = ("FROM profiles/ (not measured in this run): sum of dram__bytes_read+write over the 90 launches of "
                                f"one step in profiles/{tname}_launches_full.csv (a launch's last dirty L2 lines are written back after it "
                                "ends and are not in its count); algorithmic bytes per step %.4g" % mywork["bytes"])
        except Exception:
            pass
        roofline = {
            "bound": "fp64", "kernel": "g3_kernel_<La><Lb><Lc> (generated class-specialised 3c2e kernels; every launch of the step is one of the 90 classes, serialised on one stream)",
            "achieved": tfl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfl / fp64_peak if fp64_peak else None,
            "traffic": traffic, "traffic_source": f"profiles/{tname}_traffic.json" if traffic else None, "traffic_note": traffic_note,
            "peak_source": "DFMA microbenchmark si_fp64_peak() measured in this run (FP64 is not in MEASURED_PEAKS.json)",
            "achieved_def": "reference-graph flop model (SURVEY 8d, 88.3 flop/integral) per GPU / CUDA-event step time",
            "hbm_write_GBs": mywork["bytes"] / (ms_per_step * 1e-3) / 1e9,
            "hbm_peak_GBs": peaks.get("hbm_gbs"),
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(shells, aux, work, world),
            "note": "Boys table = the reference's table (oracle/boys.py, bit-identical to get_boys_func(14)) so that the "
                    "parity sample can be judged at 1e-12",
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "parity": parity,
            "shard": {"aux_shell_ranges": ranges, "columns_rank0": ncol},
            "allgather": allgather,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline_3c2e(
                shells, aux, per_class=args.ref_per_class, nproc=1, seed=7).items()
                if k in ("value", "unit", "cores", "kind", "sample")}
        if world == 1 and not args.no_other_keys:
            line["other_keys"] = bench_other_keys(eng, shells, sites, aux, ob, ab, fp64_peak, peaks, args)
            if AUX_ORDER == "class" and args.shard_of == 1:
                # the same tensor with the columns in the atom-major order the synthetic generator emits
                _, _, aux_atom = build_system(args.natoms, "atom")
                ab2 = eng.basis(aux_atom)
                out2 = torch.empty((ob.packed_rows, ab2.nbf_sph), dtype=torch.float64, device=eng.tdev)
                for _ in range(2):
                    eng.int3c2e(ob, ab2, layout="packed", out=out2)
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(3):
                    eng.int3c2e(ob, ab2, layout="packed", out=out2)
                a1.record()
                torch.cuda.synchronize()
                line["other_keys"]["3c2e_atom_major_aux_order"] = {"ms_per_step": a0.elapsed_time(a1) / 3,
                    "note": "same integrals, aux shells in atom-major order (column permutation of the timed tensor)"}
                del out2
        print(json.dumps(line))
        bad = (parity or {}).get("failed", 0) + ((e2e or {}).get("parity") or {}).get("failed", 0)
        bad += sum((v.get("parity") or {}).get("failed", 0) for v in (line.get("other_keys") or {}).values() if isinstance(v, dict))
        bad += ((allgather or {}).get("parity_of_received_blocks") or {}).get("failed", 0)
    else:
        bad = 0
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if bad:
        sys.stderr.write(f"PARITY FAILURE: {bad} sampled blocks outside the tolerance\n")
        sys.exit(3)


def bench_other_keys(eng, shells, sites, aux, ob, ab, fp64_peak, peaks, args, reps=20):
    """BASELINE configs 2-4 on one GPU: the 1e set (ovlp, kin, dpm, qpm), coul with 1024 charges and the 2c2e metric.
    Per key: kernel time (median of `reps` replays, CUDA events, output pre-allocated), integrals/s (unique by the
    a>=b symmetry), roofline, parity sample of the matrix the last replay wrote, end-to-end time through the host-buffer
    C-ABI entry (basis upload + compute + D2H), and the reference's generated numpy code on one host core."""
    import torch

    from sympleints_b200 import synthetic, workmodel

    res = {}
    nbf, naux = ob.nbf_sph, ab.nbf_sph
    hbm = peaks.get("hbm_gbs")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.tdev)   # > 126 MB L2

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()   # L2 flush between replays (the outputs of these keys fit into L2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return float(np.median(ts))

    def time_host(fn, n=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n

    uniq = nbf * (nbf + 1) // 2
    soa = lambda sh: int(sum(x.nbytes for x in __import__("sympleints_b200.shells", fromlist=["x"]).shells_to_soa(sh)))
    R = np.array([0.0, 0.0, 0.0])
    for key, ncomp in (("ovlp", 1), ("kin", 1), ("dpm", 3), ("qpm", 6)):
        buf = torch.empty((ncomp, nbf, nbf), dtype=torch.float64, device=eng.tdev)
        l0 = eng.launch_count()
        t = timeit(lambda: eng.int1e(key, ob, out=buf, R=R))
        nl = (eng.launch_count() - l0) // (reps + 1)
        gbs = 8 * ncomp * nbf * nbf / t / 1e9

        def host_call():
            b2 = eng.basis(shells)
            eng.int1e_host(key, b2, R=R)
            b2.close()

        th = time_host(host_call)
        res[key] = {"s": t, "integrals_per_s": ncomp * uniq / t, "gpu_launches": int(nl),
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm if hbm else None,
                                 "bytes": 8 * ncomp * nbf * nbf},
                    "e2e": {"value": ncomp * uniq / th, "unit": UNIT, "s": th, "h2d_bytes_per_step": soa(shells),
                            "d2h_bytes_per_step": 8 * ncomp * nbf * nbf, "path": "si_basis_create + si_int1e_h"}}
        if not args.no_parity:
            res[key]["parity"] = parity_2idx(buf, key, shells, R, 2, seed=23)
        if not args.no_cpu_baseline:
            res[key]["cpu_baseline"] = cpu_baseline_2idx(key, shells, ncomp)
    # one launch for the four operators (the fused 1e kernel writes all of them in one pass)
    if hasattr(eng, "int1e_set"):
        bufs = torch.empty((11, nbf, nbf), dtype=torch.float64, device=eng.tdev)
        l0 = eng.launch_count()
        t = timeit(lambda: eng.int1e_set(ob, out=bufs, R=R))
        nl = (eng.launch_count() - l0) // (reps + 1)
        gbs = 8 * 11 * nbf * nbf / t / 1e9
        res["1e_set_fused"] = {"s": t, "integrals_per_s": 11 * uniq / t, "gpu_launches": int(nl),
                               "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                                            "frac": gbs / hbm if hbm else None, "bytes": 8 * 11 * nbf * nbf},
                               "what": "ovlp + kin + dpm + qpm (1 + 1 + 3 + 6 components) in one call"}
        if not args.no_parity:
            pars = [parity_2idx(bufs[o:o + n], k, shells, R, 1, seed=29) for k, o, n in
                    (("ovlp", 0, 1), ("kin", 1, 1), ("dpm", 2, 3), ("qpm", 5, 6))]
            res["1e_set_fused"]["parity"] = {"blocks": sum(p_["blocks"] for p_ in pars), "failed": sum(p_["failed"] for p_ in pars),
                                             "worst_tol_ratio": max(p_["worst_tol_ratio"] for p_ in pars), "against": pars[0]["against"]}
    # 2c2e metric
    buf = torch.empty((naux, naux), dtype=torch.float64, device=eng.tdev)
    l0 = eng.launch_count()
    t = timeit(lambda: eng.int2c2e(ab, out=buf))
    nl = (eng.launch_count() - l0) // (reps + 1)
    gbs = 8 * naux * naux / t / 1e9

    def host_2c():
        b2 = eng.basis(aux)
        eng.int1e_host("2c2e", b2)
        b2.close()

    th = time_host(host_2c)
    uniq_a = naux * (naux + 1) // 2
    res["2c2e"] = {"s": t, "integrals_per_s": uniq_a / t, "gpu_launches": int(nl),
                   "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm if hbm else None,
                                "bytes": 8 * naux * naux},
                   "e2e": {"value": uniq_a / th, "unit": UNIT, "s": th, "h2d_bytes_per_step": soa(aux),
                           "d2h_bytes_per_step": 8 * naux * naux, "path": "si_basis_create + si_int2c2e_h"}}
    if not args.no_parity:
        res["2c2e"]["parity"] = parity_2idx(buf.reshape(1, naux, naux), "2c2e", aux, np.zeros(3), 2, seed=31)
    if not args.no_cpu_baseline:
        res["2c2e"]["cpu_baseline"] = cpu_baseline_2idx("2c2e", aux, 1, per_class=2)
    # coul, 1024 charges
    xyz, q = synthetic.point_charges(sites, n=1024)
    buf = torch.empty((nbf, nbf), dtype=torch.float64, device=eng.tdev)
    l0 = eng.launch_count()
    t = timeit(lambda: eng.coul(ob, xyz, q, out=buf))
    nl = (eng.launch_count() - l0) // (reps + 1)
    flops = 6.68e10

    def host_coul():
        b2 = eng.basis(shells)
        eng.coul_host(b2, xyz, q)
        b2.close()

    th = time_host(host_coul)
    res["coul_1024"] = {"s": t, "integrals_per_s": uniq / t, "integral_charges_per_s": uniq * 1024 / t, "gpu_launches": int(nl),
                        "roofline": {"bound": "fp64", "achieved": flops / t / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                                     "frac": flops / t / 1e12 / fp64_peak if fp64_peak else None,
                                     "note": "model flops 6.68e10 (SURVEY 8d, VRR+HRR graph); the generated "
                                             "McMurchie-Davidson kernels (codegen/gencoul.py) execute fewer"},
                        "e2e": {"value": uniq / th, "unit": UNIT, "s": th, "h2d_bytes_per_step": soa(shells) + 4 * 8 * 1024,
                                "d2h_bytes_per_step": 8 * nbf * nbf, "path": "si_basis_create + si_coul_h"}}
    if not args.no_parity:
        res["coul_1024"]["parity"] = parity_coul(buf, shells, xyz, q, seed=37)
    if not args.no_cpu_baseline:
        res["coul_1024"]["cpu_baseline"] = cpu_baseline_2idx("coul", shells, 1, per_class=2, ncharge=1024, xyz=xyz)
        res["coul_1024"]["cpu_baseline"]["unit_note"] = "unique integrals/s after the sum over 1024 charges"
    if not args.no_numba and not args.no_cpu_baseline:
        nb = numba_baseline()
        for key in ("ovlp", "kin", "dpm", "qpm"):
            res[key]["cpu_baseline_numba"] = nb.get(key, nb)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--natoms", type=int, default=50, help="50 = the BASELINE system; smaller only for debugging")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-keys", action="store_true", help="skip the 1e set, coul and 2c2e legs (N=1)")
    ap.add_argument("--other-keys", action="store_true", help="(default now; kept for compatibility)")
    ap.add_argument("--no-allgather", action="store_true", help="skip the separately timed all-gather of the shards (N > 1)")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled oracle check of the timed output")
    ap.add_argument("--no-numba", action="store_true", help="skip the Numba-renderer baseline of the 1e keys")
    ap.add_argument("--numba-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--ref-per-class", type=int, default=16)
    ap.add_argument("--aux-order", default="class", choices=["class", "atom"],
                    help="order of the auxiliary shells = column order of the tensor (class: sorted by L)")
    ap.add_argument("--shard-of", type=int, default=1, help="debug: run only rank 0's shard of an N-way partition")
    args = ap.parse_args()
    global AUX_ORDER
    AUX_ORDER = args.aux_order
    if args.numba_worker:
        numba_baseline_worker()
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
