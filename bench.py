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
def _load_ref_3c2e():
    """The reference's own generated module if it was produced in the build container."""
    import importlib.util
    p = REPO / "oracle" / "_ref" / "ref_sph_cgto_l4_laux5_python" / "int3c2e3d_sph.py"
    if not p.exists():
        return None
    spec = importlib.util.spec_from_file_location("ref_int3c2e3d_sph", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.int3c2e3d_sph


_REF = {}


def _eval_triple_cpu(a, b, c):
    """One contracted block on the CPU; returns (seconds, kind)."""
    if "fd" not in _REF:
        _REF["fd"] = _load_ref_3c2e()
    fd = _REF["fd"]
    t0 = time.perf_counter()
    if fd is not None:
        if a.L < b.L:  # the reference's own 3-index wrappers are broken: swap + transpose
            a, b = b, a
        res = np.empty((2 * a.L + 1) * (2 * b.L + 1) * (2 * c.L + 1))
        fd[(a.L, b.L, c.L)](a.exps[:, None, None], a.coeffs[:, None, None], a.center,
                            b.exps[None, :, None], b.coeffs[None, :, None], b.center,
                            c.exps[None, None, :], c.coeffs[None, None, :], c.center, np.zeros(3), res)
        kind = "reference"
    else:
        from oracle import ints
        ints.int3c2e_sph(a.L, b.L, c.L, a.exps, a.coeffs, a.center, b.exps, b.coeffs, b.center,
                         c.exps, c.coeffs, c.center)
        kind = "port"
    return time.perf_counter() - t0, kind


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
def build_system(natoms):
    from sympleints_b200 import synthetic
    shells, sites = synthetic.orbital_basis(natoms=natoms)
    aux = synthetic.aux_basis(sites)
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
        "config": config_dict(shells, aux, work, world, "reference generated numpy code (oracle/_ref) on host cores"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": last["kind"], "sample": last["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def config_dict(shells, aux, work, world, note):
    return {
        "workload": "3c2e_sph DF tensor (ab|c), lmax=4 lauxmax=5, --sph --normalize cgto, "
                    f"{len(shells)} orbital shells / {sum(2 * s.L + 1 for s in shells)} functions x "
                    f"{len(aux)} aux shells / {sum(2 * s.L + 1 for s in aux)} functions (BASELINE configs[4])",
        "integrals_per_step": work["integrals"], "shell_triples": work["triples"],
        "primitive_triples": work["prim_triples"], "model_flops_per_step": work["flops"],
        "parallelism": f"aux-shard x{world} (contiguous cost-balanced aux-shell ranges, no collective)",
        "l2": "every step writes 34 GB / N of output (>> 126 MB L2) and re-reads only ~0.1 MB of shell data and "
              "look-up tables; no explicit flush",
        "note": note,
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
    eng = sympleints_b200.get_engine(local_rank, 4, 5)
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
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=eng.tdev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": work["integrals"] / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(mywork["bytes"]), "steps": n_e2e, "s_per_step": float(dt.item()),
               "pinned_host_buffer": pinned, "cpu_affinity": numa,
               "path": "si_basis_create (H2D) + si_int3c2e_h (row-chunked compute overlapped with contiguous cudaMemcpyAsync D2H)"}
        del host

    if rank == 0:
        fp64_peak = eng.fp64_peak(0.5)
        peaks = {}
        try:
            peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        tfl = work["flops"] / (ms_per_step * 1e-3) / 1e12 / world   # per GPU
        # DRAM traffic of the kernel family from the committed ncu launch list of the same command
        # (profiles/r01_final2_launches_full.csv, summed over the 90 class launches of one N=1 step)
        traffic, traffic_note = None, None
        try:
            tj = json.loads((REPO / "profiles" / "r01_final2_traffic.json").read_text())
            if world == 1 and args.shard_of == 1 and args.natoms == 50:
                traffic = (tj["dram_read_GB_per_step"] + tj["dram_write_GB_per_step"]) * 1e9
                traffic_note = ("bytes per step = sum of dram__bytes_read+write over the 90 launches (ncu, profiles/"
                                "r01_final2_launches_full.csv); algorithmic bytes per step %.4g" % mywork["bytes"])
        except Exception:
            pass
        roofline = {
            "bound": "fp64", "kernel": "g3_kernel_<La><Lb><Lc> (generated class-specialised 3c2e kernels; every launch of the step is one of the 90 classes, serialised on one stream)",
            "achieved": tfl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfl / fp64_peak if fp64_peak else None,
            "traffic": traffic, "traffic_note": traffic_note,
            "peak_source": "DFMA microbenchmark si_fp64_peak() measured in this run (FP64 is not in MEASURED_PEAKS.json)",
            "achieved_def": "reference-graph flop model (SURVEY 8d, 88.3 flop/integral) per GPU / CUDA-event step time",
            "hbm_write_GBs": mywork["bytes"] / (ms_per_step * 1e-3) / 1e9,
            "hbm_peak_GBs": peaks.get("hbm_gbs"),
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(shells, aux, work, world, "Boys table built by the library (series + downward recursion)"),
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e,
            "shard": {"aux_shell_ranges": ranges, "columns_rank0": ncol},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline_3c2e(
                shells, aux, per_class=args.ref_per_class, nproc=1, seed=7).items()
                if k in ("value", "unit", "cores", "kind", "sample")}
        if world == 1 and args.other_keys:
            line["other_keys"] = bench_other_keys(eng, shells, sites, aux, ob, ab, fp64_peak, peaks)
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def bench_other_keys(eng, shells, sites, aux, ob, ab, fp64_peak, peaks, reps=5):
    """1e set, coul (1024 charges) and 2c2e on one GPU (BASELINE configs 1-3): integrals/s + roofline."""
    import torch

    from sympleints_b200 import synthetic

    res = {}
    nbf, naux = ob.nbf_sph, ab.nbf_sph
    hbm = peaks.get("hbm_gbs")

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return float(np.median(ts))

    uniq = nbf * (nbf + 1) // 2
    for key, ncomp in (("ovlp", 1), ("kin", 1), ("dpm", 3), ("qpm", 6)):
        buf = torch.empty((ncomp, nbf, nbf), dtype=torch.float64, device=eng.tdev)
        t = timeit(lambda: eng.int1e(key, ob, out=buf))
        gbs = 8 * ncomp * nbf * nbf / t / 1e9
        res[key] = {"s": t, "integrals_per_s": ncomp * uniq / t, "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm,
                    "unit": "GB/s", "frac": gbs / hbm if hbm else None}}
    buf = torch.empty((naux, naux), dtype=torch.float64, device=eng.tdev)
    t = timeit(lambda: eng.int2c2e(ab, out=buf))
    gbs = 8 * naux * naux / t / 1e9
    res["2c2e"] = {"s": t, "integrals_per_s": naux * (naux + 1) // 2 / t, "roofline": {"bound": "hbm", "achieved": gbs,
                   "peak": hbm, "unit": "GB/s", "frac": gbs / hbm if hbm else None}}
    xyz, q = synthetic.point_charges(sites, n=1024)
    buf = torch.empty((nbf, nbf), dtype=torch.float64, device=eng.tdev)
    t = timeit(lambda: eng.coul(ob, xyz, q, out=buf))
    flops = 6.68e10
    res["coul_1024_charges"] = {"s": t, "integrals_per_s": uniq / t, "integral_charges_per_s": uniq * 1024 / t,
                                "roofline": {"bound": "fp64", "achieved": flops / t / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                                             "frac": flops / t / 1e12 / fp64_peak if fp64_peak else None,
                                             "note": "model flops 6.68e10 (SURVEY 8d, VRR+HRR graph); the generated "
                                                     "McMurchie-Davidson kernels (codegen/gencoul.py) execute fewer"}}
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
    ap.add_argument("--other-keys", action="store_true", help="also time the 1e set, coul and 2c2e (N=1)")
    ap.add_argument("--ref-per-class", type=int, default=16)
    ap.add_argument("--shard-of", type=int, default=1, help="debug: run only rank 0's shard of an N-way partition")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
