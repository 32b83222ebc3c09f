"""Build library variants of the generated coul kernels (lanes-per-pair shift x launch bounds) for per-class
timing on the GPU box (SI_GC_PROFILE=1), and merge the timings into codegen/tuning_coul.json."""
import json, os, subprocess, sys, shutil
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
VARIANTS = [(sh, mb) for sh in (-2, -1, 0) for mb in (1, 2, 3)]

def build():
    out = REPO / "sympleints_b200" / "lib" / "variants_gc"
    out.mkdir(parents=True, exist_ok=True)
    for sh, mb in VARIANTS:
        env = dict(os.environ, SI_GC_EPT_SHIFT=str(sh), SI_GC_MB=str(mb), SI_G3_NO_TUNING="")
        env.pop("SI_G3_NO_TUNING")
        env["SI_GC_NO_TUNING"] = "1"
        subprocess.check_call([sys.executable, "-c", "from sympleints_b200 import build; build.build_library()"], cwd=REPO, env=env)
        shutil.copy(REPO / "sympleints_b200" / "lib" / "libsympleints_b200.so", out / f"lib_sh{sh}_mb{mb}.so")
        print("built", sh, mb, flush=True)

def merge(logdir):
    sys.path.insert(0, str(REPO))
    best = {}
    for sh, mb in VARIANTS:
        f = Path(logdir) / f"tunegc_sh{sh}_mb{mb}.log"
        if not f.exists():
            continue
        times = {}
        for line in open(f):
            if line.startswith("GCPROF"):
                parts = line.split()
                times.setdefault(parts[1], []).append(float(parts[2]))
        for cls, v in times.items():
            t = min(v)
            if cls not in best or t < best[cls][0]:
                best[cls] = (t, sh, mb)
    os.environ["SI_GC_NO_TUNING"] = "1"
    from sympleints_b200.codegen import gencoul as g
    tuning, tot = {}, 0
    for cls, (t, sh, mb) in sorted(best.items()):
        La, Lb = (int(c) for c in cls)
        os.environ["SI_GC_EPT_SHIFT"] = str(sh)
        g.TUNING = {}
        tuning[cls] = {"ept": g.CoulGen(La, Lb).EPT, "mb": mb, "ms": round(t, 4)}
        tot += t
    json.dump(tuning, open(REPO / "sympleints_b200" / "codegen" / "tuning_coul.json", "w"), indent=0, sort_keys=True)
    print("classes", len(tuning), "sum of best times (ms)", tot)

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        merge(sys.argv[2])
