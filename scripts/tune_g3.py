"""Build library variants of the generated 3c2e kernels (lanes-per-tuple shift x launch bounds) for
per-class timing on the GPU box (SI_G3_PROFILE=1), and merge the timings into codegen/tuning.json."""
import json, os, subprocess, sys, shutil
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
VARIANTS = [(sh, mb, 36) for sh in (-2, -1, 0, 1) for mb in (2, 3, 4)]
if os.environ.get('TUNE_ONLY_EXTRA'):
    VARIANTS = [(sh, 2, 60) for sh in (0, 1)]

def build():
    out = REPO / "sympleints_b200" / "lib" / "variants"
    out.mkdir(parents=True, exist_ok=True)
    for sh, mb, yr in VARIANTS:
        env = dict(os.environ, SI_G3_EPT_SHIFT=str(sh), SI_G3_MB=str(mb), SI_G3_YREG=str(yr), SI_G3_NO_TUNING="1")
        subprocess.check_call([sys.executable, "-c", "from sympleints_b200 import build; build.build_library()"], cwd=REPO, env=env)
        shutil.copy(REPO / "sympleints_b200" / "lib" / "libsympleints_b200.so", out / f"lib_sh{sh}_mb{mb}_y{yr}.so")
        print("built", sh, mb, yr, flush=True)

def merge(logdir):
    sys.path.insert(0, str(REPO))
    best = {}
    allv = VARIANTS
    for sh, mb, yr in allv:
        f = Path(logdir) / f"tune_sh{sh}_mb{mb}_y{yr}.log"
        if not f.exists() and yr == 36:
            f = Path(logdir) / f"tune_sh{sh}_mb{mb}.log"
        if not f.exists():
            continue
        times = {}
        for line in open(f):
            if line.startswith("G3PROF"):
                _, cls, ms = line.split()
                times.setdefault(cls, []).append(float(ms))
        for cls, v in times.items():
            t = min(v)
            if cls not in best or t < best[cls][0]:
                best[cls] = (t, sh, mb, yr)
    os.environ["SI_G3_NO_TUNING"] = "1"
    from sympleints_b200.codegen import gen3c2e as g
    tuning = {}
    tot = 0
    for cls, (t, sh, mb, yr) in sorted(best.items()):
        La, Lb, Lc = (int(c) for c in cls)
        os.environ["SI_G3_EPT_SHIFT"] = str(sh)
        ept = g.ClassGen(La, Lb, Lc).EPT
        tuning[cls] = {"ept": ept, "mb": mb, "yreg": yr, "ms": round(t, 4)}
        tot += t
    json.dump(tuning, open(REPO / "sympleints_b200" / "codegen" / "tuning.json", "w"), indent=0, sort_keys=True)
    print("classes", len(tuning), "sum of best times (ms)", tot)

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        merge(sys.argv[2])
