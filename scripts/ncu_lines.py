"""Aggregate an `ncu --set full --import-source on` capture per CUDA source line.

usage: python scripts/ncu_lines.py report.ncu-rep KERNEL_REGEX [ntop]
Prints the stall totals, the shared-memory wavefront / bank-conflict totals and the source lines with
the most warp-stall samples (needs -lineinfo at compile time)."""
import csv
import io
import subprocess
import sys


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    hdr, data, name = None, [], kern
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "Function Name":
            name = r[1]
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) >= len(hdr) - 2 and r[0] != "":
            data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    S, IE, WS, WX = ix["# Samples"], ix["Instructions Executed"], ix["L1 Wavefronts Shared"], ix["L1 Wavefronts Shared Excessive"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(f(r[S]) for r in data)
    print(f"== {name}: {tot:.0f} samples; shared wavefronts {sum(f(r[WS]) for r in data):.3g}, "
          f"excessive (bank conflicts) {sum(f(r[WX]) for r in data):.3g}")
    print("   stall samples: " + ", ".join(f"{h[6:]} {v:.0f}" for v, h in
                                           sorted(((sum(f(r[ix[h]]) for r in data), h) for h in stalls), reverse=True)[:7]))
    lsb, ssb = ix["stall_long_sb"], ix["stall_short_sb"]
    for r in sorted(data, key=lambda r: -f(r[S]))[:ntop]:
        print("%5.1f%% inst=%9d shW=%9d exc=%9d long_sb=%4d short_sb=%4d L%-5s %s" % (
            100 * f(r[S]) / tot, f(r[IE]), f(r[WS]), f(r[WX]), f(r[lsb]), f(r[ssb]), r[0], r[1].strip()[:95]))


if __name__ == "__main__":
    main()
