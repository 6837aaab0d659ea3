#!/usr/bin/env python
"""Per-device-function breakdown of an `ncu --set full --import-source on` capture of step_kernel.

usage: python tools/ncu_by_function.py <report.ncu-rep> [liblocosim_cuda.so] [kernel-substring]

The phases of the engine are __noinline__ device functions, i.e. local FUNC symbols inside the kernel's .text section;
this joins the SASS page of the report (address, instructions executed, stall samples) with the cubin's symbol table.
The report and the .so must come from the same build.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def symbols(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout
    syms, kname = [], None
    for line in out.splitlines():
        p = line.split()
        if len(p) < 8 or p[3] != "FUNC":
            continue
        name, off = p[-1], int(p[1], 16)
        size = int(p[2], 16) if p[2].startswith("0x") else int(p[2])
        if name.startswith("_Z11step_kernel") and kernel_sub in name:
            kname = name
            syms.append((off, size, "step_kernel(main)"))
        elif name.startswith("$_Z11step_kernel") and kernel_sub in name.split("$")[1]:
            syms.append((off, size, re.sub(r"^_Z\d+", "", name.split("$")[2])[:24]))
    assert kname, "kernel not found"
    syms.sort()
    # the main function's symbol spans the whole section; its own code is [0, first callee)
    return syms


def main():
    rep = sys.argv[1]
    so = sys.argv[2] if len(sys.argv) > 2 else "loco_mujoco_b200/liblocosim_cuda.so"
    ksub = sys.argv[3] if len(sys.argv) > 3 else "CfgEllEuler"
    syms = symbols(so, ksub)
    callee = [s for s in syms if s[2] != "step_kernel(main)"]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    lines = txt.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
    rows = list(rd)
    base = int(rows[0]["Address"], 16)
    stall_cols = [c for c in rd.fieldnames if c.startswith("stall_") and "Not Issued" not in c]
    agg = collections.defaultdict(lambda: collections.Counter())
    for r in rows:
        off = int(r["Address"], 16) - base
        fn = "step_kernel(main)"
        for o, sz, name in callee:
            if o <= off < o + sz:
                fn = name
                break
        a = agg[fn]
        a["inst"] += int(r["Instructions Executed"] or 0)
        a["thr"] += int(r["Thread Instructions Executed"] or 0)
        a["samples"] += int(r["# Samples"] or 0)
        for c in stall_cols:
            a[c] += int(r[c] or 0)
    tot_i = sum(a["inst"] for a in agg.values())
    tot_s = sum(a["samples"] for a in agg.values())
    print("%-26s %8s %7s %7s %6s  top stalls (samples)" % ("function", "Minst", "inst%", "time%", "thr/i"))
    for fn, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        top = sorted(((a[c], c[6:]) for c in stall_cols), reverse=True)[:4]
        print("%-26s %8.1f %6.1f%% %6.1f%% %6.1f  %s" % (
            fn, a["inst"] / 1e6, 100.0 * a["inst"] / tot_i, 100.0 * a["samples"] / max(tot_s, 1),
            a["thr"] / max(a["inst"], 1), " ".join("%s=%d" % (n, v) for v, n in top)))
    print("total warp-instructions %.1f M, samples %d" % (tot_i / 1e6, tot_s))


if __name__ == "__main__":
    main()
