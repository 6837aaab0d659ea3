#!/usr/bin/env python
"""Per-source-line table (instructions executed, stall samples) of an `ncu --set full --import-source on` capture.
usage: python tools/ncu_by_line.py report.ncu-rep locosim_core.cuh [first_line last_line] [--top N]"""
import csv
import io
import subprocess
import sys


def main():
    rep, fname = sys.argv[1], sys.argv[2]
    lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 and sys.argv[3].isdigit() else (0, 10 ** 9)
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur_file, rows, header = None, [], None
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1]
        elif r and r[0] == "Line No":
            header = r
        elif header and cur_file and cur_file.endswith(fname) and r and r[0].isdigit():
            d = dict(zip(header, r))
            ln = int(r[0])
            if lo <= ln <= hi:
                rows.append((ln, int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0), r[1].strip()[:110]))
    tot_i = sum(x[1] for x in rows) or 1
    tot_s = sum(x[2] for x in rows) or 1
    print("lines %d-%d of %s: %.1f M warp-instructions, %d samples" % (lo, hi, fname, tot_i / 1e6, tot_s))
    sel = sorted(rows, key=lambda x: -x[1])[:top] if top else rows
    for ln, ins, smp, src in sorted(sel):
        if ins or smp:
            print("%5d %9.2fM %5.1f%% %7d %5.1f%%  %s" % (ln, ins / 1e6, 100.0 * ins / tot_i, smp, 100.0 * smp / tot_s, src))


if __name__ == "__main__":
    main()
