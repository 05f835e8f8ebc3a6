"""Per-source-line instruction/sample shares of one kernel capture (ncu --import-source on).

    python tools/ncu_lines.py gpurun_out/r01q_prof_transit.ncu-rep xo_transit.cu [min_pct]
"""
import csv
import io
import os
import subprocess
import sys

rep, fname = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur, hdr, data = None, None, []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1])
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        try:
            data.append((cur, int(r[0]), int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), r[1].strip()[:110]))
        except (ValueError, IndexError):
            pass
tot = sum(d[2] for d in data) or 1
toti = sum(d[3] for d in data) or 1
print("total samples %d, warp instructions %d" % (tot, toti))
for d in sorted([d for d in data if d[0] == fname], key=lambda d: d[1]):
    if 100 * d[3] / toti >= thr or 100 * d[2] / tot >= thr:
        print("%4d %5.1f%% %5.1f%%  %s" % (d[1], 100 * d[2] / tot, 100 * d[3] / toti, d[4]))
