"""Summarise ncu captures into small text files for profiles/ (run where ncu is installed).

    python tools/ncu_summary.py gpurun_out/r01c profiles/r01_c        # <prefix>_prof_*.ncu-rep, <prefix>_launches.csv
"""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fp64.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def source_lines(rep, top=25):
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
                data.append((int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), cur, int(r[0]), r[1].strip()[:90]))
            except (ValueError, IndexError):
                pass
    if not data:
        return ["(no source-level data)"]
    tot, toti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
    byfile = collections.OrderedDict()
    for d in data:
        f = byfile.setdefault(d[2], [0, 0])
        f[0] += d[0]
        f[1] += d[1]
    lines = ["samples %d, warp instructions %d" % (tot, toti), "by file (samples%, instructions%):"]
    lines += ["  %-24s %5.1f%% %5.1f%%" % (k, 100 * v[0] / tot, 100 * v[1] / toti) for k, v in byfile.items()]
    lines.append("hottest source lines (samples%, instr%):")
    for d in sorted(data, reverse=True)[:top]:
        lines.append("  %5.1f%% %5.1f%%  %s:%d  %s" % (100 * d[0] / tot, 100 * d[1] / toti, d[2], d[3], d[4]))
    return lines


def launches(path):
    if not os.path.exists(path):
        return []
    rows = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(rows):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        u = row.get("Metric Unit", "")
        v = v / 1e3 if u.startswith("n") else v * 1e3 if u.startswith("m") else v * 1e6 if u in ("s", "second") else v
        agg.setdefault(row["Kernel Name"].split("(")[0][-70:], []).append(v)
    tot = sum(sum(v) for v in agg.values()) or 1
    out = ["launch list (ncu --metrics gpu__time_duration.sum; cold-cache, serialised -- compare SHARES):"]
    for k, v in agg.items():
        out.append("  %-72s n=%3d mean=%9.2f us share=%5.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
    return out


def main(prefix, outprefix):
    os.makedirs(os.path.dirname(outprefix) or ".", exist_ok=True)
    for rep in sorted(glob.glob(prefix + "_prof_*.ncu-rep")):
        name = os.path.basename(rep).replace(".ncu-rep", "").split("_prof_")[-1]
        rows, units = raw(rep)
        lines = ["# ncu --set full --clock-control none --import-source on : %s" % os.path.basename(rep)]
        for d in rows:
            lines.append("kernel: " + d.get("Kernel Name", "?")[:120])
            for k in KEYS:
                if k in d:
                    lines.append("  %-62s %s %s" % (k, d[k], units.get(k, "")))
            st = sorted(((float(d[k]), k) for k in d if "smsp__average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")), reverse=True)
            lines.append("  warp stalls per issue: " + ", ".join("%s=%.2f" % (k.split("stalled_")[1].split("_per_")[0], v) for v, k in st[:7]))
        lines += source_lines(rep)
        with open("%s_%s.txt" % (outprefix, name), "w") as fh:
            fh.write("\n".join(lines) + "\n")
        print("wrote", "%s_%s.txt" % (outprefix, name))
    ll = launches(prefix + "_launches.csv")
    if ll:
        with open(outprefix + "_launches.txt", "w") as fh:
            fh.write("\n".join(ll) + "\n")
        print("wrote", outprefix + "_launches.txt")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
