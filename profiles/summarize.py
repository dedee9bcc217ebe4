"""Turns what gpurun brought back (an ncu launch list and an `ncu --set full` report) into the
tracked summaries under profiles/.   python profiles/summarize.py <tag> [launches.csv] [prof.ncu-rep]"""
import collections
import csv
import subprocess
import sys


def launch_table(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        k = row["Kernel Name"].split("(")[0].replace("wg::", "")
        if k.startswith("void at::") or "at::native" in k:
            k = "(torch: bench set-up, outside the timed region)"
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total ms | avg ms | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]:.3f} | {v[1] / v[0]:.4f} | {v[1] / tot:.3f} |")
    return "\n".join(out), tot


WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 ld sectors"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long_sb stall / issue")]


def metric_table(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), lbl) for m, lbl in WANT if m in hdr]
    kn = hdr.index("Kernel Name")
    out = ["| kernel | " + " | ".join(f"{lbl} [{units[i]}]" if units[i] else lbl for i, lbl in cols) + " |",
           "|---|" + "---:|" * len(cols)]
    seen = collections.Counter()
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("wg::", "")
        seen[name] += 1
        if seen[name] > 1:
            continue
        out.append(f"| {name} | " + " | ".join(r[i] for i, _ in cols) + " |")
    return "\n".join(out)


if __name__ == "__main__":
    tag = sys.argv[1]
    if len(sys.argv) > 2:
        t, tot = launch_table(sys.argv[2])
        print(f"## {tag}: launch list (ncu --metrics gpu__time_duration.sum --clock-control none), {tot:.1f} ms of kernels\n\n{t}\n")
    if len(sys.argv) > 3:
        print(f"## {tag}: ncu --set full, first captured launch of each kernel\n\n{metric_table(sys.argv[3])}\n")
