"""Turns what profiles/run_gpu.sh brought back under gpurun_out/<tag>_* into the tracked files of
profiles/:  <tag>_summary.md (launch list table, `ncu --set full` metrics table, stall breakdown of the
source pages, the bench line), <tag>_launches.csv (the raw launch list) and traffic.json (DRAM bytes per
grid point per kernel, which bench.py reports as roofline.traffic).

  python profiles/collect.py <tag> [--points N]     # N = grid points of the --set full capture (500x500x80)
"""
import collections
import csv
import glob
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "gpurun_out")
sys.path.insert(0, HERE)
from summarize import launch_table  # noqa: E402

WANT = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "rd GB"), ("dram__bytes_write.sum", "wr GB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 ld sectors"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long_sb/issue"),
        ("smsp__inst_executed.sum", "warp inst")]
SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}


def kname(s):
    """'void wg::k_advance_uv<(bool)1>(wg::Geo, ...)' -> 'k_advance_uv<true>' (the names bench.py's kernel report uses)"""
    s = s.replace("wg::", "").replace("void ", "").replace("(bool)", "").strip()
    s = s.split("(")[0]
    s = s.replace("<1>", "<true>").replace("<0>", "<false>")
    return {"k_advect_scalar_fast<true>": "k_advect_scalar_fast<theta>", "k_advect_scalar_fast<false>": "k_advect_scalar_fast"}.get(s, s)


def raw_table(path, points):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]

    def find(m):
        for i, h in enumerate(hdr):
            if h == m or h.endswith("." + m):
                return i
        return None

    cols = [(find(m), lbl) for m, lbl in WANT if find(m) is not None]
    kn, gs = hdr.index("Kernel Name"), hdr.index("Grid Size")
    out = ["| kernel | grid | " + " | ".join(lbl for _, lbl in cols) + " |", "|---|---|" + "---:|" * len(cols)]
    seen, traffic = collections.Counter(), {}
    ird, iwr = find("dram__bytes_read.sum"), find("dram__bytes_write.sum")
    for r in rows[2:]:
        name = kname(r[kn])
        seen[name] += 1
        if seen[name] > 1:
            continue
        vals = []
        for i, _ in cols:
            v = r[i].replace(",", "")
            try:
                vals.append(f"{float(v) * SCALE.get(units[i], 1.0):.4g}")
            except ValueError:
                vals.append(v)
        out.append(f"| {name} | {r[gs].replace(' ', '')} | " + " | ".join(vals) + " |")
        try:
            tot = float(r[ird].replace(",", "")) * SCALE.get(units[ird], 1) + float(r[iwr].replace(",", "")) * SCALE.get(units[iwr], 1)
            traffic[name] = tot * 1e9 / points
        except (ValueError, TypeError):
            pass
    return "\n".join(out), traffic


def stall_table(path):
    rows = list(csv.reader(open(path)))
    title = rows[0][1] if rows and len(rows[0]) > 1 else path
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) >= len(hdr) and r[idx["# Samples"]].isdigit():
            data.append(r)
    samples = sum(int(r[idx["# Samples"]]) for r in data) or 1
    tot = {h: sum(int(r[idx[h]] or 0) for r in data) for h in st}
    out = [f"**{kname(title)}**: {samples} samples, {len(data)} SASS instructions; " +
           ", ".join(f"{h[6:]} {v / samples:.0%}" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:5]), ""]
    out.append("| samples | share | SASS | long_sb |")
    out.append("|---:|---:|---|---:|")
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:6]:
        s = int(r[idx["# Samples"]])
        out.append(f"| {s} | {s / samples:.1%} | `{r[idx['Source']].strip()[:70]}` | {r[idx['stall_long_sb']]} |")
    return "\n".join(out)


def main():
    tag = sys.argv[1]
    points = 499 * 499 * 79
    if "--points" in sys.argv:
        points = int(sys.argv[sys.argv.index("--points") + 1])
    md = [f"# {tag}\n"]
    b = os.path.join(OUT, f"{tag}_bench.json")
    if os.path.exists(b) and os.path.getsize(b):
        line = open(b).read().strip().splitlines()[-1]
        md.append("## bench line (python bench.py --steps 5 --warmup 3, 2000x2000x80, one B200)\n\n```json\n" + line + "\n```\n")
    la = os.path.join(OUT, f"{tag}_launches.csv")
    if os.path.exists(la):
        t, tot = launch_table(la)
        md.append(f"## launch list of the bench command (ncu --metrics gpu__time_duration.sum --clock-control none), "
                  f"{tot:.1f} ms of kernels\n\n{t}\n")
        shutil.copy(la, os.path.join(HERE, f"{tag}_launches.csv"))
    raw = os.path.join(OUT, f"{tag}_full_raw.csv")
    if os.path.exists(raw) and os.path.getsize(raw):
        t, traffic = raw_table(raw, points)
        md.append(f"## ncu --set full --clock-control none, 500x500x80, first captured launch of each kernel\n\n{t}\n")
        tj = os.path.join(HERE, "traffic.json")
        old = json.load(open(tj)) if os.path.exists(tj) else {}
        for k, v in traffic.items():
            old[k] = {"dram_bytes_per_point": v, "source": f"profiles/{tag}_summary.md (ncu --set full, 500x500x80)"}
        json.dump(old, open(tj, "w"), indent=1)
    srcs = sorted(glob.glob(os.path.join(OUT, f"{tag}_src_*.csv")))
    if srcs:
        md.append("## warp-stall sampling of the source pages (top SASS lines)\n")
        for s in srcs:
            if os.path.getsize(s):
                md.append(stall_table(s) + "\n")
    open(os.path.join(HERE, f"{tag}_summary.md"), "w").write("\n".join(md))
    print("wrote", os.path.join(HERE, f"{tag}_summary.md"))


if __name__ == "__main__":
    main()
