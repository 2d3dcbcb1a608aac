"""Summarise ncu outputs for profiles/.

  python tools/ncu_summarize.py launches <launches.csv> <out.md>     per-kernel launch list summary
  python tools/ncu_summarize.py raw <report.ncu-rep> <out.md>        key metrics of a --set full capture
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def launches(path, out):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            name = re.sub(r"\(.*", "", r["Kernel Name"])
            name = re.sub(r"^void ", "", name)
            rows.append((name, float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and "
                "serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"{len(rows)} launches, {total/1e6:.3f} ms total\n\n| kernel | launches | total ms | share | grid (last) | block |\n|---|---:|---:|---:|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {a[0]} | {a[1]/1e6:.3f} | {100*a[1]/total:.1f}% | {a[2]} | {a[3]} |\n")
    print(open(out).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "smsp__cycles_active.avg", "sm__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size"]


def raw(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\n")
        for row in data:
            f.write(f"## {row[idx['Kernel Name']][:100]}  (launch id {row[idx['ID']]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {row[idx[k]]} | {units[idx[k]]} |\n")
            f.write("\n")
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
