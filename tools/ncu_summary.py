"""Summarise ncu artefacts into profiles/: per-kernel launch times from a `--metrics gpu__time_duration.sum` CSV and the
key counters of a `--set full` report (read with `ncu -i … --page raw --csv`). Usage:
    python tools/ncu_summary.py launches <csv>            python tools/ncu_summary.py full <ncu-rep>"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                name = d["Kernel Name"].split("(")[0].replace("void ", "")
                agg.setdefault(name, []).append(float(d["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':58s} {'n':>4s} {'mean_us':>10s} {'sum_us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[:58]:58s} {len(v):4d} {sum(v)/len(v)/1e3:10.1f} {sum(v)/1e3:10.1f} {100*sum(v)/tot:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {d['Kernel Name'].split('(')[0]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for i, n in enumerate(hdr):
            if n in WANT:
                print(f"   {n:75s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
