#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

  python scripts/ncu_summary.py launches <launches.csv>            per-kernel time shares
  python scripts/ncu_summary.py kernel <report.ncu-rep>            key counters per captured launch
  python scripts/ncu_summary.py source <report.ncu-rep> [kernel-substr] [top-n]   hottest source lines
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return out


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare SHARES)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{t / 1e3:12.1f} us {n:5d}x {100 * t / tot:6.2f}%  avg {t / n / 1e3:10.1f} us  {k}")


def kernel(rep):
    rows = list(csv.reader(io.StringIO(ncu_csv(rep, "raw"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        print("---", d[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:86s} {d[hdr.index(k)]:>22s} {units[hdr.index(k)]}")


def source(rep, substr="", top=25):
    text = ncu_csv(rep, "source")
    blocks = text.split('"Kernel Name",')
    for blk in blocks[1:]:
        lines = blk.splitlines()
        name = lines[0]
        if substr and substr not in name:
            continue
        rows = list(csv.reader(lines[1:]))
        hdr = rows[0]
        si = hdr.index("Source")
        ci = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
        ii = hdr.index("Instructions Executed")
        ti = hdr.index("Avg. Threads Executed") if "Avg. Threads Executed" in hdr else None
        body = []
        for r in rows[1:]:
            try:
                body.append((int(r[ci] or 0), r))
            except (ValueError, IndexError):
                pass
        tot = sum(b[0] for b in body) or 1
        print("=== ", name[:120], " total samples", tot)
        for n, r in sorted(body, key=lambda x: -x[0])[:top]:
            thr = r[ti] if ti is not None else ""
            print(f"{100 * n / tot:6.2f}%  inst {r[ii]:>12s}  thr {thr:>5s}  {r[si][:110]}")
        break


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "kernel":
        kernel(sys.argv[2])
    else:
        source(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "", int(sys.argv[4]) if len(sys.argv) > 4 else 25)
