#!/usr/bin/env python
"""Summarise an `ncu --set full` report (one kernel launch) into a markdown file under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name.md "title" [warps frames]

Reads the report on a box without a GPU (`ncu -i ... --page raw|source --csv`)."""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst, title = sys.argv[1:4]
    warp_frames = int(sys.argv[4]) * int(sys.argv[5]) if len(sys.argv) > 5 else None
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    L = [f"# {title}", "", f"source: `ncu --set full --clock-control none --import-source on` → `{rep}` (one launch, B200)", "",
         "| metric | value | unit |", "|---|---|---|"]
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name":
            L.insert(2, f"kernel: `{v}`")
        try:
            stall = h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(v) > 0.05
        except ValueError:
            stall = False
        if h in WANT or stall:
            L.append(f"| {h} | {v} | {u} |")
    rows = page(rep, "source")
    hdr, data = rows[1], rows[2:]
    ia, isrc, ism = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    byop, smp, tot, ts = collections.Counter(), collections.Counter(), 0, 0
    for r in data:
        n = int(r[ia]); op = r[isrc].split()[0]
        if op.startswith('@'):
            op = r[isrc].split()[1]
        op = op.split('.')[0]
        byop[op] += n; tot += n; smp[op] += int(r[ism]); ts += int(r[ism])
    if warp_frames:
        L += ["", f"Executed SASS instructions per warp per frame ({sys.argv[4]} warps x {sys.argv[5]} frames): **{tot / warp_frames:.1f}**"]
    L += ["", "| opcode | executed (M warp-instr) | stall samples % |", "|---|---|---|"]
    for op, n in byop.most_common(24):
        L.append(f"| {op} | {n / 1e6:.2f} | {100 * smp[op] / max(ts, 1):.1f} |")
    open(dst, "w").write("\n".join(L) + "\n")
    print("\n".join(L[:40]))


if __name__ == "__main__":
    main()
