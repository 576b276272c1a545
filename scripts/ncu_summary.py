"""Summarise an .ncu-rep: key raw metrics + SASS opcode mix + top stall instructions (read on the CPU box)."""
import collections, csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'sm__cycles_elapsed.max']


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main(path, top=14):
    rows = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("kernel:", name[:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:88s} {r[i]:>16s} {units[i]}")
    rows = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv"]))))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    isrc, ist, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    f = lambda v: float(v) if v else 0.0
    tex, tst = sum(f(r[iex]) for r in data) or 1, sum(f(r[ist]) for r in data) or 1
    hist, sth = collections.Counter(), collections.Counter()
    for r in data:
        t = r[isrc].split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else t[0] if t else "?").split(".")[0]
        hist[op] += f(r[iex]); sth[op] += f(r[ist])
    print(f"SASS: {len(data)} instructions, {tex:.3g} warp-level executions")
    for op, v in hist.most_common(top):
        print(f"  {op:10s} exec {v / tex * 100:5.1f}%   stall samples {sth[op] / tst * 100:5.1f}%")
    print("top stall sites:")
    for r in sorted(data, key=lambda r: -f(r[ist]))[:10]:
        print(f"  {f(r[ist]) / tst * 100:5.1f}%  {r[isrc][:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
