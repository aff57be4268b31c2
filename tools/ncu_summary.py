"""Summarise an .ncu-rep (raw + source pages) into text for profiles/ (reads reports here, no GPU)."""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_average_branch_targets_threads_uniform.pct"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = page(rep, "raw")
    hdr = rows[0]
    print(f"# {rep}")
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[ki][:110])
    for name in WANT:
        if name in hdr:
            i = hdr.index(name)
            print(f"{name:75s} {[r[i] for r in rows[2:]]} {rows[1][i]}")
    src = page(rep, "source")
    hdr = next(r for r in src if r and r[0] == "Address")
    si, ie, smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    ops, samples, stalls = collections.Counter(), collections.Counter(), collections.Counter()
    total = tot_s = 0
    seen_kernel = 0
    for r in src:
        if r and r[0] == "Kernel Name":
            seen_kernel += 1
            if seen_kernel > 1:
                break
            continue
        if len(r) < len(hdr) or r[0] == "Address":
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
        op = m.group(2).split(".")[0] if m else "?"
        n, s = int(r[ie] or 0), int(r[smp] or 0)
        ops[op] += n; samples[op] += s; total += n; tot_s += s
        for i, h in stall_cols:
            stalls[h] += int(r[i] or 0)
    print(f"warp-instructions (first launch): {total}; sampled {tot_s}")
    print("top opcodes: " + ", ".join(f"{op} {100 * n / total:.1f}%" for op, n in ops.most_common(16)))
    ssum = sum(stalls.values()) or 1
    print("stall reasons: " + ", ".join(f"{k[6:]} {100 * v / ssum:.1f}%" for k, v in stalls.most_common(9)))


if __name__ == "__main__":
    for rep in sys.argv[1:]:
        main(rep)
        print()
