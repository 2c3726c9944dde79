"""Summarise `ncu --page raw --csv` output: one block per captured launch with the metrics DESIGN.md / profiles/ quote."""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__cluster_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    # the shared-memory data pipe: LSU wavefronts (LDS / STS) and the tensor core's operand fetches share it
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
]


def main(path, extra=()):
    rows = list(csv.reader(open(path)))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr = i
            break
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: j for j, n in enumerate(names)}
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        print(f"== {r[col['Kernel Name']][:90]}  grid {r[col.get('Grid Size', 0)]} block {r[col.get('Block Size', 0)]}")
        for k in list(KEYS) + list(extra):
            if k in col:
                print(f"   {k:<85s} {r[col[k]]:>16s} {units[col[k]]}")
        stalls = [(float(r[j].replace(',', '') or 0), n) for n, j in col.items()
                  if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued") and r[j]]
        tot = sum(v for v, _ in stalls) or 1.0
        top = sorted(stalls, reverse=True)[:6]
        print("   top stalls: " + ", ".join(f"{n.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.1f}%" for v, n in top))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
