"""Summarise an ncu --set full report into profiles/<name>_summary.txt (+ traffic.json for bench.py).

    python profiles/summarize_ncu.py gpurun_out/prof_r1.ncu-rep r1_glm_fused_bernoulli_N10M_K100 10000000 100
    python profiles/summarize_ncu.py rep name N K family G kernel_prefix     # round 2: traffic of ONE kernel only
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_gmma_per_issue_active.ratio",
]


def to_bytes(v, unit):
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * mul.get(unit, 1)


def main():
    rep, name = sys.argv[1], sys.argv[2]
    N, K = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (None, None)
    family = sys.argv[5] if len(sys.argv) > 5 else None
    G = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    kprefix = sys.argv[7] if len(sys.argv) > 7 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    here = os.environ.get("NCU_SUMMARY_DIR") or os.path.dirname(os.path.abspath(__file__))   # where the summaries go
    lines = [f"# ncu --set full --clock-control none: {os.path.basename(rep)} ({len(rows) - 2} launches captured)"]
    traffic = []
    for r in rows[2:]:
        lines.append(f"## {r[hdr.index('Kernel Name')]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k:95s} {r[i]:>18s} {units[i]}")
        i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        if kprefix is None or kprefix in r[hdr.index('Kernel Name')]:   # the dominant kernel only
            traffic.append(to_bytes(r[i_r], units[i_r]) + to_bytes(r[i_w], units[i_w]))
    with open(os.path.join(here, name + "_summary.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if N is not None:
        tp = os.path.join(here, "traffic.json")
        ents = []
        if os.path.exists(tp):
            with open(tp) as f:
                old = json.load(f)
            ents = [e for e in old.get("entries", [old])
                    if not (e.get("N") == N and e.get("K") == K and e.get("family") == family and e.get("G", 0) == G)]
        ents.append({"N": N, "K": K, "family": family, "G": G, "kernel": kprefix or "glm_fused_kernel",
                     "dram_bytes_per_launch": sum(traffic) / len(traffic),
                     "source": name + "_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum of " +
                               (kprefix or "every kernel") + ", mean over the captured launches)"})
        with open(tp, "w") as f:
            json.dump({"entries": ents}, f, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
