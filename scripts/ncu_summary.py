"""Summarises an .ncu-rep (ncu --set full) into the handful of numbers the roofline discussion needs.
    python scripts/ncu_summary.py gpurun_out/prof_attn.ncu-rep [> profiles/xxx.txt]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-memory (TMEM) active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks/SM"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks/SM"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall long_scoreboard %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier (warps/issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (warps/issue)"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar (warps/issue)"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping (warps/issue)"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle (warps/issue)"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle (warps/issue)"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of {rep}")
    for r in rows[2:]:
        print(f"\n## {r[col['Kernel Name']]}  (launch id {r[col['ID']]})")
        for key, label in KEYS:
            if key in col:
                print(f"  {label:45s} {r[col[key]]} {units[col[key]]}")
        for h in hdr:   # whatever tensor-pipe counters this ncu version exposes for sm_100
            if "pipe_tensor" in h and h not in dict(KEYS) and r[col[h]] not in ("", "n/a"):
                print(f"  {h:45s} {r[col[h]]} {units[col[h]]}")


if __name__ == "__main__":
    main()
