#!/usr/bin/env python
"""Key metrics of one kernel from an `ncu --set full` report (the lists under profiles/ncu_*.txt):
    python tools/ncu_summary.py gpurun_out/tapwgrad_r1.ncu-rep >> profiles/ncu_tapwgrad_r1.txt"""
import csv
import subprocess
import sys

WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum sm__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed lts__throughput.avg.pct_of_peak_sustained_elapsed
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum l1tex__data_pipe_tc_wavefronts_mem_shared.sum
l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum l1tex__t_requests_pipe_lsu_mem_global_op_st.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum l1tex__m_xbar2l1tex_read_bytes.sum l1tex__m_l1tex2xbar_write_bytes.sum
launch__registers_per_thread launch__shared_mem_per_block_dynamic launch__occupancy_limit_shared_mem launch__occupancy_limit_registers
sm__warps_active.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio""".split()

rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, row = rows[0], rows[1], rows[2]
d = {k: (v, u) for k, u, v in zip(hdr, units, row)}
for k in ("Kernel Name", "Grid Size", "Block Size"):
    print(f"{k:100s} {d[k][0]}")
for k in WANT:
    if k in d:
        print(f"{k:100s} {d[k][0]:>12s} {d[k][1]}")
