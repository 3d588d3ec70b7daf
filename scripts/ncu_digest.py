"""Digest of an .ncu-rep (raw page) into the handful of metrics quoted in DESIGN.md and profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','lts__t_sector_hit_rate.pct','sm__cycles_elapsed.avg.per_second']
for r in rows[2:]:
    for w in want:
        for i, n in enumerate(h):
            if n == w: print("%-80s %s %s" % (w, r[i], units[i]))
for i, n in enumerate(h):
    if n.startswith('sm__ops_path_tensor_src_fp64.sum') or n.startswith('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct'):
        print("%-80s %s %s" % (n, rows[2][i], units[i]))
