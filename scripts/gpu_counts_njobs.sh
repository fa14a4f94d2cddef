set -x; mkdir -p gpurun_out
for nj in 16 24; do
  timeout 300 ncu --metrics smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,gpu__time_duration.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
    --clock-control none -k regex:^k_stream -s 6 -c 1 --csv --log-file gpurun_out/${TAG}_counts_nj$nj.csv \
    python bench.py --steps 1 --warmup 3 --clients 8 --n-jobs $nj --no-e2e --no-variants --no-cpu-baseline > /dev/null 2>>gpurun_out/${TAG}_counts.err
done
