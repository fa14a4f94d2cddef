# instruction / wavefront counts of the encode kernel for two library builds (8 clients), then rows tests + bench with the in-tree library
set -x; mkdir -p gpurun_out
for lib in $LIBS; do
  export FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so
  timeout 300 ncu --metrics smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_alu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct \
    --clock-control none -k regex:^k_stream -s 6 -c 1 --csv --log-file gpurun_out/ab4_counts_$lib.csv \
    python bench.py --steps 1 --warmup 3 --clients 8 --no-e2e --no-variants --no-cpu-baseline > /dev/null 2>gpurun_out/ab4_$lib.err
done
unset FLASHE_B200_LIB
timeout 600 python -m pytest tests/test_rows_f_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab4_pytest_rows.log
timeout 600 python scripts/bench_rows.py > gpurun_out/ab4_rows.jsonl 2>gpurun_out/ab4_rows.err
tail -4 gpurun_out/ab4_pytest_rows.log
for lib in $LIBS; do echo $lib; grep -v "^==" gpurun_out/ab4_counts_$lib.csv | cut -d, -f5,13,15 | tail -16; done
grep -E "wire|topk" gpurun_out/ab4_rows.jsonl | cut -c1-260
