# full GPU tests (in-tree library), interleaved A/B of the headline round, ncu counts per library
set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/abc_pytest.log
for rep in 1 2; do
for lib in $LIBS; do
  export FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/abc_${lib}_r${rep}.json 2>gpurun_out/abc_${lib}.err
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --n-jobs 24 > gpurun_out/abc_${lib}_nj24_r${rep}.json 2>>gpurun_out/abc_${lib}.err
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --share-streams 1 > gpurun_out/abc_${lib}_shared_r${rep}.json 2>>gpurun_out/abc_${lib}.err
done
done
for lib in $LIBS; do
  export FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so
  timeout 300 ncu --metrics smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_uniform.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct \
    --clock-control none -k regex:^k_stream -s 6 -c 1 --csv --log-file gpurun_out/abc_counts_$lib.csv \
    python bench.py --steps 1 --warmup 3 --clients 8 --no-e2e --no-variants --no-cpu-baseline > /dev/null 2>>gpurun_out/abc_$lib.err
done
unset FLASHE_B200_LIB
tail -3 gpurun_out/abc_pytest.log
for f in gpurun_out/abc_*_r*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']/1e9,2), round(d['phases']['encode_encrypt_ms'],2), round(d['phases']['decrypt_decode_ms'],3))"; done
