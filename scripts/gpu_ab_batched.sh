# interleaved A/B of the fused lane-batched round: LIBS="a b" bash scripts/gpu_ab_batched.sh
set -x; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch" 2>&1 | tail -3
for rep in 1 2; do
for lib in $LIBS; do
  FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so python scripts/bench_batched.py > gpurun_out/abb_${lib}_r${rep}.json 2>gpurun_out/abb_${lib}.err
  echo $lib $(cat gpurun_out/abb_${lib}_r${rep}.json)
done
done
