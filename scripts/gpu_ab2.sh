# usage: LIBS="base quad" gpurun -- 'LIBS="base quad" bash scripts/gpu_ab2.sh'  -- headline round only, interleaved repetitions
set -x; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rows_f_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab2_pytest_rows.log
for rep in 1 2 3; do
for lib in $LIBS; do
  export FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants $EXTRA > gpurun_out/ab2_${lib}_r${rep}.json 2>gpurun_out/ab2_${lib}.err
done
done
unset FLASHE_B200_LIB
timeout 600 python scripts/bench_rows.py > gpurun_out/ab2_rows.jsonl 2>gpurun_out/ab2_rows.err
tail -3 gpurun_out/ab2_pytest_rows.log
for f in gpurun_out/ab2_*_r*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']/1e9,2), round(d['phases']['encode_encrypt_ms'],2))"; done
grep wire gpurun_out/ab2_rows.jsonl
