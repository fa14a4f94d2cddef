# usage: gpurun -- 'LIBS="base ctr" bash scripts/gpu_ab3.sh'  -- full GPU tests with the in-tree library, then interleaved A/B of the headline round
set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab3_pytest.log
for rep in 1 2; do
for lib in $LIBS; do
  export FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/ab3_${lib}_r${rep}.json 2>gpurun_out/ab3_${lib}.err
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --n-jobs 24 > gpurun_out/ab3_${lib}_nj24_r${rep}.json 2>>gpurun_out/ab3_${lib}.err
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --share-streams 1 > gpurun_out/ab3_${lib}_shared_r${rep}.json 2>>gpurun_out/ab3_${lib}.err
done
done
unset FLASHE_B200_LIB
tail -3 gpurun_out/ab3_pytest.log
for f in gpurun_out/ab3_*_r*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']/1e9,2), round(d['phases']['encode_encrypt_ms'],2), round(d['phases']['decrypt_decode_ms'],3))"; done
