# interleaved A/B of the headline round for $LIBS, then rows tests + rows bench with the in-tree library
set -x; mkdir -p gpurun_out
for rep in 1 2; do
for lib in $LIBS; do
  export FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/ab_${lib}_r${rep}.json 2>gpurun_out/ab_${lib}.err
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --share-streams 1 > gpurun_out/ab_${lib}_shared_r${rep}.json 2>>gpurun_out/ab_${lib}.err
done
done
unset FLASHE_B200_LIB
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_pytest.log
timeout 600 python scripts/bench_rows.py > gpurun_out/ab_rows.jsonl 2>gpurun_out/ab_rows.err
tail -3 gpurun_out/ab_pytest.log
for f in gpurun_out/ab_*_r*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']/1e9,2), round(d['phases']['encode_encrypt_ms'],2), round(d['phases']['decrypt_decode_ms'],3))"; done
grep -E "wire" gpurun_out/ab_rows.jsonl | cut -c1-120
