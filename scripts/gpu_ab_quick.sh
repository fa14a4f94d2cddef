# interleaved A/B of the round (no tests): LIBS="a b" [EXTRA="--int-bits 20 ..."] bash scripts/gpu_ab_quick.sh
set -x; mkdir -p gpurun_out
for rep in 1 2; do
for lib in $LIBS; do
  export FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/abq_${lib}_r${rep}.json 2>gpurun_out/abq_${lib}.err
  if [ -n "$EXTRA" ]; then python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-variants $EXTRA > gpurun_out/abq_${lib}_x_r${rep}.json 2>>gpurun_out/abq_${lib}.err; fi
done
done
for f in gpurun_out/abq_*_r*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']/1e9,2), round(d['phases']['encode_encrypt_ms'],3), round(d['phases']['decrypt_decode_ms'],3))"; done
