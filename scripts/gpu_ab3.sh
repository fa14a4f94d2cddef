# interleaved A/B of kernel variants on three workloads: headline (b=32, 100M x 64), b=20 (25M x 10), batched 120-bit round
# LIBS="base b3 ..." bash scripts/gpu_ab3.sh
set -x; mkdir -p gpurun_out
for rep in 1 2; do
for lib in $LIBS; do
  export FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/ab3_${lib}_b32_r${rep}.json 2>gpurun_out/ab3_${lib}.err
  python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --int-bits 20 --clients 10 --elements 25000000 > gpurun_out/ab3_${lib}_b20_r${rep}.json 2>>gpurun_out/ab3_${lib}.err
  python scripts/bench_batched.py > gpurun_out/ab3_${lib}_b120_r${rep}.json 2>>gpurun_out/ab3_${lib}.err
done
done
unset FLASHE_B200_LIB
for f in gpurun_out/ab3_*_b32_r*.json gpurun_out/ab3_*_b20_r*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']/1e9,2), round(d['phases']['encode_encrypt_ms'],3), round(d['phases']['decrypt_decode_ms'],3))"; done
for f in gpurun_out/ab3_*_b120_r*.json; do echo $f $(cat $f); done
python scripts/microbench.py --only lds
