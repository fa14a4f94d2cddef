set -x; mkdir -p gpurun_out
for rep in 1 2; do
for lib in default u4 u16; do
  L=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so; [ "$lib" = default ] && L=$PWD/flashe_b200/_lib/libflashe_b200.so
  FLASHE_B200_LIB=$L python scripts/bench_small.py --quick 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$lib', d['tag'], 'enc %.1f dec %.1f round %.1f'%(d['encode_us'],d['decode_us'],d['round_graph_us']))"
  FLASHE_B200_LIB=$L python bench.py --steps 3 --warmup 3 --int-bits 20 --clients 10 --elements 25000000 --no-e2e --no-cpu-baseline --no-variants | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib 25Mx10 b20', round(d['value']/1e9,2), round(d['phases']['encode_encrypt_ms'],4), round(d['phases']['decrypt_decode_ms'],4))"
done
for m in 2 4; do
  FLASHE_DYNAMIC_MIN_ITEMS=$m python scripts/bench_small.py --quick 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('minitems$m', d['tag'], 'enc %.1f dec %.1f round %.1f'%(d['encode_us'],d['decode_us'],d['round_graph_us']))"
done
done
