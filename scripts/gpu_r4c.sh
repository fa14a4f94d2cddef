# dynamic deal: GPU tests, small configs dynamic vs static, per-warp trace, quick headline both ways
set -x; mkdir -p gpurun_out
TAG=${TAG:-r4c}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
for rep in 1 2; do
python scripts/bench_small.py --quick > gpurun_out/${TAG}_small_dyn_$rep.jsonl 2> gpurun_out/${TAG}_small.err
FLASHE_DYNAMIC=0 python scripts/bench_small.py --quick > gpurun_out/${TAG}_small_static_$rep.jsonl 2>> gpurun_out/${TAG}_small.err
done
FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_trace.so FLASHE_TRACE_PRINT=1 python scripts/trace_small.py 2> gpurun_out/${TAG}_trace.txt
for dyn in 1 0 1 0; do
FLASHE_DYNAMIC=$dyn timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/${TAG}_bench_dyn$dyn.json 2>gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_dyn$dyn.json')); print('dyn$dyn', d['value']/1e9, d['phases']['encode_encrypt_ms'], d['phases']['decrypt_decode_ms'])"
FLASHE_DYNAMIC=$dyn python bench.py --steps 3 --warmup 3 --int-bits 20 --clients 10 --elements 25000000 --no-e2e --no-cpu-baseline --no-variants | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dyn$dyn 25Mx10 b20', d['value']/1e9, d['phases']['encode_encrypt_ms'], d['phases']['decrypt_decode_ms'])"
done
