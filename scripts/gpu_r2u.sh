# GPU tests, config timings and an ncu capture (source counters) of the int_bits 20 encode kernel
set -x; mkdir -p gpurun_out
TAG=${TAG:-r2u}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 900 python scripts/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2>gpurun_out/${TAG}_configs.err
python -c "
import json
for l in open('gpurun_out/${TAG}_configs.jsonl'):
    d=json.loads(l); print(d['config'][:60], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.startswith('ms_') or k.startswith('g_aes')})"
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stream -s 6 -c 1 -f -o gpurun_out/${TAG}_kstream_encode_b20 \
  python bench.py --steps 1 --warmup 3 --int-bits 20 --clients 10 --elements 25000000 --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${TAG}_ncu_b20.log 2>&1
fi
