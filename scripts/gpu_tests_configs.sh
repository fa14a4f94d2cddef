set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
timeout 800 python scripts/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/${TAG}_bench_quick.json 2>gpurun_out/${TAG}_bench_quick.err
tail -4 gpurun_out/${TAG}_pytest.log; python -c "
import json
for l in open('gpurun_out/${TAG}_configs.jsonl'):
    d=json.loads(l); print(d['config'][:60], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('ms_per_round','g_aes_blocks_per_s','client_elements_per_s','fill_g_aes_blocks_per_s','fill_ms','decrypt_decode_3_runs_ms')})
d=json.load(open('gpurun_out/${TAG}_bench_quick.json')); print(d['value']/1e9, d['phases'])
"
