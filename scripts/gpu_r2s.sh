set -x; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "more_clients or graph or noise" 2>&1 | tail -n 3
timeout 900 python scripts/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2>gpurun_out/${TAG}_configs.err
python -c "
import json
for l in open('gpurun_out/${TAG}_configs.jsonl'):
    d=json.loads(l); print(d['config'][:60], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.startswith('ms_') or k.startswith('g_aes') or 'fill_g' in k or 'online' in k or 'server' in k or 'client_topk' in k})"
tail -n 3 gpurun_out/${TAG}_configs.err
