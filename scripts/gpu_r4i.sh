set -x; mkdir -p gpurun_out
TAG=${TAG:-r4i}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/bench_configs.py --sparse-only > gpurun_out/${TAG}_c4.jsonl 2> gpurun_out/${TAG}_c4.err; tail -2 gpurun_out/${TAG}_c4.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_c4.jsonl').read()); print({k:round(v,4) for k,v in d.items() if k.endswith('_ms')})"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"^k_sparse|^k_stream|^k_scatter|^k_fill|^k_overlap" -c 60 --csv --log-file gpurun_out/${TAG}_c4_launches.csv python scripts/bench_configs.py --sparse-only > /dev/null 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('headline', d['value']/1e9, d['phases']['encode_encrypt_ms'], d['roofline_prf'])"
