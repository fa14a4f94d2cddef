set -x; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=3 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
timeout 900 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
tail -8 gpurun_out/${TAG}_pytest.log; python -c "
import json
for l in open('gpurun_out/${TAG}_rows.jsonl'):
    d=json.loads(l); print('%-50s %8.3f ms %7.0f GB/s %5.1f%%' % (d['kernel'][:50], d['ms'], d['gbs'], 100*d['frac_of_hbm']))
d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['phases']); print(json.dumps(d['variants'])[:1500])"
tail -3 gpurun_out/${TAG}_rows.err gpurun_out/${TAG}_bench.err
