set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rows_f_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
tail -n 4 gpurun_out/${TAG}_pytest.log
python -c "
import json
for l in open('gpurun_out/${TAG}_rows.jsonl'):
    d=json.loads(l); print('%-50s %8.3f ms %7.0f GB/s %5.1f%%' % (d['kernel'][:50], d['ms'], d['gbs'], 100*d['frac_of_hbm']))"
