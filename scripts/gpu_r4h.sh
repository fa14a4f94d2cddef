set -x; mkdir -p gpurun_out
TAG=${TAG:-r4h}
timeout 600 python -m pytest tests -m gpu -x -q -k "sparse or dynamic_masking or config4 or c4 or C4" 2>&1 | tail -5
for t in 1 0 1 0; do
FLASHE_SPARSE_TILED=$t python scripts/bench_configs.py --sparse-only > gpurun_out/${TAG}_c4_tiled$t.jsonl 2> gpurun_out/${TAG}_c4.err; tail -2 gpurun_out/${TAG}_c4.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_c4_tiled$t.jsonl').read()); print('tiled$t', {k:round(v,4) for k,v in d.items() if k.endswith('_ms')})"
done
