set -x; mkdir -p gpurun_out
TAG=${TAG:-r4f}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
for rep in 1 2; do
python scripts/bench_small.py --quick > gpurun_out/${TAG}_small_hyb_$rep.jsonl 2> gpurun_out/${TAG}_small.err
done
for nj in 16 7; do
python bench.py --steps 3 --warmup 3 --int-bits 20 --clients 10 --elements 25000000 --n-jobs $nj --no-e2e --no-cpu-baseline --no-variants | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('25Mx10 b20 njobs $nj', d['value']/1e9, d['phases']['encode_encrypt_ms'], d['phases']['decrypt_decode_ms'])"
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('headline', d['value']/1e9, d['phases'])"
