# GPU tests, then the headline round for several chunk layouts (--n-jobs)
set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
for nj in 16 24 7 13; do
  timeout 300 python bench.py --steps 3 --warmup 3 --n-jobs $nj --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${TAG}_bench_nj$nj.json 2>>gpurun_out/${TAG}_bench.err
done
tail -4 gpurun_out/${TAG}_pytest.log
for nj in 16 24 7 13; do python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_nj$nj.json')); print($nj, round(d['value']/1e9,2), round(d['phases']['encode_encrypt_ms'],2), round(d['phases']['decrypt_decode_ms'],3))"; done
