# usage: gpurun -- 'bash scripts/gpu_r1f.sh'   (misaligned-chunk fast path + ncu traffic capture)
set -x; mkdir -p gpurun_out
python -c "import os; print('cpu_count', os.cpu_count())" > gpurun_out/r1f_host.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> gpurun_out/r1f_host.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30 > gpurun_out/r1f_pytest.log
timeout 600 python bench.py > gpurun_out/r1f_bench.json 2>gpurun_out/r1f_bench.err
for nj in 16 24 7; do
  timeout 300 python bench.py --n-jobs $nj --no-e2e --no-variants --no-cpu-baseline > gpurun_out/r1f_bench_nj$nj.json 2>>gpurun_out/r1f_bench.err
done
# one full ncu capture of the timed-size encode launch (launch 6 of k_stream: 3 warm-up rounds x (encode, decode))
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stream -s 6 -c 1 -f -o gpurun_out/r1f_kstream_encode \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-variants --no-cpu-baseline > gpurun_out/r1f_ncu_full.log 2>&1
echo "ncu rc=$?"
tail -5 gpurun_out/r1f_pytest.log; head -c 1500 gpurun_out/r1f_bench.json; for nj in 16 24 7; do python -c "
import json,sys; d=json.load(open('gpurun_out/r1f_bench_nj$nj.json')); print($nj, d['value'], d['phases'], d['roofline_prf']['frac'])"; done
ls -la gpurun_out/
