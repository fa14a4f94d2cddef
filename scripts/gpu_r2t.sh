# ncu --set full capture (with source counters) of the int_bits 20 encode kernel, then the rows tests + rates
set -x; mkdir -p gpurun_out
TAG=${TAG:-r2t}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stream -s 6 -c 1 -f -o gpurun_out/${TAG}_kstream_encode_b20 \
  python bench.py --steps 1 --warmup 3 --int-bits 20 --clients 10 --elements 25000000 --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${TAG}_ncu_b20.log 2>&1
echo rc=$?
timeout 600 python -m pytest tests/test_rows_f_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_pytest_rows.log
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
tail -4 gpurun_out/${TAG}_pytest_rows.log; cut -c1-150 gpurun_out/${TAG}_rows.jsonl
