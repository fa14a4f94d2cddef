# round 2: row-f tests (bit-exact statistics, whole-model drop-in, per-layer batching) + row kernels
set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rows_f_gpu.py tests/test_gpu_parity.py -m gpu -x -q --durations=5 2>&1 | tail -25 > gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
tail -25 gpurun_out/${TAG}_pytest.log; grep -h "stats" gpurun_out/${TAG}_rows.jsonl | cut -c1-200; tail -3 gpurun_out/${TAG}_rows.err
