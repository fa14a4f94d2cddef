set -x; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rows_f_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_rows.log
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
tail -3 gpurun_out/${TAG}_pytest_rows.log; cut -c1-140 gpurun_out/${TAG}_rows.jsonl
