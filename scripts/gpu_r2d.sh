set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rows_f_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
tail -5 gpurun_out/${TAG}_pytest.log; grep -h "stats" gpurun_out/${TAG}_rows.jsonl | cut -c1-200; tail -3 gpurun_out/${TAG}_rows.err
