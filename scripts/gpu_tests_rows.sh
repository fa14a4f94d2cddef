set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
tail -4 gpurun_out/${TAG}_pytest.log; cut -c1-150 gpurun_out/${TAG}_rows.jsonl
