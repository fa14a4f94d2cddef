# round 2, first validation: GPU tests after the 128-bit packed aggregate + ADVICE fixes, row kernels
set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_rows.jsonl | cut -c1-200; tail -3 gpurun_out/${TAG}_rows.err
