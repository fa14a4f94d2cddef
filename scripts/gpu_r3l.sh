set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rows_f_gpu.py -m gpu -x -q -k "wire" 2>&1 | tail -3
timeout 300 python scripts/bench_rows.py 2>/dev/null | grep -E 'wire' | cut -c1-170
