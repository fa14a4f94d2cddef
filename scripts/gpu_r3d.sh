set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rows_f_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "stats or peer or quantizing or mirrors" 2>&1 | tail -3
timeout 300 python scripts/bench_rows.py 2>/dev/null | grep -E 'segment_stats|"decode"|topk' | cut -c1-160
