set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_rows.py 2>/dev/null | grep -E 'premasked|"decode"' | cut -c1-170
