set -x; mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_rows.py 2>/dev/null | grep -E '"decode"' | cut -c1-200
TAG=${TAG:-r3c} bash scripts/gpu_multi_check_only.sh $N 2>&1 | grep -E "all_ranks_ok|rc=" | cut -c1-700
