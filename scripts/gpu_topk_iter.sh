# top-k iteration: parity tests of row f2, event-timed rates, ncu launch list
set -x; mkdir -p gpurun_out
TAG=${TAG:-r2y}
timeout 600 python -m pytest tests/test_rows_f_gpu.py -m gpu -x -q -k "sparsify or sparse" 2>&1 | tail -4
timeout 600 python scripts/bench_rows.py 2>gpurun_out/${TAG}_rows.err | grep topk | cut -c1-120
TAG=$TAG bash scripts/gpu_ncu_topk.sh 2>/dev/null | grep -E "^(3[4-9]|4[0-9]|5[0-9]) "
