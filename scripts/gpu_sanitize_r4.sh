# compute-sanitizer over the kernels added in the last sessions: dynamic deal (tickets), edge items, tiled sparse kernels
set -x; mkdir -p gpurun_out
TAG=${TAG:-r4s}
export PYTHONUNBUFFERED=1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sparse_sum or sparse_apply_masks_batch or flashecipher_sparse or dynamic_deal or cuda_graph_replay" > gpurun_out/${TAG}_memcheck_new.log 2>&1; echo "memcheck new rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sparse_sum or sparse_apply_masks_batch or flashecipher_sparse" > gpurun_out/${TAG}_racecheck_sparse.log 2>&1; echo "racecheck sparse rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "masks_vs_oracle or full_path" > gpurun_out/${TAG}_racecheck_stream.log 2>&1; echo "racecheck stream rc=$?"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
for f in gpurun_out/${TAG}_*check_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $f | tail -4; done
