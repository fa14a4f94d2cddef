# compute-sanitizer passes over small invocations of every kernel family (memcheck + racecheck on the shared-memory paths)
set -x; mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_rows_f_gpu.py -m gpu -x -q -k "wire_vs_oracle or wire_unaligned or sparsify or stats_golden or stats_many" > gpurun_out/${TAG}_memcheck_rows.log 2>&1; echo "memcheck rows rc=$?"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sparse_sum or decode_division or packed_aggregate or masks_vs_oracle or batch120_fused or peer or full_path" > gpurun_out/${TAG}_memcheck_parity.log 2>&1; echo "memcheck parity rc=$?"
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "masks_vs_oracle" > gpurun_out/${TAG}_racecheck_masks.log 2>&1; echo "racecheck masks rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_rows_f_gpu.py -m gpu -x -q -k "sparsify_vs_oracle or sparsify_ties" > gpurun_out/${TAG}_racecheck_topk.log 2>&1; echo "racecheck topk rc=$?"
for f in gpurun_out/${TAG}_*check_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $f | tail -4; done
