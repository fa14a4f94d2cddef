set -x; mkdir -p gpurun_out
TAG=${TAG:-r4e}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
python scripts/edge_only.py
for rep in 1 2; do
python scripts/bench_small.py --quick > gpurun_out/${TAG}_small_hyb_$rep.jsonl 2> gpurun_out/${TAG}_small.err
FLASHE_DYNAMIC_MIN_ITEMS=0 python scripts/bench_small.py --quick > gpurun_out/${TAG}_small_dyn_$rep.jsonl 2> gpurun_out/${TAG}_small.err
FLASHE_DYNAMIC=0 python scripts/bench_small.py --quick > gpurun_out/${TAG}_small_static_$rep.jsonl 2>> gpurun_out/${TAG}_small.err
done
FLASHE_DYNAMIC_MIN_ITEMS=0 FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_trace.so FLASHE_TRACE_PRINT=1 python scripts/trace_small.py 2> gpurun_out/${TAG}_trace.txt
