set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch or packed or sparse_apply or many_dropout or ring" 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
timeout 900 python scripts/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2>gpurun_out/${TAG}_configs.err
tail -15 gpurun_out/${TAG}_pytest.log; cut -c1-420 gpurun_out/${TAG}_configs.jsonl; tail -5 gpurun_out/${TAG}_configs.err
