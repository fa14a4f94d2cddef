set -x; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r1d_smi.txt
timeout 900 python -m pytest tests/test_rows_f_gpu.py -m gpu -q --durations=8 2>&1 | tail -60 > gpurun_out/r1d_pytest_rows.log
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -40 > gpurun_out/r1d_pytest.log
timeout 900 python bench.py > gpurun_out/r1d_bench.json 2>gpurun_out/r1d_bench.err
timeout 600 python scripts/bench_rows.py > gpurun_out/r1d_rows.jsonl 2>gpurun_out/r1d_rows.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1d_smoke.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1d_ncu_bench.log 2>&1
tail -5 gpurun_out/r1d_pytest_rows.log gpurun_out/r1d_pytest.log; cat gpurun_out/r1d_bench.json | head -c 3000; cat gpurun_out/r1d_rows.jsonl
