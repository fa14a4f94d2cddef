# round-1 final validation on one B200: tests, smoke, bench (both arms), launch list, full-size ncu capture, row kernels
set -x; mkdir -p gpurun_out
python -c "import os; print('cpu_count', os.cpu_count())" > gpurun_out/r1t_host.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> gpurun_out/r1t_host.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/r1t_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1t_smoke.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/r1t_bench_reference_arm.json 2>gpurun_out/r1t_bench_ref.err
timeout 900 python bench.py > gpurun_out/r1t_bench.json 2>gpurun_out/r1t_bench.err
timeout 600 python scripts/bench_rows.py > gpurun_out/r1t_rows.jsonl 2>gpurun_out/r1t_rows.err
timeout 600 python scripts/bench_configs.py > gpurun_out/r1t_configs.jsonl 2>gpurun_out/r1t_configs.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/r1t_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/r1t_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stream -s 6 -c 1 -f -o gpurun_out/r1t_kstream_encode \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-variants --no-cpu-baseline > gpurun_out/r1t_ncu_full.log 2>&1
tail -4 gpurun_out/r1t_pytest.log; cat gpurun_out/r1t_smoke.log | tail -2; head -c 1200 gpurun_out/r1t_bench.json; echo; head -c 400 gpurun_out/r1t_bench_reference_arm.json
