# short validation of the final state on one B200: tests, smoke, both bench arms, ncu launch list
set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12 > gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2>gpurun_out/${TAG}_bench_ref.err
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_pytest.log; tail -1 gpurun_out/${TAG}_smoke.log; head -c 600 gpurun_out/${TAG}_bench.json
