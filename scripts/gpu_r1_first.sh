set -x; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --steps 5 --warmup 3 --share-streams 1 --no-cpu-baseline > gpurun_out/r1_bench_share.json 2>> gpurun_out/r1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 1 -c 1 -o gpurun_out/r1_prof_stream python bench.py --steps 1 --warmup 3 --elements 40000000 --clients 8 --no-e2e --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_aggregate -s 1 -c 1 -o gpurun_out/r1_prof_agg python bench.py --steps 1 --warmup 3 --elements 40000000 --clients 16 --no-e2e --no-cpu-baseline >> gpurun_out/r1_ncu_full.log 2>&1
ls -la gpurun_out
