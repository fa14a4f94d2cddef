set -x; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r1c_smi.txt
python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -40 > gpurun_out/r1c_pytest.log
python bench.py > gpurun_out/r1c_bench.json 2>gpurun_out/r1c_bench.err
python bench.py --share-streams 1 --no-e2e --no-cpu-baseline > gpurun_out/r1c_bench_shared.json 2>>gpurun_out/r1c_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1c_bench_reference.json 2>>gpurun_out/r1c_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1c_smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1c_ncu_bench.log 2>&1
