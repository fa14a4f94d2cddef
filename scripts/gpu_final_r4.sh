# round-2 (last sessions) final validation on one B200: tests, smoke, bench (both arms), rows, configs, launch list, full-size ncu capture
# of the headline encode launch and of the int_bits 20 encode launch.  TAG names the outputs (default r3f).
set -x; mkdir -p gpurun_out
TAG=${TAG:-r4z}
python -c "import os; print('cpu_count', os.cpu_count())" > gpurun_out/${TAG}_host.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> gpurun_out/${TAG}_host.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2>gpurun_out/${TAG}_bench_ref.err
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
timeout 600 python scripts/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2>gpurun_out/${TAG}_configs.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stream -s 6 -c 1 -f -o gpurun_out/${TAG}_kstream_encode \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stream -s 6 -c 1 -f -o gpurun_out/${TAG}_kstream_encode_b20 \
  python bench.py --steps 1 --warmup 3 --int-bits 20 --clients 10 --elements 25000000 --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${TAG}_ncu_b20.log 2>&1
TAG=$TAG bash scripts/gpu_ncu_topk.sh > /dev/null 2>&1
timeout 300 python scripts/bench_small.py > gpurun_out/${TAG}_small.jsonl 2>gpurun_out/${TAG}_small.err
TAG=$TAG bash scripts/gpu_ncu_c2.sh > /dev/null 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_smoke.log | tail -2; head -c 1200 gpurun_out/${TAG}_bench.json; echo; head -c 400 gpurun_out/${TAG}_bench_reference_arm.json
