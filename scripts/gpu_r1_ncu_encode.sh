set -x; mkdir -p gpurun_out
# launch list of OUR kernels only (torch's normal_ fills are excluded by the name filter)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 40 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
# encode kernel (3rd k_stream launch = encode of the 2nd warm-up round), smaller shape for replay time
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 2 -c 1 -o gpurun_out/r1_prof_encode python bench.py --steps 1 --warmup 3 --elements 40000000 --clients 8 --no-e2e --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 2 -c 1 -o gpurun_out/r1_prof_encode_share python bench.py --steps 1 --warmup 3 --elements 40000000 --clients 8 --share-streams 1 --no-e2e --no-cpu-baseline >> gpurun_out/r1_ncu_full.log 2>&1
ls -la gpurun_out
