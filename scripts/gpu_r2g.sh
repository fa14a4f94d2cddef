# round 2: micro-benchmarks (LDS ceiling, mask-only T-table kernel, bit-sliced AES-256) + full ncu captures of each,
# per-kernel times of the C1 round
set -x; mkdir -p gpurun_out
timeout 600 python scripts/microbench.py > gpurun_out/${TAG}_micro.jsonl 2>gpurun_out/${TAG}_micro.err
cat gpurun_out/${TAG}_micro.jsonl; tail -3 gpurun_out/${TAG}_micro.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lds_peak -s 2 -c 1 -f -o gpurun_out/${TAG}_ncu_lds python scripts/microbench.py --only lds > gpurun_out/${TAG}_ncu_lds.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_aes_bitslice -s 2 -c 1 -f -o gpurun_out/${TAG}_ncu_bitslice python scripts/microbench.py --only bitslice > gpurun_out/${TAG}_ncu_bitslice.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 4 -c 1 -f -o gpurun_out/${TAG}_ncu_ttable python scripts/microbench.py --only ttable > gpurun_out/${TAG}_ncu_ttable.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 60 --csv --log-file gpurun_out/${TAG}_c1_launches.csv python scripts/bench_c1.py > gpurun_out/${TAG}_c1.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_*.log; grep -c k_ gpurun_out/${TAG}_c1_launches.csv; ls -la gpurun_out/*.ncu-rep
