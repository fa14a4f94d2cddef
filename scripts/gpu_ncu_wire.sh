set -x; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_wire -c 4 -f -o gpurun_out/${TAG}_wire python scripts/bench_rows.py > gpurun_out/${TAG}_wire_ncu.log 2>&1
echo rc=$?
