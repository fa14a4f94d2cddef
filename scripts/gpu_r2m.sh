# full ncu capture of the movement kernels (second launch of each), digested ON the box (the reports are too big to bring back)
set -x; mkdir -p gpurun_out /tmp/ncu
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_decode_v4|k_wire_pack32_bulk|k_wire_unpack32_bulk|k_stats_groups|k_encode_premasked_v4|k_aggregate_vec|k_aggregate_packed|k_topk|k_add_premasked_v4' -c 60 -f -o /tmp/ncu/rows python scripts/bench_rows.py --once > gpurun_out/${TAG}_ncu_rows.log 2>&1
ncu -i /tmp/ncu/rows.ncu-rep --page raw --csv > /tmp/ncu/raw.csv 2>/dev/null
python scripts/ncu_rows_digest.py /tmp/ncu/raw.csv > gpurun_out/${TAG}_ncu_rows_digest.csv
wc -l gpurun_out/${TAG}_ncu_rows_digest.csv; cut -c1-250 gpurun_out/${TAG}_ncu_rows_digest.csv | head -50
