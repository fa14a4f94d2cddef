# rows: timings + full ncu capture of every movement kernel (second launch of each), digests to gpurun_out
set -x; mkdir -p gpurun_out
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
python -c "
import json
for l in open('gpurun_out/${TAG}_rows.jsonl'):
    d=json.loads(l); print('%-50s %8.3f ms %7.0f GB/s %5.1f%%' % (d['kernel'][:50], d['ms'], d['gbs'], 100*d['frac_of_hbm']))"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_decode_v4|k_wire_pack32_bulk|k_wire_unpack32_bulk|k_stats_groups|k_encode_premasked_v4|k_aggregate_vec|k_aggregate_packed|k_topk|k_add_premasked_v4' -c 60 -f -o gpurun_out/${TAG}_ncu_rows python scripts/bench_rows.py --once > gpurun_out/${TAG}_ncu_rows.log 2>&1
ls -la gpurun_out/${TAG}_ncu_rows.ncu-rep; tail -n 3 gpurun_out/${TAG}_ncu_rows.log
