set -x; mkdir -p gpurun_out
TAG=${TAG:-r2x}
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_topk --csv --log-file gpurun_out/${TAG}_topk_launches.csv python scripts/topk_once.py > gpurun_out/${TAG}_topk.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/${TAG}_topk_launches.csv')) if len(r) > 10 and r[0].isdigit()]
d = {}
for r in rows:
    d.setdefault(int(r[0]), [r[4][:40]]).append(r[-1])
for i in sorted(d): print(i, d[i])
PY
