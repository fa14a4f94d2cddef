# ncu launch list of the C2 round (2.5 M elements x 10 clients, int_bits 20)
set -x; mkdir -p gpurun_out
TAG=${TAG:-r3p}
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:^k_ -c 40 --csv --log-file gpurun_out/${TAG}_c2_launches.csv python bench.py --steps 2 --warmup 3 --int-bits 20 --clients 10 --elements 2500000 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/${TAG}_c2.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/${TAG}_c2_launches.csv')) if len(r) > 10 and r[0].isdigit()]
d = {}
for r in rows: d.setdefault(int(r[0]), [r[4][:34], r[7], r[8]]).append(r[-1])
for i in sorted(d)[:24]: print(i, d[i])
PY
