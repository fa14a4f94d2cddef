# A/B of library variants on the tiled sparse kernels: LIBS="a b" (ncu launch list of scripts/sparse_once.py per variant)
set -x; mkdir -p gpurun_out
for lib in $LIBS; do
  L=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so; [ "$lib" = default ] && L=$PWD/flashe_b200/_lib/libflashe_b200.so
  FLASHE_B200_LIB=$L timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_sparse" -c 24 --csv --log-file gpurun_out/ab_sparse_$lib.csv python scripts/sparse_once.py > /dev/null 2>&1
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/ab_sparse_$lib.csv')) if len(r)>10 and r[0].isdigit()]
d=collections.defaultdict(list)
order=[]
for r in rows:
    k=r[4][:32]
    d[k].append(float(r[-1])/1e3)
print('$lib', {k:[round(x,1) for x in v[:6]] for k,v in d.items()})
PY
done
