# interleaved A/B of library variants on the dense configs: LIBS="a b" bash scripts/gpu_ab_configs.sh
set -x; mkdir -p gpurun_out
for rep in 1 2; do
for lib in $LIBS; do
  FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so python scripts/bench_configs.py --dense-only > gpurun_out/abc_${lib}_r${rep}.jsonl 2>gpurun_out/abc_${lib}.err
  python - <<PY
import json
for l in open('gpurun_out/abc_${lib}_r${rep}.jsonl'):
    d = json.loads(l); print('$lib', d['config'][:44], round(d['ms_per_round'], 4), round(d.get('ms_per_round_separate_calls', 0), 4))
PY
done
done
