set -x; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=3 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
timeout 900 python scripts/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2>gpurun_out/${TAG}_configs.err
FLASHE_PDL=0 timeout 900 python scripts/bench_configs.py > gpurun_out/${TAG}_configs_nopdl.jsonl 2>gpurun_out/${TAG}_configs_nopdl.err
timeout 900 python bench.py --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
tail -8 gpurun_out/${TAG}_pytest.log; cut -c1-330 gpurun_out/${TAG}_configs.jsonl; echo; cut -c1-330 gpurun_out/${TAG}_configs_nopdl.jsonl | head -3; tail -5 gpurun_out/${TAG}_configs.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['phases'])"
