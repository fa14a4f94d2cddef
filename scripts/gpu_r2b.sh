# round 2: GPU tests after the source split / reference arm, row kernels, bench (both arms)
set -x; mkdir -p gpurun_out
python -c "import os; print('cpu_count', os.cpu_count())" > gpurun_out/${TAG}_host.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/bench_rows.py > gpurun_out/${TAG}_rows.jsonl 2>gpurun_out/${TAG}_rows.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>gpurun_out/${TAG}_bench_ref.err
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_pytest.log; grep -h "aggregate" gpurun_out/${TAG}_rows.jsonl | cut -c1-160; tail -3 gpurun_out/${TAG}_rows.err
head -c 700 gpurun_out/${TAG}_bench_reference_arm.json; echo; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['phases']); print(d.get('cpu_baseline')); print(d.get('e2e'))"
tail -3 gpurun_out/${TAG}_bench.err
