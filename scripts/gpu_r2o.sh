set -x; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
python -c "import os; print('cpu_count', os.cpu_count())" >> gpurun_out/${TAG}_topo.txt
timeout 900 python bench.py --no-variants > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e'])); print(d['cpu_baseline']['value'], d['cpu_baseline']['device_vs_reference_on_the_timed_run'])"
tail -n 3 gpurun_out/${TAG}_bench.err; head -12 gpurun_out/${TAG}_topo.txt
