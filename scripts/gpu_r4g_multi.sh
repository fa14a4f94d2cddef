# N GPUs: NCCL parity incl. peer gather, and the device-timed bench line at N and at 1 on the same box
N=${1:-2}
set -x; mkdir -p gpurun_out
TAG=${TAG:-r4g}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/${TAG}_multi_check_n$N.json 2> gpurun_out/${TAG}_multi_check_n$N.err
echo "multi_check rc=$?"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench rc=$?"
timeout 600 python bench.py --no-variants --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_n1_same_box.json 2> gpurun_out/${TAG}_bench_n1_same_box.err
cat gpurun_out/${TAG}_multi_check_n$N.json | head -c 1500; tail -3 gpurun_out/${TAG}_multi_check_n$N.err
python -c "
import json
a=json.load(open('gpurun_out/${TAG}_bench_n1_same_box.json')); b=json.load(open('gpurun_out/${TAG}_bench_n$N.json'))
print('N=1', a['value'], 'N=$N', b['value'], 'efficiency', b['value']/a['value']/$N)"
