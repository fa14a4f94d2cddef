# usage: gpurun --gpus N -- 'bash scripts/gpu_${TAG:-r2p}_multi.sh N'
N=${1:-2}
set -x; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/${TAG:-r2p}_smi_n$N.txt
nvidia-smi topo -m > gpurun_out/${TAG:-r2p}_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/${TAG:-r2p}_multi_check_n$N.json 2> gpurun_out/${TAG:-r2p}_multi_check_n$N.err
echo "multi_check rc=$?"
timeout 900 $TR --master-port 29512 bench.py --gpus $N > gpurun_out/${TAG:-r2p}_bench_n$N.json 2> gpurun_out/${TAG:-r2p}_bench_n$N.err
echo "bench rc=$?"
timeout 600 $TR --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG:-r2p}_bench_ref_n$N.json 2> gpurun_out/${TAG:-r2p}_bench_ref_n$N.err
echo "ref rc=$?"
cat gpurun_out/${TAG:-r2p}_multi_check_n$N.json; tail -3 gpurun_out/${TAG:-r2p}_multi_check_n$N.err; head -c 2500 gpurun_out/${TAG:-r2p}_bench_n$N.json; tail -3 gpurun_out/${TAG:-r2p}_bench_n$N.err; head -c 600 gpurun_out/${TAG:-r2p}_bench_ref_n$N.json
# same box, one GPU: the baseline of the scaling efficiency (the ciphertext format depends on the box's cpu_count)
timeout 600 python bench.py --no-variants --no-cpu-baseline --no-e2e > gpurun_out/${TAG:-r2p}_bench_n1_same_box.json 2> gpurun_out/${TAG:-r2p}_bench_n1_same_box.err
python -c "
import json
a=json.load(open('gpurun_out/${TAG:-r2p}_bench_n1_same_box.json')); b=json.load(open('gpurun_out/${TAG:-r2p}_bench_n$N.json'))
print('N=1', a['value'], 'N=$N', b['value'], 'efficiency', b['value']/a['value']/$N, 'e2e', b['e2e']['value'], b['e2e']['frac_of_h2d_peak'], b['e2e']['h2d_probe'])"
