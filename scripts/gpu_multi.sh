# usage: gpurun --gpus N -- 'bash scripts/gpu_r1x_multi.sh N'
N=${1:-2}
set -x; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/r1x_smi_n$N.txt
nvidia-smi topo -m > gpurun_out/r1x_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/r1x_multi_check_n$N.json 2> gpurun_out/r1x_multi_check_n$N.err
echo "multi_check rc=$?"
timeout 900 $TR --master-port 29512 bench.py --gpus $N > gpurun_out/r1x_bench_n$N.json 2> gpurun_out/r1x_bench_n$N.err
echo "bench rc=$?"
timeout 600 $TR --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r1x_bench_ref_n$N.json 2> gpurun_out/r1x_bench_ref_n$N.err
echo "ref rc=$?"
cat gpurun_out/r1x_multi_check_n$N.json; tail -3 gpurun_out/r1x_multi_check_n$N.err; head -c 2500 gpurun_out/r1x_bench_n$N.json; tail -3 gpurun_out/r1x_bench_n$N.err; head -c 600 gpurun_out/r1x_bench_ref_n$N.json
