# NCCL parity + peer-memory gather on N GPUs (no bench): gpurun --gpus N -- 'TAG=x bash scripts/gpu_multi_check_only.sh N'
N=${1:-2}
set -x; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/${TAG:-r3b}_multi_check_n$N.json 2> gpurun_out/${TAG:-r3b}_multi_check_n$N.err
echo "multi_check rc=$?"
cat gpurun_out/${TAG:-r3b}_multi_check_n$N.json; tail -5 gpurun_out/${TAG:-r3b}_multi_check_n$N.err
