set -x; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
FLASHE_E2E_TEST_SKEW=1 timeout 900 $TR --master-port 29512 bench.py --gpus 2 --no-variants > gpurun_out/${TAG}_bench_n2_skew.json 2> gpurun_out/${TAG}_bench_n2_skew.err
echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n2_skew.json')); print(d['value'], json.dumps(d['e2e'])[:1200]); print(d['roofline_prf'])"
tail -n 5 gpurun_out/${TAG}_bench_n2_skew.err
