set -x; mkdir -p gpurun_out
TAG=${TAG:-r2z}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_topk_(filter|write)" -s ${SKIP:-0} -c 2 -f -o gpurun_out/${TAG}_topk_full python scripts/topk_once.py > gpurun_out/${TAG}_topk_full.log 2>&1
echo rc=$?
