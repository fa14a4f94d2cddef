set -x; mkdir -p gpurun_out
TAG=${TAG:-r4k}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sparse_sum_tiled|k_sparse_overlap_tiled" -c 3 -f -o gpurun_out/${TAG}_sparse python scripts/sparse_once.py 2>&1 | tail -2
