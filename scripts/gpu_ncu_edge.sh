# ncu --set full of an encode and a decode launch made of edge items only (scripts/edge_only.py)
set -x; mkdir -p gpurun_out
TAG=${TAG:-r4d}
python scripts/edge_only.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stream -s 6 -c 2 -f -o gpurun_out/${TAG}_edge python scripts/edge_only.py > gpurun_out/${TAG}_ncu_edge.log 2>&1
echo rc=$?
