set -x; mkdir -p gpurun_out
TAG=${TAG:-r4j}
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_sparse|k_stream" -c 40 --csv --log-file gpurun_out/${TAG}_sparse_launches.csv python scripts/sparse_once.py 2>&1 | tail -2
