# one full ncu capture of the encode kernel at int_bits 20 (25M elements x 10 clients)
set -x; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stream -s 6 -c 1 -f -o gpurun_out/${TAG}_kstream_encode_b20 \
  python bench.py --steps 1 --warmup 3 --int-bits 20 --clients 10 --elements 25000000 --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${TAG}_ncu_b20.log 2>&1
echo rc=$?
