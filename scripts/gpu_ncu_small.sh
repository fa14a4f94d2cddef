# one full ncu capture of the encode kernel at 8 clients (in-tree library, or $FLASHE_B200_LIB)
set -x; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stream -s 6 -c 1 -f -o gpurun_out/${TAG}_kstream_encode_c8 \
  python bench.py --steps 1 --warmup 3 --clients 8 --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
echo rc=$?
