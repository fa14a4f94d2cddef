set -x; mkdir -p gpurun_out
for rep in 1 2; do
for lib in $LIBS; do
  export FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_${lib}r${rep}_unshared.json 2>gpurun_out/ab_${lib}.err
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --share-streams 1 > gpurun_out/ab_${lib}r${rep}_shared.json 2>>gpurun_out/ab_${lib}.err
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --int-bits 20 --clients 10 --elements 25000000 > gpurun_out/ab_${lib}r${rep}_b20.json 2>>gpurun_out/ab_${lib}.err
done
done
