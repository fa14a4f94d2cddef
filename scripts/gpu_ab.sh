set -x; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_pytest.log
for lib in base imad t640 t640imad t768 t768imad; do
  export FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_$lib.so
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_${lib}_unshared.json 2>gpurun_out/ab_${lib}.err
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --share-streams 1 > gpurun_out/ab_${lib}_shared.json 2>>gpurun_out/ab_${lib}.err
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --int-bits 20 --clients 10 --elements 25000000 > gpurun_out/ab_${lib}_b20.json 2>>gpurun_out/ab_${lib}.err
done
