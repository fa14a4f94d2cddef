set -x; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -40 > gpurun_out/r1b_pytest.log
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1b_unshared.json 2>gpurun_out/r1b.err
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --share-streams 1 > gpurun_out/r1b_shared.json 2>>gpurun_out/r1b.err
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --int-bits 20 --clients 10 --elements 25000000 > gpurun_out/r1b_b20.json 2>>gpurun_out/r1b.err
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 2 -c 1 -o gpurun_out/r1b_prof_encode python bench.py --steps 1 --warmup 3 --elements 40000000 --clients 8 --no-e2e --no-cpu-baseline > gpurun_out/r1b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 2 -c 1 -o gpurun_out/r1b_prof_encode_share python bench.py --steps 1 --warmup 3 --elements 40000000 --clients 8 --share-streams 1 --no-e2e --no-cpu-baseline >> gpurun_out/r1b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 2 -c 1 -o gpurun_out/r1b_prof_encode_b20 python bench.py --steps 1 --warmup 3 --elements 40000000 --clients 8 --int-bits 20 --no-e2e --no-cpu-baseline >> gpurun_out/r1b_ncu.log 2>&1
