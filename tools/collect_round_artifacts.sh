#!/bin/bash
# Runs on the GPU box (under gpurun): tests, bench, ncu launch list and full captures -> gpurun_out/
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --ba-problems 8 --cpu-frames 2 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fast|k_blur7|k_orient_describe|k_match_dir|k_select|k_resize|k_match_emit" -s 39 -c 13 -o gpurun_out/orb_full python tools/quick_bench.py 32 3 > gpurun_out/orb_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ba_step" -c 2 -o gpurun_out/ba_full python tools/ba_one_call.py > gpurun_out/ba_full.log 2>&1
tail -2 gpurun_out/pytest_gpu.log; tail -c 600 gpurun_out/bench_n1.json
