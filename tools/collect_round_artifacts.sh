#!/bin/bash
# Runs on the GPU box (under gpurun): tests, bench, ncu launch list and full captures -> gpurun_out/.
# usage: collect_round_artifacts.sh [orb|ba|all]   (gpurun brings back at most 64 MiB per call: the ORB and the BA captures go in two calls)
PART=${1:-all}
set -x
if [ "$PART" = orb ] || [ "$PART" = all ]; then
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --ba-problems 8 --cpu-frames 2 > gpurun_out/bench_under_ncu.log 2>&1
# one 32-frame extract + match step, every ORB / match kernel (the first 3 warm-up steps = 42 matching launches are skipped)
ncu --set full --clock-control none --import-source on -k regex:"k_fast|k_blur|k_orient_describe|k_match_dir|k_select|k_resize|k_match_emit" -s 42 -c 14 -o gpurun_out/orb_full python tools/quick_bench.py 32 3 > gpurun_out/orb_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_match_dir|k_match_emit" -s 6 -c 2 -o gpurun_out/match_full python tools/quick_bench.py 32 3 > gpurun_out/match_full.log 2>&1
# the register-staged variant of FAST (the TMA-staged one is the default and part of orb_full)
MAGE_FAST_TMA=0 ncu --set full --clock-control none --import-source on -k regex:"k_fast" -s 3 -c 1 -o gpurun_out/fast_reg_full python tools/quick_bench.py 32 3 > gpurun_out/fast_reg_full.log 2>&1
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1
compute-sanitizer --tool memcheck python tools/sanitize_match.py > gpurun_out/sanitizer_match.log 2>&1
tail -2 gpurun_out/pytest_gpu.log; tail -c 600 gpurun_out/bench_n1.json
fi
if [ "$PART" = ba ] || [ "$PART" = all ]; then
# bundle adjustment: the cooperative single-window kernel and the batched one-CTA-per-window kernel (296 windows, 10 LM iterations)
ncu --set full --clock-control none --import-source on -k regex:"k_ba_step_coop" -c 2 -o gpurun_out/ba_full python tools/ba_one_call.py > gpurun_out/ba_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ba_step" -s 1 -c 1 -o gpurun_out/ba_many_full python tools/ba_many_call.py 296 10 > gpurun_out/ba_many_full.log 2>&1
# global BA (config 4): the cooperative kernel at full size (second launch = first timed step) and the dense solver of the reduced system alone
ncu --set full --clock-control none --import-source on -k regex:"k_ba_step_coop" -s 1 -c 1 -o gpurun_out/ba_global_full python tools/ba_global_call.py > gpurun_out/ba_global_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_dense_debug" -s 1 -c 1 -o gpurun_out/dense_full python tools/dense_call.py > gpurun_out/dense_full.log 2>&1
python tools/ba_global_call.py > gpurun_out/ba_global_phases.log 2>&1
python tools/dense_call.py > gpurun_out/dense_phases.log 2>&1
python tools/quick_bench_ba.py > gpurun_out/quick_bench_ba.log 2>&1
python tools/global_ba_check.py > gpurun_out/global_ba_check.log 2>&1
fi
du -sh gpurun_out
