#!/bin/bash
# Round-2 profile collection (run under gpurun from the repo root; outputs land in gpurun_out/, summaries are made
# afterwards in the build container with profiles/summarize.py).  Every ncu pass follows the plain run of its command.
export PYTHONPATH=$PWD
BENCH_SHORT="python bench.py --steps 2 --warmup 1 --preheat 0 --selfplay-games 0 --train-epochs 0 --coach-eps 0 --cpu-sample 64"
$BENCH_SHORT > gpurun_out/r02_bench_short.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $BENCH_SHORT > gpurun_out/r02_ncu_launches.log 2>&1
python profiles/run_c4_forward.py > /dev/null 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"c4_trunk_tc_kernel|gemm_bf16_tc_kernel" -s 3 -c 3 -o gpurun_out/r02_main_kernels_f16f8 -f python profiles/run_c4_forward.py > gpurun_out/r02_ncu_main.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"gemm_bf16_tc_kernel|im2col3x3_image_kernel|ttt_" -s 14 -c 14 -o gpurun_out/r02_ttt_forward -f python profiles/run_ttt_forward.py > gpurun_out/r02_ncu_ttt.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"fl_forward_kernel" -s 1 -c 1 -o gpurun_out/r02_fl_forward -f python profiles/run_fl_forward.py > gpurun_out/r02_ncu_fl.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"arena_select_compact_kernel|arena_expand_backup_compact_kernel" -s 40 -c 4 -o gpurun_out/r02_arena_16k -f python profiles/run_selfplay.py 4 16384 > gpurun_out/r02_ncu_arena.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python profiles/sanitize_small.py > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python profiles/sanitize_small.py > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_memcheck.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
tail -2 gpurun_out/r02_racecheck.log gpurun_out/r02_memcheck.log
