export PYTHONPATH=$PWD
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_all6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_all6.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke2.log
python bench.py > gpurun_out/bench_r2h.log 2> gpurun_out/bench_r2h.err; echo "bench rc=$?"
BENCH_SHORT="python bench.py --steps 2 --warmup 1 --preheat 0 --selfplay-games 0 --train-epochs 0 --coach-eps 0 --cpu-sample 64"
$BENCH_SHORT > gpurun_out/r02_bench_short.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_final.csv $BENCH_SHORT > gpurun_out/r02_ncu_launches_final.log 2>&1
python profiles/run_c4_forward.py > /dev/null 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"c4_trunk|gemm_bf16_tc_kernel" -s 3 -c 3 -o gpurun_out/r02_main_kernels_final -f python profiles/run_c4_forward.py > gpurun_out/r02_ncu_main_final.log 2>&1
echo done
