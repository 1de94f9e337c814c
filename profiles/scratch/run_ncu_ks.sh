export PYTHONPATH=$PWD
AZG_RUN_PRECISION=f16f8ks python profiles/run_c4_forward.py > /dev/null 2>&1 || exit 1
AZG_RUN_PRECISION=f16f8ks timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tc_kernel" -s 2 -c 2 -o gpurun_out/r02_gemm_ks -f python profiles/run_c4_forward.py > gpurun_out/r02_ncu_ks.log 2>&1
tail -3 gpurun_out/r02_ncu_ks.log
