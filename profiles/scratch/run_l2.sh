export PYTHONPATH=$PWD
for i in 1 2; do
python profiles/time_c4.py f16f8 2>&1 | tail -1
AZG_L2_PERSIST=1 python profiles/time_c4.py f16f8 2>&1 | tail -1
done
AZG_L2_PERSIST=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"gemm_bf16_tc_kernel" -s 2 -c 2 python profiles/run_c4_forward.py 2>&1 | grep -E "gemm_bf16|duration|dram__|hit_rate|tensor" 
