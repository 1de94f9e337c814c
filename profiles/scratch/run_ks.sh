export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_nets_gpu.py tests/test_trained_gpu.py -x -q -m gpu -s > gpurun_out/ks5.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ks5.log
grep -n "trained checkpoint\|max |d pi|\|auto ->\|K-split" gpurun_out/ks5.log | cut -c1-250
timeout 200 python profiles/time_c4.py f16f8 f16f8ks f16f8ks+fold bf16x3ks > gpurun_out/ks5_time.log 2>&1; cat gpurun_out/ks5_time.log | tail -4
