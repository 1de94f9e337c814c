export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/ks4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ks4.log
grep -n "trained checkpoint\|max |d pi|\|auto ->\|K-split" gpurun_out/ks4.log | cut -c1-250
timeout 300 python profiles/time_c4.py f16f8 f16f8ks f16f8ks+fold > gpurun_out/ks4_time.log 2>&1; cat gpurun_out/ks4_time.log | tail -4
