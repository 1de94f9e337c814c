export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_all7.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/gpu_all7.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke3.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke3.log
python bench.py > gpurun_out/bench_r2i.log 2> gpurun_out/bench_r2i.err; echo "bench rc=$?"
