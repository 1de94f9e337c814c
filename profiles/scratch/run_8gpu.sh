export PYTHONPATH=$PWD
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/bench_r2i_8gpu.log 2> gpurun_out/bench_r2i_8gpu.err; echo "bench8 rc=$?"
