export PYTHONPATH=$PWD
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r2i_2gpu.log 2> gpurun_out/bench_r2i_2gpu.err; echo "bench2 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.log 2> gpurun_out/bench_ref_2gpu.err; echo "ref2 rc=$?"; tail -c 300 gpurun_out/bench_ref_2gpu.log
