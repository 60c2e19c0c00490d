set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py > gpurun_out/i_mg_check2.log 2>&1; grep -E "same|DIFFERENT|MULTI_GPU_CHECK|Error" gpurun_out/i_mg_check2.log | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_2gpu.log 2>&1; grep "^{" gpurun_out/i_bench_2gpu.log | cut -c1-1500
