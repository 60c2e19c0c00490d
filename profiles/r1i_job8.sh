set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_8gpu.log 2>&1; grep "^{" gpurun_out/i_bench_8gpu.log | cut -c1-1800; grep -iE "error|Traceback" gpurun_out/i_bench_8gpu.log | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29619 tests/multi_gpu_check.py > gpurun_out/i_mg_check8.log 2>&1; grep -E "same|DIFFERENT|MULTI_GPU_CHECK|Error" gpurun_out/i_mg_check8.log | head
