set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench_4gpu.log 2>&1; grep "^{" gpurun_out/o_bench_4gpu.log | cut -c1-1800; grep -iE "error|Traceback" gpurun_out/o_bench_4gpu.log | head -3
