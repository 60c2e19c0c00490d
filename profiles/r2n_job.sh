# r2n (2 GPUs): why the torchrun form of the fused N > 1 build differs from the oracle on hardware (r2l) while the one-process form does not (r2m)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
AC_MULTI_CHECK_CASES=d,a timeout 200 $TR --master-port 29701 tests/multi_gpu_check.py > gpurun_out/r2n_check.log 2>&1; grep -v "^\[W\|^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2n_check.log | tail -12
AC_SYNC_LAUNCHES=1 AC_MULTI_CHECK_CASES=d,a timeout 200 $TR --master-port 29702 tests/multi_gpu_check.py > gpurun_out/r2n_check_sync.log 2>&1; grep -v "^\[W\|^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2n_check_sync.log | tail -12
AC_SYNC_LAUNCHES=1 timeout 300 $TR --master-port 29703 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2n_bench_sync.json 2> gpurun_out/r2n_bench_sync.err; grep -m3 "AutocyclerGpuError\|Error" gpurun_out/r2n_bench_sync.err; head -c 600 gpurun_out/r2n_bench_sync.json
CUDA_LAUNCH_BLOCKING=1 timeout 300 $TR --master-port 29704 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2n_bench_block.json 2> gpurun_out/r2n_bench_block.err; grep -m3 "AutocyclerGpuError\|Error" gpurun_out/r2n_bench_block.err; head -c 600 gpurun_out/r2n_bench_block.json
AC_MULTI_CHECK_CASES=d timeout 250 compute-sanitizer --tool memcheck --target-processes all --print-limit 10 $TR --master-port 29705 tests/multi_gpu_check.py > gpurun_out/r2n_memcheck.log 2>&1; grep "=========" gpurun_out/r2n_memcheck.log | head -60; grep -v "=========" gpurun_out/r2n_memcheck.log | grep -v "^\[W\|^W0\|^\*\*\*\|OMP_NUM" | tail -8
