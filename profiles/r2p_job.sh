# r2p (2 GPUs): why the resident-input loop at N = 2 was slower than the host-buffer loop (r2o): every rank's stages and per-step host clocks, both loop orders
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
show() { grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity']['ok'])
for r in d['per_rank']: print('   ', r)"; }
timeout 300 $TR --master-port 29721 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2p_err_a.log | tee gpurun_out/r2p_bench_n2_a.json | show default
AC_BENCH_E2E_FIRST=1 timeout 300 $TR --master-port 29722 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2p_err_b.log | tee gpurun_out/r2p_bench_n2_b.json | show e2e_first
NCCL_DEBUG=WARN timeout 300 $TR --master-port 29723 bench.py --gpus 2 --steps 10 --warmup 6 --no-cpu-baseline 2>gpurun_out/r2p_err_c.log | tee gpurun_out/r2p_bench_n2_c.json | show warmup6
