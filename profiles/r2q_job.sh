# r2q (2 GPUs): exchange buffers kept per handle (no allocator traffic between the collectives): do the outlier steps of r2p go away?
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
show() { grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity']['ok'])
for r in d['per_rank']: print('   ', r)"; }
for rep in a b c; do
timeout 300 $TR --master-port 29731 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2q_err_$rep.log | tee gpurun_out/r2q_bench_n2_$rep.json | show rep_$rep
done
