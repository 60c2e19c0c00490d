# r2o (4 GPUs): the N > 1 forms after the stream fix (api.py: the default stream is handed to the library by its handle; dist.py orders by
# stream only when the library really runs on torch's stream): GPU multi-rank tests at 2 and 4, bench N = 1 (cfg5 first 8), 2, 4
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/r2o_pytest_multi.log 2>&1 || { echo PYTEST FAILED; tail -40 gpurun_out/r2o_pytest_multi.log | cut -c1-600; }
tail -3 gpurun_out/r2o_pytest_multi.log
show() { grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'], d['roofline']['frac'], d['stage_ms'], d.get('exchange'), d.get('limiting_stage'))"; }
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2o_err_n1.log | tee gpurun_out/r2o_bench_n1_cfg5.json | show N1cfg5 || { echo BENCH N1 FAILED; tail -5 gpurun_out/r2o_err_n1.log; }
for n in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2o_err_n$n.log | tee gpurun_out/r2o_bench_n$n.json | show N$n || { echo BENCH N$n FAILED; grep -m5 "Error" gpurun_out/r2o_err_n$n.log; }
done
ls -la gpurun_out/
