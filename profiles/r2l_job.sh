# r2l (2 GPUs): the N > 1 paths on hardware after c145997 (split P lines, sharded upload): torchrun check + one-process n_devices check
# against the oracle, then bench --gpus 2 (cfg5's first 16 assemblies, oracle SHA-256 gate) and the N=1 line beside it on the same box
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/r2l_pytest_multi.log 2>&1 || { echo PYTEST FAILED; tail -40 gpurun_out/r2l_pytest_multi.log; }
tail -3 gpurun_out/r2l_pytest_multi.log
show() { python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'], d['roofline']['frac'], d['stage_ms'], d.get('exchange'), d.get('limiting_stage'))"; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2l_err_n2.log | tee gpurun_out/r2l_bench_n2.json | show N2 || { echo BENCH N2 FAILED; tail -20 gpurun_out/r2l_err_n2.log; }
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2l_err_n1.log | tee gpurun_out/r2l_bench_n1_cfg5.json | show N1cfg5 || { echo BENCH N1 FAILED; tail -20 gpurun_out/r2l_err_n1.log; }
ls -la gpurun_out/
