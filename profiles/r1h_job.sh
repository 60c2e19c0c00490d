set -x
mkdir -p gpurun_out
AC_HOST_PROFILE=1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_cfg2.json 2> gpurun_out/h_bench_cfg2.err; python -c "
import json; d=json.load(open('gpurun_out/h_bench_cfg2.json')); print('cfg2', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])"; grep "host\]" gpurun_out/h_bench_cfg2.err | tail -14
AC_HOST_PROFILE=1 timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/h_bench_cfg4.json 2> gpurun_out/h_bench_cfg4.err; python -c "
import json; d=json.load(open('gpurun_out/h_bench_cfg4.json')); print('cfg4', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])"; grep "host\]" gpurun_out/h_bench_cfg4.err | tail -14
timeout 300 python -m pytest tests -m gpu -q -x -k "golden or larger or medium" > gpurun_out/h_pytest.log 2>&1; tail -2 gpurun_out/h_pytest.log
