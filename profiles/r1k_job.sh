set -x
mkdir -p gpurun_out
timeout 120 python profiles/pcie_probe2.py > gpurun_out/k_pcie2.log 2>&1; cat gpurun_out/k_pcie2.log
AC_HOST_PROFILE=1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_cfg2.json 2> gpurun_out/k_bench_cfg2.err; python -c "
import json; d=json.load(open('gpurun_out/k_bench_cfg2.json')); print('cfg2', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])"; grep "host\] adopt" gpurun_out/k_bench_cfg2.err | tail -1
AC_HOST_CANDIDATES=1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_cfg2_hostcands.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/k_bench_cfg2_hostcands.json')); print('cfg2 host candidates', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])"
AC_HOST_PROFILE=1 timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/k_bench_cfg4.json 2> gpurun_out/k_bench_cfg4.err; python -c "
import json; d=json.load(open('gpurun_out/k_bench_cfg4.json')); print('cfg4', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])"; grep "host\] adopt" gpurun_out/k_bench_cfg4.err | tail -1
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/k_pytest.log 2>&1; tail -2 gpurun_out/k_pytest.log
