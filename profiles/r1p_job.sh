set -x
mkdir -p gpurun_out
for mode in device host; do
if [ $mode = host ]; then export AC_HOST_RENUMBER=1; else unset AC_HOST_RENUMBER; fi
AC_HOST_PROFILE=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench_cfg2_$mode.json 2> gpurun_out/p_bench_cfg2_$mode.err; python -c "
import json; d=json.load(open('gpurun_out/p_bench_cfg2_$mode.json')); print('cfg2 renumber=$mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])"; grep "host\] renumber" gpurun_out/p_bench_cfg2_$mode.err | tail -1; grep "host\] adopt:" gpurun_out/p_bench_cfg2_$mode.err | tail -1
done
unset AC_HOST_RENUMBER
timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/p_bench_cfg4.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/p_bench_cfg4.json')); print('cfg4', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])"
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/p_pytest.log 2>&1; tail -2 gpurun_out/p_pytest.log
