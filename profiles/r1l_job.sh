set -x
mkdir -p gpurun_out
for mode in flush noflush; do
if [ $mode = noflush ]; then export AC_NO_RESULT_FLUSH=1; else unset AC_NO_RESULT_FLUSH; fi
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/l_bench_cfg2_$mode.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/l_bench_cfg2_$mode.json')); print('cfg2 $mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])"
done
unset AC_NO_RESULT_FLUSH
timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/l_bench_cfg4.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/l_bench_cfg4.json')); print('cfg4', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])"
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/l_pytest.log 2>&1; tail -2 gpurun_out/l_pytest.log
