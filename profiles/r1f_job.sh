# insert-kernel variants: resident CTAs per SM (register budget) after the probe/commit split
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest.log 2>&1; tail -2 gpurun_out/f_pytest.log
for occ in 5 6 8; do
AC_INSERT_OCC=$occ timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_occ$occ.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/f_bench_occ$occ.json')); print('occ', $occ, d['value'], d['ms_per_step'], d['stage_ms']['insert'], d['stage_ms']['adjacency'], d['roofline']['frac'])"
done
AC_INSERT_OCC=6 timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/f_bench_cfg4_occ6.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/f_bench_cfg4_occ6.json')); print('cfg4 occ6', d['value'], d['ms_per_step'], d['stage_ms'])"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:InsertLaneBody -c 1 -o gpurun_out/insert_r1f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_insert.log 2>&1; tail -1 gpurun_out/f_ncu_insert.log
for th in 32 48; do
AC_HOST_THREADS=$th timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_t$th.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/f_bench_t$th.json')); print('threads', $th, d['value'], d['ms_per_step'], d['stage_ms']['host_graph'], d['stage_ms']['host_simplify'], d['stage_ms']['host_gfa'])"
done
