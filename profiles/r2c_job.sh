# r2c: group-probed insert, centre-sampled sizing pass, cooperative level/apply kernels, shared-memory tile sort
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
for mode in host devgfa; do
  for wl in cfg2 cfg4; do
    case $mode in host) E="";; devgfa) E="AC_DEVICE_SIMPLIFY=1 AC_DEVICE_GFA=1";; esac
    env $E timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2c_bench_${wl}_${mode}.json 2> gpurun_out/r2c_bench_${wl}_${mode}.err
    python -c "
import json,sys; d=json.load(open('gpurun_out/r2c_bench_${wl}_${mode}.json')); print('$wl $mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'], d['roofline']['frac'], d['stage_ms'])"
  done
done
for occ in 5 8; do AC_INSERT_OCC=$occ timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('occ $occ', d['value'], d['stage_ms']['insert'])"; done
for load in 0.4 0.6 0.7; do AC_TABLE_LOAD=$load timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('load $load', d['value'], d['stage_ms']['insert'], d['stage_ms']['adjacency'])"; done
AC_DEVICE_SIMPLIFY=1 AC_DEVICE_GFA=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'InsertBody|SampleBody|AdjacencyBody|ApplyPass|Levels|tile_sort' -o gpurun_out/r2c_kernels_cfg2 -f python profiles/profile_build.py cfg2 51 > gpurun_out/r2c_ncu_full.log 2>&1; tail -2 gpurun_out/r2c_ncu_full.log
AC_DEVICE_SIMPLIFY=1 AC_DEVICE_GFA=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2c_launches_cfg2_devgfa.csv python profiles/profile_build.py cfg2 51 > gpurun_out/r2c_launches_cfg2.log 2>&1
