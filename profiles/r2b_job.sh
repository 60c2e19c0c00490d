# r2b: 8-byte slots + block-aligned insert: GPU suite, bench A/B, full ncu of the insert and the sizing pass
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
for mode in host devgfa; do
  for wl in cfg2 cfg4; do
    case $mode in host) E="";; devgfa) E="AC_DEVICE_SIMPLIFY=1 AC_DEVICE_GFA=1";; esac
    env $E timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2b_bench_${wl}_${mode}.json 2> gpurun_out/r2b_bench_${wl}_${mode}.err
    python -c "
import json,sys; d=json.load(open('gpurun_out/r2b_bench_${wl}_${mode}.json')); print('$wl $mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'], d['roofline']['frac'], d['stage_ms'])"
  done
done
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'InsertBody|AdjacencyBody|BoundaryBody|BloomBuildBody' -o gpurun_out/r2b_kernels_cfg2 -f python profiles/profile_build.py cfg2 51 > gpurun_out/r2b_ncu_full.log 2>&1; tail -2 gpurun_out/r2b_ncu_full.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2b_launches_cfg2.csv python profiles/profile_build.py cfg2 51 > gpurun_out/r2b_launches_cfg2.log 2>&1
