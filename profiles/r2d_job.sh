# r2d: fused ac_compress path (device simplify + full GFA text, lazy graph), insert v3 (wide fingerprint, branch-free group scan, interior flags)
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_pytest.log 2>&1; tail -3 gpurun_out/r2d_pytest.log
for mode in fused hosttail; do
  for wl in cfg2 cfg4; do
    case $mode in fused) E="";; hosttail) E="AC_HOST_SIMPLIFY=1";; esac
    env $E timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2d_bench_${wl}_${mode}.json 2> gpurun_out/r2d_bench_${wl}_${mode}.err
    python -c "
import json,sys; d=json.load(open('gpurun_out/r2d_bench_${wl}_${mode}.json')); print('$wl $mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'], d['roofline']['frac'], d['stage_ms'])"
  done
done
tail -5 gpurun_out/r2d_bench_cfg2_fused.err
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'InsertBody|SampleBody|AdjacencyBody|PackBody' -o gpurun_out/r2d_kernels_cfg2 -f python profiles/profile_build.py cfg2 51 > gpurun_out/r2d_ncu_full.log 2>&1; tail -2 gpurun_out/r2d_ncu_full.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2d_launches_cfg2.csv python profiles/profile_build.py cfg2 51 > gpurun_out/r2d_launches_cfg2.log 2>&1
timeout 200 python profiles/cli_wall.py > gpurun_out/r2d_cli.log 2>&1; grep -E "^rep|load\+repair" gpurun_out/r2d_cli.log
