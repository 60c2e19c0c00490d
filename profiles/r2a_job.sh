# r2a: A/B of the device simplify / device GFA paths (stage_ms per arm), launch list and full ncu of every kernel >= 3 % (baseline of the round)
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
for mode in host devsimp devgfa; do
  for wl in cfg2 cfg4; do
    case $mode in host) E="";; devsimp) E="AC_DEVICE_SIMPLIFY=1";; devgfa) E="AC_DEVICE_SIMPLIFY=1 AC_DEVICE_GFA=1";; esac
    env $E AC_BENCH_ALLOW_UNCHECKED=1 timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2a_bench_${wl}_${mode}.json 2> gpurun_out/r2a_bench_${wl}_${mode}.err
    python -c "
import json,sys; d=json.load(open('gpurun_out/r2a_bench_${wl}_${mode}.json')); print('$wl $mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'], d['stage_ms'])"
  done
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2a_launches_cfg2.csv python profiles/profile_build.py cfg2 51 > gpurun_out/r2a_launches_cfg2.log 2>&1; tail -1 gpurun_out/r2a_launches_cfg2.log
AC_DEVICE_SIMPLIFY=1 AC_DEVICE_GFA=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2a_launches_cfg2_devgfa.csv python profiles/profile_build.py cfg2 51 > gpurun_out/r2a_launches_cfg2_devgfa.log 2>&1; tail -1 gpurun_out/r2a_launches_cfg2_devgfa.log
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'InsertLaneBody|InsertBody|AdjacencyBody|BoundaryBody|BloomBuildBody|ChunkMinBody|EmitSeqBody|OccupiedListBody|ExportFlagBody|InitSlotsBody|PackBody' -o gpurun_out/r2a_kernels_cfg2 -f python profiles/profile_build.py cfg2 51 > gpurun_out/r2a_ncu_full.log 2>&1; tail -2 gpurun_out/r2a_ncu_full.log
ls -la gpurun_out/
