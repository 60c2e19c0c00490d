# r2h (1 GPU): expand_repeats loop in one cooperative launch, thread-per-coordinate boundaries, insert group scan on high words
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log
run() { env $1 timeout 300 python bench.py --workload ${2:-cfg2} --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r2h_err_$3.log | tee gpurun_out/r2h_bench_${2:-cfg2}_$3.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('$1 ${2:-cfg2}', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'], d['roofline']['frac'], s)"; }
run "AC_X=0" cfg2 a
run "AC_X=0" cfg4 a
run "AC_X=0" cfg1 a
run "AC_INSERT_OCC=8" cfg2 occ8
run "AC_INSERT_OCC=5" cfg2 occ5
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2h_launches_cfg2.csv python profiles/profile_build.py cfg2 51 > gpurun_out/r2h_launches_cfg2.log 2>&1; tail -1 gpurun_out/r2h_launches_cfg2.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2h_launches_cfg4.csv python profiles/profile_build.py cfg4 51 > gpurun_out/r2h_launches_cfg4.log 2>&1; tail -1 gpurun_out/r2h_launches_cfg4.log
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'InsertBody|AdjacencyBody|BoundaryBody|SimplifyCoop' -o gpurun_out/r2h_kernels_cfg2 -f python profiles/profile_build.py cfg2 51 > gpurun_out/r2h_ncu_full.log 2>&1; tail -2 gpurun_out/r2h_ncu_full.log
