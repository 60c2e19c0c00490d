# r2k (1 GPU): the pipelined insert loop with the home group used for the first look only (r2j hung on the earlier form); fail fast
set -x
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2k_smoke.log 2>&1 || { echo SMOKE FAILED; tail -5 gpurun_out/r2k_smoke.log; exit 1; }
tail -1 gpurun_out/r2k_smoke.log
run() { env $1 timeout 120 python bench.py --workload ${2:-cfg2} --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r2k_err_$3.log | tee gpurun_out/r2k_bench_${2:-cfg2}_$3.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('$1 ${2:-cfg2}', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'], d['roofline']['frac'], s)" || { echo BENCH FAILED $1 $2; tail -3 gpurun_out/r2k_err_$3.log; exit 1; }; }
run "AC_X=0" cfg2 a
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2k_pytest.log 2>&1 || { echo PYTEST FAILED; tail -15 gpurun_out/r2k_pytest.log; exit 1; }
tail -3 gpurun_out/r2k_pytest.log
run "AC_INSERT_OCC=3" cfg2 occ3
run "AC_INSERT_OCC=5" cfg2 occ5
run "AC_INSERT_PLAIN=1 AC_INSERT_OCC=6" cfg2 plain6
run "AC_X=0" cfg4 a
run "AC_INSERT_PLAIN=1 AC_INSERT_OCC=6" cfg4 plain6
run "AC_X=0" cfg1 a
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2k_launches_cfg2.csv python profiles/profile_build.py cfg2 51 > gpurun_out/r2k_launches_cfg2.log 2>&1; tail -1 gpurun_out/r2k_launches_cfg2.log
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'ac_insert_kernel' -c 1 -o gpurun_out/r2k_insert_cfg2 -f python profiles/profile_build.py cfg2 51 > gpurun_out/r2k_ncu_full.log 2>&1; tail -2 gpurun_out/r2k_ncu_full.log
