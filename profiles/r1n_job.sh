# closing evidence with the round's final code (1 GPU)
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/n_smoke.log 2>&1; tail -1 gpurun_out/n_smoke.log
AC_HOST_PROFILE=1 timeout 400 python bench.py > gpurun_out/n_bench_default.json 2> gpurun_out/n_bench_default.err; cut -c1-2600 gpurun_out/n_bench_default.json; grep "host\] adopt" gpurun_out/n_bench_default.err | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 700 --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/n_ncu_launches.log 2>&1; tail -1 gpurun_out/n_ncu_launches.log | cut -c1-200
timeout 200 python profiles/cli_wall.py > gpurun_out/n_cli.log 2>&1; grep -E "^rep|load\+repair" gpurun_out/n_cli.log
