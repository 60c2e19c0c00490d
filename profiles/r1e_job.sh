# round-1 evidence job (1 GPU): GPU parity suite, bench with both insert bodies, ncu of the insert kernel, launch list, CLI wall
set -x
mkdir -p gpurun_out
nproc > gpurun_out/e_nproc.txt
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; tail -3 gpurun_out/e_pytest.log
AC_HOST_PROFILE=1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_lane.json 2> gpurun_out/e_bench_lane.err; cut -c1-1600 gpurun_out/e_bench_lane.json
AC_INSERT_CHUNKED=1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_chunk.json 2> gpurun_out/e_bench_chunk.err; cut -c1-1600 gpurun_out/e_bench_chunk.json
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:InsertLaneBody -c 1 -o gpurun_out/insert_r1e python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/e_ncu_insert.log 2>&1; tail -2 gpurun_out/e_ncu_insert.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 400 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/e_ncu_launches.log 2>&1; tail -1 gpurun_out/e_ncu_launches.log | cut -c1-300
timeout 400 python profiles/cli_wall.py > gpurun_out/e_cli.log 2>&1; tail -30 gpurun_out/e_cli.log
AC_TABLE_LOAD=0.65 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_load65.json 2> /dev/null; cut -c1-1600 gpurun_out/e_bench_load65.json
AC_HOST_PROFILE=1 timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/e_bench_cfg4_t16.json 2> gpurun_out/e_bench_cfg4_t16.err; cut -c1-1600 gpurun_out/e_bench_cfg4_t16.json
AC_HOST_THREADS=48 AC_HOST_PROFILE=1 timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/e_bench_cfg4_t48.json 2> gpurun_out/e_bench_cfg4_t48.err; cut -c1-1600 gpurun_out/e_bench_cfg4_t48.json
AC_HOST_POOL=0 AC_HOST_PROFILE=1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_nopool.json 2> gpurun_out/e_bench_nopool.err; cut -c1-1600 gpurun_out/e_bench_nopool.json
