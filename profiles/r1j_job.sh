# round-1 closing evidence on one B200: smoke, GPU parity suite, the default bench line, the reference arm, ncu of the insert kernel, launch list
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/j_smoke.log 2>&1; tail -2 gpurun_out/j_smoke.log
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/j_pytest.log 2>&1; tail -2 gpurun_out/j_pytest.log
AC_HOST_PROFILE=1 timeout 400 python bench.py > gpurun_out/j_bench_default.json 2> gpurun_out/j_bench_default.err; cut -c1-2500 gpurun_out/j_bench_default.json; grep "host\] adopt" gpurun_out/j_bench_default.err | tail -1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/j_bench_reference.json 2> gpurun_out/j_bench_reference.err; cut -c1-1200 gpurun_out/j_bench_reference.json
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:InsertLaneBody -c 1 -o gpurun_out/insert_r1j python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_insert.log 2>&1; tail -1 gpurun_out/j_ncu_insert.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 600 --csv --log-file gpurun_out/j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_launches.log 2>&1; tail -1 gpurun_out/j_ncu_launches.log | cut -c1-200
timeout 300 python profiles/cli_wall.py > gpurun_out/j_cli.log 2>&1; grep -E "^rep|load\+repair" gpurun_out/j_cli.log
timeout 120 python profiles/pcie_probe.py > gpurun_out/j_pcie.log 2>&1; cat gpurun_out/j_pcie.log
