# r2e: GPU suite on the fixed fetch order + k up to 511; L2 persistence / table load sweep for the insert
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2e_pytest.log 2>&1; tail -3 gpurun_out/r2e_pytest.log
run() { env $1 timeout 300 python bench.py --workload ${2:-cfg2} --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r2e_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('$1 ${2:-cfg2}', d['value'], d['ms_per_step'], d['parity']['ok'], 'insert', s['insert'], 'adj', s['adjacency'], 'links', s['links'], 'simp', s['device_simplify'], 'gfa', s['device_gfa'])"; }
run "AC_X=0"
run "AC_L2_PERSIST=1 AC_HOST_PROFILE=1"; grep "persisting" gpurun_out/r2e_err.log | head -1
run "AC_TABLE_LOAD=0.6"
run "AC_TABLE_LOAD=0.7"
run "AC_L2_PERSIST=1 AC_TABLE_LOAD=0.6"
run "AC_L2_PERSIST=1 AC_TABLE_LOAD=0.7"
run "AC_L2_PERSIST=1 AC_TABLE_LOAD=0.8"
run "AC_X=0" cfg4
run "AC_L2_PERSIST=1" cfg4
run "AC_L2_PERSIST=1 AC_TABLE_LOAD=0.7" cfg4
