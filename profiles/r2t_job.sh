# r2t (1 GPU): L2 persistence for the 2-bit sequence store (or the table) while the table is probed, on the inputs whose table exceeds the L2
set -x
mkdir -p gpurun_out
run() { env $1 timeout 200 python bench.py --workload ${2:-cfg2} --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r2t_err_$3.log | grep '^{' | tee gpurun_out/r2t_bench_${2:-cfg2}_$3.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('$1 ${2:-cfg2}', d['value'], d['ms_per_step'], d['parity']['ok'], 'ins', s['insert_kernel'], 'adj', s['adjacency'], 'bnd', s['boundaries'], 'links', s['links'])" || { echo BENCH FAILED $1 $2; tail -3 gpurun_out/r2t_err_$3.log; }; }
run "AC_X=0" cfg4 a
run "AC_L2_PERSIST=packed" cfg4 packed
run "AC_L2_PERSIST=table" cfg4 table
run "AC_X=0" cfg3 a
run "AC_L2_PERSIST=packed" cfg3 packed
run "AC_L2_PERSIST=packed" cfg2 packed
