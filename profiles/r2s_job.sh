# r2s (1 GPU): expand_repeats passes over per-level candidate lists (one candidate per warp at a time) with 1 / 2 / 4 CTAs per SM against the
# old neighbour walk; Bloom build with four entries per thread; GPU suite
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2s_pytest.log 2>&1 || { echo PYTEST FAILED; tail -30 gpurun_out/r2s_pytest.log | cut -c1-400; }
tail -3 gpurun_out/r2s_pytest.log
run() { env $1 timeout 200 python bench.py --workload ${2:-cfg2} --no-cpu-baseline --steps 20 --warmup 3 2>gpurun_out/r2s_err_$3.log | grep '^{' | tee gpurun_out/r2s_bench_${2:-cfg2}_$3.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('$1 ${2:-cfg2}', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], 'adj', s['adjacency'], 'simp', s['device_simplify'])" || { echo BENCH FAILED $1 $2; tail -3 gpurun_out/r2s_err_$3.log; }; }
run "AC_X=0" cfg2 a
run "AC_SIMPLIFY_CTAS_PER_SM=1" cfg2 ctas1
run "AC_SIMPLIFY_CTAS_PER_SM=4" cfg2 ctas4
run "AC_SIMPLIFY_BY_NEIGHBOURS=1" cfg2 neighbours
run "AC_X=0" cfg4 a
run "AC_SIMPLIFY_CTAS_PER_SM=4" cfg4 ctas4
run "AC_SIMPLIFY_BY_NEIGHBOURS=1" cfg4 neighbours
run "AC_X=0" cfg1 a
run "AC_X=0" cfg3 a
