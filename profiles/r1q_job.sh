# last check of the round's final tree on one B200: GPU suite, memcheck of a small build, default bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/q_pytest.log 2>&1; tail -2 gpurun_out/q_pytest.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/q_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|SMOKE" gpurun_out/q_memcheck.log | tail -2
timeout 400 python bench.py > gpurun_out/q_bench_default.json 2> gpurun_out/q_bench_default.err; cut -c1-400 gpurun_out/q_bench_default.json; python -c "
import json; d=json.load(open('gpurun_out/q_bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['stage_ms'])"
