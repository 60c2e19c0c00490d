set -x
mkdir -p gpurun_out
lscpu | grep -E "NUMA|Socket|^CPU\(s\)" > gpurun_out/s_lscpu.txt; nvidia-smi topo -m 2>/dev/null | head -12 >> gpurun_out/s_lscpu.txt; cat gpurun_out/s_lscpu.txt | head -14
for mode in bind nobind bind nobind; do
if [ $mode = nobind ]; then export AC_BENCH_NO_NUMA_BIND=1; else unset AC_BENCH_NO_NUMA_BIND; fi
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench_$mode.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/s_bench_$mode.json')); s=d['stage_ms']; print('$mode', 'node', d['config']['numa_node'], d['value'], d['ms_per_step'], d['e2e']['value'], 'd2h', s['d2h'], 'graph', s['host_graph'], 'simplify', s['host_simplify'], 'gfa', s['host_gfa'])"
done
