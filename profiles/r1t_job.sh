set -x
mkdir -p gpurun_out
timeout 200 python profiles/cli_wall.py > gpurun_out/t_cli.log 2>&1; grep -E "^rep|load\+repair|NUMA" gpurun_out/t_cli.log
timeout 300 python bench.py > gpurun_out/t_bench_default.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/t_bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['config']['numa_node'], d['clocks']['samples'], d['cpu_baseline']['value'], d['stage_ms'])"
