# r2g (1 GPU): re-measure of the round-2 tree after the container was replaced: GPU suite, bench cfg2/cfg4/cfg1, launch list, full ncu of the hot kernels
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
run() { env $1 timeout 300 python bench.py --workload ${2:-cfg2} --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r2g_err_$3.log | tee gpurun_out/r2g_bench_${2:-cfg2}_$3.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('$1 ${2:-cfg2}', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'], d['roofline']['frac'], s)"; }
run "AC_X=0" cfg2 a
run "AC_X=0" cfg4 a
run "AC_X=0" cfg1 a
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2g_launches_cfg2.csv python profiles/profile_build.py cfg2 51 > gpurun_out/r2g_launches_cfg2.log 2>&1; tail -1 gpurun_out/r2g_launches_cfg2.log
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'InsertBody|SampleBody|AdjacencyBody|BoundaryBody|BloomBuildBody|ChunkMinBody|EmitSeqBody|ApplyPass|tile_sort|Levels|PackBody|GfaSequence' -o gpurun_out/r2g_kernels_cfg2 -f python profiles/profile_build.py cfg2 51 > gpurun_out/r2g_ncu_full.log 2>&1; tail -2 gpurun_out/r2g_ncu_full.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err; tail -c 600 gpurun_out/r2g_bench_default.json
ls -la gpurun_out/
