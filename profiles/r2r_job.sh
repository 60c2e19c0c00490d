# r2r (1 GPU): GPU suite; A/B of the level skipping, the cooperative grid size and the table load; bench cfg2 (default line with cpu_baseline and cli_wall), cfg4, cfg1;
# launch list and `ncu --set full` of every kernel >= 3 % of the device time on cfg2, insert + adjacency on cfg4 at k = 31 / 51 / 91
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2r_pytest.log 2>&1 || { echo PYTEST FAILED; tail -30 gpurun_out/r2r_pytest.log | cut -c1-400; }
tail -3 gpurun_out/r2r_pytest.log
run() { env $1 timeout 200 python bench.py --workload ${2:-cfg2} --no-cpu-baseline --steps 20 --warmup 3 2>gpurun_out/r2r_err_$3.log | grep '^{' | tee gpurun_out/r2r_bench_${2:-cfg2}_$3.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('$1 ${2:-cfg2}', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['gpu_launches'], d['roofline']['frac'], s)" || { echo BENCH FAILED $1 $2; tail -3 gpurun_out/r2r_err_$3.log; }; }
run "AC_X=0" cfg2 a
run "AC_SIMPLIFY_ALL_LEVELS=1" cfg2 all_levels
run "AC_COOP_CTAS=64" cfg2 coop64
run "AC_COOP_CTAS=32" cfg2 coop32
run "AC_TABLE_LOAD=0.4" cfg2 load40
run "AC_TABLE_LOAD=0.6" cfg2 load60
run "AC_X=0" cfg2 b
run "AC_X=0" cfg4 a
run "AC_SIMPLIFY_ALL_LEVELS=1" cfg4 all_levels
run "AC_X=0" cfg1 a
run "AC_X=0" cfg3 a
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2r_launches_cfg2.csv python profiles/profile_build.py cfg2 51 > gpurun_out/r2r_launches_cfg2.log 2>&1; tail -1 gpurun_out/r2r_launches_cfg2.log
timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'InsertBody|SampleBody|AdjacencyBody|BloomBuildBody|BoundaryBody|ClaimedListBody|ChunkMinBody|EmitSeqBody|SimplifyCoopBody|LevelsCoopBody|tile_sort|GfaSequenceBody|GfaLinkBody|MergeWaysBody' -c 24 -o gpurun_out/r2r_kernels_cfg2 -f python profiles/profile_build.py cfg2 51 > gpurun_out/r2r_ncu_full.log 2>&1; tail -2 gpurun_out/r2r_ncu_full.log
python profiles/kernel_table.py gpurun_out/r2r_kernels_cfg2.ncu-rep --json gpurun_out/r2r_kernels_cfg2.json > gpurun_out/r2r_kernels_cfg2.md 2>&1; head -30 gpurun_out/r2r_kernels_cfg2.md
for k in 31 51 91; do
timeout 300 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled -k regex:'InsertBody|AdjacencyBody|BoundaryBody|SimplifyCoopBody' -c 4 -o gpurun_out/r2r_kernels_cfg4_k$k -f python profiles/profile_build.py cfg4 $k > gpurun_out/r2r_ncu_cfg4_k$k.log 2>&1; tail -1 gpurun_out/r2r_ncu_cfg4_k$k.log
python profiles/kernel_table.py gpurun_out/r2r_kernels_cfg4_k$k.ncu-rep --json gpurun_out/r2r_kernels_cfg4_k$k.json > gpurun_out/r2r_kernels_cfg4_k$k.md 2>&1; cat gpurun_out/r2r_kernels_cfg4_k$k.md
rm -f gpurun_out/r2r_kernels_cfg4_k$k.ncu-rep
done
python bench.py --steps 20 --warmup 3 > gpurun_out/r2r_bench_default.json 2> gpurun_out/r2r_bench_default.err; tail -c 1500 gpurun_out/r2r_bench_default.json
ls -la gpurun_out/ | head -60
