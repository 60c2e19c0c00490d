set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r_pytest.log 2>&1; tail -3 gpurun_out/r_pytest.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -x -k "reference_fixed_seqs or loaded_graphs" > gpurun_out/r_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r_memcheck.log | tail -2
