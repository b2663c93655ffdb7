timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/r2_call4_suite.log
tail -3 gpurun_out/r2_call4_suite.log
B="timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B --workload c3 > gpurun_out/r2_c4_c3_cheb3.json 2> gpurun_out/r2_c4_err.log; echo "rc=$?" >> gpurun_out/r2_c4_err.log
$B --workload c3 --precond 0 > gpurun_out/r2_c4_c3_jacobi.json 2>> gpurun_out/r2_c4_err.log; echo "rc=$?" >> gpurun_out/r2_c4_err.log
$B > gpurun_out/r2_c4_c4_cheb3.json 2>> gpurun_out/r2_c4_err.log; echo "rc=$?" >> gpurun_out/r2_c4_err.log
$B --precond 0 > gpurun_out/r2_c4_c4_jacobi.json 2>> gpurun_out/r2_c4_err.log; echo "rc=$?" >> gpurun_out/r2_c4_err.log
$B --cheb-degree 4 > gpurun_out/r2_c4_c4_cheb4.json 2>> gpurun_out/r2_c4_err.log; echo "rc=$?" >> gpurun_out/r2_c4_err.log
timeout 200 python profiles/spmv_probe.py 7 200 > gpurun_out/r2_c4_probe_sell.json 2>> gpurun_out/r2_c4_err.log
tail -12 gpurun_out/r2_c4_err.log
