timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r2_call3_suite.log
tail -5 gpurun_out/r2_call3_suite.log
B="timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B --workload c3 > gpurun_out/r2_c3_c3_cheb3.json 2> gpurun_out/r2_c3_err.log; echo "rc=$?" >> gpurun_out/r2_c3_err.log
$B --workload c3 --precond 0 > gpurun_out/r2_c3_c3_jacobi.json 2>> gpurun_out/r2_c3_err.log; echo "rc=$?" >> gpurun_out/r2_c3_err.log
$B > gpurun_out/r2_c3_c4_cheb3.json 2>> gpurun_out/r2_c3_err.log; echo "rc=$?" >> gpurun_out/r2_c3_err.log
$B --precond 0 > gpurun_out/r2_c3_c4_jacobi.json 2>> gpurun_out/r2_c3_err.log; echo "rc=$?" >> gpurun_out/r2_c3_err.log
PE_PCG2=0 $B > gpurun_out/r2_c3_c4_cheb3_multikernel.json 2>> gpurun_out/r2_c3_err.log; echo "rc=$?" >> gpurun_out/r2_c3_err.log
$B --cheb-degree 2 > gpurun_out/r2_c3_c4_cheb2.json 2>> gpurun_out/r2_c3_err.log; echo "rc=$?" >> gpurun_out/r2_c3_err.log
$B --cheb-degree 4 > gpurun_out/r2_c3_c4_cheb4.json 2>> gpurun_out/r2_c3_err.log; echo "rc=$?" >> gpurun_out/r2_c3_err.log
tail -12 gpurun_out/r2_c3_err.log
