python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2_call1_suite.log
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r2_c1_jacobi.json 2> gpurun_out/r2_c1_err.log
$B --precond 1 --cheb-degree 3 > gpurun_out/r2_c1_cheb3.json 2>> gpurun_out/r2_c1_err.log
PE_CHEB_FP32=1 $B --precond 1 --cheb-degree 3 > gpurun_out/r2_c1_cheb3_fp32.json 2>> gpurun_out/r2_c1_err.log
PE_CHEB_FP32=1 $B --precond 1 --cheb-degree 2 > gpurun_out/r2_c1_cheb2_fp32.json 2>> gpurun_out/r2_c1_err.log
PE_CHEB_FP32=1 $B --precond 1 --cheb-degree 4 > gpurun_out/r2_c1_cheb4_fp32.json 2>> gpurun_out/r2_c1_err.log
PE_CHEB_FP32=1 $B --precond 1 --cheb-degree 3 --eig-ratio 15 > gpurun_out/r2_c1_cheb3_fp32_r15.json 2>> gpurun_out/r2_c1_err.log
PE_PCG_MAX_NNZ=999999999999 $B > gpurun_out/r2_c1_jacobi_pcg.json 2>> gpurun_out/r2_c1_err.log
tail -3 gpurun_out/r2_call1_suite.log
