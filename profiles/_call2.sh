timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_call2_suite.log
tail -3 gpurun_out/r2_call2_suite.log
timeout 200 python profiles/spmv_probe.py 7 200 > gpurun_out/r2_c2_probe_sell.json 2> gpurun_out/r2_c2_probe.err
PE_FORMAT=bsr timeout 200 python profiles/spmv_probe.py 7 200 > gpurun_out/r2_c2_probe_bsr.json 2>> gpurun_out/r2_c2_probe.err
timeout 200 python profiles/spmv_probe.py 6 400 > gpurun_out/r2_c2_probe_sell_r6.json 2>> gpurun_out/r2_c2_probe.err
cat gpurun_out/r2_c2_probe_sell.json gpurun_out/r2_c2_probe_bsr.json
B="timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r2_c2_jacobi.json 2> gpurun_out/r2_c2_err.log
$B --precond 1 --cheb-degree 3 > gpurun_out/r2_c2_cheb3_fp32.json 2>> gpurun_out/r2_c2_err.log
$B --precond 1 --cheb-degree 2 > gpurun_out/r2_c2_cheb2_fp32.json 2>> gpurun_out/r2_c2_err.log
PE_CHEB_FP32=0 $B --precond 1 --cheb-degree 3 > gpurun_out/r2_c2_cheb3_fp64.json 2>> gpurun_out/r2_c2_err.log
tail -5 gpurun_out/r2_c2_err.log
