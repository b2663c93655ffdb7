nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_call5_multi.log
tail -5 gpurun_out/r2_call5_multi.log
export PE_SETUP_TIMING=1
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3"
$T > gpurun_out/r2_c5_n2_c4_cheb3.json 2> gpurun_out/r2_c5_err.log; echo "rc=$?" >> gpurun_out/r2_c5_err.log
$T --precond 0 > gpurun_out/r2_c5_n2_c4_jacobi.json 2>> gpurun_out/r2_c5_err.log; echo "rc=$?" >> gpurun_out/r2_c5_err.log
$T --workload c3 > gpurun_out/r2_c5_n2_c3_cheb3.json 2>> gpurun_out/r2_c5_err.log; echo "rc=$?" >> gpurun_out/r2_c5_err.log
B="timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/r2_c5_n1_c4_cheb3.json 2>> gpurun_out/r2_c5_err.log; echo "rc=$?" >> gpurun_out/r2_c5_err.log
$B --workload c3 > gpurun_out/r2_c5_n1_c3_cheb3.json 2>> gpurun_out/r2_c5_err.log; echo "rc=$?" >> gpurun_out/r2_c5_err.log
grep -v "^\[W\|Warning\|warn" gpurun_out/r2_c5_err.log | tail -30
