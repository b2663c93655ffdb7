# 4 GPUs: multi-rank parity tests (2 and 4 ranks) + C4 strong scaling point + C5 weak scaling point
nvidia-smi -L | wc -l
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_call7_multi.log
tail -5 gpurun_out/r2_call7_multi.log
export PE_SETUP_TIMING=1
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4"
$T --steps 6 --warmup 3 > gpurun_out/r2_c7_n4_c4_cheb3.json 2> gpurun_out/r2_c7_err.log; echo "rc=$?" >> gpurun_out/r2_c7_err.log
$T --steps 3 --warmup 3 --workload c5 > gpurun_out/r2_c7_n4_c5.json 2>> gpurun_out/r2_c7_err.log; echo "rc=$?" >> gpurun_out/r2_c7_err.log
grep -v "^\[W\|Warning\|warn\|^\*\*\*\|OMP_NUM" gpurun_out/r2_c7_err.log | tail -20
