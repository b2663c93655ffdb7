# 8 GPUs: C4 strong scaling point (default + variants) + C5 weak scaling point
nvidia-smi -L | wc -l
export PE_SETUP_TIMING=1
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8"
$T --steps 6 --warmup 3 > gpurun_out/r2_c8_n8_c4_cheb3.json 2> gpurun_out/r2_c8_err.log; echo "rc=$?" >> gpurun_out/r2_c8_err.log
$T --steps 6 --warmup 3 --cheb-degree 4 > gpurun_out/r2_c8_n8_c4_cheb4.json 2>> gpurun_out/r2_c8_err.log; echo "rc=$?" >> gpurun_out/r2_c8_err.log
$T --steps 20 --warmup 5 > gpurun_out/r2_c8_n8_c4_cheb3_k20.json 2>> gpurun_out/r2_c8_err.log; echo "rc=$?" >> gpurun_out/r2_c8_err.log
$T --steps 3 --warmup 3 --workload c5 > gpurun_out/r2_c8_n8_c5.json 2>> gpurun_out/r2_c8_err.log; echo "rc=$?" >> gpurun_out/r2_c8_err.log
grep -v "^\[W\|Warning\|warn\|^\*\*\*\|OMP_NUM" gpurun_out/r2_c8_err.log | tail -40
