# 8 GPUs: C4 strong-scaling point after the cooperative halo push, small inner-pass chunks, balanced ownership default, Chebyshev(4) default
nvidia-smi -L | wc -l
export PE_SETUP_TIMING=
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --no-e2e"
PE_PCG_TRACE=gpurun_out/r2_c16_trace_n8 $T --steps 6 --warmup 3 > gpurun_out/r2_c16_n8_c4_cheb4.json 2> gpurun_out/r2_c16_err.log; echo "rc=$?" >> gpurun_out/r2_c16_err.log
$T --steps 6 --warmup 3 --cheb-degree 3 > gpurun_out/r2_c16_n8_c4_cheb3.json 2>> gpurun_out/r2_c16_err.log; echo "rc=$?" >> gpurun_out/r2_c16_err.log
$T --steps 20 --warmup 5 > gpurun_out/r2_c16_n8_c4_cheb4_k20.json 2>> gpurun_out/r2_c16_err.log; echo "rc=$?" >> gpurun_out/r2_c16_err.log
grep -v "^\[W\|Warning\|warn\|^\*\*\*\|OMP_NUM\|pe rank\|^kes)" gpurun_out/r2_c16_err.log | tail -12
python - <<'P'
import json
for f in ('r2_c16_n8_c4_cheb4','r2_c16_n8_c4_cheb3','r2_c16_n8_c4_cheb4_k20'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['ms_per_step'],2), 'parity', d['parity']['parity_max_rel'], 'fp64 pass', r['avg_launch_ms'], r['frac'], 'inner', r['preconditioner_pass'] and (r['preconditioner_pass']['avg_ms'], r['preconditioner_pass']['frac']), 'p', r['pressure_spmv'])
        print('   ', r['phase_ms_per_step'])
        print('   ', d['iterations_per_step']['cg_displacement_per_step'], d['iterations_per_step']['cg_pressure'], 'init', d['init_s'], d['setup_ms'])
    except Exception as e:
        print(f, 'failed', e)
P
python profiles/pcg_trace_report.py gpurun_out/r2_c16_trace_n8_rank0.bin
rm -f gpurun_out/r2_c16_trace_n8_rank[1-6].bin
