# 4 GPUs: final code: one acquire per CTA and halo, release atomics — 4-rank parity cases + Gmsh + C4 strong-scaling point
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider -k "4 or gmsh" 2>&1 | tail -15 > gpurun_out/r2_call18_multi.log
tail -5 gpurun_out/r2_call18_multi.log
export PE_SETUP_TIMING=
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --no-e2e"
$T --steps 6 --warmup 3 > gpurun_out/r2_c18_n4_c4_cheb3.json 2> gpurun_out/r2_c18_err.log; echo "rc=$?" >> gpurun_out/r2_c18_err.log
grep -v "^\[W\|Warning\|warn\|^\*\*\*\|OMP_NUM\|pe rank\|^kes)" gpurun_out/r2_c18_err.log | tail -8
python - <<'P'
import json
for f in ('r2_c18_n4_c4_cheb3',):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['ms_per_step'],2), 'parity', d['parity']['parity_max_rel'], 'fp64 pass', r['avg_launch_ms'], r['frac'], 'inner', r['preconditioner_pass'] and (r['preconditioner_pass']['avg_ms'], r['preconditioner_pass']['frac']), 'p', r['pressure_spmv'])
        print('   ', r['phase_ms_per_step'])
        print('   ', d['iterations_per_step']['cg_displacement_per_step'], d['iterations_per_step']['cg_pressure'], 'init', d['init_s'], d['setup_ms'])
    except Exception as e:
        print(f, 'failed', e)
P
