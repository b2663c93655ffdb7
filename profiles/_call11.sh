# gather-pipelined stream: GPU suite + C4 / C3 bench lines (1 GPU)
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2_call11_suite.log
tail -4 gpurun_out/r2_call11_suite.log
B="timeout 600 python bench.py --no-cpu-baseline"
$B --steps 6 --warmup 3 > gpurun_out/r2_c11_c4_cheb3.json 2> gpurun_out/r2_c11_err.log; echo "rc=$?" >> gpurun_out/r2_c11_err.log
$B --steps 6 --warmup 3 --workload c3 > gpurun_out/r2_c11_c3_cheb3.json 2>> gpurun_out/r2_c11_err.log; echo "rc=$?" >> gpurun_out/r2_c11_err.log
grep -v "^\[W\|Warning\|warn\|^\*\*\*\|OMP_NUM" gpurun_out/r2_c11_err.log | tail -12
python - <<'P'
import json
for f in ('r2_c11_c4_cheb3','r2_c11_c3_cheb3'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['ms_per_step'],2), 'parity', d['parity']['parity_max_rel'], 'fp64 pass', r['avg_launch_ms'], r['frac'], 'inner', r['preconditioner_pass'] and (r['preconditioner_pass']['avg_ms'], r['preconditioner_pass']['frac']), 'p', r['pressure_spmv'])
        print('   ', r['phase_ms_per_step'])
        print('   ', d['iterations_per_step']['cg_displacement_per_step'], d['iterations_per_step']['cg_pressure'])
    except Exception as e:
        print(f, 'failed', e)
P
