# 1 GPU: suite + C4 / C3 lines with the pressure-solve phase clock + per-warp trace of the last passes of a C3 displacement solve
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2_call14_suite.log
tail -3 gpurun_out/r2_call14_suite.log
B="timeout 600 python bench.py --no-cpu-baseline --no-e2e"
$B --steps 6 --warmup 3 > gpurun_out/r2_c14_c4_cheb3.json 2> gpurun_out/r2_c14_err.log; echo "rc=$?" >> gpurun_out/r2_c14_err.log
PE_PCG_TRACE=gpurun_out/r2_c14_trace_c3 $B --steps 6 --warmup 3 --workload c3 > gpurun_out/r2_c14_c3_cheb3.json 2>> gpurun_out/r2_c14_err.log; echo "rc=$?" >> gpurun_out/r2_c14_err.log
PE_CHUNK_KB=64 $B --steps 6 --warmup 3 --workload c3 > gpurun_out/r2_c14_c3_chunk64.json 2>> gpurun_out/r2_c14_err.log; echo "rc=$?" >> gpurun_out/r2_c14_err.log
PE_CHUNK_KB=16 $B --steps 6 --warmup 3 --workload c3 > gpurun_out/r2_c14_c3_chunk16.json 2>> gpurun_out/r2_c14_err.log; echo "rc=$?" >> gpurun_out/r2_c14_err.log
grep -v "^\[W\|Warning\|warn\|^\*\*\*\|OMP_NUM" gpurun_out/r2_c14_err.log | tail -12
python - <<'P'
import json
for f in ('r2_c14_c4_cheb3','r2_c14_c3_cheb3','r2_c14_c3_chunk64','r2_c14_c3_chunk16'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['ms_per_step'],2), 'parity', d['parity']['parity_max_rel'], 'fp64 pass', r['avg_launch_ms'], r['frac'], 'inner', r['preconditioner_pass'] and (r['preconditioner_pass']['avg_ms'], r['preconditioner_pass']['frac']), 'p', r['pressure_spmv'])
        print('   ', r['phase_ms_per_step'])
    except Exception as e:
        print(f, 'failed', e)
P
