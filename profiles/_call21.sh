# the round's last 24 GPU-seconds: the reference-side binding (integration/_build/fss_gpu = GpuBackend.h + the reference's run() on pe_*,
# linked with libporoel.so) on the shipped input.data, four time steps; fields come back in gpurun_out/c21/solution/
# (as run, the parameter file was a copy on disk of the record's "input" entry; it is taken from the record here)
mkdir -p gpurun_out/c21/solution && cd gpurun_out/c21 && python -c "import json; open('input.data','w').write(json.load(open('../../tests/golden/reference_run_shipped_4steps.json'))['input'])"
timeout 14 ../../integration/_build/fss_gpu input.data 1 20000 > out.log 2> err.log; echo "rc=$?"
tail -4 out.log; tail -2 err.log
