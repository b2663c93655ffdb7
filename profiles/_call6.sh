# round-2 evidence: launch list + ncu --set full of every kernel family at C4 (1 GPU)
export PE_SETUP_TIMING=
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
# (a) multi-kernel path: the matrix passes and vector kernels exist as separate launches
PE_PCG2=0 PE_PCG=0 timeout 300 $CMD > gpurun_out/r2_c6_plain_mk.log 2>&1 &&
PE_PCG2=0 PE_PCG=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r2_launches_multikernel.csv $CMD > gpurun_out/r2_c6_ncu_a.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_multikernel.csv > gpurun_out/r2_launches_multikernel_summary.txt
PE_PCG2=0 PE_PCG=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_spmv_sell --launch-skip 400 -c 24 -o gpurun_out/r2_spmv_sell $CMD > gpurun_out/r2_c6_ncu_b.log 2>&1
PE_PCG2=0 PE_PCG=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_cg_update|k_cg_direction|k_cheb_first|k_dot|k_jacobi_dot|k_cg_start" --launch-skip 600 -c 12 -o gpurun_out/r2_vector_kernels $CMD > gpurun_out/r2_c6_ncu_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_u_rhs|k_projection_rhs|k_elasticity|k_pressure_residual|k_pressure_matrices|k_residual_t1|k_fill_panels|csr_to_bsr|row_pattern" -c 70 -o gpurun_out/r2_cell_kernels $CMD > gpurun_out/r2_c6_ncu_d.log 2>&1
# (b) default path: launch list + the persistent kernel itself (8th k_pcg2 launch = displacement solve of time step 1)
timeout 300 $CMD > gpurun_out/r2_c6_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_default.csv $CMD > gpurun_out/r2_c6_ncu_e.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_default.csv > gpurun_out/r2_launches_default_summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_pcg2 --launch-skip 7 -c 1 -o gpurun_out/r2_pcg2 $CMD > gpurun_out/r2_c6_ncu_f.log 2>&1
ls -la gpurun_out/*.ncu-rep; head -30 gpurun_out/r2_launches_default_summary.txt; head -30 gpurun_out/r2_launches_multikernel_summary.txt; tail -3 gpurun_out/r2_c6_ncu_f.log
