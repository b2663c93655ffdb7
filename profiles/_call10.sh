# round-2 ncu evidence at C4 (1 GPU): --set full of every kernel family, reduced to CSV / a table ON THE BOX (the reports are too big to return)
export PE_SETUP_TIMING=
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
mkdir -p /tmp/rep
# (a) multi-kernel path: matrix passes and CG vector kernels as separate launches (a window in the middle of the solves)
PE_PCG2=0 PE_PCG=0 timeout 300 $CMD > gpurun_out/r2_c10_plain_mk.log 2>&1 &&
PE_PCG2=0 PE_PCG=0 timeout 900 $NCU -k regex:"k_spmv_sell|k_cg_update|k_cg_direction|k_cheb_first|k_dot|k_jacobi_dot|k_cg_start" --launch-skip 500 -c 100 -o /tmp/rep/r2_solver_kernels $CMD > gpurun_out/r2_c10_ncu_a.log 2>&1
# (b) default path: cell / setup kernels, then the persistent kernel (u solve of time step 1 and the solves after it)
timeout 300 $CMD > gpurun_out/r2_c10_plain.log 2>&1 &&
timeout 900 $NCU -k regex:"k_u_rhs|k_projection_rhs|k_elasticity|k_pressure_residual|k_pressure_matrices|k_residual_t1|k_fill_panels|csr_to_bsr|row_pattern|k_neumann|k_axpy|k_axpby|k_distribute|k_invdiag" -c 90 -o /tmp/rep/r2_cell_kernels $CMD > gpurun_out/r2_c10_ncu_b.log 2>&1
timeout 1200 $NCU -k regex:k_pcg2 --launch-skip 7 -c 3 -o /tmp/rep/r2_pcg2 $CMD > gpurun_out/r2_c10_ncu_c.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/rep/r2_launches_default.csv $CMD > gpurun_out/r2_c10_ncu_d.log 2>&1
python profiles/summarize_launches.py /tmp/rep/r2_launches_default.csv > gpurun_out/r2_launches_default_summary.txt
ls -la /tmp/rep
python profiles/ncu_table.py /tmp/rep/r2_solver_kernels.ncu-rep /tmp/rep/r2_cell_kernels.ncu-rep /tmp/rep/r2_pcg2.ncu-rep > gpurun_out/r2_ncu_table.md 2> gpurun_out/r2_ncu_table.err
for r in r2_solver_kernels r2_cell_kernels r2_pcg2; do
  ncu -i /tmp/rep/$r.ncu-rep --page raw --csv 2>/dev/null | python profiles/slim_raw_csv.py > gpurun_out/${r}_raw_slim.csv
done
ncu -i /tmp/rep/r2_pcg2.ncu-rep --page details --csv 2>/dev/null > gpurun_out/r2_pcg2_details.csv
# hottest source lines of the persistent kernel's first captured launch (SASS-level sampling folded to source lines)
ncu -i /tmp/rep/r2_pcg2.ncu-rep --page source --csv --print-source cuda 2>/dev/null | head -c 3000000 > gpurun_out/r2_pcg2_source.csv
[ $(stat -c %s /tmp/rep/r2_pcg2.ncu-rep) -lt 30000000 ] && cp /tmp/rep/r2_pcg2.ncu-rep gpurun_out/
cat gpurun_out/r2_ncu_table.md; head -12 gpurun_out/r2_launches_default_summary.txt; tail -3 gpurun_out/r2_c10_ncu_c.log; du -sh gpurun_out
