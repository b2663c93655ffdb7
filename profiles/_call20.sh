# 1 GPU, the round's last 90 GPU-seconds: per-kernel metrics (the columns of ncu_table.py, explicit --metrics instead of --set full
# to keep the replay count low) of the cell-loop and CG vector kernels — C3 first (safe), C4 if the clock allows
mkdir -p gpurun_out
export PE_PCG2=0 PE_PCG=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed
K='regex:k_elasticity|k_u_rhs|k_projection_rhs|k_pressure_residual|k_residual_t1|k_cg_update|k_cheb_first|k_cg_direction'
timeout 30 python profiles/kernel_probe.py 6 > gpurun_out/r2_c20_probe_c3.txt 2> gpurun_out/r2_c20_err.log; echo "probe rc=$? t=$SECONDS"
cat gpurun_out/r2_c20_probe_c3.txt
timeout 40 ncu --metrics $M --clock-control none -k "$K" -c 40 -f -o gpurun_out/r2_kernels_c3 python profiles/kernel_probe.py 6 > gpurun_out/r2_c20_ncu_c3.log 2>&1; echo "ncu c3 rc=$? t=$SECONDS"
ncu -i gpurun_out/r2_kernels_c3.ncu-rep --page raw --csv > gpurun_out/r2_kernels_c3_raw.csv 2>/dev/null
python profiles/ncu_table.py gpurun_out/r2_kernels_c3.ncu-rep > gpurun_out/r2_kernels_c3_table.md 2>/dev/null
cat gpurun_out/r2_kernels_c3_table.md; echo "t=$SECONDS"
if [ $SECONDS -lt 34 ]; then
  timeout $((76 - SECONDS)) ncu --metrics $M --clock-control none -k "$K" -c 40 -f -o gpurun_out/r2_kernels_c4 python profiles/kernel_probe.py 7 > gpurun_out/r2_c20_ncu_c4.log 2>&1; echo "ncu c4 rc=$? t=$SECONDS"
  ncu -i gpurun_out/r2_kernels_c4.ncu-rep --page raw --csv > gpurun_out/r2_kernels_c4_raw.csv 2>/dev/null
  python profiles/ncu_table.py gpurun_out/r2_kernels_c4.ncu-rep > gpurun_out/r2_kernels_c4_table.md 2>/dev/null
  cat gpurun_out/r2_kernels_c4_table.md
fi
