# 1 GPU, last call of the round (2.8 GPU-minutes left): the stand-alone FP64 pass over the sliced block-ELL copies of A and J at C4
# (k_spmv_sell, the same stream code as the persistent kernel's CG pass), plain run first, then ncu --set full of 2 + 2 launches
mkdir -p gpurun_out
timeout 50 python profiles/spmv_probe.py 7 20 > gpurun_out/r2_c19_probe.json 2> gpurun_out/r2_c19_err.log; echo "probe rc=$?"
cat gpurun_out/r2_c19_probe.json
timeout 100 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 40 -c 4 -f -o gpurun_out/r2_spmv_sell python profiles/spmv_probe.py 7 20 > gpurun_out/r2_c19_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_c19_ncu.log
ncu -i gpurun_out/r2_spmv_sell.ncu-rep --page raw --csv > gpurun_out/r2_spmv_sell_raw.csv 2>/dev/null
python profiles/ncu_table.py gpurun_out/r2_spmv_sell.ncu-rep > gpurun_out/r2_spmv_sell_table.md 2>/dev/null
cat gpurun_out/r2_spmv_sell_table.md
