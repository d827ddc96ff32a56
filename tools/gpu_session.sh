mkdir -p gpurun_out
DP_WAIT_MODE=1 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 3000 --csv --log-file gpurun_out/r03c_dense4m_inst_ncu.csv python tools/ncu_dense.py 1 > /dev/null 2>&1
echo done
