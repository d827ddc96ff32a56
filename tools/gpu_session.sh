mkdir -p gpurun_out
python bench.py > gpurun_out/r02t_bench_all.json 2> gpurun_out/r02t_bench_all.err; echo "bench rc=$?"
python tools/sc_rounds.py 20 1 b3 > gpurun_out/r02t_sc_rounds_nu20.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_sc_ -c 80 --csv --log-file gpurun_out/r02t_sumcheck20_rounds_ncu.csv python tools/sc_rounds.py 20 0 b3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02t_dense4m_launches_ncu.csv python tools/ncu_dense.py 1 > /dev/null 2>&1
tail -3 gpurun_out/r02t_bench_all.err | grep -v zkml; cat gpurun_out/r02t_sc_rounds_nu20.log | tail -30
