mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:'k_merkle_l1|k_merkle_up' -c 60 --csv --log-file gpurun_out/r03f_dense_merkle_traffic.csv python tools/ncu_dense.py 1 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03f_smoke.log 2>&1; tail -2 gpurun_out/r03f_smoke.log
