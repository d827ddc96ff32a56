mkdir -p gpurun_out
L=gpurun_out/r03b.log; : > $L
python -m pytest tests/test_zkml.py tests/test_gpu_mle_sumcheck.py tests/test_gpu_basefold.py tests/test_gpu_cnn.py -m gpu -x -q 2>&1 | tail -2 >> $L
DP_HOST_PROF=1 DP_WAIT_MODE=1 python tools/throughput_probe.py 48 2> gpurun_out/r03b_hostprof.err | grep workers >> $L
grep "worker0" gpurun_out/r03b_hostprof.err | tail -3 >> $L; grep "hostprof.*dp_sc_\(create\|destroy\)" gpurun_out/r03b_hostprof.err >> $L
DP_WAIT_MODE=1 python tools/throughput_probe.py 32 48 64 2>/dev/null | grep workers >> $L
DP_WAIT_SPINNERS=0 DP_WAIT_MODE=1 python tools/throughput_probe.py 48 64 2>/dev/null | grep workers >> $L
cat $L
