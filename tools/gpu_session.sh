mkdir -p gpurun_out
L=gpurun_out/r03g.log; : > $L
python -m pytest tests/test_gpu_mle_sumcheck.py tests/test_zkml.py tests/test_gpu_cnn.py tests/test_gpu_matmul.py tests/test_gpu_golden.py tests/test_gpu_basefold.py tests/test_gpu_baseline_size.py -m gpu -x -q 2>&1 | tail -3 >> $L
python tools/sc_rounds.py 10 1 logup 2>&1 | tail -1 >> $L
python tools/sc_rounds.py 12 1 e3 2>&1 | tail -1 >> $L
python tools/sc_rounds.py 20 1 b3 2>&1 | tail -1 >> $L
python tools/throughput_probe.py 1 2>/dev/null | grep workers | cut -c1-110 >> $L
DP_WAIT_MODE=1 python tools/throughput_probe.py 48 48 2>/dev/null | grep workers | cut -c1-220 >> $L
cat $L
