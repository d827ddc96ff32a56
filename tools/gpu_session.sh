mkdir -p gpurun_out
L=gpurun_out/r02d_streams.log; : > $L
for cs in 1 2 4; do echo "== DP_COMMIT_STREAMS=$cs" >> $L; DP_COMMIT_STREAMS=$cs python tools/throughput_probe.py 12 16 20 24 32 2>/dev/null | grep workers >> $L; done
echo "== DP_SC_NO_TAIL=1 DP_COMMIT_STREAMS=1" >> $L; DP_SC_NO_TAIL=1 DP_COMMIT_STREAMS=1 python tools/throughput_probe.py 16 24 32 48 2>/dev/null | grep workers >> $L
echo "== DP_SC_NO_TAIL=1 DP_COMMIT_STREAMS=4" >> $L; DP_SC_NO_TAIL=1 python tools/throughput_probe.py 16 24 32 48 2>/dev/null | grep workers >> $L
cat $L
