mkdir -p gpurun_out
L=gpurun_out/r02p.log; : > $L
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
echo "== block" >> $L; DP_WAIT_MODE=1 python tools/throughput_probe.py 32 48 64 2>/dev/null | grep workers >> $L
DP_WAIT_MODE=1 python tools/trace_api.py 48 288 2>/dev/null | grep -v zkml >> $L
python bench.py --only > gpurun_out/r02p_bench_dense.json 2> gpurun_out/r02p_bench_dense.err; tail -3 gpurun_out/r02p_bench_dense.err >> $L
python - >> $L <<'PY'
import json
d=json.load(open("gpurun_out/r02p_bench_dense.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","parity_checked","clocks")}); print(d["run"]); print(d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["bound"])
print(d["cpu_baseline"]["value"])
PY
cat $L
