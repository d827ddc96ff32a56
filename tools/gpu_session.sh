mkdir -p gpurun_out
L=gpurun_out/r03i.log; : > $L
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $L
DP_WAIT_MODE=1 python -m pytest tests/test_zkml.py tests/test_gpu_cnn.py tests/test_gpu_mle_sumcheck.py tests/test_gpu_matmul.py -m gpu -x -q 2>&1 | tail -2 >> $L
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> $L
python bench.py > gpurun_out/r03i_bench_all.json 2> gpurun_out/r03i_bench_all.err; echo "bench rc=$?" >> $L
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r03i_bench_reference.json 2>> $L; echo "ref rc=$?" >> $L
cat $L
python - <<'PY'
import json
for f in ("gpurun_out/r03i_bench_reference.json","gpurun_out/r03i_bench_all.json"):
    d=json.loads([x for x in open(f) if x.startswith("{")][0])
    print(f, d["value"], d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("cores"), d.get("cpu_baseline",{}).get("value"), d.get("parity_checked"))
    for k,v in d.get("workloads",{}).items(): print("   ",k, v["value"], v.get("e2e",{}).get("value"), (v.get("cpu_baseline") or {}).get("value"), v.get("parity_checked"))
    if "run" in d: print("   latency", d["run"]["single_stream_latency_ms"]); r=d["roofline"]; print("   roofline", r["kernel"], r["bound"], r["achieved"], r["peak"], r["frac"], r["traffic"], r.get("whole_step"))
PY
