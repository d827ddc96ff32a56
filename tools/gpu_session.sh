mkdir -p gpurun_out
L=gpurun_out/r03k.log; : > $L
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $L
python bench.py --only --steps 2 > gpurun_out/r03k_bench_dense_short.json 2> gpurun_out/r03k_bench.err; echo "bench rc=$?" >> $L
python - >> $L <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r03k_bench_dense_short.json") if x.startswith("{")][0])
print({k:d.get(k) for k in ("value","e2e","parity_checked","gpu_launches")}); print(d["run"])
PY
cat $L
