mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r03j_bench_n2.json 2> gpurun_out/r03j_bench_n2.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r03j_bench_n2.json") if x.startswith("{")][0])
print({k:d[k] for k in ("value","n_gpus","ms_per_step","e2e")}); print(d["run"]["parallelism"]); s=d["sharded"]; print({k:{q:v.get(q) for q in ("sharded_ms","single_gpu_ms","speedup","bit_identical_to_single_gpu_proof","error")} for k,v in s.items()}); print({k:(v["value"], v["e2e"]["value"]) for k,v in d["workloads"].items()})
PY
