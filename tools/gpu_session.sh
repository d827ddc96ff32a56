mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r02z_bench_n4.json 2> gpurun_out/r02z_bench_n4.err
echo "rc=$?"; grep -v "zkml\|^$\|\*\*\*\|OMP" gpurun_out/r02z_bench_n4.err | tail -5; cat /sys/fs/cgroup/cpu.max
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r02z_bench_n4.json") if x.startswith("{")][0])
print({k:d[k] for k in ("value","n_gpus","ms_per_step","e2e")}); print(d["run"]["parallelism"]); print(json.dumps(d["sharded"], indent=0)[:1500]); print({k:(v["value"], v["e2e"]["value"]) for k,v in d["workloads"].items()})
PY
