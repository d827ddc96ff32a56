#!/usr/bin/env python3
"""bench.py -- one JSON line per run (contract in the task statement).

A "step" is one pass of the hot path over one batch of synthetic input.  Workloads:
  sumcheck20  BASELINE.json configs[0] shape: IOPProverState::prove_parallel, nu=20, degree 3, three Base
              MLEs (splitmix64 seeds per SURVEY.md 8d), 20 rounds with host Poseidon2 Fiat-Shamir.
`value`  : proofs/s with the MLEs already resident in HBM (rotating input sets larger than L2).
`e2e`    : proofs/s through the host-facing API with HOST buffers: upload (H2D) + prove + proof (D2H).
`--impl reference` times the CPU path (the oracle port -- the reference itself is Rust with un-vendored
dependencies and cannot be built in this image) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200"))

P = np.uint64(0xFFFFFFFF00000001)
MASK = (1 << 64) - 1


def splitmix_f(seed, n):
    """n splitmix64 draws mod p (vectorised; identical stream to oracle/field.hpp SplitMix64)."""
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & MASK) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return z % P


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self.proc = None

    def __enter__(self):
        # ONE long-running nvidia-smi (-lms 200) for the whole timed region, as in the profiling recipe
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.strip().splitlines():
            self.samples.append([x.strip() for x in line.split(",")])

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[1])); mx.append(float(s[2]))
                for nm, v in zip(names, s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
class SumcheckWorkload:
    name = "sumcheck prove_parallel nu=20 deg=3 3xBase (BASELINE configs[0] shape), host Poseidon2 FS"
    NV = 20
    NSETS = 8   # 8 x 24 MiB = 192 MiB of inputs > 126 MB L2: a set is evicted before it is reused

    def __init__(self):
        self.products = [((1, 0), [0, 1, 2])]
        n = 1 << self.NV
        self.host_sets = [[splitmix_f(3 * k + j + 1, n) for j in range(3)] for k in range(self.NSETS)]
        self.h2d = 3 * n * 8
        self.d2h = self.NV * 4 * 16 + 3 * 16 + self.NV * 16
        # algorithmic bytes of one proof (48 n per Base MLE) and field ops per proof
        self.alg_bytes = 3 * 48 * n
        # K1 per pair at degree 3: 8 mul + 19 add; fold adds 3 mul + 6 add per pair of the NEXT round
        self.field_ops = sum((1 << (self.NV - 1 - r)) * (8 + 19 + (9 if r > 0 else 0)) for r in range(self.NV))

    def setup_device(self, dp):
        self.dp = dp
        self.dev_sets = [[dp.Mle.upload(a, False) for a in s] for s in self.host_sets]
        dp.lib().dp_synchronize()

    def step_resident(self, i):
        return self.dp.sumcheck_prove_parallel(self.dev_sets[i % self.NSETS], self.products, self.NV)

    def step_e2e(self, i):
        ms = [self.dp.Mle.upload(a, False) for a in self.host_sets[i % self.NSETS]]
        out = self.dp.sumcheck_prove_parallel(ms, self.products, self.NV)
        for m in ms:
            m.free()
        return out

    def cpu_step(self, O, i):
        mles = [(a, False) for a in self.host_sets[i % self.NSETS]]
        return O.sumcheck_prove(mles, self.products, self.NV)

    cpu_sample = "1 full proof (same workload) per step"


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sumcheck20")
    args = ap.parse_args()
    K, W = args.steps, max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = SumcheckWorkload()

    if args.impl == "reference":
        if rank != 0:
            return
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_py as O   # bench.py's reference/cpu_baseline leg is one of the places allowed to run oracle/
        for i in range(min(W, 1)):
            wl.cpu_step(O, i)
        t0 = time.perf_counter()
        for i in range(K):
            wl.cpu_step(O, i)
        dt = time.perf_counter() - t0
        v = K / dt
        print(json.dumps({
            "impl": "reference", "metric": "proofs/sec", "value": v, "unit": "proofs/s", "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 (Goldilocks / GoldilocksExt2 modular integers)", "data": "synthetic",
            "config": {"workload": wl.name},
            "cpu_baseline": {"value": v, "unit": "proofs/s", "cores": 1, "kind": "port", "sample": wl.cpu_sample},
            "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import dpb200 as dp
    if not torch.cuda.is_available() or dp.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dp.init(local_rank)
    dp.use_torch_stream()
    wl.setup_device(dp)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for i in range(warm):
            fn(i)
        barrier()
        l0 = dp.lib().dp_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warm + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = dp.lib().dp_kernel_launches() - l0
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    with ClockSampler(local_rank) as clk:
        ms, launches = timed(wl.step_resident, K, max(W, 3))
    clocks = clk.summary()
    ms_e2e, _ = timed(wl.step_e2e, K, 2)

    # roofline leg: per-kernel CUDA-event timing of the dominant kernel over the same steps
    dp.profile_reset(); dp.profile_enable(True)
    for i in range(K):
        wl.step_resident(i)
    torch.cuda.synchronize()
    prof = dp.profile_read()
    dp.profile_enable(False)
    peaks, peak_kind = load_peaks()
    roof = None
    if prof:
        name = max(prof, key=lambda k: prof[k][1])
        cnt, tot_ms, tot_bytes = prof[name]
        ach = tot_bytes / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
        roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_kind": peak_kind,
                "launches": cnt, "avg_us": 1e3 * tot_ms / max(cnt, 1), "alg_bytes_per_launch": tot_bytes / max(cnt, 1),
                "all_kernels": {k: {"launches": v[0], "ms": v[1], "GBps": (v[2] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0)}
                                for k, v in prof.items()}}

    cpu = None
    if rank == 0 and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_py as O   # cpu_baseline leg: the oracle is the checker/baseline, never the measured product
        t0 = time.perf_counter()
        n = 0
        while n < 2 or time.perf_counter() - t0 < 10.0:
            wl.cpu_step(O, n); n += 1
            if n >= 50:
                break
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "proofs/s", "cores": 1, "kind": "port", "sample": "%d proofs of the same workload" % n}

    if rank == 0:
        total = K * world
        v = total / (ms * 1e-3)
        out = {
            "metric": "proofs/sec", "value": v, "unit": "proofs/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 (Goldilocks / GoldilocksExt2 modular integers)", "data": "synthetic",
            "config": {"workload": wl.name, "l2": "rotating input sets (%d x %.0f MiB > L2)" % (wl.NSETS, wl.h2d / 2**20),
                       "parallelism": "replicas x%d (one independent proof stream per GPU, no data-path collective)" % world},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "proofs/s", "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "field_ops_per_s": wl.field_ops * v,
            "alg_GBps_whole_step": wl.alg_bytes * v / 1e9,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
