#!/usr/bin/env python3
"""bench.py -- one JSON line per run (contract in the task statement).

A "step" is one pass of the hot path over one batch of synthetic input.  Workloads (--workload):
  dense4m     (default) BASELINE.json configs[1]: full zkml proof of the Dense-4M MLP of SURVEY.md 8(d) Cfg 2 --
              4 x [Dense 1024x1024 + bias -> Requant -> ReLU], BIT_LEN 8, synthetic weights/input -- i.e.
              Prover::prove(trace): witness commits, LogUp-GKR lookups, per-layer sumchecks, table proofs and the
              Basefold batch opening, with host Poseidon2 Fiat-Shamir.  Context::generate (weight commits) is
              setup and is not timed, exactly as zkml/src/bin/bench.rs:390-408 times it.
  sumcheck20  BASELINE.json configs[0] shape: IOPProverState::prove_parallel, nu=20, degree 3, three Base MLEs.
`value`  : proofs/s with everything the proof reads already resident in HBM (model, commitments, tables;
           for sumcheck20 the MLEs), L2 flushed between steps / rotating inputs.
`e2e`    : proofs/s through the host-facing call with HOST buffers: for dense4m inference from the host input
           vector + witness upload + prove + the serialised proof copied back; for sumcheck20 upload + prove.
`--impl reference` times the CPU path (the oracle port -- the reference itself is Rust with un-vendored
dependencies and cannot be built in this image) on the same config.
"""
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before torch/CUDA initialise: 16 proof streams + commit pools need more than 8 hardware queues
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
        # ONE long-running nvidia-smi (-lms 200) for the whole timed region, as in the profiling recipe; rank 0 only (its
        # line is the one printed, and N pollers taking driver locks perturb a launch-heavy workload)
        if int(os.environ.get("RANK", "0")) != 0:
            return self
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
    cpu_returns_seconds = False
    l2_note = "rotating input sets (8 x 24 MiB > L2)"
    flush = False


class BasefoldWorkload:
    """BASELINE.json configs[3] on one GPU: Basefold commit + open of one Base polynomial with 2^24 evaluations
    (splitmix64 mod p), Poseidon2 Merkle, point of 24 E challenges from a fixed seed (SURVEY.md 8d Cfg 4)."""
    NV = 24

    def __init__(self):
        self.name = "Basefold commit+open, 2^%d Base evaluations, RS rate 1/2, Poseidon2 Merkle, 200 queries" % self.NV
        n = 1 << self.NV
        self.evals = splitmix_f(1, n)
        self.point = splitmix_f(4, 2 * self.NV).reshape(self.NV, 2)
        self.h2d = 8 * n
        self.d2h = None
        # commit ~64 n, open ~240 n bytes
        self.alg_bytes = (64 + 240) * n

    def setup_device(self, dp):
        self.dp = dp
        self.mle = dp.Mle.upload(self.evals, False)
        _, flat = dp.pcs_open(self.mle, self.NV, self.point, cap=1 << 23)
        self.d2h = int(flat.size * 8)

    def step_resident(self, i):
        self.dp.pcs_open(self.mle, self.NV, self.point, cap=1 << 23)

    def step_e2e(self, i):
        m = self.dp.Mle.upload(self.evals, False)
        out = self.dp.pcs_open(m, self.NV, self.point, cap=1 << 23)
        m.free()
        return out

    def cpu_step(self, O, i):
        nv = 20      # bounded sample: 2^20 (1/16 of the workload), scaled below
        ev = self.evals[: 1 << nv]
        t0 = time.perf_counter()
        O.pcs_open(ev, False, nv, self.point[:nv], cap=1 << 23)
        return (time.perf_counter() - t0) * (1 << (self.NV - nv))

    cpu_sample = "commit+open of the first 2^20 evaluations (1/16 of the workload), time scaled x16"
    cpu_returns_seconds = True
    l2_note = "256 MiB L2 flush between steps; the 128 MiB input + 256 MiB codeword + 512 MiB oracle exceed L2 anyway"
    flush = True


def splitmix_raw(seed, n, start=0):
    with np.errstate(over="ignore"):
        i = np.arange(start + 1, start + n + 1, dtype=np.uint64)
        z = np.uint64(seed & MASK) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


class DenseWorkload:
    """n_layers x [Dense(width x width) + bias -> Requant -> ReLU]; the tensors are the same splitmix64 streams the
    oracle's synthetic_mlp()/synthetic_input() draw, so both arms prove the identical model and input."""
    NL, W = 4, 1024
    SEED_MODEL, SEED_INPUT = 1, 2

    def __init__(self):
        nl, w = self.NL, self.W
        self.name = "Dense-4M: %d x [Dense %dx%d + bias -> Requant -> ReLU] (%.1fM weights), BIT_LEN 8, full zkml Prover::prove" % (nl, w, w, nl * w * w / 1e6)
        per = w * w + w
        draws = splitmix_raw(self.SEED_MODEL, nl * per).astype(np.int64)
        vals = (draws.view(np.uint64) % np.uint64(255)).astype(np.int64) - 127
        self.weights = np.concatenate([vals[l * per: l * per + w * w] for l in range(nl)])
        self.bias = np.concatenate([vals[l * per + w * w: (l + 1) * per] for l in range(nl)])
        lw = w.bit_length() - 1
        fp = ((lw + 24 + 7) // 8) * 8 - lw
        self.rq = np.array([[lw, fp, 3 << (fp - 2), 2 * 7 + lw + 1]] * nl, dtype=np.int64)
        self.x = (splitmix_raw(self.SEED_INPUT, w) % np.uint64(128)).astype(np.int64)
        # algorithmic bytes of the big streaming parts of one proof: fix_high over the
        # 4 weight matrices (8 B/elt, one pass) + batch_open over 4 x 2^20 codewords/evals
        n = w * w
        self.alg_bytes = nl * 8 * n + (nl * (16 * n + 8 * n) + 240 * n // 4)
        self.h2d = None
        self.d2h = None

    def setup_device(self, dp):
        self.dp = dp
        self.ctx = dp.ZkmlContext(self.NL, self.W, self.weights, self.bias, self.rq)
        proof = self.ctx.prove(self.x)            # warm everything once; also sizes the proof
        self.d2h = int(proof.size * 8)
        nl, w = self.NL, self.W
        ncols = nl * (2 + (int(self.rq[0][0] + self.rq[0][1]) // 8) + 2)
        self.h2d = int(8 * w + ncols * 8 * w + 8 * (256 + 256 + (1 << 15)))
        self.ctx.run_inference(self.x)
        dp.lib().dp_synchronize()

    STREAMS = 16    # concurrent independent proofs per GPU (one host thread + CUDA stream + device arena each)
    units_per_step = STREAMS   # one bench step = one batch of STREAMS inference inputs, each proved independently

    def step_resident(self, i):
        self.ctx.prove_trace()

    def step_e2e(self, i):
        return self.ctx.prove(self.x)

    def run_resident(self, k, device):
        self.ctx.prove_concurrent(host_workers(self.STREAMS), k * self.units_per_step, device=device, e2e=False)

    def run_e2e(self, k, device):
        self.ctx.prove_concurrent(host_workers(self.STREAMS), k * self.units_per_step, device=device, e2e=True)

    def cpu_step(self, O, i):
        _, ms = O.zkml_prove(self.NL, self.W, self.SEED_MODEL, self.SEED_INPUT, want_proof=False)
        return ms[1] * 1e-3     # Prover::prove only; Context::generate (ms[0]) is setup

    cpu_sample = "1 full proof of the same model and input per CPU step, i.e. 1/16 of a GPU step (Context::generate not counted)"
    cpu_returns_seconds = True
    l2_note = ("%d proofs in flight per GPU: aggregate working set (~130 MB of weights/codewords/oracles/trees per proof) "
               "is >> the 126 MB L2; single-stream latency is measured with a 256 MiB L2 flush between proofs" % STREAMS)
    flush = True


class CnnWorkload:
    """CNN-264k (SURVEY.md 8(d) Cfg 3) in its padded form: conv5x5 -> requant -> relu -> maxpool (x2) -> fc x3 on a
    3x32x32 input; the arrays come from deep-prove_b200/models.py and both arms prove the identical model and input."""
    STREAMS = 16
    units_per_step = STREAMS

    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200"))
        import models
        self.desc, self.data, self.x, self.n_params = models.cnn(seed=1)
        self.name = ("CNN-264k: conv5x5(12) -> requant -> relu -> maxpool -> conv5x5(33) -> requant -> relu -> maxpool -> fc 247 -> fc 173 -> fc 10 "
                     "on 3x32x32 (%d parameters, padded to powers of two), BIT_LEN 8, full zkml Prover::prove" % self.n_params)
        # big streaming parts of one proof: fix_high over the fc1 matrix (2^20 Base) + Basefold batch_open over the 2^20-sized commitments
        n = 1 << 20
        self.alg_bytes = 8 * n + (16 * n + 8 * n) + 240 * n
        self.h2d = None
        self.d2h = None

    def setup_device(self, dp):
        self.dp = dp
        self.ctx = dp.ModelContext(self.desc, self.data, self.x.size)
        proof = self.ctx.prove(self.x)
        self.d2h = int(proof.size * 8)
        self.h2d = int(8 * self.x.size + 8 * 600_000)     # input + the witness columns uploaded per proof (requant/relu/pool columns)
        self.ctx.run_inference(self.x)
        dp.lib().dp_synchronize()

    def step_resident(self, i):
        self.ctx.prove_trace()

    def step_e2e(self, i):
        return self.ctx.prove(self.x)

    def run_resident(self, k, device):
        self.ctx.prove_concurrent(host_workers(self.STREAMS), k * self.units_per_step, device=device, e2e=False)

    def run_e2e(self, k, device):
        self.ctx.prove_concurrent(host_workers(self.STREAMS), k * self.units_per_step, device=device, e2e=True)

    def cpu_step(self, O, i):
        _, ms = O.model_prove(self.desc, self.data, self.x, want_proof=False)
        return ms[1] * 1e-3

    cpu_sample = "1 full proof of the same model and input per CPU step, i.e. 1/16 of a GPU step (Context::generate not counted)"
    cpu_returns_seconds = True
    l2_note = ("%d proofs in flight per GPU (aggregate working set >> the 126 MB L2); single-stream latency is measured with a 256 MiB L2 flush between proofs" % STREAMS)
    flush = True


def host_workers(streams):
    """host threads per GPU: every in-flight proof has a thread that hashes and spin-waits, so never oversubscribe the box
    (8 ranks x 16 threads would take all 128 hardware threads of the measurement host and starve everything else)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpus = os.cpu_count() or 16
    if world * (streams + 2) <= cpus:
        return streams
    return max(4, cpus // world - 4)


def pin_to_gpu_numa_node(torch, local_rank):
    """Every round trip of the prover crosses PCIe twice (mapped-memory message, challenge mailbox): keep this rank's
    host threads on the CPUs local to its GPU (sysfs local_cpulist of the GPU's PCI function).  Best effort."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def max_over_ranks(ms, dist, device="cuda"):
    """device-timed milliseconds -> the slowest rank's value (the job is only done when every rank is)"""
    if dist is None:
        return float(ms)
    import torch
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_value(steps_per_rank, world, ms):
    """replicas: every rank proves `steps_per_rank` independent proofs; value = all proofs / slowest rank's time"""
    return steps_per_rank * world / (ms * 1e-3)


def cpu_arm(wl, O, steps, warm):
    """time the CPU path (oracle port, all host threads it can use): returns (proofs/s, seconds per step, cores)"""
    cores = O.lib().dpo_num_threads()
    for i in range(warm):
        wl.cpu_step(O, i)
    tot = 0.0
    for i in range(steps):
        t0 = time.perf_counter()
        r = wl.cpu_step(O, i)
        dt = time.perf_counter() - t0
        tot += r if wl.cpu_returns_seconds else dt
    return steps / tot, tot / steps, cores


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu captures (profiles/)
NCU_TRAFFIC = {
    ("sumcheck20", "k_sc_round"): (7.69e6, "profiles/r01c_sumcheck20_rounds_ncu.csv: rounds 2-11 of one nu=20 proof, mean per launch (algorithmic 6.6 MB)"),
    ("dense4m", "k_sc_round"): (7.1e4, "profiles/r01_ncu_full_summary.csv: one small round (latency-bound)"),
    ("dense4m", "k_merkle_x8"): (None, None),
    ("basefold24", "k_merkle"): (1.87e8, "profiles/r01b_ncu_full_summary.csv: k_merkle_up levels 2-3 of the 2^25-leaf tree, mean per launch (269+105 MB and 134+37 MB; algorithmic 3 x 32 B x hashes = 403 / 201 MB incl. L2-absorbed writes)"),
}

# "Dense 4M proving time 2335 ms" (README.md:18; hardware and exact architecture NOT stated)
PUBLISHED = {"dense4m": 1.0 / 2.335, "cnn264k": 1.0 / 1.242}   # and "CNN 264k ... proving time 1242 ms" (README.md:17)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dense4m", choices=["dense4m", "cnn264k", "sumcheck20", "basefold24"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = {"dense4m": DenseWorkload, "cnn264k": CnnWorkload, "sumcheck20": SumcheckWorkload, "basefold24": BasefoldWorkload}[args.workload]()
    W = max(args.warmup, 0)
    dtype = "u64 (Goldilocks / GoldilocksExt2 modular integers)"

    if args.impl == "reference":
        if rank != 0:
            return
        K = args.steps if args.steps is not None else 2
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_py as O   # the reference arm is one of the two places allowed to execute oracle/
        v, sec, cores = cpu_arm(wl, O, K, min(W, 1))
        print(json.dumps({
            "impl": "reference", "metric": "proofs/sec", "value": v, "unit": "proofs/s", "n_gpus": args.gpus, "steps": K,
            "warmup": min(W, 1), "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": {"workload": wl.name},
            "cpu_baseline": {"value": v, "unit": "proofs/s", "cores": cores, "kind": "port", "sample": wl.cpu_sample},
            "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    K = args.steps if args.steps is not None else {"dense4m": 6, "cnn264k": 6, "sumcheck20": 20, "basefold24": 5}[args.workload]
    import torch
    import dpb200 as dp
    if not torch.cuda.is_available() or dp.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    pin_to_gpu_numa_node(torch, local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dp.init(local_rank)
    dp.use_torch_stream()
    wl.setup_device(dp)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if wl.flush else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        """K steps, each bracketed by CUDA events on the launch stream; the L2 flush sits between steps outside the
        events; barrier + synchronize on both sides; max over ranks."""
        for i in range(warm):
            fn(i)
        barrier()
        l0 = dp.lib().dp_kernel_launches()
        evs = []
        for i in range(steps):
            if flush_buf is not None:
                flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(warm + i)
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = dp.lib().dp_kernel_launches() - l0
        return max_over_ranks(ms, dist), launches

    def timed_many(run, steps, warm):
        """K steps issued as one batch (concurrent proof streams): bracketed by CUDA events recorded on the main stream
        with barrier + synchronize on both sides (the worker streams are drained before the call returns)"""
        if warm:
            run(warm, local_rank)
        barrier()
        l0 = dp.lib().dp_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps, local_rank)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dist), dp.lib().dp_kernel_launches() - l0

    latency_ms = None
    if hasattr(wl, "run_resident"):
        with ClockSampler(local_rank) as clk:
            ms, launches = timed_many(wl.run_resident, K, max(W, 3))
        clocks = clk.summary()
        ms_e2e, _ = timed_many(wl.run_e2e, K, 1)
        lat, _ = timed(wl.step_resident, min(K, 10), 2)     # one proof at a time, L2 flushed between proofs
        latency_ms = lat / min(K, 10)
    else:
        with ClockSampler(local_rank) as clk:
            ms, launches = timed(wl.step_resident, K, max(W, 3))
        clocks = clk.summary()
        ms_e2e, _ = timed(wl.step_e2e, K, 2)

    # roofline leg: CUDA events around every hot kernel launch (dp_profile_*), same steps
    dp.profile_reset(); dp.profile_enable(True)
    for i in range(min(K, 5)):
        if flush_buf is not None:
            flush_buf.zero_()
        wl.step_resident(i)
    torch.cuda.synchronize()
    prof = dp.profile_read()
    dp.profile_enable(False)
    peaks, peak_kind = load_peaks()
    roof = None
    if prof:
        # the resident tail kernel spans many rounds and its event time includes the host's Fiat-Shamir between them:
        # it is reported separately and is not a candidate for the dominant kernel
        tail = {k: v for k, v in prof.items() if k.startswith("k_sc_tail")}
        rest = {k: v for k, v in prof.items() if not k.startswith("k_sc_tail")}
        prof = rest if rest else prof           # a workload made of resident rounds only: fall back to the tail itself
        name = max(prof, key=lambda k: prof[k][1])
        cnt, tot_ms, tot_bytes = prof[name]
        ach = tot_bytes / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
        all_ms = sum(v[1] for v in prof.values())
        traffic = NCU_TRAFFIC.get((args.workload, name.split("(")[0]))
        roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None,
                "resident_tail": {k: {"launches": v[0], "ms_including_host_waits": round(v[1], 4)} for k, v in tail.items()} or None,
                "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "launches": cnt, "avg_us": 1e3 * tot_ms / max(cnt, 1), "alg_bytes_per_launch": tot_bytes / max(cnt, 1),
                "share_of_kernel_time": tot_ms / all_ms if all_ms > 0 else None,
                # context for a latency-bound dominant kernel: the kernels that actually stream, by algorithmic bytes
                "streaming_kernels": {k: {"GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 1), "frac": round(v[2] / (v[1] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4), "ms": round(v[1], 4)}
                                      for k, v in sorted(prof.items(), key=lambda kv: -kv[1][2])[:3] if v[1] > 0},
                "all_kernels": {k: {"launches": v[0], "ms": round(v[1], 4), "GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 2) if v[1] > 0 else 0.0}
                                for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_py as O   # cpu_baseline leg: the oracle is the checker/baseline, never the measured product
        v, sec, cores = cpu_arm(wl, O, 5 if args.workload == "sumcheck20" else 1, 0)
        cpu = {"value": v, "unit": "proofs/s", "cores": cores, "kind": "port",
               "sample": wl.cpu_sample + " (C++ restatement of the reference algorithm, not the Rust reference)"}

    if rank == 0:
        ups = getattr(wl, "units_per_step", 1)
        total = K * ups * world
        v = whole_job_value(K * ups, world, ms)
        pub = PUBLISHED.get(args.workload)
        out = {
            "metric": "proofs/sec" if args.workload != "basefold24" else "commit+open/sec", "value": v, "unit": "proofs/s" if args.workload != "basefold24" else "openings/s",
            "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": (v / pub) if (pub and world == 1) else None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": wl.name, "l2": wl.l2_note,
                       "parallelism": "replicas x%d GPUs (no data-path collective)%s" % (world, (", %d concurrent independent proofs per GPU" % host_workers(wl.STREAMS)) if hasattr(wl, "STREAMS") else ""),
                       "proofs_per_step": ups,
                       "single_stream_latency_ms": latency_ms,
                       "baseline_note": "vs_baseline divides by the reference README's proving time (Dense 4M 2335 ms / CNN 264k 1242 ms; hardware and exact architecture not stated)"},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "proofs/s", "h2d_bytes_per_step": wl.h2d * ups, "d2h_bytes_per_step": wl.d2h * ups},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "alg_GBps_whole_step": wl.alg_bytes * v / world / 1e9,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
