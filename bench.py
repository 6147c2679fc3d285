#!/usr/bin/env python
"""Benchmark of the hot path: correspondence-sets/sec through PointDSC.forward (testing mode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one testing-mode forward over one batch of B synthetic correspondence sets (default: the
configuration BASELINE.json's metric is quoted on, N=1000, B=256, 3DMatch snapshot).  Batches shard
across ranks with no data-path collective (weak scaling: B sets per GPU); NCCL only reduces the timing.

value   : sets/s with inputs resident in HBM (CUDA events around exactly K steps, max over ranks)
e2e     : sets/s through the reference-facing module call with pinned HOST tensors — the H2D copy of the
          step's inputs and the D2H copy of (final_trans, final_labels) are inside the timed region
roofline: the dominant kernel (per-layer SC-weighted attention) — algorithmic FLOPs per launch / its mean
          launch duration measured live with CUDA events on the launch stream (pdsc_profile_*)
cpu_baseline: the CPU oracle (a torch-CPU restatement of the reference, "port") on the box's host cores,
          bounded sample, rank 0 at N=1 only.  `--impl reference` times that same CPU path as its own arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "correspondence-sets/sec (PointDSC.forward, N=1000, B=256)"   # BASELINE.json's metric (the default configuration)


def metric_of(args):
    """The metric label of this run: BASELINE.json's string at the default configuration, the same label with the actual
    N and B otherwise (parity-test sized runs must not carry the headline label)."""
    return METRIC if (args.n == 1000 and args.batch == 256) else f"correspondence-sets/sec (PointDSC.forward, N={args.n}, B={args.batch})"
UNIT = "sets/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--dataset", default="3dmatch", choices=["3dmatch", "kitti"])
    ap.add_argument("--precision", default=os.environ.get("POINTDSC_PRECISION", "fp16x3"))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config_of(args, world):
    return {"workload": f"{args.dataset}-like synthetic correspondences, N={args.n}, B={args.batch} sets per GPU per step, "
                        f"k=40, S={int(args.n * 0.1)} seeds, 12 SCNonlocal layers, released {args.dataset} snapshot",
            "n": args.n, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "precision": args.precision, "parallelism": f"dp{world} (sets sharded, no data-path collective)",
            "l2": "per-step working set exceeds L2 (SC matrix alone is 4*N*NS*B bytes = "
                  f"{4 * args.n * ((args.n + 63) // 64 * 64) * args.batch / 1e6:.0f} MB vs 126 MB L2); no flush needed"}


def load_snapshot(dataset):
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, "tests", "golden", f"snapshot_{dataset}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def make_inputs(args, rank, world=1):
    """This rank's shard of the global batch (B sets per GPU, weak scaling): global set g has seed g and an inlier
    ratio cycling through 0.5 / 0.3 / 0.2 / 0.4, so any sharding of the same global batch sees the same sets."""
    import torch
    from pointdsc_b200.shard import shard_bounds
    from pointdsc_b200.synth import make_pair
    ratios = [0.5, 0.3, 0.2, 0.4]
    lo, hi = shard_bounds(args.batch * world, rank, world)
    pairs = [make_pair(g, args.n, args.dataset, ratios[g % 4]) for g in range(lo, hi)]
    return {k: torch.stack([p[k] for p in pairs], 0).contiguous() for k in pairs[0]}


class ClockSampler:
    """nvidia-smi clocks/throttle-reason sampler running during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 <= t <= t1 + 0.1:
                    sm.append(float(f[0]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def pick_cpu_threads(args, sets):
    """torch's CPU ops on small tensors slow down badly when oversubscribed (128 threads: 26 s per N=1000 forward,
    8 threads: 0.17 s), so "all the host threads it can use" is found by timing one forward per candidate count
    (bounded: a candidate that takes > 4x the best so far ends the search) and keeping the fastest."""
    import torch
    from oracle import pointdsc_oracle as O
    sd = load_snapshot(args.dataset)
    cfg = O.default_config(args.dataset)
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        O.forward_testing(sd, cfg, sets["corr_pos"][0], sets["src_keypts"][0], sets["tgt_keypts"][0])  # warm this pool size
        t0 = time.perf_counter()
        O.forward_testing(sd, cfg, sets["corr_pos"][0], sets["src_keypts"][0], sets["tgt_keypts"][0])
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
        elif dt > 4 * best_t:
            break
    return best


def cpu_oracle_rate(args, sets, budget_s, threads):
    """sets/s of the CPU oracle on a bounded sample of the same workload."""
    import torch
    from oracle import pointdsc_oracle as O
    torch.set_num_threads(threads)
    sd = load_snapshot(args.dataset)
    cfg = O.default_config(args.dataset)
    O.forward_testing(sd, cfg, sets["corr_pos"][0], sets["src_keypts"][0], sets["tgt_keypts"][0])  # warm-up
    done, t0 = 0, time.perf_counter()
    while done < sets["corr_pos"].shape[0] and (time.perf_counter() - t0 < budget_s or done < 2):
        O.forward_testing(sd, cfg, sets["corr_pos"][done], sets["src_keypts"][done], sets["tgt_keypts"][done])
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path, as restated by the oracle (the reference is Python and
    cannot travel to the GPU box), all host threads, loop of bs=1 testing forwards."""
    if rank != 0:
        return
    import torch
    from oracle import pointdsc_oracle as O
    sd = load_snapshot(args.dataset)
    cfg = O.default_config(args.dataset)
    probe = make_inputs(argparse.Namespace(**{**vars(args), "batch": 4}), 0)
    threads = pick_cpu_threads(args, probe)
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    O.forward_testing(sd, cfg, probe["corr_pos"][1], probe["src_keypts"][1], probe["tgt_keypts"][1])
    one = max(time.perf_counter() - t0, 1e-3)
    # a step = a bounded sample of the batch, sized so that steps + warmup stay within ~2 minutes
    per_step = int(max(1, min(args.batch, 120.0 / (one * (args.steps + args.warmup)))))
    need = per_step * (args.steps + args.warmup)
    small = argparse.Namespace(**{**vars(args), "batch": min(args.batch, need)})
    sets = make_inputs(small, 0)
    nsets = sets["corr_pos"].shape[0]

    def step(i):
        for j in range(per_step):
            b = (i * per_step + j) % nsets
            O.forward_testing(sd, cfg, sets["corr_pos"][b], sets["src_keypts"][b], sets["tgt_keypts"][b])
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = (f"{per_step} sets per step (loop of bs=1 testing forwards) x {args.steps} steps of the N={args.n} workload, "
              f"{threads} host threads (fastest of the candidate counts on {os.cpu_count()} cores)")
    print(json.dumps({
        "impl": "reference", "metric": metric_of(args), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_engine(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pointdsc_b200 import PointDSC
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfgd = {"3dmatch": dict(inlier_threshold=0.10, sigma_d=0.10, nms_radius=0.10),
            "kitti": dict(inlier_threshold=0.6, sigma_d=1.2, nms_radius=0.6)}[args.dataset]
    model = PointDSC(in_dim=6, num_layers=12, num_channels=128, num_iterations=10, ratio=0.1, k=40,
                     precision=args.precision, **cfgd)
    res = model.load_state_dict(load_snapshot(args.dataset), strict=False)
    assert res.missing_keys == [], res
    model = model.to(dev).eval()
    host = make_inputs(args, rank, world)
    pinned = {k: host[k].pin_memory() for k in ("corr_pos", "src_keypts", "tgt_keypts")}
    d = {k: host[k].to(dev) for k in ("corr_pos", "src_keypts", "tgt_keypts")}
    B, N = args.batch, args.n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from pointdsc_b200.shard import gather_counters, output_checksum
    from pointdsc_b200.shard import max_over_ranks as _max_over_ranks

    def max_over_ranks(x):
        return _max_over_ranks(x, dev)

    # ---- device-resident throughput ----------------------------------------------------------------------
    for _ in range(args.warmup):
        out = model.run(d["corr_pos"], d["src_keypts"], d["tgt_keypts"])
    model.profile(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out = model.run(d["corr_pos"], d["src_keypts"], d["tgt_keypts"])
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = max_over_ranks(e0.elapsed_time(e1))
    prof = model.profile_read()
    model.profile(False)
    clocks = sampler.stop(t0, t1)
    value = B * world * args.steps / (ms * 1e-3)

    # sanity: the timed work produced registrations (not a skipped / cached forward)
    err = (out["final_trans"].cpu() - host["gt_trans"]).abs().amax(dim=(1, 2))
    scale = 0.05 if args.dataset == "3dmatch" else 0.5
    registered = float((err < scale).float().mean())

    # ---- end to end: pinned host tensors in, host tensors out, copies inside the timed region -------------
    for _ in range(min(2, args.warmup)):
        model.run(pinned["corr_pos"], pinned["src_keypts"], pinned["tgt_keypts"])
    barrier()
    th0 = time.perf_counter()
    for _ in range(args.steps):
        ho = model.run(pinned["corr_pos"], pinned["src_keypts"], pinned["tgt_keypts"])
    torch.cuda.synchronize()
    th1 = time.perf_counter()
    e2e_s = max_over_ranks(th1 - th0)
    e2e_value = B * world * args.steps / e2e_s
    h2d = sum(pinned[k].numel() * 4 for k in pinned)
    d2h = ho["final_trans"].numel() * 4 + ho["final_labels"].numel() * 4

    counters = output_checksum(out["final_trans"].cpu(), out["final_labels"].cpu())
    counters["registered"] = registered
    per_rank = gather_counters(counters)     # NCCL is used for timing / counters only: there is no data-path collective
    if rank != 0:
        return
    registered = sum(c["registered"] * c["sets"] for c in per_rank) / max(1.0, sum(c["sets"] for c in per_rank))
    # ---- roofline of the dominant kernel --------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "1400 TFLOP/s sustained (of fallback)"
    attn_ms, attn_launches = prof["attention"]
    flops_per_launch = 4.0 * 128 * N * N * B          # QK^T + PV of one layer over the rank's B sets (algorithmic)
    achieved = flops_per_launch / (attn_ms / max(attn_launches, 1) * 1e-3) / 1e12 if attn_ms > 0 else None
    ns = (N + 63) // 64 * 64
    hbm_bytes = 4.0 * N * ns * B + 4.0 * 128 * N * B * 3 + 4.0 * 128 * N * B   # SC tiles + Q,K,V images (hi+lo) + msg
    hbm_gbs = hbm_bytes / (attn_ms / max(attn_launches, 1) * 1e-3) / 1e9 if attn_ms > 0 else None
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "attention_traffic.json")))
        key = f"{args.precision}_N{N}_B{B}"
        traffic = tj.get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    executed = {"bf16x3": 3, "fp16x3": 3, "bf16": 1, "fp32": 1}[args.precision]
    roofline = {"kernel": "tc_attention_persistent_kernel" if args.precision != "fp32" else "attention_simt_kernel",
                "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_flops_per_launch": flops_per_launch, "launch_ms": attn_ms / max(attn_launches, 1),
                "launches_timed": attn_launches,
                # the same launch seen from the memory side: SC (re-read by every layer) + Q/K/V images + msg
                "hbm_algorithmic_bytes_per_launch": hbm_bytes, "hbm_achieved_gbs": hbm_gbs,
                "hbm_frac_of_copy_peak": (hbm_gbs / float(peaks.get("hbm_gbs", 6568.4))) if hbm_gbs else None,
                "note": f"algorithmic FLOPs = 4*C*N^2*B per layer; {args.precision} executes {executed}x that on the tensor pipe"}
    total_ms = prof["total"][0]
    stages = {k: {"ms_per_step": v[0] / args.steps, "share": (v[0] / total_ms if total_ms > 0 else None)}
              for k, v in prof.items() if k != "total"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = pick_cpu_threads(args, host)
        rate, done, dt = cpu_oracle_rate(args, host, args.cpu_seconds, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {done} sets of the step's batch, loop of bs=1 testing forwards, {dt:.1f} s on {threads} host threads "
                         f"(fastest of the candidate thread counts on {os.cpu_count()} cores; torch {torch.__version__} CPU fp32)"}
    launches = model.launches_per_forward(B, N) * args.steps
    print(json.dumps({
        "metric": metric_of(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16 hi/lo split (3 products) with f32 accumulate", "bf16": "bf16 with f32 accumulate",
                  "fp16x3": "fp16 hi/lo split (3 products) with f32 accumulate", "fp32": "f32"}[args.precision],
        "data": "synthetic", "config": config_of(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "stages": stages,
        "registered_fraction": registered,
        "rank_checksums": [{k: c[k] for k in ("sets", "trans_abs", "inliers")} for c in per_rank]}))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import __graft_entry__ as g
    g.build()   # no-op unless the sources changed since the in-tree library was built
    run_engine(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
