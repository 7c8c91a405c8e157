#!/usr/bin/env python
"""bench.py -- loop-candidate keyframes/sec through desc -> search -> PnP (BASELINE.json metric).

A step = one batch of B synthetic keyframes per GPU through the hot path:
  NetVLAD forward (480x640x3 u8 -> 8192-D)  ->  top-5 inner-product search of every new descriptor
  against a 100k x 8192 fp32 descriptor DB (sharded round-robin over the ranks, per-shard top-k
  all-gathered over NCCL and merged)  ->  DLS-PnP RANSAC (reference parameters: <= 50 hypotheses of
  15 points, 200 correspondences, 20 % outliers) for every keyframe treated as a loop candidate.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference] [--workload default|config2|pnp5]

--workload default is the driver's line (BASELINE config 4 sizes on one or more GPUs).  `config2` is BASELINE config 2 (480x640x1
-> 4096-D gray model, 10k-keyframe DB, one B200) through the same pipeline; `pnp5` is BASELINE config 5 (1024 candidates x
4096 hypotheses x 200 correspondences, adaptive termination off) on the verifier alone.

`value` = keyframes/s with inputs resident in HBM; `e2e` = the same through the C-ABI host calls
(pinned host images in, host results out, every step).  `--impl reference` times the CPU restatement
of the reference path (oracle/) on the box's host cores.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is ONE process that may use every host core,
    # so lift the cap before numpy / torch load their BLAS and OpenMP runtimes (and again at run time, below)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "loop_candidate_keyframes_per_sec"
UNIT = "keyframes/s"
ROWS, COLS, CHNLS = 480, 640, 3
DB_ROWS, DIM = 100_000, 8192
N_CORR, HYP = 200, 50
GOLD_W = os.path.join(ROOT, "tests", "golden", "keras_raw_mobilenet_conv7.npz")
GOLD_W_GRAY = os.path.join(ROOT, "tests", "golden", "keras_raw_gray_conv6.npz")
# counted-flop model of one DLS-PnP hypothesis: quartic cost + gradient coefficients from 15 points 0.03,
# block-triangular elimination 0.20, Hessenberg reduction 10/3 n^3 = 0.07, Francis QR (~2 double-shift sweeps of 10 n^2 per
# eigenvalue) 0.39, inverse iteration / back-substitution of the real roots 0.05, scoring 30 flop x 200 points 0.006
PNP_MFLOP_PER_HYP = 0.75


def load_net(workload="default"):
    """Shipped default model's weights if the converted fixture is present, else random-init
    weights of the same architecture (there is no network for checkpoints)."""
    from cerebro_b200 import keras_weights

    if workload == "config2":
        z = np.load(GOLD_W_GRAY)
        raw = {k.replace("__", "/"): z[k] for k in z.files}
        return keras_weights.fold_mobilenet_netvlad(raw), raw, "Apr2019/gray_conv6_K16__centeredinput (shipped weights)"
    if os.path.exists(GOLD_W):
        z = np.load(GOLD_W)
        raw = {k.replace("__", "/"): z[k] for k in z.files}
        return keras_weights.fold_mobilenet_netvlad(raw), raw, "mobilenet_conv7_allpairloss (shipped weights)"
    return keras_weights.random_mobilenet_netvlad(CHNLS, 7, 16, seed=0), None, "mobilenet_conv7 architecture, random init"


def workload_config(batch, world):
    return {
        "workload": "desc(480x640x%d->%d-D NetVLAD) + top-5 search over %dk x %d fp32 DB + DLS-PnP RANSAC "
        "(200 corr, <=50 hyp) per keyframe" % (CHNLS, DIM, DB_ROWS // 1000, DIM),
        "keyframes_per_step_per_gpu": batch,
        "db_rows": DB_ROWS,
        "descriptor_dim": DIM,
        "db_sharding": "round-robin over %d rank(s), NCCL all-gather of per-shard top-5" % world,
        "pnp": {"correspondences": N_CORR, "max_hypotheses": HYP, "outlier_frac": 0.2, "adaptive": True},
        "l2": "DB (%.2f GB / ranks) and activations (%d keyframes x ~66 MB) exceed the 126 MB L2; no explicit flush" % (DB_ROWS * DIM * 4 / 1e9, batch),
    }


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle/), timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_pipeline_setup(raw, seed=0):
    import torch

    from cerebro_b200 import synthetic
    from oracle import search as osearch

    cores = os.cpu_count() or 1
    torch.set_num_threads(min(cores, 32))  # batch-1 convolutions do not scale past a few dozen threads
    try:  # numpy's BLAS (the fp32 search GEMV) may have been capped by an inherited OMP_NUM_THREADS
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=cores)
    except Exception:
        pass
    g = torch.Generator().manual_seed(seed)
    db = torch.randn((DB_ROWS, DIM), generator=g, dtype=torch.float32)
    db /= db.norm(dim=1, keepdim=True)
    index = osearch.IndexFlatIP(DIM)
    index.x = db.numpy()
    rng = np.random.default_rng(seed)
    return dict(raw=raw, index=index, rng=rng, cores=cores, synthetic=synthetic)


def cpu_pipeline_run(ctx, n_keyframes):
    """desc (torch-CPU fp32, all cores) -> fp32 BLAS search top-5 -> oracle RANSAC (1 thread, as the
    reference's loopcandidate_consumer_th).  Returns seconds."""
    from oracle import dls_pnp, netvlad

    imgs = ctx["synthetic"].band_limited_images(n_keyframes, ROWS, COLS, CHNLS, seed=11)
    cands = [ctx["synthetic"].loop_candidate(ctx["rng"], n=N_CORR) for _ in range(n_keyframes)]
    tabs = [dls_pnp.sample_table(7, c, HYP, N_CORR) for c in range(n_keyframes)]
    t0 = time.perf_counter()
    for i in range(n_keyframes):
        d = netvlad.describe(imgs[i : i + 1], ctx["raw"], dtype="float32")
        ctx["index"].search(d, 5, accumulate="float32")
        dls_pnp.ransac_pnp(cands[i][0], cands[i][1], tabs[i])
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    net, raw, model_name = load_net()
    if raw is None:
        print(json.dumps({"impl": "reference", "unavailable": "raw Keras weight fixture tests/golden/keras_raw_mobilenet_conv7.npz missing"}))
        return
    ctx = cpu_pipeline_setup(raw)
    kf = 2  # keyframes per step: ~1 s of CPU work
    for _ in range(args.warmup):
        cpu_pipeline_run(ctx, 1)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_pipeline_run(ctx, kf)
    v = kf * args.steps / t
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": v,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": t / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32 descriptor / f32 search / f64 pnp (CPU)",
        "data": "synthetic",
        "config": workload_config(args.batch, 1),  # the GPU arm's workload; each CPU step is a bounded sample of it (below)
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ctx["cores"], "kind": "port",
                         "sample": "%d keyframes per step x %d steps (of the %d-keyframe step), torch-CPU NetVLAD (all cores) + fp32 BLAS search of the full 100k x 8192 DB + oracle DLS-PnP RANSAC (1 thread)" % (kf, args.steps, args.batch)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.p = None
        self.device = device
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = samples in the upper half (the first/last samples may be idle)
        sm_sorted = sorted(sm)
        med = sm_sorted[len(sm_sorted) // 2] if sm_sorted else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def descriptor_algorithmic(net, rows, cols, chnls):
    """Algorithmic work of one frame's forward pass on the fused path: dense FLOPs (stem + pointwise + VLAD head) and the
    bytes that MUST cross HBM when every depthwise output stays on chip (image in, every block's input and output once,
    the final feature map read twice by the VLAD head, descriptor out)."""
    h, w = (rows + 1 - 3) // 2 + 1, (cols + 1 - 3) // 2 + 1
    flops = 2.0 * h * w * 9 * chnls * 32
    byts = rows * cols * chnls + h * w * 32 * 2  # image in, stem out
    c = 32
    for blk in net["blocks"]:
        byts += h * w * c * 2  # block input, read once
        if blk["stride"] == 2:
            h, w = (h + 1 - 3) // 2 + 1, (w + 1 - 3) // 2 + 1
        flops += 2.0 * h * w * 9 * c  # depthwise (CUDA cores; counted for completeness)
        if blk["pw_w"] is not None:
            cout = blk["pw_w"].shape[1]
            flops += 2.0 * h * w * c * cout
            c = cout
        byts += h * w * c * 2  # block output, written once
    d = c
    flops += 2.0 * h * w * d * 16 * 2  # soft-assignment + aggregation
    byts += 2 * h * w * d * 2 + 16 * d * 4
    return flops, float(byts)


def measure_h2d_gbs(dev, nbytes=256 << 20, reps=4):
    """Host -> device bandwidth from pinned memory, all ranks at the same time when called between barriers."""
    import torch

    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def run_pnp5(args, dev, local_rank):
    """BASELINE config 5: 1024 candidates x 4096 hypotheses x 200 correspondences, adaptive termination off."""
    import torch

    from cerebro_b200 import synthetic
    from cerebro_b200.pnp import PnpBatch, default_params
    from oracle import dls_pnp as D

    n_cand, H = 1024, 4096
    rng = np.random.default_rng(3)
    cands = [synthetic.loop_candidate(rng, n=N_CORR) for _ in range(n_cand)]
    offsets = torch.tensor(np.arange(n_cand + 1, dtype=np.int32) * N_CORR, device=dev)
    X = torch.tensor(np.concatenate([c[0] for c in cands]), device=dev)
    uv = torch.tensor(np.concatenate([c[1] for c in cands]), device=dev)
    pb = PnpBatch(max_candidates=n_cand, max_points_total=n_cand * N_CORR, max_hypotheses=H, device=local_rank)
    prm = default_params(seed=1, max_iterations=H, adaptive=0)
    out = pb.solve_device(offsets, X, uv, prm)
    torch.cuda.synchronize()
    # parity spot-check: 8 candidates against the oracle RANSAC on the same counter-based sample table
    T = out["T"].cpu().numpy()
    bh = out["best_hyp"].cpu().numpy()
    ni = out["n_inliers"].cpu().numpy()
    p2 = D.RansacParameters(adaptive=False, max_iterations=H)
    checked = 0
    t0 = time.perf_counter()
    for c in (0, 1, 2, 3, 511, 512, 1022, 1023):
        r = D.ransac_pnp(cands[c][0], cands[c][1], D.sample_table(1, c, H, N_CORR), p2)
        e = D.pose_error(T[c], r["T"])
        assert int(bh[c]) == r["best_hyp"] and int(ni[c]) == r["n_inliers"], (c, int(bh[c]), r["best_hyp"], int(ni[c]), r["n_inliers"])
        assert e[0] < 1e-3 and e[1] < 1e-2, (c, e)
        checked += 1
    cpu_s = time.perf_counter() - t0
    steps = max(1, min(args.steps, 3))
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        pb.solve_device(offsets, X, uv, prm, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    clocks = sampler.stop()
    hyps = n_cand * H
    props = torch.cuda.get_device_properties(dev)
    fp64_peak = props.multi_processor_count * 64 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
    tf = hyps * PNP_MFLOP_PER_HYP * 1e6 / (ms * 1e-3) / 1e12
    print(json.dumps({
        "metric": "dls_pnp_ransac_hypotheses_per_sec", "value": hyps / ms * 1e3, "unit": "hypotheses/s", "n_gpus": 1, "steps": steps,
        "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 5: 1024 candidates x 4096 hypotheses x 200 correspondences, 15-point DLS-PnP samples, "
                   "sigma 1e-3 noise, 20 % outliers, error_thresh 0.03, MLE cost, adaptive termination off"},
        "candidates_per_s": n_cand / ms * 1e3,
        "roofline": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak, "traffic": None,
                     "flop_model_mflop_per_hypothesis": PNP_MFLOP_PER_HYP,
                     "peak_source": "SMs x 64 FP64 FMA lanes x 2 x max SM clock (no measured FP64 peak in MEASURED_PEAKS.json)"},
        "parity": {"candidates_checked_against_oracle": checked, "tolerance": "identical best hypothesis + inlier count; pose < 1e-3 rad / 1e-2 m"},
        "cpu_baseline": {"value": 8 * H / cpu_s, "unit": "hypotheses/s", "cores": 1, "kind": "port",
                         "sample": "8 candidates x 4096 hypotheses, oracle/dls_pnp.py (numpy, 1 thread)"},
        "clocks": clocks, "gpu_launches": 6 * ((hyps + 16383) // 16384) + 1,
    }))
    pb.close()


def main():
    global ROWS, COLS, CHNLS, DB_ROWS, DIM
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="keyframes per step per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="default", choices=["default", "config2", "pnp5"])
    ap.add_argument("--cpu-baseline-keyframes", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.workload == "config2":
        CHNLS, DB_ROWS, DIM = 1, 10_000, 4096

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG=VERSION makes NCCL printf its version banner to stdout (NCCL_DEBUG_FILE does not move it); keep stdout
        # for the one JSON line.  Any other level (INFO, WARN ...) is the caller's choice and stays, routed to stderr.
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        # every rank drives its GPU from its own slice of the host cores (descriptor / search / verifier threads + the copy
        # engines' pinned buffers stay put instead of migrating across all cores of the box)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local_rank * per : (local_rank + 1) * per]))
        except Exception:
            pass

    from cerebro_b200 import build as cb_build
    from cerebro_b200 import synthetic
    from cerebro_b200.loop_detector import LoopPipeline

    cb_build.build()  # no-op when the in-tree .so is up to date
    if args.workload == "pnp5":
        if rank == 0:
            run_pnp5(args, dev, local_rank)
        if world > 1:
            dist.destroy_process_group()
        return
    net, raw, model_name = load_net(args.workload)
    B = args.batch
    rows_local = (DB_ROWS + world - 1) // world
    pipe = LoopPipeline(net, ROWS, COLS, CHNLS, B, rows_local + 256, device=local_rank, sharded=(world > 1), n_corr=N_CORR, hypotheses=HYP)

    # ---- synthetic DB shard, generated on the device (rows are unit-norm N(0,1) vectors)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_mine = len(range(rank, DB_ROWS, world))
    chunk = 12_500
    local_index = pipe.index.local if world > 1 else pipe.index
    first_rows = None
    for a in range(0, n_mine, chunk):
        nrow = min(chunk, n_mine - a)
        x = torch.randn((nrow, DIM), generator=g, device=dev, dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True)
        if first_rows is None:
            first_rows = x[:B].clone()  # local rows 0..B-1 of this shard = global labels j * world + rank
        local_index.add_local(x)  # every rank generates and bulk-loads only its own shard
        del x
    torch.cuda.synchronize()

    # ---- parity inside the warm-up (BASELINE configs 3 / 4): every rank queries noisy copies of rows that live on the NEXT
    # rank's shard; the merged global top-1 must be exactly the planted label, and the score the fp64 dot product.
    gq = torch.Generator(device=dev).manual_seed(99 + rank)
    if world > 1:
        all_first = torch.empty((world,) + tuple(first_rows.shape), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(all_first, first_rows)
        owner = (rank + 1) % world
        target_rows = all_first[owner]
    else:
        owner, target_rows = 0, first_rows
    noise = torch.randn(target_rows.shape, generator=gq, device=dev, dtype=torch.float32)
    noise /= noise.norm(dim=1, keepdim=True)
    planted = 0.9 * target_rows + (1.0 - 0.81) ** 0.5 * noise
    planted /= planted.norm(dim=1, keepdim=True)
    planted = planted.contiguous()
    if world > 1:
        ps, pl = pipe.index.search_sharded_device(planted, 5)
    else:
        ps, pl = pipe.index.search_device(planted, 5)
    torch.cuda.synchronize()
    want = torch.arange(B, device=dev, dtype=torch.int64) * world + owner
    exact = (planted.double() * target_rows.double()).sum(1)
    if not torch.equal(pl[:, 0], want) or not torch.allclose(ps[:, 0], exact, rtol=0, atol=1e-9):
        raise SystemExit("bench.py: sharded search parity FAILED on rank %d (planted neighbours on rank %d's shard)" % (rank, owner))
    parity_note = "%d planted neighbours per rank on the next rank's shard: top-1 labels identical, fp64 scores within 1e-9" % B

    # ---- per-step inputs
    imgs_host = torch.from_numpy(synthetic.band_limited_images(B, ROWS, COLS, CHNLS, seed=100 + rank)).pin_memory()
    rng = np.random.default_rng(5 + rank)
    cands = [synthetic.loop_candidate(rng, n=N_CORR) for _ in range(B)]
    X_host = torch.from_numpy(np.concatenate([c[0] for c in cands])).pin_memory()
    uv_host = torch.from_numpy(np.concatenate([c[1] for c in cands])).pin_memory()
    offsets_np = np.arange(B + 1, dtype=np.int32) * N_CORR
    imgs_dev = imgs_host.to(dev)
    X_dev, uv_dev = X_host.to(dev), uv_host.to(dev)
    offsets_dev = torch.from_numpy(offsets_np).to(dev)
    bufs = {
        "desc": torch.empty((B, DIM), dtype=torch.float32, device=dev),
        "search_out": (torch.empty((B, 5), dtype=torch.float64, device=dev), torch.empty((B, 5), dtype=torch.int64, device=dev)),
        "pnp": None,
    }
    bufs["pnp"] = pipe.pnp.solve_device(offsets_dev, X_dev, uv_dev, pipe.params)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        return pipe.step_device(imgs_dev, offsets_dev, X_dev, uv_dev, bufs)

    # ---- device-resident throughput (value)
    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        labels, scores, pout = step_dev()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = B * world / ms_per_step * 1e3

    # ---- stage breakdown (device events, same inputs), rank 0 reports
    def time_stage(fn, iters=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    nq = B * world  # queries every shard sweeps per step
    q_all = torch.empty((nq, DIM), dtype=torch.float32, device=dev)
    q_all[:B] = bufs["desc"]
    if world > 1:
        q_all[B:] = bufs["desc"].repeat(world - 1, 1)
    st_desc = time_stage(lambda: pipe.desc.compute_device(imgs_dev, out=bufs["desc"]))
    st_search = time_stage(lambda: local_index.search_device(q_all, 5))
    st_pnp = time_stage(lambda: pipe.pnp.solve_device(offsets_dev, X_dev, uv_dev, pipe.params, out=bufs["pnp"]))
    # dominant-kernel roofline: the search sweep (HBM-bound): algorithmic bytes = local rows * D * 4 per sweep
    # tensor-core sweep: one pass over the DB serves up to 128 queries (64-query tile when no more than 64 wait)
    sweeps = (nq + 127) // 128
    q_per_sweep = 64 if nq <= 64 else 128
    sweep_ms, n_sw = _sweep_timing(local_index, q_all, nq=min(nq, q_per_sweep))
    sweep3_ms, n_sw3 = _sweep_timing(local_index, q_all, nq=3)  # the reference's natural batch: v, vm, vmm
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tensor_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1350.0)))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6.65 TB/s / 1350 TFLOP/s"
    alg_bytes = float(n_mine) * DIM * 4
    sweep_kernel = "scores_tc_kernel" if os.environ.get("CB_TC_V1") == "1" else "scores_tc2_kernel"
    achieved = alg_bytes / (sweep_ms / max(n_sw, 1) * 1e-3) / 1e9 if n_sw else None
    # whole-step view: every stage against the roofline that bounds it
    d_flops, d_bytes = descriptor_algorithmic(net, ROWS, COLS, CHNLS)
    d_tf = d_flops * B / (st_desc * 1e-3) / 1e12
    d_gbs = d_bytes * B / (st_desc * 1e-3) / 1e9
    props = torch.cuda.get_device_properties(dev)
    fp64_peak = props.multi_processor_count * 64 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    p_tf = B * HYP * PNP_MFLOP_PER_HYP * 1e6 / (st_pnp * 1e-3) / 1e12
    s_gbs = sweeps * alg_bytes / (st_search * 1e-3) / 1e9
    roofline_stages = {
        "descriptor": {"ms": st_desc, "gflop_per_keyframe": d_flops / 1e9, "hbm_mbytes_per_keyframe": d_bytes / 1e6,
                       "tensor": {"achieved": d_tf, "peak": tensor_peak, "unit": "TFLOP/s", "frac": d_tf / tensor_peak},
                       "hbm": {"achieved": d_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": d_gbs / hbm_peak},
                       "bound": "hbm (floor %.2f ms per %d keyframes; arithmetic intensity %.0f flop/B is below the ~207 flop/B ridge)" % (d_bytes * B / hbm_peak / 1e6, B, d_flops / d_bytes)},
        "search": {"ms": st_search, "bound": "hbm", "sweeps_per_step": sweeps, "achieved": s_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": s_gbs / hbm_peak,
                   "note": "whole stage (query split + sweep(s) + top-k + fp64 re-rank) against the sweeps' algorithmic bytes"},
        "pnp": {"ms": st_pnp, "bound": "fp64", "hypotheses_per_step": B * HYP, "achieved": p_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": p_tf / fp64_peak,
                "flop_model_mflop_per_hypothesis": PNP_MFLOP_PER_HYP},
    }

    # ---- end to end through the host C-ABI calls (pinned host buffers in, host results out)
    from concurrent.futures import ThreadPoolExecutor

    pool = ThreadPoolExecutor(max_workers=1)
    desc_pool = ThreadPoolExecutor(max_workers=3)
    Xs = [X_host.numpy()[i * N_CORR : (i + 1) * N_CORR] for i in range(B)]
    uvs = [uv_host.numpy()[i * N_CORR : (i + 1) * N_CORR] for i in range(B)]
    # Three descriptor handles (same weights), each driven by one call at a time: batches i+1 and i+2 are uploaded on their
    # handles' copy streams while batch i runs its forward pass -- every handle owns its streams and device buffers.  (With
    # two handles the upload of batch i+2 started only after the search of batch i had returned: ~0.5 ms of idle GPU per step.)
    from cerebro_b200.descriptor import NetvladDescriptor

    ND = 3  # descriptor handles in flight: the upload of batch i+2 must not wait for the search of batch i
    descs = [pipe.desc] + [NetvladDescriptor(net, ROWS, COLS, CHNLS, max_batch=B, device=local_rank) for _ in range(ND - 1)]
    d_out = [torch.empty((B, DIM), dtype=torch.float32).pin_memory().numpy() for _ in range(ND + 1)]

    def run_host(n_steps):
        """n_steps keyframe batches through the host C-ABI calls, organised like the reference node: descriptor
        thread(s) (desc_th), the search thread (this one) and the verifier thread (loopcandidate_consumer_th) -- ctypes
        releases the GIL inside the calls and every handle owns its streams, so batch i+1 is uploaded and described
        while batch i is searched and verified.  Every batch's upload and read-back happens inside this function.
        With a sharded DB the search is ONE host call, cb_index_search_sharded (own descriptors up, the query and top-k
        all-gathers on the handle's stream, merged lists down): no torch tensor, no extra device round trip."""
        from collections import deque

        futs = deque(desc_pool.submit(descs[j % ND].compute, imgs_host.numpy(), d_out[j % (ND + 1)]) for j in range(min(ND, n_steps)))
        pnp_futs = deque()  # the verifier thread always has the next batch queued (its results are read one step later)
        res = None
        dbg = os.environ.get("BENCH_E2E_DEBUG", "")  # diagnosis only: drop one stage to see which dependency leaves the GPU idle
        for i in range(n_steps):
            if "nopnp" not in dbg:
                pnp_futs.append(pool.submit(pipe.pnp.solve, Xs, uvs, pipe.params))
            d = futs.popleft().result()
            if i + ND < n_steps:  # handle i % ND is free again; its next batch goes to another output buffer than the one search(i) reads
                futs.append(desc_pool.submit(descs[i % ND].compute, imgs_host.numpy(), d_out[(i + ND) % (ND + 1)]))
            if "nosearch" in dbg:
                res = None
            elif world > 1:
                res = pipe.index.search_sharded(d, 5)
            else:
                res = pipe.index.search(d, 5)
            if len(pnp_futs) > 1:
                res = (res, pnp_futs.popleft().result())
        while pnp_futs:  # every batch's verification finishes inside the timed region
            res = (res, pnp_futs.popleft().result())
        return res

    run_host(4)
    barrier()
    e2e_steps = max(3, min(4 * args.steps, 80))  # its own timed region: long enough (0.25 s) to average over the host threads' jitter
    t0 = time.perf_counter()
    run_host(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / float(t_e.item())
    h2d = B * ROWS * COLS * CHNLS + B * DIM * 4 + B * N_CORR * 5 * 8 + (B + 1) * 4
    d2h = B * DIM * 4 + B * 5 * (4 + 8 + 8) + B * (16 * 8 + 4 + 12)
    # the host -> device wall: pinned-memory upload bandwidth with every rank copying at the same time
    barrier()
    h2d_gbs = measure_h2d_gbs(dev)
    t_bw = torch.tensor([h2d_gbs], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_bw, op=dist.ReduceOp.MIN)
    h2d_gbs_min = float(t_bw.item())
    barrier()

    # ---- latency of ONE keyframe through the host calls from pageable memory (the reference handles one keyframe at a time)
    latency = None
    if rank == 0 and world == 1 and not args.no_latency:
        latency = measure_latency(pipe, synthetic, cands)

    if rank == 0:
        cpu_base = None
        if not args.no_cpu_baseline and raw is not None:
            ctx = cpu_pipeline_setup(raw)
            cpu_pipeline_run(ctx, 1)
            secs = cpu_pipeline_run(ctx, args.cpu_baseline_keyframes)
            cpu_base = {"value": args.cpu_baseline_keyframes / secs, "unit": UNIT, "cores": ctx["cores"], "kind": "port",
                        "sample": "%d keyframes: torch-CPU fp32 NetVLAD (all cores) + fp32 BLAS top-5 search of the full %dk x %d DB + oracle DLS-PnP RANSAC (1 thread)" % (args.cpu_baseline_keyframes, DB_ROWS // 1000, DIM)}
        # stem + fused blocks (+ an unfused depthwise for a model cut after one) + VLAD head (3); per sweep: query split + tcgen05
        # sweep; per <=128 queries: top-k (two-pass from 4096 rows: bound + filter) + finalize; sharded: + merge; PnP (6)
        launches = (1 + len(net["blocks"]) + 3) + (2 * sweeps + (3 if n_mine >= 4096 else 2) * ((nq + 127) // 128)) + (1 if world > 1 else 0) + 6
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "q15 activations + f16 hi/lo operands, f32 accumulate (descriptor); f16 hi+lo planes / f32 accumulate sweep + f64 re-rank (search); f64 (pnp)",
            "data": "synthetic",
            "model": model_name,
            "config": workload_config(B, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "h2d_gbs_all_ranks_copying": h2d_gbs_min,
                    "pcie_ceiling_keyframes_per_s": h2d_gbs_min * 1e9 / (h2d / B) * world,
                    "note": "host C-ABI calls only (cb_descriptor_compute, cb_index_search[_sharded], cb_pnp_solve_batch); the ceiling is "
                            "the measured pinned upload bandwidth (slowest rank, all ranks copying) over the bytes one keyframe uploads"},
            "gpu_launches": launches,
            "stages_ms": {"descriptor": st_desc, "search": st_search, "pnp": st_pnp},
            "roofline": {"kernel": "%s (tcgen05 search sweep, %d queries per pass, %d launch(es) per step)" % (sweep_kernel, q_per_sweep, sweeps), "bound": "hbm",
                         "achieved": achieved,
                         "peak": hbm_peak, "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None,
                         "traffic": _traffic_from_profile(world, sweep_kernel) if args.workload == "default" else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": sweep_ms / max(n_sw, 1)},
            "roofline_streaming": {"kernel": "scores_kernel<4,8> (3-query sweep: the reference's v, vm, vmm batch)", "bound": "hbm",
                                   "achieved": alg_bytes / (sweep3_ms / max(n_sw3, 1) * 1e-3) / 1e9 if n_sw3 else None,
                                   "peak": hbm_peak, "unit": "GB/s",
                                   "frac": (alg_bytes / (sweep3_ms / max(n_sw3, 1) * 1e-3) / 1e9 / hbm_peak) if n_sw3 else None},
            "roofline_stages": roofline_stages,
            "parity_in_warmup": parity_note,
            "latency": latency,
            "cpu_baseline": cpu_base,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        pipe.comm.close()
        dist.destroy_process_group()


def measure_latency(pipe, synthetic, cands, reps=15):
    """Wall time of ONE keyframe through the host C-ABI calls from pageable memory, the way the reference node drives the path:
    cb_descriptor_compute (1 image) -> cb_index_add -> cb_index_naive_candidate (the 3-query v / vm / vmm rule over the whole DB,
    Cerebro.cpp:1019-1056) -> cb_pnp_solve_batch (1 candidate).  Medians over `reps` keyframes, in ms."""
    img = np.array(synthetic.band_limited_images(1, ROWS, COLS, CHNLS, seed=4242))  # plain pageable numpy memory
    X, uv = np.array(cands[0][0]), np.array(cands[0][1])
    t = {"descriptor": [], "add+naive_candidate(3 queries)": [], "top5_search(1 query)": [], "pnp(1 candidate)": [], "total": []}
    for i in range(reps + 2):
        t0 = time.perf_counter()
        d = pipe.desc.compute(img)
        t1 = time.perf_counter()
        pipe.index.add(d)
        l = pipe.index.ntotal
        pipe.index.naive_candidate(l, 50, 12, 0.85)
        t2 = time.perf_counter()
        pipe.index.search(d, 5)
        t3 = time.perf_counter()
        pipe.pnp.solve([X], [uv], pipe.params)
        t4 = time.perf_counter()
        if i >= 2:
            for k, v in zip(t, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)):
                t[k].append(v * 1e3)
    med = {k: float(np.median(v)) for k, v in t.items()}
    med["note"] = "one keyframe, pageable host buffers, blocking C-ABI calls back to back, median of %d; the 3-query rule and the 1-query top-5 each sweep the whole %dk x %d DB" % (reps, DB_ROWS // 1000, DIM)
    med["unit"] = "ms"
    return med


def _traffic_from_profile(world, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the sweep kernel, from the newest committed
    ncu --set full capture of this same workload (profiles/*_traffic.json, written by tools/summarize_ncu.py);
    only valid for the single-GPU 100k x 8192 DB."""
    import glob

    if world != 1:
        return None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            v = json.load(open(path)).get(kernel + "_dram_bytes_per_launch")
        except Exception:
            v = None
        if v:
            return v
    return None


def _sweep_timing(index, q, nq=16):
    """Summed duration and count of sweep-kernel launches over 10 searches of <= nq queries, from
    the library's CUDA events recorded on the launching stream around exactly that kernel."""
    import torch

    q16 = q[:nq].contiguous() if q.shape[0] >= nq else q
    out = index.search_device(q16, 5)
    torch.cuda.synchronize()
    index.set_timing(True)
    for _ in range(10):
        index.search_device(q16, 5, out=out)
    torch.cuda.synchronize()
    ms, n = index.sweep_timing()
    index.set_timing(False)
    return ms, n


if __name__ == "__main__":
    main()
