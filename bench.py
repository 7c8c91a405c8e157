#!/usr/bin/env python
"""bench.py -- loop-candidate keyframes/sec through desc -> search -> PnP (BASELINE.json metric).

A step = one batch of B synthetic keyframes per GPU through the hot path:
  NetVLAD forward (480x640x3 u8 -> 8192-D)  ->  top-5 inner-product search of every new descriptor
  against a 100k x 8192 fp32 descriptor DB (sharded round-robin over the ranks, per-shard top-k
  all-gathered over NCCL and merged)  ->  DLS-PnP RANSAC (reference parameters: <= 50 hypotheses of
  15 points, 200 correspondences, 20 % outliers) for every keyframe treated as a loop candidate.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

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


def load_net():
    """Shipped default model's weights if the converted fixture is present, else random-init
    weights of the same architecture (there is no network for checkpoints)."""
    from cerebro_b200 import keras_weights

    if os.path.exists(GOLD_W):
        z = np.load(GOLD_W)
        raw = {k.replace("__", "/"): z[k] for k in z.files}
        return keras_weights.fold_mobilenet_netvlad(raw), raw, "mobilenet_conv7_allpairloss (shipped weights)"
    return keras_weights.random_mobilenet_netvlad(CHNLS, 7, 16, seed=0), None, "mobilenet_conv7 architecture, random init"


def workload_config(batch, world):
    return {
        "workload": "desc(480x640x3->8192-D NetVLAD) + top-5 search over 100k x 8192 fp32 DB + DLS-PnP RANSAC "
        "(200 corr, <=50 hyp) per keyframe",
        "keyframes_per_step_per_gpu": batch,
        "db_rows": DB_ROWS,
        "descriptor_dim": DIM,
        "db_sharding": "round-robin over %d rank(s), NCCL all-gather of per-shard top-5" % world,
        "pnp": {"correspondences": N_CORR, "max_hypotheses": HYP, "outlier_frac": 0.2, "adaptive": True},
        "l2": "DB (3.28 GB / ranks) and activations exceed the 126 MB L2; no explicit flush",
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
        "config": workload_config(kf, 1),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ctx["cores"], "kind": "port",
                         "sample": "%d keyframes per step x %d steps, torch-CPU NetVLAD (all cores) + fp32 BLAS search of the full 100k x 8192 DB + oracle DLS-PnP RANSAC (1 thread)" % (kf, args.steps)},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="keyframes per step per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-baseline-keyframes", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

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

    from cerebro_b200 import build as cb_build
    from cerebro_b200 import synthetic
    from cerebro_b200.loop_detector import LoopPipeline

    cb_build.build()  # no-op when the in-tree .so is up to date
    net, raw, model_name = load_net()
    B = args.batch
    rows_local = (DB_ROWS + world - 1) // world
    pipe = LoopPipeline(net, ROWS, COLS, CHNLS, B, rows_local + 8, device=local_rank, sharded=(world > 1), n_corr=N_CORR, hypotheses=HYP)

    # ---- synthetic DB shard, generated on the device (rows are unit-norm N(0,1) vectors)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_mine = len(range(rank, DB_ROWS, world))
    chunk = 12_500
    local_index = pipe.index.local if world > 1 else pipe.index
    for a in range(0, n_mine, chunk):
        nrow = min(chunk, n_mine - a)
        x = torch.randn((nrow, DIM), generator=g, device=dev, dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True)
        local_index.add_local(x)  # every rank generates and bulk-loads only its own shard
        del x
    torch.cuda.synchronize()

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
        "queries": torch.empty((B * world, DIM), dtype=torch.float32, device=dev),
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

    q_all = bufs["queries"] if world > 1 else bufs["desc"]
    st_desc = time_stage(lambda: pipe.desc.compute_device(imgs_dev, out=bufs["desc"]))
    st_search = time_stage(lambda: local_index.search_device(q_all, 5))
    st_pnp = time_stage(lambda: pipe.pnp.solve_device(offsets_dev, X_dev, uv_dev, pipe.params, out=bufs["pnp"]))
    # dominant-kernel roofline: the search sweep (HBM-bound): algorithmic bytes = local rows * D * 4 per sweep
    nq = q_all.shape[0]
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
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6.65 TB/s"
    alg_bytes = float(n_mine) * DIM * 4
    sweep_kernel = "scores_tc_kernel" if os.environ.get("CB_TC_V1") == "1" else "scores_tc2_kernel"
    achieved = alg_bytes / (sweep_ms / max(n_sw, 1) * 1e-3) / 1e9 if n_sw else None

    # ---- end to end through the host C-ABI calls (pinned host buffers in, host results out)
    from concurrent.futures import ThreadPoolExecutor

    pool = ThreadPoolExecutor(max_workers=1)
    desc_pool = ThreadPoolExecutor(max_workers=2)
    Xs = [X_host.numpy()[i * N_CORR : (i + 1) * N_CORR] for i in range(B)]
    uvs = [uv_host.numpy()[i * N_CORR : (i + 1) * N_CORR] for i in range(B)]
    # Two descriptor handles (same weights), each driven by one call at a time: batch i+1 is uploaded on one handle's copy
    # stream while batch i runs its forward pass on the other's -- every handle owns its streams and device buffers.
    from cerebro_b200.descriptor import NetvladDescriptor

    descs = [pipe.desc, NetvladDescriptor(net, ROWS, COLS, CHNLS, max_batch=B, device=local_rank)]
    d_out = [torch.empty((B, DIM), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]

    def run_host(n_steps):
        """n_steps keyframe batches through the host C-ABI calls, organised like the reference node: descriptor
        thread(s) (desc_th), the search thread (this one) and the verifier thread (loopcandidate_consumer_th) -- ctypes
        releases the GIL inside the calls and every handle owns its streams, so batch i+1 is uploaded and described
        while batch i is searched and verified.  Every batch's upload and read-back happens inside this function."""
        from collections import deque

        futs = deque(desc_pool.submit(descs[j % 2].compute, imgs_host.numpy(), d_out[j % 2]) for j in range(min(2, n_steps)))
        res = None
        for i in range(n_steps):
            fut_p = pool.submit(pipe.pnp.solve, Xs, uvs, pipe.params)
            d = futs.popleft().result()
            if world > 1:
                dd = torch.from_numpy(d).to(dev)
                dist.all_gather_into_tensor(bufs["queries"], dd)
                s, l = pipe.index.search_device(bufs["queries"], 5)
                res = (s.cpu(), l.cpu())
            else:
                res = pipe.index.search(d, 5)
            if i + 2 < n_steps:  # handle i % 2 and its output buffer are free again: batch i has been searched
                futs.append(desc_pool.submit(descs[i % 2].compute, imgs_host.numpy(), d_out[i % 2]))
            res = (res, fut_p.result())
        return res

    run_host(2)
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    run_host(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / float(t_e.item())
    h2d = B * ROWS * COLS * CHNLS + B * DIM * 4 * (1 if world == 1 else 1) + B * N_CORR * 5 * 8 + (B + 1) * 4
    d2h = B * DIM * 4 + B * world * 5 * (4 + 8 + 8) // max(world, 1) + B * (16 * 8 + 4 + 12)

    if rank == 0:
        cpu_base = None
        if not args.no_cpu_baseline and raw is not None:
            ctx = cpu_pipeline_setup(raw)
            cpu_pipeline_run(ctx, 1)
            secs = cpu_pipeline_run(ctx, args.cpu_baseline_keyframes)
            cpu_base = {"value": args.cpu_baseline_keyframes / secs, "unit": UNIT, "cores": ctx["cores"], "kind": "port",
                        "sample": "%d keyframes: torch-CPU fp32 NetVLAD (all cores) + fp32 BLAS top-5 search of the full 100k x 8192 DB + oracle DLS-PnP RANSAC (1 thread)" % args.cpu_baseline_keyframes}
        # stem + 7 fused blocks + VLAD head (3); per sweep: query split + tcgen05 sweep; per <=128 queries: top-k + finalize; PnP (6)
        # (two-pass top-k from 4096 rows: bound + filter + finalize per group of <= 128 queries)
        launches = (1 + 7 + 3) + (2 * sweeps + (3 if n_mine >= 4096 else 2) * ((nq + 127) // 128)) + 6
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
            "dtype": "f16 activations/f32 accumulate (descriptor), f16 hi+lo planes/f32 accumulate sweep + f64 re-rank (search), f64 (pnp)",
            "data": "synthetic",
            "model": model_name,
            "config": workload_config(B, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": launches,
            "stages_ms": {"descriptor": st_desc, "search": st_search, "pnp": st_pnp},
            "roofline": {"kernel": "%s (tcgen05 search sweep, %d queries per pass, %d launch(es) per step)" % (sweep_kernel, q_per_sweep, sweeps), "bound": "hbm",
                         "achieved": achieved,
                         "peak": hbm_peak, "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None,
                         "traffic": _traffic_from_profile(world, sweep_kernel), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": sweep_ms / max(n_sw, 1)},
            "roofline_streaming": {"kernel": "scores_kernel<4,8> (3-query sweep: the reference's v, vm, vmm batch)", "bound": "hbm",
                                   "achieved": alg_bytes / (sweep3_ms / max(n_sw3, 1) * 1e-3) / 1e9 if n_sw3 else None,
                                   "peak": hbm_peak, "unit": "GB/s",
                                   "frac": (alg_bytes / (sweep3_ms / max(n_sw3, 1) * 1e-3) / 1e9 / hbm_peak) if n_sw3 else None},
            "cpu_baseline": cpu_base,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
