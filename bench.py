#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric; workload at N=1 = configs[1]: 64 SAM2.1-tiny crops of 1024² +
node analysis of the resulting wire masks, per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...] [--no-extras]

One step = one pass of the hot path over one batch of synthetic schematics (seeded generator,
circuitvision_b200/synth.py).  Inputs are resident in HBM for `value`; `e2e` goes through the public Python API
with pinned HOST buffers (H2D + D2H inside the timed region).  Images are independent: under torchrun every
rank processes its own batch (weak scaling, no data-path collective); time = max over ranks.

The default run (N=1) also emits time-boxed, driver-run sub-records in the same JSON line (`extra`):
  cfg3        SAM2.1-base+ on 256 crops (chunks of 64)                       BASELINE configs[2]
  cfg4_nodes  drop-in node analysis on dense 4096² masks, resident and e2e   BASELINE configs[3]
  cfg4_ccl    native-resolution CCL kernel (cv_ccl_label) GB/s vs HBM peak   BASELINE configs[3], north_star's 60 % target
`--workload cfg5` is BASELINE configs[4]: a FIXED total of schematics sharded image-wise over the ranks (strong
scaling), per-image results gathered on the host of rank 0 by index.

`--impl reference` times the reference's CPU implementation of the same path on this box's host cores (the oracle
port: /root/reference does not travel and its sam2 dependency is not installable) and never loads the product .so.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SAM2.1 crops/sec + node-analysis images/sec at 1024^2"
UNIT = "images/s"
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r2_ncu_traffic_pipeline_b64.json")


def _args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pipeline", choices=["auto", "pipeline", "nodes", "nodes4096", "sam2", "cfg5"])
    ap.add_argument("--batch", type=int, default=None,
                    help="images per GPU per step (default 64; 256 for nodes4096, whose border walks are latency-bound "
                         "chains that only a larger batch amortises)")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--variant", default="tiny")
    ap.add_argument("--chunk", type=int, default=64, help="crops per SAM 2.1 engine pass (workspace is sized for this many)")
    ap.add_argument("--total", type=int, default=8192, help="cfg5: schematics in the whole job (sharded over the ranks)")
    ap.add_argument("--operands", default="fp16", choices=["fp16", "bf16"],
                    help="16-bit tensor-core operand format of the SAM 2.1 path (DESIGN.md section 2)")
    ap.add_argument("--depth", type=int, default=2, help="batches in flight in the e2e pipeline (CropPipeline / MaskPipeline slots)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg3 / cfg4 sub-records of the default run")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernels with events in the timed region")
    a = ap.parse_args()
    if a.workload == "auto":
        a.workload = "pipeline"
    if a.workload == "nodes4096":
        a.size = 4096
    if a.batch is None:
        a.batch = 256 if a.workload == "nodes4096" else 64
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    return a


def workload_name(workload, a):
    if workload == "pipeline":
        return f"cfg2: SAM2.1-{a.variant} {a.batch} crops {a.size}^2 + node analysis of the {a.batch} masks, per GPU"
    if workload == "sam2":
        return f"SAM2.1-{a.variant} {a.batch} crops {a.size}^2 (segmentation only), per GPU"
    if workload == "nodes4096":
        return f"cfg4: drop-in node analysis (cv_nodes_analyze) on {a.batch} dense 4096^2 masks, per GPU"
    if workload == "cfg5":
        return (f"cfg5: {a.total} schematics {a.size}^2, SAM2.1-{a.variant} + node analysis, sharded image-wise over the ranks, "
                f"per-image results gathered on the host of rank 0")
    return f"node analysis only: {a.batch} wire masks {a.size}^2 (SAM2 stage excluded), per GPU"


def make_config(workload, a):
    """`config` of the JSON line — built from the arguments only, so both arms print the same object."""
    use_sam2 = workload in ("pipeline", "sam2", "cfg5")
    n_pool = 4 if a.size <= 1024 else 2
    return {"workload": workload_name(workload, a), "images_per_gpu_per_step": a.batch, "size": a.size, "variant": a.variant,
            "l2_policy": f"inputs rotate over a pool of {n_pool} batches "
                         f"({n_pool * a.batch * a.size * a.size * (3 if use_sam2 else 1) >> 20} MiB) larger than L2",
            "sharding": "image-wise, no collective",
            "stage_overlap": "node analysis of batch i on a second stream under the SAM 2.1 forward of batch i+1"
            if workload in ("pipeline", "cfg5") else "none"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------ CPU side (oracle port)
def _cpu_nodes_worker(task):
    """One worker process: generate its share of schematics (untimed), then time the oracle over them."""
    seeds, size, reps = task
    import cv2
    from circuitvision_b200 import synth
    from oracle import node_oracle
    cv2.setNumThreads(1)  # one image per core; the pool supplies the parallelism
    data = [synth.make_schematic(s, size)[:2] for s in seeds]
    node_oracle.get_node_connections(*data[0])  # warm
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for mask, boxes in data:
            node_oracle.get_node_connections(mask, boxes)
        out.append(time.perf_counter() - t0)
    return out


def cpu_nodes_throughput(n_images: int, size: int, procs: int, reps: int = 1):
    """Oracle node analysis (cv2 — the reference's own arithmetic) over n_images on `procs` worker processes,
    inputs generated before the clock starts.  Returns (images/s of the slowest worker's mean rep, seconds, per-rep)."""
    import multiprocessing as mp
    procs = max(1, min(procs, n_images))
    shares = [list(range(10_000 + k, 10_000 + n_images, procs)) for k in range(procs)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        per = pool.map(_cpu_nodes_worker, [(sh, size, reps) for sh in shares])
    rep_s = [max(w[r] for w in per) for r in range(reps)]  # a rep ends when the slowest worker ends
    dt = float(np.mean(rep_s))
    return n_images / dt, dt, rep_s


class CpuPipeline:
    """The reference's two hot calls on the host, one image at a time as the reference runs them
    (analysis_pipeline.py:206 segment_with_sam2 -> :234 get_node_connections): fp32 SAM 2.1 restatement with every torch
    thread, then the cv2 node analysis of the mask it produced."""

    def __init__(self, variant: str, size: int, threads: int, with_nodes: bool = True):
        import torch
        from oracle import sam2_oracle
        torch.set_num_threads(threads)
        self.model = sam2_oracle.build_oracle(variant, seed=0)
        self.size, self.with_nodes, self.k = size, with_nodes, 0

    def step(self, n_images: int) -> float:
        from circuitvision_b200 import synth
        from oracle import node_oracle, sam2_oracle
        data = []
        for _ in range(n_images):
            _, boxes, rgb = synth.make_schematic(10_000 + self.k, self.size, render_rgb=True)
            data.append((rgb, boxes))
            self.k += 1
        t0 = time.perf_counter()
        for rgb, boxes in data:
            mask, _, _ = sam2_oracle.segment(self.model, rgb)
            if self.with_nodes:
                node_oracle.get_node_connections(mask, boxes)
        return time.perf_counter() - t0


def run_reference(a, workload):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if workload in ("pipeline", "sam2", "cfg5"):
        # >= 8 SAM 2.1 images over the timed steps; a step is a bounded sample of the batch
        n_img = max(1, math.ceil(8 / max(1, a.steps)))
        cpu = CpuPipeline(a.variant, a.size, cores, with_nodes=workload != "sam2")
        for _ in range(min(a.warmup, 2)):
            cpu.step(1)
        secs = [cpu.step(n_img) for _ in range(a.steps)]
        ms = float(np.mean(secs)) * 1e3
        value = n_img / (ms / 1e3)
        sample = (f"{n_img} image(s) of {a.size}^2 per step x {a.steps} steps: SAM2.1-{a.variant} fp32 restatement on {cores} torch "
                  f"threads" + ("" if workload == "sam2" else ", then the cv2 node analysis of its mask (same process)"))
    else:
        per_step = min(a.batch, 64) if a.size <= 1024 else min(a.batch, 16)
        _, _, rep_s = cpu_nodes_throughput(per_step, a.size, cores, reps=a.warmup + a.steps)
        ms = float(np.mean(rep_s[a.warmup:])) * 1e3
        value = per_step / (ms / 1e3)
        sample = f"{per_step} images of {a.size}^2 per step, oracle node analysis (cv2) on {min(cores, per_step)} worker processes"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if workload == "cfg5" else "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
        "config": make_config(workload, a),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ roofline helpers
def _gemm_bytes(name):
    m = re.match(r"gemm M(\d+) N(\d+) K(\d+) bn\d+(.*)", name)
    if not m:
        return None
    M, N, K = (int(x) for x in m.groups()[:3])
    tag = m.group(4)
    return M * K * 2 + N * K * 2 + M * N * (2 if "->b" in tag else 4) + (M * N * 4 if "+r" in tag else 0), 2.0 * M * N * K


def _fused_bytes(name):
    """Algorithmic HBM bytes per launch of the non-GEMM tensor kernels, from the shape tag: attention reads
    Q, K, V and writes O once (16-bit); the fused MLP reads and writes the fp32 rows once; the patch embedding reads the
    raw pixels and writes X0 (fp32) + norm1 (16-bit)."""
    m = re.match(r"attn(?:_global)? Mq(\d+) Wq(\d+) Wkv(\d+) h(\d+) d(\d+)", name)
    if m:
        Mq, Wq, Wkv, h, d = (int(x) for x in m.groups())
        return 2.0 * d * h * (2 * Mq + 2 * (Mq // Wq) * Wkv)
    m = re.match(r"mlp M(\d+) C(\d+)", name)
    if m:
        return 8.0 * int(m.group(1)) * int(m.group(2))
    m = re.match(r"patch_embed M(\d+) E(\d+)", name)
    if m:
        return float(m.group(1)) * (12 + 6 * int(m.group(2)))
    return None


def _shape_row(r, steps, tens_peak, hbm_peak):
    """One launch group of the timed region: achieved rate and its fraction of BOTH rooflines (algorithmic bytes follow
    from the shape tag): the windowed-attention and stage-1/2 shapes are HBM-bound, not tensor-bound."""
    row = {"name": r["name"], "launches": r["launches"], "ms_per_step": r["ms"] / steps,
           "rate_T_per_s": r["work"] / max(r["ms"], 1e-9) / 1e9}
    gb = _gemm_bytes(r["name"])
    sec = r["ms"] / 1e3 / max(1, r["launches"])
    if gb:
        row["tensor_frac"] = gb[1] / sec / 1e12 / tens_peak
        row["hbm_frac"] = gb[0] / sec / 1e9 / hbm_peak
    elif r["name"].startswith(("attn", "mlp", "patch_embed")):
        row["tensor_frac"] = row["rate_T_per_s"] / tens_peak
        fb = _fused_bytes(r["name"])
        if fb:
            row["hbm_frac"] = fb / sec / 1e9 / hbm_peak
    return row


TENSOR_KERNELS = ("k_gemm_tc", "k_attn_tc", "k_attn_global", "k_mlp_fused", "k_patch_embed")
ALIAS = {"gemm ": "k_gemm_tc", "attn_global ": "k_attn_global", "attn ": "k_attn_tc", "ln_rows ": "k_ln_rows", "mlp ": "k_mlp_fused",
         "patch_embed ": "k_patch_embed"}


def fold_table(table):
    folded = {}
    for r in table:
        name = next((v for k, v in ALIAS.items() if r["name"].startswith(k)), r["name"])
        f = folded.setdefault(name, {"name": name, "launches": 0, "ms": 0.0, "work": 0.0, "bytes": 0.0})
        f["launches"] += r["launches"]
        f["ms"] += r["ms"]
        f["work"] += r["work"]
        gb = _gemm_bytes(r["name"])
        if gb:
            f["bytes"] += r["launches"] * gb[0]
    return list(folded.values())


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p.get("hbm_gbs", 6650.0), p.get("bf16_tflops_sustained", 1400.0), p.get("bf16_tflops", 1650.0), True
    except Exception:
        return 6650.0, 1400.0, 1650.0, False


def make_roofline(table, steps, workload, B, variant, S):
    """Roofline of the dominant kernel from live CUDA-event durations of the timed region.  Tensor-core kernels are
    reported against the tensor peak; the byte view of the GEMM shapes is a secondary field."""
    hbm_peak, tens_peak, _, measured = load_peaks()
    shapes = sorted(table, key=lambda r: -r["ms"])[:16]
    folded = fold_table(table)
    tot = sum(r["ms"] for r in folded) or 1.0
    kern_rows = [{"name": r["name"], "launches": r["launches"], "ms_per_step": r["ms"] / steps, "share": r["ms"] / tot,
                  "work_per_launch": r["work"] / max(1, r["launches"])} for r in sorted(folded, key=lambda r: -r["ms"])]
    top = max(folded, key=lambda r: r["ms"])
    avg_s = top["ms"] / 1e3 / max(1, top["launches"])
    wpl = top["work"] / max(1, top["launches"])
    src = "MEASURED_PEAKS.json" if measured else "fallback (B200_PROFILING.md)"
    if top["name"] in TENSOR_KERNELS:
        ach = wpl / avg_s / 1e12
        roof = {"kernel": top["name"], "bound": "tensor", "achieved": ach, "peak": tens_peak, "unit": "TFLOP/s",
                "frac": ach / tens_peak, "traffic": None, "peak_source": src + " bf16_tflops_sustained (kernel timed inside a long step)"}
        if top["bytes"]:
            roof["hbm_frac"] = top["bytes"] / (top["ms"] / 1e3) / 1e9 / hbm_peak
            roof["hbm_note"] = "byte view: unfused A + W + C (+ residual) per launch over all shapes against the HBM copy peak"
    else:
        ach = wpl / avg_s / 1e9
        roof = {"kernel": top["name"], "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": None, "peak_source": src + " hbm_gbs"}
    roof["avg_launch_us"] = avg_s * 1e6
    roof["share_of_step"] = top["ms"] / tot
    # every tensor-core kernel of the step together (encoder GEMMs + fused MLPs + attention): flops / summed time
    tk = [r for r in folded if r["name"] in TENSOR_KERNELS]
    if tk:
        t_ms, t_w = sum(r["ms"] for r in tk), sum(r["work"] for r in tk)
        roof["all_tensor_kernels"] = {"ms_per_step": t_ms / steps, "TFLOPs": t_w / (t_ms / 1e3) / 1e12,
                                      "frac": t_w / (t_ms / 1e3) / 1e12 / tens_peak}
    try:
        tr = json.load(open(TRAFFIC_JSON))
        k = tr["kernels"].get(top["name"])
        if k and workload == "pipeline" and B == 64 and variant == "tiny" and S == 1024:
            roof["traffic"] = k["traffic_bytes_per_launch"]
            roof["traffic_source"] = os.path.relpath(TRAFFIC_JSON, ROOT) + " (ncu dram__bytes_read+write, per launch)"
    except Exception:
        pass
    return roof, kern_rows, [_shape_row(r, steps, tens_peak, hbm_peak) for r in shapes]


# ------------------------------------------------------------------------------------------ GPU side
class Ctx:
    pass


def _timed(ctx, fn, steps, after=None):
    import torch
    import torch.distributed as dist
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    if after is not None:
        after()
    e1.record()
    ctx.barrier()
    ms = e0.elapsed_time(e1)
    if ctx.world > 1:
        t = torch.tensor([ms], device=ctx.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def make_pool(rank, B, S, n_pool, want_rgb, seed0=0, dense=None):
    """Synthetic inputs, distinct per rank: `uniq` distinct schematics per pool slot, tiled to the batch."""
    from circuitvision_b200 import synth
    uniq = min(B, 16)
    pool_masks, pool_boxes, pool_rgb = [], [], []
    for p in range(n_pool):
        ms_, bx_, rgb_ = [], [], []
        for i in range(uniq):
            m, b, rgb = synth.make_schematic(seed0 + 1_000_000 * rank + 1000 * p + i, S, render_rgb=want_rgb)
            ms_.append(m)
            bx_.append(b)
            rgb_.append(rgb)
        idx = [i % uniq for i in range(B)]
        pool_masks.append(np.stack([ms_[i] for i in idx]))
        pool_boxes.append([bx_[i] for i in idx])
        if want_rgb:
            pool_rgb.append(np.stack([rgb_[i] for i in idx]))
    return pool_masks, pool_boxes, pool_rgb


def run_main_workload(ctx, a, workload):
    import torch
    from circuitvision_b200 import _lib
    from circuitvision_b200.circuit_analyzer import CircuitAnalyzer
    lib, dev, rank, world = ctx.lib, ctx.dev, ctx.rank, ctx.world
    B, S = a.batch, a.size
    n_pool = 4 if S <= 1024 else 2
    use_sam2 = workload in ("pipeline", "sam2")
    use_nodes = workload in ("pipeline", "nodes", "nodes4096")
    t_gen = time.perf_counter()
    pool_masks, pool_boxes, pool_rgb = make_pool(rank, B, S, n_pool, use_sam2)
    gen_s = time.perf_counter() - t_gen

    A = CircuitAnalyzer(use_sam2=False, debug=False, device=ctx.local, render_debug_images=False)
    na = A._na()
    sam = None
    if use_sam2:
        from circuitvision_b200 import sam2_infer
        sam = sam2_infer.build_random_init(a.variant, device=dev, seed=0, max_batch=min(B, a.chunk),
                                           operand_dtype=torch.float16 if a.operands == "fp16" else torch.bfloat16)
    d_masks = [torch.from_numpy(m).to(dev) for m in pool_masks] if not use_sam2 else None
    h_masks = [torch.from_numpy(m).pin_memory() for m in pool_masks] if not use_sam2 else None
    d_boxes = [na.upload_boxes(bx, S, S) for bx in pool_boxes]
    d_rgb = [torch.from_numpy(x).to(dev) for x in pool_rgb] if use_sam2 else None
    h_rgb = [torch.from_numpy(x).pin_memory() for x in pool_rgb] if use_sam2 else None

    launches_per_step = [0]
    s_main = torch.cuda.current_stream(dev)
    s_nodes = torch.cuda.Stream(dev) if (use_sam2 and use_nodes) else None
    ev_mask = torch.cuda.Event()

    def step_resident(i):
        p = i % n_pool
        n = 0
        if use_sam2:
            masks = sam.segment_batch_u8(d_rgb[p])  # [B,S,S] u8 {0,255} on device
            n += sam.last_launches
        else:
            masks = d_masks[p]
        if use_nodes:
            rec, off, rb, mx = d_boxes[p]
            if s_nodes is not None:
                ev_mask.record(s_main)
                with torch.cuda.stream(s_nodes):
                    s_nodes.wait_event(ev_mask)
                    r = na.run(masks, rec, off, mx, rb)
                    masks.record_stream(s_nodes)
            else:
                r = na.run(masks, rec, off, mx, rb)
            n += r.launches
        launches_per_step[0] = n

    def join():
        if s_nodes is not None:
            s_main.wait_stream(s_nodes)  # the last batch's node analysis ends inside the timed region

    for i in range(a.warmup):
        step_resident(i)
    torch.cuda.synchronize()
    prof = not a.no_profile
    lib.cv_profile_reset()
    lib.cv_profile_enable(1 if prof else 0)
    ctx.barrier()  # every rank has finished its warm-up: the sampler below sees rank 0's GPU under load, not waiting for the others
    clocks = ClockSampler(ctx.local) if rank == 0 else None
    total_ms = _timed(ctx, step_resident, a.steps, join)
    clk = clocks.stop() if clocks else None
    lib.cv_profile_enable(0)
    table = _lib.profile_table() if prof else []
    value = B * a.steps * world / (total_ms / 1e3)

    # ---- e2e through the public batch API with pinned host buffers
    # the same K steps as the resident region (capped at 20): with 5 the pipeline's fill and drain — the first upload before any
    # compute, the last batch's node analysis + download + host lists after it — were a tenth of the measurement
    e2e_steps = max(2, min(a.steps, 20))
    if workload == "pipeline":
        from circuitvision_b200.pipeline import CropPipeline
        pipe = CropPipeline(sam, B, depth=a.depth)
        sink = [0]

        def consume(res):
            # what a caller reads back: every image's node list in the reference's format
            for b in range(B):
                sink[0] += len(res.nodes(b))

        def run_e2e(steps):
            for i in range(steps):
                if pipe.inflight == pipe.depth:
                    consume(pipe.collect())
                p = i % n_pool
                pipe.submit(h_rgb[p], pool_boxes[p])
            while pipe.inflight:
                consume(pipe.collect())

        run_e2e(max(2, a.depth))  # warm: every slot allocates its pinned result buffers on first use
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        run_e2e(e2e_steps)  # collect() waits for each batch's device->host copies
        e1.record()
        ctx.barrier()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3 * 0.0)
        h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
        del pipe
    elif use_nodes:
        from circuitvision_b200.pipeline import MaskPipeline
        pipe = MaskPipeline(ctx.local, B, S, S, depth=2)
        sink = [0]

        def run_e2e(steps):
            for i in range(steps):
                if pipe.inflight == pipe.depth:
                    sink[0] += int(pipe.collect().n_nodes.sum())
                p = i % n_pool
                pipe.submit(h_masks[p], pool_boxes[p])
            while pipe.inflight:
                sink[0] += int(pipe.collect().n_nodes.sum())

        run_e2e(2)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(e2e_steps)
        e1.record()
        ctx.barrier()
        e2e_ms = e0.elapsed_time(e1)
        h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
        del pipe
    else:
        h_out = torch.empty((B, S, S), dtype=torch.uint8, pin_memory=True)
        hb = [0, 0]

        def step_e2e(i):
            x = h_rgb[i % n_pool].to(dev, non_blocking=True)
            masks = sam.segment_batch_u8(x)
            h_out.copy_(masks, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            hb[0], hb[1] = x.numel(), h_out.numel()

        step_e2e(0)
        e2e_ms = _timed(ctx, step_e2e, e2e_steps)
        h2d, d2h = hb
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = B * e2e_steps * world / (e2e_ms / 1e3)
    out = dict(value=value, total_ms=total_ms, clk=clk, table=table, e2e_value=e2e_value, e2e_ms=e2e_ms, e2e_steps=e2e_steps,
               h2d=int(h2d), d2h=int(d2h), launches=int(launches_per_step[0] * a.steps), gen_s=gen_s)
    del sam, d_rgb, h_rgb, d_masks, h_masks, A, na
    torch.cuda.empty_cache()
    return out


# ---- driver-run sub-records of the default line (time-boxed; N = 1 only)
def extra_cfg3(ctx, a):
    """BASELINE configs[2]: SAM2.1-base+ encoder + decoder on 256 crops of 1024² (engine passes of 64)."""
    import torch
    from circuitvision_b200 import _lib, sam2_infer
    lib, dev = ctx.lib, ctx.dev
    B, chunk = 256, 64
    _, _, pool_rgb = make_pool(ctx.rank, B, 1024, 1, True, seed0=50_000)
    sam = sam2_infer.build_random_init("base_plus", device=dev, seed=0, max_batch=chunk,
                                       operand_dtype=torch.float16 if a.operands == "fp16" else torch.bfloat16)
    d = torch.from_numpy(pool_rgb[0]).to(dev)
    h = torch.from_numpy(pool_rgb[0]).pin_memory()
    for _ in range(2):
        sam.segment_batch_u8(d)
    torch.cuda.synchronize()
    steps = 3
    lib.cv_profile_reset()
    lib.cv_profile_enable(1)
    ms = _timed(ctx, lambda i: sam.segment_batch_u8(d), steps)
    lib.cv_profile_enable(0)
    roof, kern, shapes = make_roofline(_lib.profile_table(), steps, "cfg3", B, "base_plus", 1024)
    # end to end: upload of batch i+1 and download of batch i-1 on their own streams under the forward of batch i
    h_out = [torch.empty((B, 1024, 1024), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    x_dev = [torch.empty_like(d) for _ in range(2)]
    s_main, s_in, s_out = torch.cuda.current_stream(), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_fwd = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    e_steps = 4

    def e2e_run(n):
        for i in range(n):
            k = i % 2
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_fwd[k])  # the forward that read this input buffer two batches ago
                x_dev[k].copy_(h, non_blocking=True)
                ev_in[k].record(s_in)
            s_main.wait_event(ev_in[k])
            masks = sam.segment_batch_u8(x_dev[k])
            ev_fwd[k].record(s_main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_fwd[k])
                if i >= 2:
                    ev_out[k].synchronize()  # the host buffer's previous download has landed (a caller would read it here)
                h_out[k].copy_(masks, non_blocking=True)
                masks.record_stream(s_out)
                ev_out[k].record(s_out)
        s_main.wait_stream(s_out)

    e2e_run(2)
    torch.cuda.synchronize()
    e_ms = _timed(ctx, lambda i: e2e_run(e_steps) if i == 0 else None, 1)
    rec = {"workload": "cfg3: SAM2.1-base_plus encoder+decoder, 256 crops 1024^2 per step (4 engine passes of 64)",
           "value": B * steps / (ms / 1e3), "unit": "crops/s", "ms_per_step": ms / steps, "steps": steps,
           "e2e": {"value": B * e_steps / (e_ms / 1e3), "unit": "crops/s", "steps": e_steps, "h2d_bytes_per_step": int(h.numel()),
                   "d2h_bytes_per_step": int(h_out[0].numel())},
           "flops_per_image": 0.645e12 + 5.7e9,
           "whole_step_tensor_frac": (0.645e12 + 5.7e9) * B * steps / (ms / 1e3) / 1e12 / load_peaks()[1],
           "roofline": roof, "kernels": kern[:6], "gpu_launches": int(sam.last_launches * steps * (B // chunk))}
    del sam, d, h, h_out, x_dev
    torch.cuda.empty_cache()
    if not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu = CpuPipeline("base_plus", 1024, cores, with_nodes=False)
        cpu.step(1)
        dt = cpu.step(2)
        rec["cpu_baseline"] = {"value": 2 / dt, "unit": "crops/s", "cores": cores, "kind": "port",
                               "sample": f"2 images, SAM2.1-base_plus fp32 restatement, {cores} torch threads, {dt:.1f} s"}
    return rec


def extra_cfg4_nodes(ctx, a):
    """BASELINE configs[3] (i): the drop-in node analysis on dense 4096² masks, resident and end to end."""
    import torch
    from circuitvision_b200 import _lib
    from circuitvision_b200.nodes import NodeAnalyzer
    from circuitvision_b200.pipeline import MaskPipeline
    lib, dev = ctx.lib, ctx.dev
    B, S = 128, 4096
    pool_masks, pool_boxes, _ = make_pool(ctx.rank, B, S, 2, False, seed0=60_000)
    na = NodeAnalyzer(ctx.local)
    d_masks = [torch.from_numpy(m).to(dev) for m in pool_masks]
    d_boxes = [na.upload_boxes(bx, S, S) for bx in pool_boxes]
    launches = [0]

    def step(i):
        rec, off, rb, mx = d_boxes[i % 2]
        launches[0] = na.run(d_masks[i % 2], rec, off, mx, rb).launches

    for i in range(3):
        step(i)
    steps = 6
    lib.cv_profile_reset()
    lib.cv_profile_enable(1)
    ms = _timed(ctx, step, steps)
    lib.cv_profile_enable(0)
    roof, kern, _ = make_roofline(_lib.profile_table(), steps, "cfg4", B, "", S)
    hbm_peak = load_peaks()[0]
    alg = 2.0 * S * S * B  # read mask + write emptied
    rec = {"workload": f"cfg4: drop-in node analysis (cv_nodes_analyze) on {B} dense 4096^2 masks per step",
           "value": B * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "algorithmic_GBps": alg * steps / (ms / 1e3) / 1e9, "hbm_frac_whole_call": alg * steps / (ms / 1e3) / 1e9 / hbm_peak,
           "roofline": roof, "kernels": kern[:6], "gpu_launches": int(launches[0] * steps)}
    del d_masks
    torch.cuda.empty_cache()
    h_masks = [torch.from_numpy(m).pin_memory() for m in pool_masks]
    pipe = MaskPipeline(ctx.local, B, S, S, depth=2)
    sink = [0]

    def run_e2e(n):
        for i in range(n):
            if pipe.inflight == pipe.depth:
                sink[0] += int(pipe.collect().n_nodes.sum())
            pipe.submit(h_masks[i % 2], pool_boxes[i % 2])
        while pipe.inflight:
            sink[0] += int(pipe.collect().n_nodes.sum())

    run_e2e(2)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e_steps = 16  # 2048 masks: with 4 steps the first upload and the last download (2 GB each, nothing to overlap with) were a third of it
    run_e2e(e_steps)
    e1.record()
    ctx.barrier()
    e_ms = e0.elapsed_time(e1)
    rec["e2e"] = {"value": B * e_steps / (e_ms / 1e3), "unit": UNIT, "steps": e_steps, "h2d_bytes_per_step": int(pipe.h2d_bytes),
                  "d2h_bytes_per_step": int(pipe.d2h_bytes), "pcie_GBps": (pipe.h2d_bytes + pipe.d2h_bytes) * e_steps / (e_ms / 1e3) / 1e9,
                  "note": "pinned H2D of batch i+1 / analysis of i / compacted D2H of i-1 on separate streams (MaskPipeline)"}
    del pipe, h_masks
    torch.cuda.empty_cache()
    if not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = 2 * cores
        ips, dt, _ = cpu_nodes_throughput(n, S, cores, reps=2)
        rec["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{n} dense 4096^2 masks, oracle node analysis (cv2) on {cores} worker processes, {dt:.1f} s per pass"}
    return rec


def _cpu_ccl_worker(task):
    import cv2
    from circuitvision_b200 import synth
    cv2.setNumThreads(1)
    seeds, reps = task
    masks = [synth.make_schematic(s, 4096)[0] for s in seeds]
    t0 = time.perf_counter()
    for _ in range(reps):
        for m in masks:
            cv2.connectedComponents(m, connectivity=8, ltype=cv2.CV_32S)
    return (time.perf_counter() - t0) / reps


def extra_cfg4_ccl(ctx, a):
    """BASELINE configs[3] (ii): native-resolution CCL kernel, 1 B/px read + 4 B/px label write = 5 B/px."""
    import torch
    from circuitvision_b200 import _lib, synth
    lib, dev = ctx.lib, ctx.dev
    S = 4096
    base = torch.from_numpy(np.stack([synth.make_schematic(900 + i, S)[0] for i in range(8)])).to(dev)  # 8 distinct masks
    st = torch.cuda.current_stream().cuda_stream
    hbm_peak = load_peaks()[0]

    def measure(B, steps, table):
        # two input pools (> L2 each from 8 images on), assembled on the device from the 8 distinct masks
        pool = [base[(torch.arange(B, device=dev) + p) % 8].contiguous() for p in range(2)]
        labels = torch.empty((B, S, S), dtype=torch.int32, device=dev)
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        ws_bytes = lib.cv_ccl_workspace_bytes(B, S, S)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

        def run(i):
            m = pool[i % 2]
            _lib.check(lib.cv_ccl_label(m.data_ptr(), B, S, S, 8, labels.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws_bytes, st),
                       "cv_ccl_label")

        for i in range(4):
            run(i)
        if table:
            lib.cv_profile_reset()
            lib.cv_profile_enable(1)
        ms = _timed(ctx, run, steps)
        tab = []
        if table:
            lib.cv_profile_enable(0)
            tab = sorted(_lib.profile_table(), key=lambda r: -r["ms"])
        del pool, labels, ws, counts
        torch.cuda.empty_cache()
        return ms, tab

    # a step = one cv_ccl_label call over a quarter of cfg 4's 1024 masks; the 16-mask call of round 1 is reported beside it
    # (the scan and write passes each end in a tail of half a wave of long CTAs, which a small batch does not amortise)
    B, steps = 256, 6
    ms, tab = measure(B, steps, True)
    ms16, _ = measure(16, 20, False)
    alg = 5.0 * B * S * S
    gbps = alg * steps / (ms / 1e3) / 1e9
    gbps16 = 5.0 * 16 * S * S * 20 / (ms16 / 1e3) / 1e9
    rec = {"workload": f"cfg4: native-resolution 8-connected CCL (cv_ccl_label) on {B} dense 4096^2 masks per step",
           "value": B * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "roofline": {"kernel": "cv_ccl_label (all passes)", "bound": "hbm", "achieved": gbps, "peak": hbm_peak, "unit": "GB/s",
                        "frac": gbps / hbm_peak, "frac_of_8TBps_nominal": gbps / 8000.0, "traffic": None,
                        "algorithmic_bytes_per_px": 5},
           "batch16": {"ms_per_step": ms16 / 20, "achieved": gbps16, "frac": gbps16 / hbm_peak},
           "kernels": [{"name": r["name"], "launches": r["launches"], "ms_per_step": r["ms"] / steps} for r in tab[:6]],
           "gpu_launches": int(sum(r["launches"] for r in tab))}
    del base
    torch.cuda.empty_cache()
    if not a.no_cpu_baseline:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        with mp.get_context("spawn").Pool(cores) as p:
            per = p.map(_cpu_ccl_worker, [([900 + k, 901 + k], 2) for k in range(cores)])
        dt = max(per)
        rec["cpu_baseline"] = {"value": 2 * cores / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "GBps_at_5B_per_px": 2 * cores * 5.0 * S * S / dt / 1e9,
                               "sample": f"{2 * cores} masks, cv2.connectedComponents(8) one per core on {cores} processes, {dt:.2f} s"}
    return rec


# ---- cfg5: a fixed job sharded over the ranks, results gathered on the host of rank 0
def run_cfg5(ctx, a):
    import torch
    import torch.distributed as dist
    from circuitvision_b200 import sam2_infer, sharding, synth
    from circuitvision_b200.pipeline import CropPipeline
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    B, S, total = a.batch, a.size, a.total
    if total % B:
        total = (total // B) * B
    host_group = dist.new_group(backend="gloo") if world > 1 else None
    # every rank generates the SAME job description (seed = image index); pixel data of 16 distinct schematics is tiled
    uniq = 16
    proto = [synth.make_schematic(7000 + i, S, render_rgb=True) for i in range(uniq)]
    items = list(range(total))
    sam = sam2_infer.build_random_init(a.variant, device=dev, seed=0, max_batch=min(B, a.chunk),
                                       operand_dtype=torch.float16 if a.operands == "fp16" else torch.bfloat16)
    pipe = CropPipeline(sam, B, depth=2, want_images=False)
    n_pool = 4
    h_pool, b_pool = [], []
    for p in range(n_pool):
        idx = [(p * 5 + i) % uniq for i in range(B)]
        h_pool.append(torch.from_numpy(np.stack([proto[j][2] for j in idx])).pin_memory())
        b_pool.append([proto[j][1] for j in idx])
    launches = [0]

    def process(shard, rng):
        """One result per image of this rank's shard: the reference's node list (ids, component uids, contour) — yielded batch
        by batch, so that run_sharded can send finished pieces to rank 0 while the device works on the next batches."""
        nb = len(rng) // B
        k = 0

        def consume(res):
            return [res.nodes(b) for b in range(B)]

        for i in range(nb):
            if pipe._inflight == 2:
                yield consume(pipe.collect())
            pipe.submit(h_pool[k % n_pool], b_pool[k % n_pool])
            launches[0] += pipe.last_launches
            k += 1
        while pipe._inflight:
            yield consume(pipe.collect())

    stream_chunk = 4 * B  # results travel to rank 0 in pieces of four device batches

    def job():
        if world > 1:
            return sharding.run_sharded(items, process, group=host_group, dst=0, stream_chunk=stream_chunk)
        return sharding.run_sharded(items, process)

    # warm-up: W small jobs
    small = list(range(B * world * 2))
    for _ in range(a.warmup):
        if world > 1:
            sharding.run_sharded(small, process, group=host_group, dst=0, stream_chunk=stream_chunk)
        else:
            sharding.run_sharded(small, process)
    launches[0] = 0
    clocks = ClockSampler(ctx.local) if rank == 0 else None
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    steps = max(1, min(a.steps, 3))
    n_results = 0
    for _ in range(steps):
        full = job()
        if rank == 0:
            assert len(full) == total and all(r is not None for r in full)
            n_results += sum(len(r) for r in full)
    e1.record()
    ctx.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop() if clocks else None
    if rank != 0:
        return None
    value = total * steps / (ms / 1e3)
    per_rank = len(sharding.shard_range(total, world, 0))
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": a.warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": ("fp16" if a.operands == "fp16" else "bf16") + " tensor-core operands, fp32 accumulate + u8/int32",
        "data": "synthetic", "config": dict(make_config("cfg5", a), total_images=total, images_per_rank=per_rank),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(pipe.h2d_bytes * (per_rank // B)),
                "d2h_bytes_per_step": int(pipe.d2h_bytes * (per_rank // B)),
                "note": "the job IS end to end: pinned host crops in, per-image node lists built on every rank and gathered on rank 0 "
                        "(gloo object gathers of four device batches each, streamed under the remaining batches) inside the timed region; "
                        "a step is one pass over the whole job"},
        "gpu_launches": int(launches[0]), "clocks": clk, "wall_ms_per_step": wall_ms / steps,
        "nodes_gathered_per_step": n_results // steps,
    }


def main():
    a = _args()
    workload = a.workload
    if a.impl == "reference":
        return run_reference(a, workload)  # never touches the product library

    import torch
    import torch.distributed as dist
    from circuitvision_b200 import _lib

    ctx = Ctx()
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(ctx.local)
    ctx.dev = torch.device("cuda", ctx.local)
    if ctx.world > 1:
        dist.init_process_group("nccl", device_id=ctx.dev)
    ctx.lib = _lib.load()
    _lib.require_device(ctx.local)

    def barrier():
        torch.cuda.synchronize()
        if ctx.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx.barrier = barrier

    if workload == "cfg5":
        line = run_cfg5(ctx, a)
        if ctx.rank == 0:
            print(json.dumps(line))
        if ctx.world > 1:
            dist.destroy_process_group()
        return

    r = run_main_workload(ctx, a, workload)
    if ctx.rank != 0:
        if ctx.world > 1:
            dist.destroy_process_group()
        return
    use_sam2 = workload in ("pipeline", "sam2")
    use_nodes = workload in ("pipeline", "nodes", "nodes4096")
    B, S = a.batch, a.size
    roofline, kern_rows, top_shapes = (None, [], [])
    if r["table"]:
        roofline, kern_rows, top_shapes = make_roofline(r["table"], a.steps, workload, B, a.variant, S)

    # ---- CPU baseline on this box's host cores (bounded sample; N = 1 only)
    cpu = None
    if not a.no_cpu_baseline and ctx.world == 1:
        cores = os.cpu_count() or 1
        if use_sam2:
            cp = CpuPipeline(a.variant, S, cores, with_nodes=use_nodes)
            cp.step(1)
            n = 8
            dt = cp.step(n)
            cpu = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n} images of {S}^2: SAM2.1-{a.variant} fp32 restatement on {cores} torch threads"
                             + (", then the cv2 node analysis of each mask" if use_nodes else "") + f", {dt:.1f} s"}
            if use_nodes:
                ips_nodes, dtn, _ = cpu_nodes_throughput(128, S, cores, reps=2)
                cpu["nodes_only_images_per_s"] = ips_nodes
        else:
            n_cpu = 128 if S <= 1024 else 32
            ips_nodes, dt, _ = cpu_nodes_throughput(n_cpu, S, cores, reps=3)
            cpu = {"value": ips_nodes, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n_cpu} images of {S}^2, oracle node analysis (cv2) on {cores} worker processes, {dt:.1f} s"}

    extra = None
    if workload == "pipeline" and ctx.world == 1 and not a.no_extras and a.variant == "tiny" and B == 64:
        extra = {}
        for name, fn in (("cfg4_ccl", extra_cfg4_ccl), ("cfg4_nodes", extra_cfg4_nodes), ("cfg3", extra_cfg3)):
            t0 = time.perf_counter()
            try:
                extra[name] = fn(ctx, a)
            except Exception as e:  # a sub-record must not take the headline down with it
                extra[name] = {"error": f"{type(e).__name__}: {e}"}
            extra[name]["wall_s"] = round(time.perf_counter() - t0, 1)
            torch.cuda.empty_cache()

    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": r["total_ms"] / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": (("fp16" if a.operands == "fp16" else "bf16") + " tensor-core operands, fp32 accumulate + u8/int32") if use_sam2 else "u8/int32",
        "data": "synthetic", "config": make_config(workload, a), "input_gen_s": round(r["gen_s"], 2),
        "e2e": {"value": r["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                "steps": r["e2e_steps"], "ms_per_step": r["e2e_ms"] / r["e2e_steps"]},
        "gpu_launches": r["launches"], "clocks": r["clk"], "roofline": roofline, "cpu_baseline": cpu,
        "kernels": kern_rows[:12], "top_shapes": top_shapes, "extra": extra,
    }
    print(json.dumps(line))
    if ctx.world > 1:
        dist.destroy_process_group()


def _only_json_on_stdout():
    """Libraries print to the process's stdout on their own (NCCL's version banner on the first collective, for one):
    everything the run prints goes to stderr, and only the JSON line reaches the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _only_json_on_stdout()
    main()
