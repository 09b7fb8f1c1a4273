#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric; workload = configs[1]: 64 SAM2.1-tiny crops of 1024² +
node analysis of the resulting wire masks, per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

One step = one pass of the hot path over one batch of synthetic schematics (seeded generator,
circuitvision_b200/synth.py).  Inputs are resident in HBM for `value`; `e2e` goes through the public Python API
with pinned HOST buffers (H2D + D2H inside the timed region).  Images are independent: under torchrun every
rank processes its own batch (weak scaling, no data-path collective); time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SAM2.1 crops/sec + node-analysis images/sec at 1024^2"
UNIT = "images/s"


def _args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "pipeline", "nodes", "nodes4096", "sam2"])
    ap.add_argument("--batch", type=int, default=None,
                    help="images per GPU per step (default 64; 256 for nodes4096, whose border walks are latency-bound "
                         "chains that only a larger batch amortises)")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--variant", default="tiny")
    ap.add_argument("--chunk", type=int, default=64, help="crops per SAM 2.1 engine pass (workspace is sized for this many)")
    ap.add_argument("--operands", default="fp16", choices=["fp16", "bf16"],
                    help="16-bit tensor-core operand format of the SAM 2.1 path (DESIGN.md section 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernels with events in the timed region")
    return ap.parse_args()


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


# ------------------------------------------------------------------------------------------ CPU side
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


def cpu_sam2_throughput(n_images: int, size: int, variant: str, threads: int):
    """fp32 CPU restatement of SAM2ImageWrapper.forward (oracle/sam2_oracle.py) with all host threads."""
    import torch
    from oracle import sam2_oracle
    torch.set_num_threads(threads)
    model = sam2_oracle.build_oracle(variant, seed=0)
    from circuitvision_b200 import synth
    xs = []
    for i in range(n_images):
        _, _, rgb = synth.make_schematic(10_000 + i, size, render_rgb=True)
        xs.append(sam2_oracle.preprocess_rgb(rgb))
    with torch.no_grad():
        model(xs[0][None])  # warm
        t0 = time.perf_counter()
        for x in xs:
            model(x[None])
        dt = time.perf_counter() - t0
    return n_images / dt, dt


def have_sam2():
    try:
        from circuitvision_b200 import sam2_infer  # noqa: F401
        from circuitvision_b200 import _lib
        return hasattr(_lib.load(), "cv_sam2_forward")
    except Exception:
        return False


def run_reference(a, workload):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores
    (oracle port: /root/reference does not exist on the GPU box and its SAM2 dependency is not installable)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step_nodes = min(a.batch, 64) if a.size <= 1024 else min(a.batch, 16)
    sam2_s = None
    _, _, rep_s = cpu_nodes_throughput(per_step_nodes, a.size, cores, reps=a.warmup + a.steps)
    ms = [r * 1e3 for r in rep_s[a.warmup:]]
    sample = f"{per_step_nodes} images of {a.size}^2 per step, oracle node analysis (cv2) on {min(cores, per_step_nodes)} worker processes"
    ips_nodes = per_step_nodes / (np.mean(ms) / 1e3)
    value = ips_nodes
    if workload in ("pipeline", "sam2"):
        n_sam = 2
        ips_sam, _ = cpu_sam2_throughput(n_sam, a.size, a.variant, cores)
        sam2_s = 1.0 / ips_sam
        sample += f"; SAM2.1-{a.variant} fp32 restatement on {n_sam} images with {cores} torch threads"
        value = 1.0 / (1.0 / ips_sam + (1.0 / ips_nodes if workload == "pipeline" else 0.0))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
        "config": {"workload": workload_name(workload, a), "images_per_gpu_per_step": a.batch, "size": a.size},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "sam2_s_per_image": sam2_s, "nodes_images_per_s": ips_nodes},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(workload, a):
    if workload == "pipeline":
        return f"cfg2: SAM2.1-{a.variant} {a.batch} crops {a.size}^2 + node analysis of the {a.batch} masks, per GPU"
    if workload == "sam2":
        return f"SAM2.1-{a.variant} {a.batch} crops {a.size}^2 (segmentation only), per GPU"
    if workload == "nodes4096":
        return f"cfg4: node analysis + native CCL on {a.batch} dense 4096^2 masks, per GPU"
    return f"node analysis only: {a.batch} wire masks {a.size}^2 (SAM2 stage excluded), per GPU"


def _shape_row(r, steps, tens_peak, hbm_peak):
    """One launch group of the timed region: achieved rate and — for GEMM shapes, whose algorithmic bytes follow from the
    shape tag — its fraction of BOTH rooflines (most encoder GEMMs at K <= 192 are HBM-bound, not tensor-bound)."""
    import re
    row = {"name": r["name"], "launches": r["launches"], "ms_per_step": r["ms"] / steps,
           "rate_T_per_s": r["work"] / max(r["ms"], 1e-9) / 1e9}
    m = re.match(r"gemm M(\d+) N(\d+) K(\d+) bn\d+(.*)", r["name"])
    if m:
        M, N, K = (int(x) for x in m.groups()[:3])
        tag = m.group(4)
        byts = M * K * 2 + N * K * 2 + M * N * (2 if "->b" in tag else 4) + (M * N * 4 if "+r" in tag else 0)
        sec = r["ms"] / 1e3 / max(1, r["launches"])
        row["tensor_frac"] = 2.0 * M * N * K / sec / 1e12 / tens_peak
        row["hbm_frac"] = byts / sec / 1e9 / hbm_peak
        row["bound"] = "hbm" if row["hbm_frac"] > row["tensor_frac"] else "tensor"
    elif r["name"].startswith("attn"):
        row["tensor_frac"] = row["rate_T_per_s"] / tens_peak
    return row


# ------------------------------------------------------------------------------------------ GPU side
def main():
    a = _args()
    workload = a.workload
    if workload == "auto":
        workload = "pipeline" if have_sam2() else "nodes"
    if workload == "nodes4096":
        a.size = 4096
    if a.batch is None:
        a.batch = 256 if workload == "nodes4096" else 64
    if a.impl == "reference":
        return run_reference(a, workload)

    import torch
    import torch.distributed as dist
    from circuitvision_b200 import _lib, synth
    from circuitvision_b200.circuit_analyzer import CircuitAnalyzer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.require_device(local)

    B, S = a.batch, a.size
    n_pool = 4 if S <= 1024 else 2  # rotating input pool, larger than the 126 MB L2 in total
    use_sam2 = workload in ("pipeline", "sam2")
    use_nodes = workload in ("pipeline", "nodes", "nodes4096")

    # ---- synthetic inputs (distinct per rank), resident in HBM and mirrored in pinned host memory
    t_gen = time.perf_counter()
    uniq = min(B, 16)  # generator is host NumPy: build `uniq` distinct schematics per pool slot and tile them
    pool_masks, pool_boxes, pool_rgb = [], [], []
    for p in range(n_pool):
        seeds = [1_000_000 * rank + 1000 * p + i for i in range(uniq)]
        ms_, bx_, rgb_ = [], [], []
        for s in seeds:
            m, b, rgb = synth.make_schematic(s, S, render_rgb=use_sam2)
            ms_.append(m)
            bx_.append(b)
            rgb_.append(rgb)
        idx = [i % uniq for i in range(B)]
        pool_masks.append(np.stack([ms_[i] for i in idx]))
        pool_boxes.append([bx_[i] for i in idx])
        if use_sam2:
            pool_rgb.append(np.stack([rgb_[i] for i in idx]))
    gen_s = time.perf_counter() - t_gen

    A = CircuitAnalyzer(use_sam2=False, debug=False, device=local, render_debug_images=False)
    na = A._na()
    sam = None
    if use_sam2:
        from circuitvision_b200 import sam2_infer
        import torch as _t
        sam = sam2_infer.build_random_init(a.variant, device=dev, seed=0, max_batch=min(B, a.chunk),
                                           operand_dtype=_t.float16 if a.operands == "fp16" else _t.bfloat16)
    d_masks = [torch.from_numpy(m).to(dev) for m in pool_masks]
    h_masks = [torch.from_numpy(m).pin_memory() for m in pool_masks]
    d_boxes = [na.upload_boxes(bx, S, S) for bx in pool_boxes]
    d_rgb = [torch.from_numpy(x).to(dev) for x in pool_rgb] if use_sam2 else None
    h_rgb = [torch.from_numpy(x).pin_memory() for x in pool_rgb] if use_sam2 else None

    launches_per_step = [0]
    # Two stages of consecutive batches overlap on the device: the node analysis of batch i (small grids, its border
    # tracers are latency-bound) runs on a second stream while the SAM 2.1 forward of batch i+1 fills the SMs.
    s_main = torch.cuda.current_stream(dev)
    s_nodes = torch.cuda.Stream(dev) if (use_sam2 and use_nodes) else None
    ev_mask = torch.cuda.Event()

    def step_resident(i):
        p = i % n_pool
        n = 0
        masks = d_masks[p]
        if use_sam2:
            masks = sam.segment_batch_u8(d_rgb[p])  # [B,S,S] u8 {0,255} on device
            n += sam.last_launches
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if s_nodes is not None:
            s_main.wait_stream(s_nodes)  # the last batch's node analysis ends inside the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(a.warmup):
        step_resident(i)
    torch.cuda.synchronize()

    # Every launch of the timed region is bracketed by CUDA events on its own stream (cv_profile_*): the kernel table and
    # roofline.achieved come from the SAME K steps that give `value` (the brackets cost < 1 % of the step: 50.9 vs 51.4 ms
    # measured with and without them, profiles/README.md).
    prof = not a.no_profile
    lib.cv_profile_reset()
    lib.cv_profile_enable(1 if prof else 0)
    clocks = ClockSampler(local) if rank == 0 else None
    total_ms = timed(step_resident, a.steps)
    clk = clocks.stop() if clocks else None
    lib.cv_profile_enable(0)
    table = _lib.profile_table() if prof else []
    imgs = B * a.steps * world
    value = imgs / (total_ms / 1e3)

    # ---- e2e through the public API with host buffers
    h2d = d2h = 0
    pinned_out = {}

    def step_e2e(i):
        nonlocal h2d, d2h
        p = i % n_pool
        if use_sam2:
            x = h_rgb[p].to(dev, non_blocking=True)
            masks = sam.segment_batch_u8(x)
            bi = h_rgb[p].numel()
        else:
            masks = h_masks[p].to(dev, non_blocking=True)
            bi = h_masks[p].numel()
        bo = 0
        if use_nodes:
            r = na.analyze(masks, pool_boxes[p], grow=False)  # packs + uploads the boxes, runs, syncs on the tables
            host = r.tables_to_host()
            # the two result images go to pinned host buffers (allocated on the first, untimed call)
            if "emp" not in pinned_out:
                pinned_out["emp"] = torch.empty(tuple(r.emptied.shape), dtype=torch.uint8, pin_memory=True)
                pinned_out["enh"] = torch.empty(tuple(r.enhanced.shape), dtype=torch.uint8, pin_memory=True)
            pinned_out["emp"].copy_(r.emptied, non_blocking=True)
            pinned_out["enh"].copy_(r.enhanced, non_blocking=True)
            torch.cuda.synchronize()
            bo = sum(v.nbytes for v in host.values()) + pinned_out["emp"].numel() + pinned_out["enh"].numel()
            bi += sum(len(b) for b in pool_boxes[p]) * 48 + 4 * (B + 1)
        else:
            out = masks.cpu()
            bo = out.numel()
        h2d, d2h = bi, bo

    e2e_steps = max(2, min(a.steps, 5))
    if workload == "pipeline":
        # the batch API a user calls: pinned host crops in, host node tables + images out, copies software-pipelined
        from circuitvision_b200.pipeline import CropPipeline
        pipe = CropPipeline(sam, B, depth=2)
        n_nodes = [0]

        def run_e2e(steps):
            for i in range(steps):
                if pipe._inflight == 2:
                    n_nodes[0] += int(pipe.collect().nodes_table.tables_to_host()["results"]["n_nodes"].sum())
                p = i % n_pool
                pipe.submit(h_rgb[p], pool_boxes[p])
            while pipe._inflight:
                n_nodes[0] += int(pipe.collect().nodes_table.tables_to_host()["results"]["n_nodes"].sum())

        run_e2e(2)  # warm (allocates the pinned result buffers)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(e2e_steps)  # collect() waits for each batch's device->host copies
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    else:
        step_e2e(0)
        e2e_ms = timed(step_e2e, e2e_steps)
    e2e_value = B * e2e_steps * world / (e2e_ms / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event durations from the timed region)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    roofline = None
    kern_rows = []
    if table:
        # the library tags GEMM / attention / LayerNorm launches with their shapes; fold them back per kernel
        alias = {"gemm ": "k_gemm_tc", "attn_global ": "k_attn_global", "attn ": "k_attn_tc", "ln_rows ": "k_ln_rows"}
        folded = {}
        for r in table:
            name = next((v for k, v in alias.items() if r["name"].startswith(k)), r["name"])
            f = folded.setdefault(name, {"name": name, "launches": 0, "ms": 0.0, "work": 0.0})
            f["launches"] += r["launches"]
            f["ms"] += r["ms"]
            f["work"] += r["work"]
        shapes_all = list(table)
        shapes = sorted(table, key=lambda r: -r["ms"])[:16]
        table = list(folded.values())
        tot = sum(r["ms"] for r in table) or 1.0
        for r in sorted(table, key=lambda r: -r["ms"]):
            kern_rows.append({"name": r["name"], "launches": r["launches"], "ms_per_step": r["ms"] / a.steps,
                              "share": r["ms"] / tot, "work_per_launch": r["work"] / max(1, r["launches"])})
        top = max(table, key=lambda r: r["ms"])
        avg_s = top["ms"] / 1e3 / max(1, top["launches"])
        wpl = top["work"] / max(1, top["launches"])
        tensor = top["name"] in ("k_gemm_tc", "k_attn_tc", "k_attn_global")
        gemm_bytes = 0.0
        if top["name"] == "k_gemm_tc":
            # k_gemm_tc is one kernel over ~40 shapes: the K <= 192 ones (stages 1-2, most of its time) are bound by HBM,
            # the K >= 384 ones by the tensor pipe.  Both fractions are computed over ALL its launches (algorithmic bytes
            # from the shape tags) and the binding roofline — the larger fraction — is the one reported as `bound`.
            import re as _re
            for r in shapes_all:
                m = _re.match(r"gemm M(\d+) N(\d+) K(\d+) bn\d+(.*)", r["name"])
                if m:
                    M_, N_, K_ = (int(x) for x in m.groups()[:3])
                    tag = m.group(4)
                    gemm_bytes += r["launches"] * (M_ * K_ * 2 + N_ * K_ * 2 + M_ * N_ * (2 if "->b" in tag else 4) +
                                                   (M_ * N_ * 4 if "+r" in tag else 0))
        hbm_frac_gemm = (gemm_bytes / (top["ms"] / 1e3) / 1e9 / hbm_peak) if gemm_bytes else 0.0
        tens_frac = (wpl / avg_s / 1e12 / tens_peak) if tensor else 0.0
        if tensor and hbm_frac_gemm > tens_frac:
            ach = gemm_bytes / (top["ms"] / 1e3) / 1e9
            roofline = {"kernel": top["name"], "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": None,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                        "tensor_frac": tens_frac, "tensor_achieved_TFLOPs": wpl / avg_s / 1e12,
                        "note": "aggregate over all GEMM shapes of the step; algorithmic bytes = A + W + C (+ residual) per launch"}
        elif tensor:
            ach = wpl / avg_s / 1e12
            roofline = {"kernel": top["name"], "bound": "tensor", "achieved": ach, "peak": tens_peak, "unit": "TFLOP/s",
                        "frac": ach / tens_peak, "traffic": None,
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback"}
        else:
            ach = wpl / avg_s / 1e9
            roofline = {"kernel": top["name"], "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": None,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"}
        roofline["avg_launch_us"] = avg_s * 1e6
        roofline["share_of_step"] = top["ms"] / tot
        # DRAM bytes per launch of that kernel from the committed ncu pass of this same command (profiles/), when the
        # workload matches the one that was profiled
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic_pipeline_b64.json")))
            k = tr["kernels"].get(top["name"])
            if k and workload == "pipeline" and B == 64 and a.variant == "tiny" and S == 1024:
                roofline["traffic"] = k["traffic_bytes_per_launch"]
                roofline["traffic_source"] = "profiles/r1_ncu_traffic_pipeline_b64.json (ncu dram__bytes_read+write, per launch)"
                roofline["traffic_GBps"] = k["traffic_bytes_per_launch"] / avg_s / 1e9
        except Exception:
            pass

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if not a.no_cpu_baseline and world == 1:  # the CPU leg runs at N = 1 only (rank 0 is the only rank there)
        cores = os.cpu_count() or 1
        n_cpu = 128 if S <= 1024 else 32
        ips_nodes, dt, _ = cpu_nodes_throughput(n_cpu, S, cores, reps=3)
        sample = f"{n_cpu} images of {S}^2, oracle node analysis (cv2) on {cores} worker processes, {dt:.1f} s"
        v = ips_nodes
        extra = {"nodes_images_per_s": ips_nodes}
        if use_sam2:
            ips_sam, dts = cpu_sam2_throughput(3, S, a.variant, cores)
            sample += f"; SAM2.1-{a.variant} fp32 restatement on 3 images, {cores} torch threads, {dts:.1f} s"
            extra["sam2_images_per_s"] = ips_sam
            v = 1.0 / (1.0 / ips_sam + (1.0 / ips_nodes if use_nodes else 0.0))
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, **extra}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": (("fp16" if a.operands == "fp16" else "bf16") + " tensor-core operands, fp32 accumulate + u8/int32") if use_sam2 else "u8/int32",
        "data": "synthetic",
        "config": {"workload": workload_name(workload, a), "images_per_gpu_per_step": B, "size": S,
                   "l2_policy": f"inputs rotate over a pool of {n_pool} batches ({n_pool * B * S * S * (13 if use_sam2 else 1) >> 20} MiB) larger than L2",
                   "sharding": "image-wise, no collective", "input_gen_s": round(gen_s, 2),
                   "stage_overlap": "node analysis of batch i on a second stream under the SAM 2.1 forward of batch i+1" if s_nodes is not None else "none"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps},
        "gpu_launches": int(launches_per_step[0] * a.steps),
        "clocks": clk,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "kernels": kern_rows[:12],
        "top_shapes": [_shape_row(r, a.steps, tens_peak, hbm_peak) for r in shapes] if table else [],
    }
    print(json.dumps(line))
    if world > 1:
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
