#!/usr/bin/env python
"""Benchmark of the volumetric hot path (BASELINE.json metric: voxels/sec, 256x256x32 grid,
20-class SSC) -- one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2]

A step = one pass of the volumetric forward (stereo cost volume + aggregation, depth_net, MIE,
lift (x) splat, 3-D encoder + neck, occupancy head, x2 trilinear + argmax) over one synthetic
stereo pair per GPU (features after the 2-D image backbone).
  value  : inputs resident in HBM, device-timed (CUDA events), max over ranks.
  e2e    : the same call with HOST (pinned) feature buffers: H2D of the two feature maps and D2H of
           the uint8 label volume inside the timed region.
  roofline: the dominant kernel (occupancy-head conv 384->192 k3 on the 128x128x16 grid, 1.04 TFLOP
           in one launch) timed alone with CUDA events; FLOP/s against the measured tensor peak.
  cpu_baseline / --impl reference: the oracle port (oracle/restatement.py, PyTorch CPU fp32, all
           host threads) on one forward of the same workload.
Multi-GPU: the path shards by sample (one stereo pair per rank, no data-path collective) -> weak
scaling; NCCL is used for the barrier and the max-over-ranks reduction only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# roofline.traffic of the dominant kernel comes from the ncu --set full capture committed this round; tools/prof_kernels.sh
# writes the two DRAM counters of that capture into profiles/r02_head_conv_traffic.json next to the raw page
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r02_head_conv_traffic.json")          # {"tf32": {...}, "f16": {...}}

VOXELS = {"config1": 128 * 128 * 16, "config2": 256 * 256 * 32, "config0": 64 * 64 * 8, "tiny": 32 * 32 * 8,
          "config4": 512 * 512 * 64}


def workload_config(workload: str, B: int = 1) -> str:
    """The ONE description of the workload both arms print (so the driver sees the same config on both lines)."""
    from stereoscene_b200 import presets
    occ = presets.WORKLOADS[workload][0]
    return (f"{workload}: synthetic 1242x375 stereo -> 384x1280 -> 48x160x112 frustum -> {'x'.join(map(str, occ))} logits, "
            f"20 classes, B={B} stereo pair per step per worker (features after the 2-D image backbone)")


def policy_text(ops, name: str) -> str:
    names = {ops.SS_MATH_TF32: "tf32", ops.SS_MATH_TF32X3: "tf32x3 (compensated, tcgen05)", ops.SS_MATH_3XTF32: "3xtf32 (compensated, mma.sync)",
             ops.SS_MATH_F16: "f16 (fp16 operands = TF32's 11-bit significand, fp32 accumulate; TF32 where a kernel has no fp16 path)"}
    pol = ops.MATH_POLICIES[name]
    return ", ".join(f"{g}={names[m]}" for g, m in ((g, pol.get(g, ops.SS_MATH_TF32)) for g in ("stereo", "depthnet", "mie", "voxel")))


def measure_tf32_peak(dev) -> float:
    """cuBLAS TF32 GEMM throughput (TFLOP/s) on this GPU, measured here: the denominator of the tensor roofline of a
    kernel that multiplies in TF32 (MEASURED_PEAKS.json only holds the bf16 figure)."""
    n = 8192
    a = torch.randn((n, n), device=dev)
    b = torch.randn((n, n), device=dev)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(3):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 10 * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return best


def golden_parity(workload: str, out: dict, seed: int):
    """Error of the benchmarked forward's logits against the committed fixture of the REFERENCE's own forward on the same
    seeded inputs (tests/golden/golden_<workload>.npz, written by oracle/make_golden_full.py): max|d|/max|ref| and
    rms(d)/rms(ref) on the fixture's strided sample, label agreement on the sample.  None if there is no fixture."""
    import numpy as np
    js = os.path.join(ROOT, "tests", "golden", f"golden_{workload}.json")
    if not os.path.exists(js):
        return None
    with open(js) as f:
        meta = json.load(f)
    if meta["seed"] != seed:
        return None
    gold = np.load(os.path.join(ROOT, "tests", "golden", f"golden_{workload}.npz"))
    res = {}
    for key, t in (("logits_up", out["output_voxels"]), ("logits", out["logits_lowres"]), ("depth_prob", out["depth"])):
        sl = tuple(slice(*x) for x in meta["samplers"][key])
        d = t[:1][sl].detach().cpu().double().numpy() - gold[key].astype(np.float64)
        res[key] = {"max_rel": float(np.abs(d).max() / meta["stats"][key]["absmax"]),
                    "rms_rel": float(np.sqrt((d * d).mean()) / meta["stats"][key]["rms"])}
    sl = tuple(slice(*x) for x in meta["samplers"]["logits_up"])
    lab = out["output_voxels"][:1][sl].argmax(1).cpu().numpy().astype(np.uint8)
    res["label_agreement"] = float((lab == gold["labels_up_sample"]).mean())
    res["against"] = f"tests/golden/golden_{workload}.npz (reference's own forward, seed {seed})"
    return res


def image_encoder_extra(model, mc, dev, left, right, calib, occ, seed):
    """SURVEY.md section 8 row N2, outside the metric (BASELINE.json's path starts from backbone features): the 2-D image
    encoder (EfficientNet-B7 + SECONDFPN) on one stereo pair -- time alone, time of the whole detector from images, parity of
    its output features against the reference's own efficientnet.py (tests/golden/golden_image_full.npz, same seed)."""
    import json as _json
    import numpy as np
    from stereoscene_b200 import cabi, presets, synth
    from stereoscene_b200.registry import build_backbone, build_neck
    cfg = presets.model_config("config2", image_encoder=True)["model"]
    enc = torch.nn.ModuleDict(dict(img_backbone=build_backbone(cfg["img_backbone"]), img_neck=build_neck(cfg["img_neck"])))
    synth.randomize_weights_(enc, seed)
    enc = enc.to(dev).eval()
    model.img_backbone, model.img_neck = enc["img_backbone"], enc["img_neck"]
    il, ir = synth.stereo_images(1, mc["input_size"], seed=seed, device=dev)
    pair = torch.cat([il, ir], 0)

    def timed_graph(fn, n=10):
        with torch.no_grad():
            fn(); fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    out = {"model": "CustomEfficientNet-b7 + SECONDFPN (stereoscene.py:59-74), 2 x 3 x %d x %d" % tuple(mc["input_size"]),
           "in_metric": False}
    with torch.no_grad():
        c0 = cabi.launch_count()
        feat = model.image_encoder_cl(pair)
        out["gpu_launches"] = int(cabi.launch_count() - c0)
    out["ms_per_pair"] = timed_graph(lambda: model.image_encoder_cl(pair))
    out["from_images_ms_per_step"] = timed_graph(lambda: model.forward_images(il, ir, left, right, calib, occ_size=occ, want_labels=True))
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden")
    meta_p = os.path.join(gdir, "golden_image_full.json")
    if os.path.exists(meta_p) and tuple(mc["input_size"]) == (384, 1280):
        meta = _json.load(open(meta_p))
        if meta["seed"] == seed:
            gold = np.load(os.path.join(gdir, "golden_image_full.npz"))["img_feat"].astype(np.float64)
            sl = tuple(slice(*x) for x in meta["samplers"]["img_feat"])
            got = feat.squeeze(1).permute(0, 3, 1, 2)[sl].double().cpu().numpy()
            d = got - gold
            st = meta["stats"]["img_feat"]
            out["parity_img_feat"] = {"max_rel": float(np.abs(d).max() / st["absmax"]), "rms_rel": float(np.sqrt((d * d).mean()) / st["rms"]),
                                      "against": "tests/golden/golden_image_full.npz (reference's own efficientnet.py, seed 0)"}
    model.img_backbone = model.img_neck = None
    return out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(VOXELS))
    ap.add_argument("--math", default="mixed", choices=["mixed", "mixed_tf32stereo", "mixed16", "tf32", "f16", "tf32x3", "3xtf32"],
                    help="per-stage math policy (stereoscene_b200.ops.MATH_POLICIES); 'mixed' = plain TF32 tensor-core math with "
                         "the error-compensated TF32x3 mode on depth_net and the MIE block: the cheapest policy whose logits "
                         "are within 1e-3 of the reference's forward")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the extra timing + error lines of the other math policies")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the forward in a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0)
    ap.add_argument("--pairs-per-step", type=int, default=1, help="stereo pairs per GPU per step (batch of one forward); with --shard x: pairs per step of the whole job")
    ap.add_argument("--xshard-comm", default="peer", choices=["peer", "nccl"],
                    help="--shard x: halo exchange / statistics all-reduce as NVLink peer-memory kernels (default) or torch.distributed calls")
    ap.add_argument("--shard", default="sample", choices=["sample", "x"],
                    help="sample: one stereo pair per rank, no data-path collective (throughput mode, default); x: the X-slab sharded "
                         "latency mode of BASELINE.json configs[3] (stereoscene_b200/xshard.py): one all-gather of depth_prob||img_feat at "
                         "the MIE boundary, halo exchange + GroupNorm all-reduce in the voxel stack; total work fixed -> strong scaling")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU baseline = the oracle port (the one place bench.py may execute oracle/)
# ------------------------------------------------------------------------------------------
def cpu_forward_factory(workload: str, seed: int = 0):
    from oracle import restatement as O
    from stereoscene_b200 import presets, synth
    model, mc = presets.build(workload)
    synth.randomize_weights_(model, seed)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    del model
    xl, xr = synth.stereo_features(1, mc["input_size"], 8, seed=seed)
    left, right, calib = synth.kitti_calibration(1, mc["input_size"])
    gc = mc["model"]["img_view_transformer"]["grid_config"]

    def run():
        with torch.no_grad():
            return O.volumetric_forward(sd, xl, xr, left, right, calib, gc, mc["input_size"], mc["occ_size"])
    return run


def time_cpu(workload: str, steps: int, warmup: int, budget_s: float):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run = cpu_forward_factory(workload)
    t0 = time.perf_counter()
    run()                                           # first pass doubles as warm-up and cost estimate
    est = time.perf_counter() - t0
    done_warm = 1
    while done_warm < warmup and (done_warm + steps) * est < budget_s:
        run(); done_warm += 1
    n = max(1, min(steps, int((budget_s - done_warm * est) / max(est, 1e-6))))
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
    per = statistics.median(ts)
    return dict(value=VOXELS[workload] / per, unit="voxels/s", cores=cores, kind="port",
                sample=f"{n} timed forward(s) of {workload} (B=1, fp32, torch {torch.__version__} CPU, "
                       f"oracle/restatement.py), {per:.2f} s each", seconds_per_forward=per, steps_run=n)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = time_cpu(args.workload, args.steps, args.warmup, args.cpu_budget_s)
    line = {
        "impl": "reference", "metric": "voxels/sec", "value": cb["value"], "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_forward"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_config(args.workload, 1),
                   "note": "reference has no native/GPU-independent build; its CPU path = PyTorch CPU fp32 "
                           "(oracle port pinned to the reference's own forward), all host threads"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "steps_run": cb["steps_run"], "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_xshard(args):
    """--shard x: latency of a FIXED job (args.pairs_per_step stereo pairs per step, default 1) on N ranks."""
    import torch.distributed as dist
    from stereoscene_b200 import cabi, ops, presets, sharding, synth, xshard
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    cabi.load()
    ops.set_math_policy(args.math)
    model, mc = presets.build(args.workload)
    synth.randomize_weights_(model, 0)
    model = model.to(dev).eval()
    P = max(1, args.pairs_per_step)
    counts = [len(range(r, P, world)) for r in range(world)]
    b = max(counts)
    xl, xr = synth.stereo_features(max(b, 1), mc["input_size"], 8, seed=rank, device=dev)      # only the owner's pair is used when P = 1
    left, right, calib = synth.kitti_calibration(max(b, 1), mc["input_size"], device=dev)
    occ = mc["occ_size"]
    pipe = xshard.XShardedPipeline(model, world, rank, peer_memory=(args.xshard_comm == "peer"))
    pipe.use_graph = not args.no_graph and args.xshard_comm == "peer"

    def step():
        with torch.no_grad():
            return pipe.forward(xl, xr, left, right, calib, occ, counts)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n0 = cabi.launch_count()
    outs = step()
    torch.cuda.synchronize()
    launches = cabi.launch_count() - n0
    coll = dict(pipe.path.collectives)
    gathered = pipe.gathered_bytes
    for _ in range(max(4, args.warmup - 1)):          # >= 4: the voxel graph is captured on the third call
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = float(sharding.max_over_ranks(torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev))[0])
    if rank == 0:
        line = {
            "metric": "voxels/sec", "value": VOXELS[args.workload] * P * args.steps / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": workload_config(args.workload, P), "math": args.math, "math_policy": policy_text(ops, args.math),
                       "cuda_graph": "voxel-space path (frustum stages eager)" if pipe.use_graph else False,
                       "frustum_split": bool(pipe.split_frustum and world > 1 and P == 1),
                       "parallelism": f"x-slab x{world}: frustum stages of pair s on rank s % {world}, ONE all-gather of "
                       "depth_prob||img_feat at the MIE boundary, halo send/recv + GroupNorm all-reduce in the voxel stack",
                       "l2": "no flush: one step streams > 9 GB of activations, >> 126 MB L2"},
            "shard": {"mode": "x-slab", "comm": "NVLink peer-memory kernels (csrc/peer.cu)" if args.xshard_comm == "peer" else "torch.distributed (NCCL)", "pairs_per_step": P, "slab_planes": pipe.plan.xs, "per_sample": coll,
                      "allgather_bytes_per_step": gathered, "output": f"rank r holds planes [{2 * pipe.plan.xs} r, {2 * pipe.plan.xs} (r+1)) of every label volume"},
            "gpu_launches": int(launches * args.steps), "gpu_launches_per_step": int(launches), "clocks": clocks,
            "latency_ms": ms / args.steps,
        }
        if xshard._TIMING:
            line["shard"]["serialised_ms_per_step"] = {k: v / (args.steps + max(2, args.warmup - 1) + 1) for k, v in xshard.TIMES.items()}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_ours(args):
    if args.shard == "x":
        return run_xshard(args)
    import torch.distributed as dist
    from stereoscene_b200 import cabi, ops, presets, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout = the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    cabi.load()
    ops.set_math_policy(args.math)

    seed = 0
    model, mc = presets.build(args.workload)
    synth.randomize_weights_(model, seed)
    model = model.to(dev).eval()
    B = max(1, args.pairs_per_step)                         # stereo pairs per rank per step (weak scaling)
    xl_h, xr_h = synth.stereo_features(B, mc["input_size"], 8, seed=seed + rank, pin=True)
    left, right, calib = synth.kitti_calibration(B, mc["input_size"], device=dev)
    xl_d, xr_d = xl_h.to(dev), xr_h.to(dev)
    occ = mc["occ_size"]
    labels_h = torch.empty((B, *occ), dtype=torch.uint8).pin_memory()

    def forward(xl, xr):
        return model.forward_features(xl, xr, left, right, calib, occ_size=occ, want_labels=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also fills the splat-index / packed-weight caches) and launch census -------
    with torch.no_grad():
        forward(xl_d, xr_d)
        torch.cuda.synchronize()
        n0 = cabi.launch_count()
        out = forward(xl_d, xr_d)
        torch.cuda.synchronize()
        launches_per_step = cabi.launch_count() - n0
        for _ in range(max(0, args.warmup - 2)):
            forward(xl_d, xr_d)
    torch.cuda.synchronize()

    # ---- optional CUDA graph of the device-resident step ---------------------------------------
    graph = None
    if not args.no_graph:
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s), torch.no_grad():
                forward(xl_d, xr_d)                       # warm this stream's arena / caches
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g):
                gout = forward(xl_d, xr_d)
            g.replay()
            torch.cuda.synchronize()
            if not torch.equal(gout["labels"], out["labels"]) and \
                    float((gout["labels"] != out["labels"]).float().mean()) > 1e-3:
                raise RuntimeError("graph replay disagrees with eager run")
            graph = g
        except Exception as e:                             # fall back to eager launches, say so
            if rank == 0:
                print(f"[bench] CUDA graph capture unavailable ({type(e).__name__}: {e}); timing eager launches",
                      file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def step_device():
        if graph is not None:
            graph.replay()
        else:
            with torch.no_grad():
                forward(xl_d, xr_d)

    for _ in range(3):
        step_device()
    # ---- timed region: device-resident inputs ----------------------------------------------------
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)

    # ---- timed region: end to end from host buffers ---------------------------------------------
    # public serving API (stereoscene_b200.runtime.VolumetricEngine): pinned host features -> H2D on a copy
    # stream -> CUDA-graph forward -> D2H of the uint8 label volume, double-buffered so the copies of
    # neighbouring pairs overlap the compute.  Falls back to the eager public call if graphs are disabled.
    e2e_mode = "engine(graph+overlapped copies)"
    eng = None
    if not args.no_graph:
        try:
            from stereoscene_b200.runtime import VolumetricEngine
            eng = VolumetricEngine(model, left, right, calib, occ, tuple(xl_h.shape), device=dev)
        except Exception as e:
            if rank == 0:
                print(f"[bench] VolumetricEngine unavailable ({type(e).__name__}: {e}); e2e uses eager calls", file=sys.stderr)
            eng = None

    def step_e2e_eager():
        with torch.no_grad():
            a = xl_h.to(dev, non_blocking=True)
            b = xr_h.to(dev, non_blocking=True)
            o = forward(a, b)
            labels_h.copy_(o["labels"], non_blocking=True)

    def run_e2e(n):
        if eng is not None:
            for _ in eng.stream((xl_h, xr_h) for _ in range(n)):
                pass
            cur = torch.cuda.current_stream()
            for ev in eng.d2h_done:
                cur.wait_event(ev)
        else:
            for _ in range(n):
                step_e2e_eager()

    if eng is None:
        e2e_mode = "eager forward_features()"
    run_e2e(3)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    if eng is not None:                      # make the engine's streams start after the start event
        eng.copy_stream.wait_event(f0)
        eng.d2h_stream.wait_event(f0)
        eng.compute_stream.wait_event(f0)
    run_e2e(args.steps)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop()
    if eng is not None and rank == 0:
        # the engine's labels must be the eager forward's labels
        same = float((eng.infer(xl_h, xr_h).to(dev) != out["labels"]).float().mean())
        if same > 1e-3:
            raise SystemExit(f"engine output disagrees with the eager forward ({same:.3e} of labels differ)")

    # ---- per-stage split under the reference's stage names (bevdepth_occupancy.py:63-79, 103-122) ----
    stage_ms = stage_breakdown(model, xl_d, xr_d, left, right, calib, occ)

    from stereoscene_b200 import sharding
    t = sharding.max_over_ranks(torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev))
    ms_dev, ms_e2e = float(t[0]), float(t[1])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    value = sharding.whole_job_voxels_per_s(VOXELS[args.workload], B, world, args.steps, ms_dev)
    e2e = sharding.whole_job_voxels_per_s(VOXELS[args.workload], B, world, args.steps, ms_e2e)

    # ---- roofline of the dominant kernel, timed alone (CUDA events on the launching stream) -------
    roof, kernels = dominant_kernel_roofline(model, mc, dev, pk)

    # ---- parity of the benchmarked mode (and, at N=1, time + parity of the other math policies) -------------------
    parity = None
    if B == 1:
        with torch.no_grad():
            parity = golden_parity(args.workload, forward(xl_d, xr_d), seed)
    other_modes = {}
    if world == 1 and not args.no_other_modes and args.workload in ("config1", "config2"):
        for name in ("tf32", "mixed", "tf32x3"):
            if name == args.math:
                continue
            ops.set_math_policy(name)
            try:
                with torch.no_grad():
                    o = forward(xl_d, xr_d)
                    torch.cuda.synchronize()
                    g2 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g2):
                        forward(xl_d, xr_d)
                for _ in range(3):
                    g2.replay()
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(10):
                    g2.replay()
                a1.record()
                torch.cuda.synchronize()
                msm = a0.elapsed_time(a1) / 10
                other_modes[name] = {"ms_per_step": msm, "value": VOXELS[args.workload] * B / (msm * 1e-3),
                                     "math_policy": policy_text(ops, name), "parity": golden_parity(args.workload, o, seed)}
                del g2
            finally:
                ops.set_math_policy(args.math)

    image_enc = None
    if world == 1 and B == 1 and not args.no_other_modes:
        image_enc = image_encoder_extra(model, mc, dev, left, right, calib, occ, seed)

    # the cuBLAS TF32 GEMM peak is measured LAST: 60 back-to-back 8192^3 GEMMs push the chip to its power cap and would
    # slow everything timed after them
    time.sleep(1.0)
    pk["tf32_tflops_cublas"] = measure_tf32_peak(dev)
    if roof["peak"] is None:                       # TF32 kernel: against the TF32 GEMM peak measured just now
        roof["peak"] = pk["tf32_tflops_cublas"]
        roof["frac"] = roof["achieved"] / roof["peak"]
    roof["tf32_gemm_tflops_measured_in_run"] = pk["tf32_tflops_cublas"]

    line = {
        "metric": "voxels/sec", "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
        "config": {"workload": workload_config(args.workload, B),
                   "storage": "fp32 channels-last", "math": args.math, "math_policy": policy_text(ops, args.math),
                   "cuda_graph": graph is not None,
                   "l2": "no flush: one step streams > 9 GB of activations, >> 126 MB L2",
                   "parallelism": f"sample-sharded x{world} (no data-path collective)"},
        "e2e": {"value": e2e, "unit": "voxels/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(xl_h.numel() * 4 + xr_h.numel() * 4), "d2h_bytes_per_step": int(labels_h.numel())},
        "e2e_mode": e2e_mode,
        "stage_ms": stage_ms,
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "clocks": clocks,
        "roofline": roof,
        "kernels": kernels,
        "peaks": pk,
        "parity": parity,
        "other_math_policies": other_modes,
        "image_encoder": image_enc,
    }
    if args.workload == "config2":
        # whole-step algorithmic totals of SURVEY.md section 8(a) (B=1, fp32 storage, incl. depth_net): 3,986 GFLOP and 9,139 MB
        # of compulsory traffic per stereo pair -- the step is tensor-bound in TF32 (floor 3986/826 = 4.8 ms vs 1.4 ms of HBM)
        sec = ms_dev / args.steps * 1e-3 / B
        line["step_roofline"] = {"algorithmic_gflop": 3986.0, "algorithmic_mb": 9139.0, "tflops": 3986.0 / sec / 1e3,
                                 "tflops_frac_of_bf16_peak": 3986.0 / sec / 1e3 / pk["tflops"], "hbm_gbs": 9139.0 / sec / 1e3,
                                 "hbm_frac": 9139.0 / sec / 1e3 / pk["hbm_gbs"], "bound": "tensor (TF32 = half the bf16 peak)"}
    if world == 1 and not args.no_cpu_baseline:
        cb = time_cpu(args.workload, 1, 1, 90.0)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def stage_breakdown(model, xl, xr, left, right, calib, occ, iters=7):
    """Eager per-stage device times (CUDA events, median over `iters` runs after one warm-up), reference stage names."""
    from stereoscene_b200 import ops
    vt = model.img_view_transformer
    keys = ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    samples = [[] for _ in range(4)]
    with torch.no_grad():
        for it in range(iters + 1):
            ml = vt.get_mlp_input(*[left[k] for k in keys]); mr = vt.get_mlp_input(*[right[k] for k in keys])
            torch.cuda.synchronize()
            ev[0].record()
            bev, _ = vt([xl] + [left[k] for k in keys] + [ml] + [xr] + [right[k] for k in keys] + [mr] + [calib, None, None])
            ev[1].record()
            with ops.math_scope("voxel.encoder"):
                levels = model.img_bev_encoder_backbone.forward_vol(ops.Vol(bev.permute(0, 2, 3, 4, 1)))
            ev[2].record()
            with ops.math_scope("voxel.neck"):
                neck = model.img_bev_encoder_neck.forward_vol(levels)
            ev[3].record()
            with ops.math_scope("voxel.head"):
                logits = model.pts_bbox_head.forward_voxel_vol([neck])[0]
            ops.trilinear(logits, occ, want_labels=True)
            ev[4].record()
            torch.cuda.synchronize()
            if it:
                for i in range(4):
                    samples[i].append(ev[i].elapsed_time(ev[i + 1]))
    names = ("view_transformer", "bev_encoder", "bev_neck", "occ_head+upsample")
    return {n: statistics.median(a) for n, a in zip(names, samples)}


def _time_launches(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def dominant_kernel_roofline(model, mc, dev, pk):
    """Times individual kernels of the step alone with CUDA events (inputs are > L2 or re-streamed,
    see notes) and relates ALGORITHMIC flops/bytes to the measured peaks."""
    from stereoscene_b200 import ops
    from stereoscene_b200.ops import Vol
    vt = model.img_view_transformer
    nx = [int(round(float(v))) for v in vt.nx.detach().cpu()]
    kernels = []
    torch.manual_seed(0)
    # (1) head conv 384->192 k3 on the LSS grid: the single largest launch of the step
    head = model.pts_bbox_head.occ_convs[0][0]
    x = torch.randn((1, nx[0], nx[1], nx[2], head.in_channels), device=dev)
    sc = torch.rand((1, head.in_channels), device=dev) + 0.5
    sh = torch.randn((1, head.in_channels), device=dev) * 0.1
    y = torch.empty((1, nx[0], nx[1], nx[2], head.out_channels), device=dev)
    vin = Vol(x, sc, sh, ops.SS_ACT_RELU)

    with ops.math_scope("voxel.head"):
        head_mode = ops.default_math()                    # the mode the active policy runs this layer in
    if head_mode not in (ops.SS_MATH_TF32, ops.SS_MATH_F16):
        head_mode = ops.SS_MATH_TF32

    def head_conv():
        ops.arena(dev).reset()
        ops.conv(vin, head, out=y, want_stats=True, math_mode=head_mode)
    t = _time_launches(head_conv)
    V = nx[0] * nx[1] * nx[2]
    flops = 2.0 * V * 27 * head.in_channels * head.out_channels
    ach = flops / t / 1e12
    traffic, traffic_src = None, "no ncu capture of this round found (profiles/r02_head_conv_traffic.json)"
    if os.path.exists(TRAFFIC_JSON):
        with open(TRAFFIC_JSON) as f:
            tj = json.load(f).get("f16" if head_mode == ops.SS_MATH_F16 else "tf32")
        if tj:
            traffic, traffic_src = tj["dram_bytes_read"] + tj["dram_bytes_write"], tj["source"]
    f16 = head_mode == ops.SS_MATH_F16
    roof = {"bound": "tensor",
            "kernel": ("conv_halo_kernel<192, 2> (OccHead conv 384->192 k3 on the 128x128x16 grid; TMA halo planes, fp32 planes rewritten "
                       "in place as fp16 by the fix-up warps, tcgen05.mma kind::f16 with SWIZZLE_64B weight tiles, TMEM accumulators)") if f16 else
                      ("conv_halo_kernel<192, 0> (OccHead conv 384->192 k3 on the 128x128x16 grid; TMA halo planes + tcgen05.mma kind::tf32, "
                       "TMEM accumulators)"),
            "math": "fp16 operands (11-bit significand = TF32's), fp32 accumulate" if f16 else "tf32",
            "achieved": ach, "peak": pk["tflops"] if f16 else None, "unit": "TFLOP/s", "frac": ach / pk["tflops"] if f16 else None,
            "peak_bf16": pk["tflops"], "frac_of_bf16_peak": ach / pk["tflops"],
            "traffic": traffic, "traffic_source": traffic_src,
            "launch_ms": t * 1e3, "algorithmic_flops": flops,
            "algorithmic_bytes": (x.numel() + y.numel()) * 4.0,
            "peak_source": (f"peak = MEASURED_PEAKS.json dense bf16 burst ({pk['source']}): kind::f16 and kind::bf16 share the 16-bit tensor rate"
                            if f16 else "peak = cuBLAS TF32 GEMM (8192^3) measured at the end of this run on this GPU (the kernel multiplies in "
                                        f"TF32); peak_bf16 = MEASURED_PEAKS.json bf16 burst ({pk['source']}) beside it")}
    kernels.append({"name": "occ_head conv3d 384->192 k3", "bound": "tensor", "ms": t * 1e3, "tflops": ach,
                    "frac": ach / pk["tflops"]})
    # (2) full-res 32->32 k3 frustum conv (HBM-bound in the algorithmic accounting: in + out once)
    D, H, W = vt.D, vt.frustum.shape[1], vt.frustum.shape[2]
    c32 = vt.stereo_volume_net.dres0[0][0]
    xv = torch.randn((1, D, H, W, 32), device=dev)
    yv = torch.empty_like(xv)

    def frustum_conv():
        ops.arena(dev).reset()
        ops.conv(Vol(xv), c32, out=yv, want_stats=True, math_mode=ops.SS_MATH_TF32)
    t = _time_launches(frustum_conv)
    byts = 2.0 * xv.numel() * 4
    kernels.append({"name": "frustum conv3d 32->32 k3 (112x48x160)", "bound": "hbm", "ms": t * 1e3,
                    "gbs": byts / t / 1e9, "frac": byts / t / 1e9 / pk["hbm_gbs"],
                    "tflops": 2.0 * D * H * W * 27 * 32 * 32 / t / 1e12})

    def frustum_conv_x3():
        ops.arena(dev).reset()
        ops.conv(Vol(xv), c32, out=yv, want_stats=True, math_mode=ops.SS_MATH_TF32X3)
    t = _time_launches(frustum_conv_x3)
    kernels.append({"name": "frustum conv3d 32->32 k3, compensated TF32x3 (3 accumulating launches)", "bound": "hbm", "ms": t * 1e3,
                    "gbs": byts / t / 1e9, "frac": byts / t / 1e9 / pk["hbm_gbs"],
                    "tflops": 2.0 * D * H * W * 27 * 32 * 32 / t / 1e12})
    # (3) gwc + warp (write-bound)
    fea = torch.randn((2, 1, H, W, 64), device=dev)
    cal = torch.full((1, 1), 380.3, device=dev)
    t = _time_launches(lambda: ops.gwc_warp(fea, cal, D, 32))
    byts = (fea.numel() + D * H * W * 32) * 4.0
    kernels.append({"name": "gwc_warp", "bound": "hbm", "ms": t * 1e3, "gbs": byts / t / 1e9,
                    "frac": byts / t / 1e9 / pk["hbm_gbs"]})
    # (4) lift (x) splat (write-bound)
    left, _, _ = __import__("stereoscene_b200.synth", fromlist=["x"]).kitti_calibration(1, mc["input_size"], device=dev)
    idx = vt.splat_index(*[left[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")])
    dp = torch.softmax(torch.randn((1, D, H, W), device=dev), 1)
    ft = torch.randn((1, H, W, vt.numC_Trans), device=dev)
    t = _time_launches(lambda: ops.lift_splat(dp, ft, idx))
    byts = (dp.numel() + ft.numel() + V * vt.numC_Trans) * 4.0 + idx.order.numel() * 4.0
    kernels.append({"name": "lift_splat", "bound": "hbm", "ms": t * 1e3, "gbs": byts / t / 1e9,
                    "frac": byts / t / 1e9 / pk["hbm_gbs"]})
    # (5) x2 trilinear + argmax (write-bound)
    lg = torch.randn((1, nx[0], nx[1], nx[2], 20), device=dev)
    occ = mc["occ_size"]
    t = _time_launches(lambda: ops.trilinear(lg, occ, want_labels=True))
    byts = (lg.numel() + occ[0] * occ[1] * occ[2] * 20) * 4.0 + occ[0] * occ[1] * occ[2]
    kernels.append({"name": "trilinear_x2+argmax", "bound": "hbm", "ms": t * 1e3, "gbs": byts / t / 1e9,
                    "frac": byts / t / 1e9 / pk["hbm_gbs"]})
    # (6) BRI attention (tensor-bound: 2 x 13.2 GFLOP of QK^T (two passes) + 13.2 GFLOP PV per call)
    q = torch.softmax(torch.randn((1, D, H, W), device=dev) * 2, 1)
    kvv = torch.softmax(torch.randn((1, D, H, W), device=dev) * 2, 1)
    both = torch.empty((1, D, H, W, 2), device=dev)
    prm = vt.volume_interaction.lss2stereo.packed()
    t = _time_launches(lambda: ops.bri_attention(q, kvv, prm, both[..., 0], 2))
    fl = 2.0 * (H * W) * (H * W) * D * 2
    kernels.append({"name": "bri_attention 7680 tokens x 112 (prep + tcgen05 kernel + split combine)", "bound": "tensor", "ms": t * 1e3,
                    "tflops": fl / t / 1e12, "frac": fl / t / 1e12 / pk["tflops"]})
    # (7) DepthNet 2-D conv 640->640 k3 on the 48x160 map (tensor-bound, 56.6 GFLOP, one wave of 120 CTAs)
    dc = vt.depth_net.depth_conv[0].conv1
    xd = torch.randn((1, 1, H, W, dc.in_channels), device=dev)
    t = _time_launches(lambda: ops.conv(Vol(xd), dc, math_mode=ops.SS_MATH_TF32))
    fl = 2.0 * H * W * 9 * dc.in_channels * dc.out_channels
    kernels.append({"name": "depth_net conv2d 640->640 k3 (48x160)", "bound": "tensor", "ms": t * 1e3, "tflops": fl / t / 1e12,
                    "frac": fl / t / 1e12 / pk["tflops"]})
    t = _time_launches(lambda: ops.conv(Vol(xd), dc, math_mode=ops.SS_MATH_TF32X3))
    kernels.append({"name": "depth_net conv2d 640->640 k3, compensated TF32x3", "bound": "tensor", "ms": t * 1e3,
                    "tflops": fl / t / 1e12, "frac": fl / t / 1e12 / pk["tflops"]})
    return roof, kernels


def _json_only_stdout():
    """Keep stdout = the ONE JSON line: libraries (NCCL prints its version banner with printf) write to fd 1 behind Python's
    back, so fd 1 is pointed at stderr for the life of the process and the line goes to a private copy of the real stdout."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


_REAL_STDOUT = None


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    args = parse()
    _REAL_STDOUT = _json_only_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
