#!/usr/bin/env python
"""bench.py - rays/s and seconds per DTU-shaped depth map of the per-ray rendering hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode tc16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one full depth map of the workload (BASELINE.json configs[1]: 1600x1216 ray grid, 3 source
views, unfavourable set 1/16/36, 64+64 samples per ray) rendered by ``ufo_render_rays`` - i.e. everything
``UFORecon.infer`` does for every pixel ray (code1/model.py:393-478) - from synthetic DTU-format inputs
(``uforecon_b200.synthetic``) and a synthetic checkpoint with the reference's state-dict layout.

* ``value``  rays/s, whole job, inputs (scene tensors AND sampler uniforms) resident in HBM.
* ``e2e``    the same metric through ``ufo_render_rays_host``: sampler uniforms start in pinned HOST memory,
             depth/rgb end in pinned HOST memory, copies inside the timed region.
* ``roofline``      the dominant kernel of the timed region against its roofline.
* ``cpu_baseline``  the CPU restatement of the reference (oracle/) timed on this box's host cores on a bounded
                    sample of the same workload (N=1 only).
* N>1: STRONG scaling is the headline - one depth map per step (BASELINE configs[2]: favourable view set), contiguous row
  blocks of the ray grid per rank, the NCCL gather of depth/rgb to rank 0 inside the timed region; ``weak_scaling`` (one full
  map per rank, no communication) is an extra.  ``--sweep 49`` runs BASELINE configs[4] (images round-robin over the ranks).

``--impl reference`` times the reference's own CPU implementation of the same path on the host cores - the UNMODIFIED
reference staged under baseline/_ref by ``baseline/reference_arm.py`` (the oracle port only if that copy is absent) - and
prints the same JSON line shape.  ``reference_cuda`` (N=1 extra) times the same unmodified call with model and scene on cuda:0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "rays_per_sec"
UNIT = "rays/s"


# --------------------------------------------------------------------------------------------------
# algorithmic work per unit
# --------------------------------------------------------------------------------------------------
def flops_per_point(nv: int) -> dict:
    """Algorithmic FLOPs (2 x MAC, true K, no padding) per sample point of the transformer stage."""
    d, dr = 80, 88
    loftr = lambda dd: 2 * (3 * dd * dd + dd * dd + 4 * dd * dd + 2 * dd * dd)      # qkv, merge, mlp0, mlp2
    attn = lambda dd: 2 * 2 * dd * (dd // 8)                                          # K^T V and Q.KV per token
    view = (loftr(d) + attn(d)) * (nv + 1)
    ray = loftr(dr) + attn(dr)
    presim = 2 * (8 * 32 + 32 * 32 + 32 * 16)
    density = 2 * (88 * 32 + 32 * 16 + 16)
    radiance = 2 * (83 * 16 + 16 * 8 + 8) * nv
    return {"view": view, "ray": ray, "presim": presim, "density": density, "radiance": radiance,
            "total": view + ray + presim + density + radiance}


def tap_bytes_per_point(nv: int) -> int:
    """Bytes of texel/voxel taps one point reads (L2/L1 level)."""
    return 512 * nv * (nv - 1) + 864 * nv + 512 * nv + 48 * nv + 16 * nv


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
VIEW_SETS = {"unfavorable": [1, 16, 36], "favorable": [23, 24, 33]}


def view_ids(name: str, nv: int):
    from uforecon_b200 import synthetic
    if name in VIEW_SETS and nv == 3:
        return VIEW_SETS[name]
    return synthetic.TEN_VIEW_LIST[:nv]


def build_workload(args):
    from uforecon_b200 import checkpoint, synthetic
    W, H = args.width, args.height
    views = view_ids(args.views, args.nv)
    sd, src = checkpoint.load_hot_path_state(os.path.join(ROOT, "pretrained", "uforecon.ckpt"))
    batch = synthetic.make_batch(views, (W, H))
    scene = synthetic.make_scene(batch)
    batch["depth_info"] = scene["depth_info"]
    return batch, scene, sd, src, views


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "tf_burst": j["bf16_tflops"], "tf_sustained": j["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


# --------------------------------------------------------------------------------------------------
# CPU leg (oracle port of the reference) - used by cpu_baseline and by --impl reference
# --------------------------------------------------------------------------------------------------
_REF_MODELS = {}


def cpu_chunks(batch, scene, sd, n_chunks: int, chunk: int = 800, warm: int = 1):
    """Times ``n_chunks`` calls of the reference's ``infer`` on ``chunk`` rays each (the reference's own chunking:
    --test_ray_num 800, script/eval_dtu_unfavorable.sh) on the host cores.  Runs the UNMODIFIED reference staged under
    baseline/_ref (``kind`` "reference"); only when that copy is absent, the oracle restatement (``kind`` "port").
    Returns (rays/s, seconds, rays, kind)."""
    from baseline import reference_arm
    if reference_arm.available():
        nv = batch["source_imgs"].shape[1]
        if nv not in _REF_MODELS:
            _REF_MODELS[nv] = reference_arm.load_model(nv, sd, "cpu")
        v, secs, rays, _ = reference_arm.infer_chunks(_REF_MODELS[nv], batch, scene, n_chunks, chunk, warm, "cpu")
        return v, secs, rays, "reference"
    from oracle import uforecon_oracle as orc      # CPU baseline only, never the product path
    from uforecon_b200 import synthetic
    H, W = batch["source_imgs"].shape[-2:]
    total = H * W
    times = []
    with torch.no_grad():
        for i in range(warm + n_chunks):
            # chunks spread over the image so that the sample sees centre and border rays
            begin = (total // (warm + n_chunks + 1)) * (i + 1)
            ray_idx = torch.arange(begin, min(begin + chunk, total))
            u_c, u_f = synthetic.sampler_uniforms(len(ray_idx), seed=100 + i)
            t0 = time.perf_counter()
            orc.infer(batch, scene, sd, ray_idx, u_c, u_f)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append((dt, len(ray_idx)))
    secs = sum(t for t, _ in times)
    rays = sum(n for _, n in times)
    return rays / secs, secs, rays, "port"


def reference_on_cuda(batch, scene, sd, dev, chunks=(800, 4000, 16000), n_chunks=3):
    """Same-device bar: the UNMODIFIED reference's ``UFORecon.infer`` with the model and the scene
    tensors on the B200 - its ATen / cuBLAS op sequence - per chunk size, TF32 off (torch 1.13 default) and on."""
    from baseline import reference_arm
    if not reference_arm.available():
        return {"unavailable": "baseline/_ref not staged"}
    nv = batch["source_imgs"].shape[1]
    model = reference_arm.load_model(nv, sd, dev)
    b = reference_arm.to_device(batch, dev)
    s = reference_arm.to_device(scene, dev)
    H, W = batch["source_imgs"].shape[-2:]
    rows = []
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            for ch in chunks:
                try:
                    v, secs, rays, _ = reference_arm.infer_chunks(model, b, s, n_chunks, ch, 1, dev)
                    rows.append({"chunk_rays": ch, "tf32": tf32, "rays_per_s": v, "sec_per_depth_map": H * W / v, "rays_timed": rays})
                except torch.OutOfMemoryError:
                    rows.append({"chunk_rays": ch, "tf32": tf32, "error": "out of memory"})
                    torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    del model, b, s
    torch.cuda.empty_cache()
    return {"what": "unmodified reference UFORecon.infer (code1/model.py:393-478) with model + scene on cuda:0, "
                    "wall clock around each call with synchronize, sampler uniforms drawn on the CPU as the reference does",
            "rows": rows, "best_rays_per_s": max((r.get("rays_per_s", 0.0) for r in rows), default=0.0)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (baseline/_ref, unmodified) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    batch, scene, sd, src, views = build_workload(args)
    chunk = 800
    per_step = max(1, args.ref_chunks)
    # warm-up steps and timed steps are both bounded samples (per_step chunks of 800 rays)
    for _ in range(min(args.warmup, 1)):
        cpu_chunks(batch, scene, sd, 1, chunk, warm=0)
    t_all, r_all, kind = 0.0, 0, "port"
    for _ in range(args.steps):
        _, secs, rays, kind = cpu_chunks(batch, scene, sd, per_step, chunk, warm=0)
        t_all += secs
        r_all += rays
    val = r_all / t_all
    H, W = batch["source_imgs"].shape[-2:]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "sec_per_depth_map": H * W / val,
        "config": workload_config(args, views, src, "cpu-fp32"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{args.steps} steps x {per_step} chunks of {chunk} rays of the same workload "
                                   f"(reference chunking --test_ray_num 800), extrapolated; "
                                   + ("unmodified reference staged under baseline/_ref" if kind == "reference"
                                      else "oracle restatement (baseline/_ref absent)")},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, views, ckpt_src, mode_name):
    n = max(args.gpus, 1)
    cfg = "BASELINE configs[1]" if n == 1 else "BASELINE configs[2]: rays sharded over the GPUs"
    par = ("1 GPU, one full depth map per step" if n == 1 else
           f"{n} GPUs, ONE depth map per step: contiguous row blocks of the ray grid per rank (no collective in the render), "
           f"one NCCL gather of depth/rgb to rank 0 inside the timed region")
    return {"workload": f"DTU-shaped {args.width}x{args.height} ray grid, {args.nv} source views {views} "
                        f"({args.views}), 64 coarse + 64 importance samples/ray, full depth-map render ({cfg})",
            "rays_per_depth_map": args.width * args.height, "n_views": args.nv, "mode": mode_name,
            "checkpoint": ckpt_src,
            "l2": "inputs larger than L2 (scene tensors 4.2 GB at 1600x1216 vs 126 MB L2); no explicit flush",
            "timing": "value/ms_per_step: K steps between CUDA events, no per-kernel brackets; roofline: the same K steps repeated with every launch bracketed by CUDA events",
            "parallelism": par}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from uforecon_b200 import _lib, dist as ufodist
    from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"bench.py: warning --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    import __graft_entry__ as ge
    ge.build()
    lib = _lib.load()
    mode = {"fp32": _lib.UFO_MODE_FP32, "tc16": _lib.UFO_MODE_TC_F16}[args.mode]

    t0 = time.time()
    batch, scene, sd, ckpt_src, views = build_workload(args)
    weights = HotPathWeights(sd, dev)
    sc = Scene(batch, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"], dev)
    H, W = sc.H, sc.W
    n_rays = H * W if args.rays <= 0 else min(args.rays, H * W)
    setup_s = time.time() - t0

    # sampler uniforms (FixedSampler / ImportanceSampler draws of the reference, sampler.py:42,86):
    # pinned host copies for the e2e leg, device copies for the resident leg
    gen = torch.Generator().manual_seed(1234 + rank)
    u_c_host = torch.rand(64, n_rays, generator=gen).pin_memory()
    u_f_host = torch.rand(64, n_rays, generator=gen).pin_memory()
    u_c, u_f = u_c_host.to(dev), u_f_host.to(dev)
    out_depth = torch.empty(n_rays, device=dev)
    out_depthz = torch.empty(n_rays, device=dev)
    out_rgb = torch.empty(n_rays, 3, device=dev)
    out = _lib.UfoRenderOut()
    out.depth, out.depth_z, out.rgb = out_depth.data_ptr(), out_depthz.data_ptr(), out_rgb.data_ptr()
    stream = torch.cuda.current_stream(dev)

    def step_resident(begin=0, n=n_rays):
        _lib.check(lib.ufo_render_rays(sc.handle, weights.handle, None, begin, n, u_c.data_ptr() + 4 * begin,
                                       u_f.data_ptr() + 4 * begin, n_rays, mode, C.byref(out), None, stream.cuda_stream))

    depth_host = torch.empty(n_rays).pin_memory()
    rgb_host = torch.empty(n_rays, 3).pin_memory()

    def step_e2e():
        _lib.check(lib.ufo_render_rays_host(sc.handle, weights.handle, 0, n_rays, u_c_host.data_ptr(), u_f_host.data_ptr(),
                                            mode, depth_host.data_ptr(), rgb_host.data_ptr(), stream.cuda_stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    begin, n_mine = ufodist.shard_rows(H, W, world, rank)
    counts = ufodist.shard_counts(H, W, world)
    if args.rays > 0:
        begin, n_mine, counts = 0, n_rays, [n_rays] * world

    def step_sharded():
        """one depth map over all ranks: this rank's row block, then the gather of depth/rgb to rank 0"""
        step_resident(begin, n_mine)
        return ufodist.gather_depth_rgb(out_depthz[:n_mine], out_rgb[:n_mine], counts) if world > 1 else None

    step_main = step_resident if world == 1 else step_sharded
    # ---- warm-up
    for _ in range(args.warmup):
        step_main()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream.  N = 1: one full depth map per step.  N > 1: ONE
    #      depth map per step, rows sharded over the ranks, gather to rank 0 included (strong scaling)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = lib.ufo_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_main()
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.ufo_launch_count() - launches0
    ms_step = ms_total / args.steps
    value = n_rays * args.steps / (ms_total * 1e-3)
    # ---- the same K steps again with every launch bracketed by CUDA events (ufo_profile_*): per-kernel device time
    #      for the roofline entries; kept out of the region above because the 2 event records per launch cost ~3 %
    _lib.profile_begin()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_main()
    e1.record(stream)
    barrier()
    ms_prof_total = max_over_ranks(e0.elapsed_time(e1))
    prof = _lib.profile_end(256)
    clk = clocks.stop() if rank == 0 else None

    # ---- e2e: host buffers, copies inside the timed region.  N = 1: ufo_render_rays_host.  N > 1: every rank uploads the
    #      uniforms of its row block from pinned host memory, renders, the blocks are gathered to rank 0 and rank 0 copies the
    #      assembled depth map to pinned host memory
    if world > 1:      # this rank's block of the uniforms as its own contiguous pinned host / device buffers
        u_c_hs = u_c_host[:, begin:begin + n_mine].contiguous().pin_memory()
        u_f_hs = u_f_host[:, begin:begin + n_mine].contiguous().pin_memory()

    def step_e2e_sharded():
        # the same entry point as N = 1: uniforms of this rank's row block from pinned host memory, uploaded in column blocks on the
        # library's copy stream under the render; the outputs stay on the device for the gather
        _lib.check(lib.ufo_render_rays_host(sc.handle, weights.handle, begin, n_mine, u_c_hs.data_ptr(), u_f_hs.data_ptr(), mode,
                                            out_depthz.data_ptr(), out_rgb.data_ptr(), stream.cuda_stream))
        got = ufodist.gather_depth_rgb(out_depthz[:n_mine], out_rgb[:n_mine], counts)
        if rank == 0:
            depth_host.copy_(got[0], non_blocking=True)
            rgb_host.copy_(got[1], non_blocking=True)
            stream.synchronize()

    step_e2e_main = step_e2e if world == 1 else step_e2e_sharded
    step_e2e_main()
    barrier()
    k_e2e = max(1, min(args.steps, args.e2e_steps))
    e0.record(stream)
    for _ in range(k_e2e):
        step_e2e_main()
    e1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = n_rays * k_e2e / (ms_e2e * 1e-3)
    e2e_h2d = 2 * 64 * 4 * n_rays if world == 1 else 2 * 64 * 4 * n_rays    # all ranks together upload the whole map's uniforms
    e2e_d2h = 16 * n_rays

    # ---- accuracy of this mode at THIS workload (outside every timed region): a band of rows rendered in the measured
    #      mode and in fp32 mode (the 1e-5 parity path) with the same uniforms -> the north-star tolerance figures
    accuracy = None
    if rank == 0 and args.mode != "fp32" and not args.no_accuracy:
        try:
            accuracy = measure_accuracy(lib, sc, weights, batch, mode, u_c, u_f, n_rays, W, H, stream)
        except Exception as ex:  # reported extra: never fail the headline line
            accuracy = {"error": str(ex)}

    # ---- weak-scaling extra (N > 1): every rank renders its own full depth map, no communication
    weak = None
    if world > 1:
        step_resident()
        barrier()
        e0.record(stream)
        for _ in range(2):
            step_resident()
        e1.record(stream)
        barrier()
        ms_w = max_over_ranks(e0.elapsed_time(e1))
        weak = {"value": world * n_rays * 2 / (ms_w * 1e-3), "unit": UNIT, "ms_per_step": ms_w / 2,
                "what": "one full depth map per rank per step, no collective (image-sharded sweeps, BASELINE configs[4])"}

    # ---- roofline of the dominant kernel
    pk = peaks()
    roofs = roofline_entries(prof, args, n_rays if world == 1 else n_mine, pk)
    roof = roofs[0] if roofs else None
    costvol = None
    if rank == 0 and world == 1 and not args.no_costvolume:
        try:
            costvol = bench_costvolume(args, dev)
        except Exception as ex:  # measurement extra: never fail the headline line
            costvol = {"error": str(ex)}
    tsdf = None
    if rank == 0 and world == 1 and not args.no_costvolume:
        try:
            tsdf = bench_tsdf(args, dev, None, None)
        except Exception as ex:
            tsdf = {"error": str(ex)}

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        v, secs, rays, kind = cpu_chunks(batch, scene, sd, args.cpu_chunks)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"{args.cpu_chunks} chunks of 800 rays of the same workload after 1 warm-up chunk "
                         f"({secs:.1f} s of CPU work, "
                         + ("unmodified reference UFORecon.infer from baseline/_ref" if kind == "reference" else "oracle restatement of the reference")
                         + ", torch CPU fp32)",
               "sec_per_depth_map_extrapolated": n_rays / v}
    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_ref_cuda:
        try:
            ref_cuda = reference_on_cuda(batch, scene, sd, dev)
        except Exception as ex:  # reported extra: never fail the headline line
            ref_cuda = {"error": repr(ex)}

    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            extras = bench_extras(batch, scene, sd, dev, mode)
        except Exception as ex:  # reported extra: never fail the headline line
            extras = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tc16": "f16"}[args.mode], "data": "synthetic",
            "sec_per_depth_map": ms_step * 1e-3,
            "config": dict(workload_config(args, views, ckpt_src, args.mode), accuracy=accuracy,
                           rays_per_rank=[int(c) for c in counts] if world > 1 else [int(n_rays)]),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d,
                    "d2h_bytes_per_step": e2e_d2h, "steps": k_e2e,
                    "api": ("ufo_render_rays_host (pinned host uniforms in, pinned host depth/rgb out)" if world == 1 else
                            "per rank: ufo_render_rays_host on its row block (pinned host uniforms in, outputs kept on the device), NCCL gather to rank 0, assembled depth/rgb to pinned host")},
            "gpu_launches": int(launches),
            "accuracy": accuracy,
            "weak_scaling": weak,
            "clocks": clk,
            "roofline": roof,
            "rooflines": roofs,
            "costvolume": costvol,
            "tsdf": tsdf,
            "extras": extras,
            "profiled_ms_per_step": ms_prof_total / args.steps,
            "kernels": [{"name": n, "launches": c, "ms": round(ms, 3)} for n, c, ms in sorted(prof, key=lambda x: -x[2])[:12]],
            "cpu_baseline": cpu,
            "reference_cuda": ref_cuda,
            "setup_s": round(setup_s, 1),
            "scene_device_bytes": sc.device_bytes,
        }
        print(json.dumps(line), flush=True)
    sc.close()
    weights.close()
    if world > 1:
        dist.destroy_process_group()


def bench_extras(batch, scene, sd, dev, mode):
    """What the drop-in API costs around the render itself (wall clock with synchronize, one repetition after a warm-up):
    ``ufo_scene_create`` (the repack of the view set's tensors) from device-resident and from host inputs, and
    ``UFOReconRenderer.render_depth_map`` with the reference's CPU uniform stream and with device-side uniforms."""
    from baseline import reference_arm      # only its to_device helper
    from uforecon_b200.renderer import Scene, UFOReconRenderer
    out = {}
    b_dev, s_dev = reference_arm.to_device(batch, dev), reference_arm.to_device(scene, dev)
    for name, (b, sc_) in (("scene_create_device_inputs_s", (b_dev, s_dev)), ("scene_create_host_inputs_s", (batch, scene))):
        for it in range(2):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            x = Scene(b, sc_["source_imgs_feat"], sc_["feature_volume"], sc_["match_feature"], dev)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            nbytes = x.device_bytes
            x.close()
        out[name] = dt
    out["scene_device_bytes"] = int(nbytes)
    ren = UFOReconRenderer(sd, dev, mode=mode)
    ren.begin_scene(b_dev, s_dev["source_imgs_feat"], s_dev["feature_volume"], s_dev["match_feature"])
    for name, kw in (("render_depth_map_device_rng_s", dict(device_rng=True)), ("render_depth_map_cpu_rng_s", dict(device_rng=False))):
        for it in range(2 if kw["device_rng"] else 1):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            d, c = ren.render_depth_map(b_dev, s_dev["source_imgs_feat"], s_dev["feature_volume"], s_dev["match_feature"], **kw)
            d, c = d.cpu(), c.cpu()
            dt = time.perf_counter() - t0
        out[name] = dt
    ren.close()
    out["what"] = ("wall seconds; render_depth_map = the chunk loop of extract_geometry (model.py:814-831) for one 1600x1216 map incl. the "
                   "device->host copy of depth/rgb; cpu_rng draws the reference's 2 x 64 uniforms per ray from torch's CPU generator")
    return out


def sweep_sources(n_images: int):
    """Source views of every render view of a sweep: the 3 nearest cameras of the synthetic rig (what dtu_pairs.txt encodes
    for the real rig: view-selection scores, best first)."""
    import numpy as np
    from uforecon_b200 import synthetic
    rig = synthetic.make_rig(n_images)
    eyes = np.stack([-m[:3, :3].T @ m[:3, 3] for m in rig])
    out = []
    for v in range(n_images):
        d = np.linalg.norm(eyes - eyes[v], axis=1)
        d[v] = np.inf
        out.append([int(i) for i in np.argsort(d)[:3]])
    return out


def run_sweep(args):
    """BASELINE configs[4]: a sweep of ``--sweep`` depth maps (49 = one DTU scan, the TSDF-fusion input), whole images
    round-robin over the ranks (``dist.shard_images``).  Per image, inside the timed region: host->device copy of the image's
    batch tensors (rays, source images, MVS depth), ``ufo_scene_create`` (the repack of its view set), the render of all
    H*W rays, device->host copy of depth and colour.  The encoder outputs (features, volumes, match maps) are device-resident
    synthetic tensors shared by all images - producing them is the PyTorch encoder's job, outside this path."""
    import torch.distributed as dist
    from baseline import reference_arm      # only its to_device helper
    from uforecon_b200 import _lib, checkpoint, dist as ufodist, synthetic
    from uforecon_b200.renderer import HotPathWeights, Scene, render_rays
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    ge.build()
    mode = {"fp32": _lib.UFO_MODE_FP32, "tc16": _lib.UFO_MODE_TC_F16}[args.mode]
    W, H = args.width, args.height
    sd, ckpt_src = checkpoint.load_hot_path_state(os.path.join(ROOT, "pretrained", "uforecon.ckpt"))
    weights = HotPathWeights(sd, dev)
    mine = ufodist.shard_images(args.sweep, world, rank)
    srcs = sweep_sources(args.sweep)
    # host-side inputs of this rank's images (the dataloader's job), pinned; encoder outputs once, on the device
    batches = []
    scene = None
    for v in mine:
        b = synthetic.make_batch([v] + srcs[v][:args.nv - 1], (W, H))
        if scene is None:
            scene = reference_arm.to_device(synthetic.make_scene(b), dev)
        b["depth_info"] = scene["depth_info"].cpu()
        batches.append({k: (t.pin_memory() if torch.is_tensor(t) and t.numel() > 1024 else t) for k, t in b.items()})
    n = H * W
    u_c = torch.rand(64, n, device=dev)
    u_f = torch.rand(64, n, device=dev)
    depth_host = torch.empty(n).pin_memory()
    rgb_host = torch.empty(n, 3).pin_memory()

    def one_image(b):
        bd = {k: (t.to(dev, non_blocking=True) if torch.is_tensor(t) else t) for k, t in b.items() if k not in ("proj_matrices",)}
        sc = Scene(bd, scene["source_imgs_feat"], scene["feature_volume"], scene["match_feature"], dev)
        r = render_rays(sc, weights, None, n, u_c, u_f, mode, ray_begin=0, want=("depth_z", "rgb"))
        depth_host.copy_(r["depth_z"], non_blocking=True)
        rgb_host.copy_(r["rgb"], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        sc.close()

    if batches:
        one_image(batches[0])                      # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for b in batches:
        one_image(b)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total = float(tt.item())
    if rank == 0:
        counts = [len(ufodist.shard_images(args.sweep, world, r)) for r in range(world)]
        line = {"metric": METRIC, "value": args.sweep * n / total, "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1,
                "ms_per_step": total * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": {"fp32": "f32", "tc16": "f16"}[args.mode], "data": "synthetic",
                "config": {"workload": f"{args.sweep}-view depth-map sweep of one synthetic scan at {W}x{H}, {args.nv} source views per image "
                                       f"(3 nearest cameras), 64+64 samples/ray, images round-robin over {world} GPU(s) (BASELINE configs[4])",
                           "per_image": "H2D of the image's batch tensors + ufo_scene_create + full-map render + D2H of depth/rgb",
                           "checkpoint": ckpt_src, "mode": args.mode},
                "sweep": {"images": args.sweep, "total_s": total, "images_per_rank": counts, "busiest_rank_images": max(counts),
                          "s_per_image_on_busiest_rank": total / max(counts)},
                "e2e": {"value": args.sweep * n / total, "unit": UNIT, "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in batches[0].values() if torch.is_tensor(t)) * args.sweep),
                        "d2h_bytes_per_step": 16 * n * args.sweep}}
        print(json.dumps(line), flush=True)
    weights.close()
    if world > 1:
        dist.destroy_process_group()


def measure_accuracy(lib, sc, weights, batch, mode, u_c, u_f, n_rays, W, H, stream, rows=16):
    """Tensor-core mode vs fp32 mode on `rows` image rows from the middle of the map: p99 depth error as a fraction of the
    depth interval and colour PSNR (BASELINE.json north star: <= 0.005 and >= 50 dB).  Rays whose in-image test is decided
    by the last bits of the projection are left out of the PSNR (see tests/test_gpu_tc.py)."""
    import ctypes as C
    from uforecon_b200 import _lib
    k = min(n_rays, rows * W)
    begin = ((n_rays - k) // 2 // W) * W
    dev = u_c.device
    res = {}
    for name, m in (("tc", mode), ("ref", _lib.UFO_MODE_FP32)):
        d, r, z = torch.empty(k, device=dev), torch.empty(k, 3, device=dev), torch.empty(k, 128, device=dev)
        o = _lib.UfoRenderOut()
        o.depth, o.rgb, o.z = d.data_ptr(), r.data_ptr(), z.data_ptr()
        _lib.check(lib.ufo_render_rays(sc.handle, weights.handle, None, begin, k, u_c.data_ptr() + 4 * begin,
                                       u_f.data_ptr() + 4 * begin, n_rays, m, C.byref(o), None, stream.cuda_stream))
        res[name] = (d, r, z)
    torch.cuda.synchronize()
    near, far = float(batch["near_fars"][0, 0, 0]), float(batch["near_fars"][0, 0, 1])
    de = (res["tc"][0] - res["ref"][0]).abs() / (far - near)
    P = batch["source_poses"][0].to(dev)
    dirs = batch["ray_d"][0][:, begin:begin + k].t().to(dev)
    pts = batch["ray_o"][0].to(dev)[None, None] + res["ref"][2][:, :, None] * dirs[:, None, :]
    q = torch.einsum("vij,rsj->vrsi", P[:, :3, :3], pts) + P[:, None, None, :3, 3]
    uv = q[..., :2] / q[..., 2:3]
    amb = ((uv.abs() - 1).abs() < 2e-5).any(-1).any(0).any(1)
    mse = float(((res["tc"][1] - res["ref"][1])[~amb] ** 2).mean())
    return {"against": "fp32 mode of this library (1e-5 parity path) on the same rays and uniforms", "rays": int(k),
            "depth_err_p99_frac_of_interval": float(de.quantile(0.99)), "depth_err_p50_frac_of_interval": float(de.median()),
            "rgb_psnr_db": 10 * math.log10(1.0 / max(mse, 1e-20)), "mask_ambiguous_rays_excluded": int(amb.sum()),
            "tolerance": {"depth_err_p99_frac_of_interval": 0.005, "rgb_psnr_db": 50.0}}


def _ncu_traffic():
    """Per kernel family: DRAM bytes per launch and unit utilisations from the committed ncu --set full capture at the default
    chunk size (profiles/ncu_traffic.json, written by tools/ncu_traffic.py), or {}."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def roofline_entries(prof, args, n_rays, pk):
    """One roofline entry per kernel family of the timed region, largest share first.

    Tensor kernels: algorithmic FLOPs of the points the launches process / live CUDA-event time, against the sustained bf16
    peak.  The gather: the DRAM bytes it really moves (ncu capture at the same chunk size) / live time against the HBM copy
    peak - a small fraction, because its limiter is not HBM: the texel / voxel tap bytes it pulls through L1 (``tap_gbs``) and
    ncu's L1 / issue utilisation are reported beside it."""
    if not prof:
        return []
    total_ms = sum(ms for _, _, ms in prof)
    nv = args.nv
    fl = flops_per_point(nv)
    # sample points a launch really processes: the ray stage sees 64 + 128 per ray like the reference; the gathers and
    # the view stage see 64 + 64 in the tensor-core modes (coarse results are reused) and 64 + 128 in fp32 mode
    pts_ray = n_rays * 192 * args.steps
    pts_pt = n_rays * (128 if args.mode != "fp32" else 192) * args.steps
    pts = pts_ray
    traffic = _ncu_traffic()
    same_cfg = args.nv == 3 and args.width == 1600 and args.height == 1216 and args.rays <= 0
    out = []
    for name, cnt, ms in sorted(prof, key=lambda x: -x[2]):
        bound, unit, work, note = "tensor", "TFLOP/s", None, None
        fam = name.split("<")[0]
        t = traffic.get(fam) if isinstance(traffic.get(fam), dict) and same_cfg else None
        avg_s = ms * 1e-3 / cnt
        extra = {}
        if name.startswith("k_view_tc"):
            work = pts_pt * (fl["view"] + fl["radiance"])
        elif name.startswith("k_ray_tc"):
            work = pts * (fl["ray"] + fl["density"])
        elif name.startswith("k_linear<"):
            k, n = [int(x) for x in name[len("k_linear<"):-1].split(",")]
            rows = {80: pts * (nv + 1), 160: pts * (nv + 1), 88: pts, 176: pts}[k]
            work = 2.0 * rows * k * n
        elif name.startswith("k_gather"):
            bound, unit = "hbm", "GB/s"
            tap = pts_pt * tap_bytes_per_point(nv)
            extra["tap_gbs"] = tap / cnt / avg_s / 1e9
            extra["tap_bytes_per_launch"] = tap / cnt
            if t:
                work = t["bytes_per_launch"] * cnt
                note = ("achieved = DRAM bytes the launch moves (ncu, same chunk size: the scene tensors it reads and the token / colour / "
                        "direction rows it writes for the view stage - 28 % of the bytes in the coarse pass, 85 % in the fine pass) / live launch time; the limiter is L1 + issue, not HBM: see tap_gbs "
                        "(texel / voxel tap bytes pulled through L1) and the ncu utilisations")
            else:
                # no ncu capture at this view count / size: a lower bound of the launch's DRAM traffic that follows from the
                # data layout alone - the 16-bit token rows, (r,g,b,mask) and direction rows it must write for the view stage
                work = pts_pt * nv * (80 * 2 + 16 + 16)
                note = ("no ncu capture for this configuration: achieved = the bytes the launch WRITES for the view stage (token, colour and "
                        "direction rows; a lower bound of its DRAM traffic) / live launch time; the limiter is L1 + issue, see tap_gbs")
        if work is None:
            continue
        per_launch = work / cnt
        achieved = per_launch / avg_s / (1e12 if bound == "tensor" else 1e9)
        peak = pk["tf_sustained"] if bound == "tensor" else pk["hbm_gbs"]
        e = {"kernel": name, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
             "traffic": t["bytes_per_launch"] if t else None,
             "peak_source": pk["src"] + (" (sustained bf16)" if bound == "tensor" else " (copy)"),
             "launches": cnt, "avg_launch_ms": avg_s * 1e3, "share_of_step": ms / total_ms,
             "algorithmic_work_per_launch": per_launch,
             "work_basis": "algorithmic work of the sample points the launches process (ray stage 192 per ray; gathers and "
                           "view stage 128 per ray in the tensor-core modes, where the fine pass reuses the coarse results; "
                           "the reference evaluates 192)"}
        if t:
            e["ncu"] = {k: round(t[k], 2) for k in ("l1_throughput_pct", "l2_throughput_pct", "issue_active_pct", "tensor_pipe_active_pct",
                                                    "dram_throughput_pct") if k in t}
        e.update(extra)
        if note:
            e["note"] = note
        out.append(e)
    return out


def bench_costvolume(args, dev):
    """Kernel 1 (cost-volume build, TransMVSNet.py:76-100) at the workload's size: 3 cascade stages, N = V = n_views."""
    from uforecon_b200 import _lib, checkpoint
    from uforecon_b200.costvolume import similarity_volume
    from uforecon_b200 import synthetic
    nv, W, H = args.nv, args.width, args.height
    sd = checkpoint.synthetic_state_dict(0)
    batch = synthetic.make_batch(view_ids(args.views, nv), (W, H))
    comb = [list(range(i, nv)) + list(range(0, i)) for i in range(nv)]
    g = torch.Generator(device=dev).manual_seed(3)
    res = []
    vw = None
    for si, (stage, D, C) in enumerate((("stage1", 48, 32), ("stage2", 32, 16), ("stage3", 8, 8))):
        sc = synthetic.STAGE_SCALE[stage]
        hs, ws = H // sc, W // sc
        feats = [torch.randn(nv, C, hs, ws, device=dev, generator=g) for _ in range(nv)]
        proj = batch["proj_matrices"][stage][0][torch.tensor(comb)].contiguous()
        base = 425.0 + 2.65 * 192 * (0.3 + 0.4 * torch.rand(nv, 1, hs, ws, device=dev, generator=g))
        hyp = (base + (torch.arange(D, device=dev).view(1, D, 1, 1) - D / 2) * 2.65 * (4 / (si + 1)) * (4.0 if si == 0 else 1.0)).contiguous()
        if vw is not None:
            vw = torch.nn.functional.interpolate(vw, scale_factor=2, mode="nearest").contiguous()
        similarity_volume(feats, proj, hyp, sd, view_weights=vw, device=dev)          # warm-up
        _lib.profile_begin()
        for _ in range(3):
            sim, vw_new = similarity_volume(feats, proj, hyp, sd, view_weights=vw, device=dev)
        prof = _lib.profile_end(64)
        ms = sum(m for n, c, m in prof if n.startswith("k_costvol")) / 3
        ms_repack = sum(m for n, c, m in prof if n.startswith("k_nchw")) / 3
        vox = nv * D * hs * ws
        tap = vox * (nv - 1) * (4 * C * 4 + C * 4 / (nv - 1)) + vox * 4
        compulsory = nv * nv * C * hs * ws * 4 + vox * 4 * 2
        res.append({"stage": stage, "voxels": vox, "kernel_ms": ms, "repack_ms": ms_repack,
                    "tap_gbs": tap / (ms * 1e-3) / 1e9, "compulsory_gbs": compulsory / (ms * 1e-3) / 1e9})
        vw = vw_new
        del feats, sim
    return res


def bench_tsdf(args, dev, depth_z, sc_batch):
    """Next row N3: TSDF integration of depth maps (ufo_tsdf_integrate) - 16 synthetic views into a 512^3 volume."""
    from uforecon_b200 import _lib
    from uforecon_b200.tsdf import TSDFVolume
    import numpy as np
    H, W = 1216, 1600
    n_views = 16
    K = np.array([[2892.33, 0, 823.2], [0, 2883.18 * 1216 / 1200, 619.07 * 1216 / 1200], [0, 0, 1]], dtype=np.float32)
    depths, poses = [], []
    g = torch.Generator(device=dev).manual_seed(5)
    for v in range(n_views):
        th = 0.12 * v - 0.9
        eye = 650.0 * np.array([np.sin(th), 0.05 * v, -np.cos(th)])
        z = -eye / np.linalg.norm(eye)
        x = np.cross(np.array([0, 1.0, 0]), z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        c2w = np.eye(4, dtype=np.float32); c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, z, eye
        poses.append(c2w)
        depths.append(600.0 + 60.0 * torch.rand(H, W, device=dev, generator=g))
    vol = TSDFVolume(np.array([[-192.0, 192.0]] * 3), voxel_size=0.75, margin=3, device=dev)      # 512^3 voxels
    vol.integrate_many(depths, [K] * n_views, poses)                                             # warm-up
    _lib.profile_begin()
    for _ in range(3):
        vol.integrate_many(depths, [K] * n_views, poses)
    prof = _lib.profile_end(16)
    ms = sum(m for n, c, m in prof if n.startswith("k_tsdf")) / 3
    vox = int(np.prod(vol._vol_dim))
    upd = int((vol.device_volumes()[1] > 0).sum())
    # algorithmic bytes of the fused launch: read + write (tsdf, weight) once per touched voxel, one 4-byte depth tap per
    # voxel and view; the reference's kernel moves (tsdf, weight) once per VIEW
    fused = upd * 16 + vox * n_views * 4
    per_view = n_views * (upd * 16) + vox * n_views * 4
    res = {"voxels": vox, "views": n_views, "updated_voxels": upd, "kernel_ms": ms, "voxel_views_per_s": vox * n_views / (ms * 1e-3),
           "fused_gbs": fused / (ms * 1e-3) / 1e9, "reference_schedule_bytes": per_view, "hbm_peak_gbs": peaks()["hbm_gbs"]}
    # iso-surface extraction of the fused volume on the device (ufo_tsdf_mesh_*; the reference copies the volumes to the
    # host and runs scikit-image).  Algorithmic bytes per voxel: classify reads 4 and writes 1 (case) + 24/32 (group record
    # and counts), the two scans and the pack move 24/32, emit reads 24/32 - 7.25 in all - plus 24 B per vertex and 12 B per
    # face written
    from uforecon_b200.tsdf import marching_cubes
    v, f, n = marching_cubes(vol.device_volumes()[0])                                             # warm-up
    nv, nf = int(v.shape[0]), int(f.shape[0])
    del v, f, n
    _lib.profile_begin()
    for _ in range(3):
        out = marching_cubes(vol.device_volumes()[0])
        del out
    prof = _lib.profile_end(16)
    mesh_ms = sum(m for nme, c, m in prof if nme.startswith(("k_mc", "cub_"))) / 3
    mesh_bytes = int(vox * 7.25) + nv * 24 + nf * 12
    res["mesh"] = {"verts": nv, "faces": nf, "device_ms": mesh_ms, "voxels_per_s": vox / (mesh_ms * 1e-3),
                   "algorithmic_gbs": mesh_bytes / (mesh_ms * 1e-3) / 1e9, "frac_of_hbm_peak": mesh_bytes / (mesh_ms * 1e-3) / 1e9 / peaks()["hbm_gbs"],
                   "kernels": [{"name": nme, "launches": c, "ms": round(m / 3, 3)} for nme, c, m in prof]}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-accuracy", action="store_true", help="skip the tensor-core vs fp32 accuracy leg (outside the timed region)")
    ap.add_argument("--mode", default=os.environ.get("UFO_BENCH_MODE", "tc16"), choices=["fp32", "tc16"],
                    help="tc16 = tcgen05 with fp16 operands (default: meets the north-star tolerance at full size); fp32 = the 1e-5 parity path")
    ap.add_argument("--width", type=int, default=int(os.environ.get("UFO_BENCH_W", "1600")))
    ap.add_argument("--height", type=int, default=int(os.environ.get("UFO_BENCH_H", "1216")))
    ap.add_argument("--nv", type=int, default=3)
    ap.add_argument("--views", default=None, choices=["unfavorable", "favorable"],
                    help="source view set at NV=3; default: unfavorable (1,16,36 = BASELINE configs[1]) at N=1, favorable "
                         "(23,24,33 = configs[2], rays sharded over the GPUs) at N>1")
    ap.add_argument("--sweep", type=int, default=0, help="BASELINE configs[4]: render this many depth maps (49), images round-robin over the ranks")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-chunks", type=int, default=4)
    ap.add_argument("--ref-chunks", type=int, default=2, help="--impl reference: 800-ray chunks per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-costvolume", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the scene-create / render_depth_map timings (N=1 extra)")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the same-device extra (unmodified reference on cuda:0)")
    ap.add_argument("--rays", type=int, default=0, help="profiling aid: render only the first N rays of the map per step")
    args = ap.parse_args()
    if args.views is None:
        args.views = "unfavorable" if max(args.gpus, int(os.environ.get("WORLD_SIZE", "1"))) == 1 else "favorable"
    if args.sweep > 0 and args.impl != "reference":
        run_sweep(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
