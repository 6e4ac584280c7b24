#!/usr/bin/env python
"""bench.py — Quadfield render hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]

One "step" = one 800x800 frame (640 000 rays) of BASELINE.json configs[1] ("c2"): rays -> BVH first-K hits
(K=8, 20 480-triangle quadrature mesh) -> hash grid (16 levels, T=2^19) + 64-wide MLPs at the hits ->
composite.  Prints ONE JSON line (rank 0).

  value     rays/s with the frame's rays already resident in HBM (200 precomputed views, 3.07 GB > L2, cycled over 3 streams)
  e2e       rays/s through the reference-facing call from HOST inputs, copies inside the timed region: a 3x4 camera pose in
            host memory per frame (the reference's loader builds the rays on the device, nerf_synthetic.py:289-378) ->
            MeshRenderer.render_pose -> rgb + depth copied to pinned host memory (train_finetune.py:620-626)
  e2e_rays_from_host   round 1's flavour: pinned rays host->device (24 B/ray) + image device->host (20 B/ray) every step
  roofline  dominant kernel (ngp_forward_tc_kernel: hash-grid gather + tcgen05 MLPs): algorithmic 512 B per hit sample
            / its CUDA-event time measured live (kernel timed alone on its stream), vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (C / numpy / PyTorch port of the reference path) on every ray of one frame, rank 0, N=1
  c2_dense  the same frame from camera radius 2.2 (5 hits / ray instead of 1.7), N=1
  c4, c5    BASELINE configs[3] / [4] (1080p neural K=32 on a 1.15 M-triangle mesh; 4K baked SG textures): ONE frame per
            step, ray-sharded over the ranks in 4-row bands dealt round-robin, one NCCL gather per frame in the timed region
  train, field_train, field_train_occgrid   BASELINE configs[2]: fwd+bwd training steps on 2^18 rays, gradient all-reduce

N > 1 (torchrun): every rank renders its own c2 frames (views are the independent units; weak scaling, no data-path
collective), max-over-ranks device time; one NCCL gather of the last frame at the end of the timed region.
--impl reference times the reference's CPU path (the oracle port: the reference itself is CUDA-only and its native
dependencies are not installable offline) on the host cores, every ray of the frame per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rays_per_sec_render_fwd"
UNIT = "rays/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--views", type=int, default=200, help="resident ray sets cycled through (c2: 200 x 15.4 MB)")
    ap.add_argument("--cpu-sample", type=int, default=800, help="cpu_baseline renders a sample x sample sub-grid of one frame "
                    "(800 = every ray of a c2 frame, 3-6 s on 8-16 host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the fwd+bwd training-step measurement (extra key 'train')")
    ap.add_argument("--no-big", action="store_true", help="skip the ray-sharded configs[3] / configs[4] legs (extra keys 'c4', 'c5')")
    ap.add_argument("--train-rays", type=int, default=1 << 18, help="rays per training batch per GPU (BASELINE configs[2])")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="qf_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------- CPU arm (oracle)
def cpu_render_sample(sample: int, threads: int, config: str = "c2", steps: int = 1, warmup: int = 0):
    """Times the CPU oracle on a sample x sample sub-grid of one frame of the workload.  -> (rays/s, detail)."""
    import numpy as np
    import torch
    from oracle import quadfield_oracle as O
    from quadraturefields_b200 import scene as S
    torch.set_num_threads(threads)
    O.PREFER_C = True       # the C restatement (oracle/bruteforce.c, all host threads) instead of the numpy one ...
    O.PREFER_BVH = True     # ... through its CPU BVH (bit-identical to the brute force; the reference's Embree shape)
    O.set_c_threads(threads)   # torchrun exports OMP_NUM_THREADS=1; `cores` in the JSON line is what really runs
    cfg = S.CONFIGS[config]
    vertices, faces = O.shell_mesh(cfg["radii"], cfg["sub"], jitter=1e-3, seed=42)
    f, cx, cy, W, H = O.pinhole_intrinsics(cfg["W"], cfg["H"], S.CAMERA_ANGLE_X)
    poses = S.spiral_poses(cfg["views"], cfg.get("cam_radius", 4.03))
    meta = O.make_grid_meta(log2_hashmap_size=cfg["log2_T"])
    table, base_w, head_w = S.random_field_params(42, meta.n_entries)
    params = O.NGPParams(torch.tensor([-1.5] * 3 + [1.5] * 3), meta, table, base_w, head_w)
    times, detail = [], {}
    for it in range(warmup + steps):
        o, d = O.generate_rays(poses[it % len(poses)], W, H, f, cx, cy)
        ys = np.linspace(0, H - 1, sample).round().astype(np.int64)
        xs = np.linspace(0, W - 1, sample).round().astype(np.int64)
        idx = (ys[:, None] * W + xs[None, :]).reshape(-1)
        o, d = np.ascontiguousarray(o[idx]), np.ascontiguousarray(d[idx])
        tm = {}
        t0 = time.perf_counter()
        out = O.render_mesh_ngp(o, d, vertices, faces, params, K=cfg["K"], timings=tm, threads=threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            detail = dict(tm, hits=int(out["index_ray"].shape[0]), rays=int(idx.size))
    n = sample * sample
    return n / (sum(times) / len(times)), detail, sum(times) / len(times)


def run_reference(args):
    """Reference arm: the reference's own path on the host CPU (oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is rank 0 alone and uses all host threads (the
    # variable is read when the OpenMP runtimes start, i.e. before torch is imported below, and also seeds worker threads)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["MKL_NUM_THREADS"] = str(threads)
    # every step renders EVERY ray of one frame of the workload the line names (800x800 for c2: 1.4-6 s per step on 8-16 host
    # cores); steps / warm-up are capped so that the whole arm stays within a few minutes
    import quadraturefields_b200.scene as S_
    sample = max(S_.CONFIGS[args.config]["W"], S_.CONFIGS[args.config]["H"]) if args.config in ("c1", "c2") else 256
    requested = args.steps
    _, _, probe = cpu_render_sample(sample, threads, args.config, steps=1, warmup=0)           # also the (single) warm-up step
    args.steps = max(1, min(args.steps, int(200.0 / max(probe, 1e-3))))                        # the whole arm stays under ~4 minutes
    rps, detail, sec = cpu_render_sample(sample, threads, args.config, steps=args.steps, warmup=0)
    desc = (f"{sample}x{sample} rays (the whole frame) of one {args.config} frame per step; intersect {detail.get('intersect_s', 0):.2f}s "
            f"(OpenMP CPU BVH traversal), field {detail.get('field_s', 0):.2f}s, composite {detail.get('composite_s', 0):.3f}s; "
            f"{detail.get('hits', 0)} hits")
    line = {"impl": "reference", "metric": METRIC, "value": rps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "steps_requested": requested, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.config), "note": "reference path is CUDA-only with un-installable native "
                       "dependencies; timed here as its CPU port in oracle/ (OpenMP C BVH intersector + PyTorch field and compositing)"},
            "cpu_baseline": {"value": rps, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": rps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(config: str) -> str:
    return {"c2": "BASELINE configs[1]: NeRF-synthetic-shaped 800x800 frame per step (640000 rays), 200 views, 20480-tri "
                  "quadrature mesh, hash grid 16 lvls T=2^19 + 64-wide MLPs, K=8 hits/ray",
            "c1": "BASELINE configs[0]: 100x100 frame, 20480-tri mesh, T=2^19, K=8",
            "c4": "BASELINE configs[3]: Shelly-shaped 1920x1080 frame, 1.15M-tri shell mesh, T=2^21, K=32"}.get(config, config)


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from quadraturefields_b200.parallel import bind_to_gpu_numa
        bind_to_gpu_numa(local)                      # host buffers of the e2e leg land on the GPU's NUMA node
        dist.init_process_group("nccl", device_id=dev)
    entry.build()
    from quadraturefields_b200 import _lib, scene as S
    lib = _lib.load()

    sc = S.make_scene(args.config, device=dev)
    N = sc.n_rays
    n_views = min(args.views, len(sc.poses))
    # resident inputs: precomputed ray sets for n_views poses (rank r starts at view r so ranks render different frames)
    rays = [sc.rays(v) for v in range(n_views)]
    from quadraturefields_b200.utils import FramePipeline
    # frames rotate over three streams: the trace (ALU bound) of one frame overlaps the shading (L1-gather bound) of another;
    # measured 1 / 2 / 3 / 4 streams: 0.446 / 0.376 / 0.370 / 0.370 ms per frame
    NS = int(os.environ.get("QF_BENCH_STREAMS", "3"))
    pipe = FramePipeline(sc.renderer, NS)
    outs = [dict(rgb=torch.empty((N, 3), device=dev), opacity=torch.empty((N, 1), device=dev), depth=torch.empty((N, 1), device=dev))
            for _ in range(NS)]
    out = outs[0]
    from quadraturefields_b200 import parallel as P
    view_of = lambda step: P.view_for_step(step, rank, world, n_views)

    hit_slots = torch.zeros((args.warmup + args.steps, 1), dtype=torch.int32, device=dev)   # one slot per step, written by the kernel

    # N > 1, the final image gather (north star): the LAST frame of every rank goes to rank 0.  Its composite kernel stores
    # the pixels straight into rank 0's memory over NVLink (parallel.PeerFrame, one slot per rank) instead of an NCCL gather
    # after the loop (r2i at N=8, 20 steps: the gather of 7 x 12.8 MB was 7 % of the timed region).  Falls back to the
    # collective when the mapping or the self-check fails.
    final_peer, final_note = None, None
    if world > 1 and os.environ.get("QF_BENCH_PEER_FRAME", "1") != "0":
        ok = torch.ones((1,), device=dev)
        try:
            final_peer = P.PeerFrame(sc.H, sc.W, dev, dst=0, slots=world)
        except Exception as e:                           # noqa: BLE001
            final_note = f"peer frame unavailable ({type(e).__name__}: {e})"
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) < 1.0:
            final_peer = None
        else:
            o_, d_ = rays[view_of(0)]
            sc.renderer.render(o_, d_, image_width=sc.W, frame_out=final_peer.pointers(rank))
            final_peer.sync()
            chk = sc.renderer.render(o_, d_, image_width=sc.W)
            flat = P.gather_frame(torch.cat([chk["rgb"], chk["opacity"], chk["depth"]], dim=1), [N] * world, dst=0)
            if rank == 0:
                got = torch.cat([torch.cat(final_peer.frame(r), dim=1) for r in range(world)])
                if not torch.equal(flat, got):
                    ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) < 1.0:
                final_note = "peer-frame self-check against the NCCL gather FAILED: collective used"
                final_peer.close()
                final_peer = None
            else:
                final_note = "last frame of every rank stored into rank 0's memory over NVLink by its composite kernel (bit-identical to the NCCL gather on the check frame)"

    def step_resident(i, last=False):
        o, d = rays[view_of(i)]
        if last and final_peer is not None:
            pipe.submit(o, d, frame_out=final_peer.pointers(rank), hits_out=hit_slots[i], image_width=sc.W)
        else:
            pipe.submit(o, d, out=outs[i % NS], hits_out=hit_slots[i], image_width=sc.W)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    pipe.begin()
    for i in range(args.warmup):
        step_resident(i)
    pipe.join()
    if world > 1:  # NCCL sets up its send/recv channels lazily: do it outside the timed region
        P.gather_frame(torch.cat([out["rgb"], out["opacity"], out["depth"]], dim=1), [N] * world, dst=0)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    lib.qf_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    pipe.begin()
    for i in range(args.steps):
        step_resident(args.warmup + i, last=(i == args.steps - 1))
    pipe.join()
    if world > 1:  # the final image gather (north star): last frame of every rank to rank 0
        if final_peer is not None:
            final_peer.sync()
        else:
            P.gather_frame(torch.cat([out["rgb"], out["opacity"], out["depth"]], dim=1), [N] * world, dst=0)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    import ctypes as C
    ms3o = (C.c_double * 3)()
    ncho = C.c_int64()
    _lib.check(lib.qf_profile_read(ms3o, C.byref(ncho)), "qf_profile_read")      # stage times inside the (overlapped) timed region
    # the dominant kernel timed alone (single stream, same process, same inputs) for the roofline
    iso_steps = min(args.steps, 10)
    iso_hits = torch.zeros((iso_steps, 1), dtype=torch.int32, device=dev)
    for i in range(iso_steps):
        o_, d_ = rays[view_of(args.warmup + i)]
        sc.render(o_, d_, out=out, hits_out=iso_hits[i], image_width=sc.W)
    ms3 = (C.c_double * 3)()
    nch = C.c_int64()
    _lib.check(lib.qf_profile_read(ms3, C.byref(nch)), "qf_profile_read")
    lib.qf_profile_enable(0)
    hits_acc = hit_slots[args.warmup:].to(torch.int64).sum().reshape(1)
    total_hits = hits_acc.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_hits, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    value = N * world * args.steps / (ms_total * 1e-3)
    if final_peer is not None:
        final_peer.close()

    # ---- e2e: reference-facing call with HOST buffers (pinned rays in, image out), copies inside the timed region.
    # H2D, D2H and NB compute streams over NB slots: H2D of frame i+1 and D2H of frame i-1 overlap the render of frame i.
    n_host = min(n_views, 8)
    host_rays = [(rays[v][0].cpu().pin_memory(), rays[v][1].cpu().pin_memory()) for v in range(n_host)]
    NB = 4          # slots in flight: with 3 a render waited for the D2H of its slot (r2i: 0.267 -> 0.238 ms per frame, = render only)
    host_out = [dict(rgb=torch.empty((N, 3)).pin_memory(), opacity=torch.empty((N, 1)).pin_memory(),
                     depth=torch.empty((N, 1)).pin_memory()) for _ in range(NB)]
    d_in = [(torch.empty((N, 3), device=dev), torch.empty((N, 3), device=dev)) for _ in range(NB)]
    d_out = [dict(rgb=torch.empty((N, 3), device=dev), opacity=torch.empty((N, 1), device=dev),
                  depth=torch.empty((N, 1), device=dev)) for _ in range(NB)]
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    s_cmps = [torch.cuda.Stream(dev) for _ in range(NB)]
    ev_in = [torch.cuda.Event() for _ in range(NB)]       # H2D of slot done
    ev_cmp = [torch.cuda.Event() for _ in range(NB)]      # render of slot done (inputs free, outputs ready)
    ev_out = [torch.cuda.Event() for _ in range(NB)]      # D2H of slot done (device outputs free)

    def run_e2e(n_steps, first):
        for j in range(n_steps):
            i, b = first + j, j % NB
            ho, hd = host_rays[view_of(i) % n_host]
            with torch.cuda.stream(s_h2d):
                s_h2d.wait_event(ev_cmp[b])               # previous render from this slot finished reading it
                d_in[b][0].copy_(ho, non_blocking=True)
                d_in[b][1].copy_(hd, non_blocking=True)
                ev_in[b].record(s_h2d)
            s_cmp = s_cmps[b]
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[b])
                s_cmp.wait_event(ev_out[b])               # previous D2H from this slot finished
                sc.render(d_in[b][0], d_in[b][1], out=d_out[b], image_width=sc.W)
                ev_cmp[b].record(s_cmp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_cmp[b])
                for k in ("rgb", "opacity", "depth"):
                    host_out[b][k].copy_(d_out[b][k], non_blocking=True)
                ev_out[b].record(s_d2h)
        for st in (s_h2d, s_d2h, *s_cmps):
            torch.cuda.current_stream(dev).wait_stream(st)

    run_e2e(max(args.warmup, 3), 0)
    barrier()
    ev0.record()
    run_e2e(args.steps, args.warmup)
    ev1.record()
    barrier()
    ms_e = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    e2e_value = N * world * args.steps / (float(ms_e.item()) * 1e-3)
    # ---- e2e, the reference's real eval input: a 48-byte camera pose on the HOST per frame (its loader builds the rays on
    # the device, nerf_synthetic.py:289-378), image (rgb + depth, what the eval loop takes back, train_finetune.py:620-626)
    # device->host into pinned memory every step.  NB slots / compute streams so the D2H of frame i overlaps the next frames.
    pose_host = [dict(rgb=torch.empty((N, 3)).pin_memory(), depth=torch.empty((N, 1)).pin_memory()) for _ in range(NB)]
    from quadraturefields_b200.utils import MeshRenderer
    pose_renderers = [MeshRenderer(sc.mesh_intersect, radiance_field=sc.radiance_field) for _ in range(NB)]   # private ray scratch per slot

    def run_e2e_pose(n_steps, first):
        for j in range(n_steps):
            i, b = first + j, j % NB
            s_cmp = s_cmps[b]
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_out[b])               # previous D2H from this slot finished
                pose_renderers[b].render_pose(sc.poses[view_of(i) % len(sc.poses)], sc.W, sc.H, sc.focal, sc.cx, sc.cy, out=d_out[b])
                ev_cmp[b].record(s_cmp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_cmp[b])
                for k in ("rgb", "depth"):
                    pose_host[b][k].copy_(d_out[b][k], non_blocking=True)
                ev_out[b].record(s_d2h)
        for st in (s_d2h, *s_cmps):
            torch.cuda.current_stream(dev).wait_stream(st)

    run_e2e_pose(max(args.warmup, 3), 0)
    barrier()
    ev0.record()
    run_e2e_pose(args.steps, args.warmup)
    ev1.record()
    barrier()
    ms_p = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_p, op=dist.ReduceOp.MAX)
    e2e_pose_value = N * world * args.steps / (float(ms_p.item()) * 1e-3)
    # ---- e2e, PNG-ready flavour: what the reference's eval loop finally keeps of a frame on the host are the uint8 colour and
    # depth images it writes (train_finetune.py:639-646); made on the device (MeshRenderer.frame_to_uint8) they are 4 B / ray
    # instead of 16.  Secondary key: with 8 ranks the fp32 images of the headline e2e saturate the host's PCIe ingest.
    u8_dev = [(torch.empty((N, 3), dtype=torch.uint8, device=dev), torch.empty((N,), dtype=torch.uint8, device=dev)) for _ in range(NB)]
    u8_host = [(torch.empty((N, 3), dtype=torch.uint8).pin_memory(), torch.empty((N,), dtype=torch.uint8).pin_memory()) for _ in range(NB)]

    def run_e2e_u8(n_steps, first):
        for j in range(n_steps):
            i, b = first + j, j % NB
            s_cmp = s_cmps[b]
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_out[b])               # previous D2H from this slot finished
                r = pose_renderers[b]
                r.render_pose(sc.poses[view_of(i) % len(sc.poses)], sc.W, sc.H, sc.focal, sc.cx, sc.cy, out=d_out[b])
                r.frame_to_uint8(d_out[b], rgb8=u8_dev[b][0], depth8=u8_dev[b][1])
                ev_cmp[b].record(s_cmp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_cmp[b])
                u8_host[b][0].copy_(u8_dev[b][0], non_blocking=True)
                u8_host[b][1].copy_(u8_dev[b][1], non_blocking=True)
                ev_out[b].record(s_d2h)
        for st in (s_d2h, *s_cmps):
            torch.cuda.current_stream(dev).wait_stream(st)

    run_e2e_u8(max(args.warmup, 3), 0)
    barrier()
    ev0.record()
    run_e2e_u8(args.steps, args.warmup)
    ev1.record()
    barrier()
    ms_u = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_u, op=dist.ReduceOp.MAX)
    e2e_u8_value = N * world * args.steps / (float(ms_u.item()) * 1e-3)
    del pose_renderers, u8_dev, u8_host
    train = None if args.no_train else run_train_steps(args, sc, dev, rank, world, barrier)
    field_train = None if args.no_train else run_field_train_steps(args, sc, dev, rank, world, barrier)
    field_train_occ = None if args.no_train else run_field_train_occgrid_steps(args, sc, dev, rank, world, barrier)
    up2 = run_upsample2_frames(args, sc, dev) if args.config == "c2" else None
    clk = clocks.stop() if clocks else None
    dense = run_dense_c2(args, dev) if (args.config == "c2" and world == 1 and not args.no_big) else None
    # free the c2 scene's resident ray sets (3 GB) before the large configs are built
    del rays, host_rays, pipe, outs, d_in, d_out, host_out
    torch.cuda.empty_cache()
    c4 = None if args.no_big else run_sharded_frame_leg("c4", args, dev, rank, world, barrier, baked=False)
    c5 = None if args.no_big else run_sharded_frame_leg("c5", args, dev, rank, world, barrier, baked=True)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        shade_ms = ms3[1] / max(nch.value, 1)
        hits_per_launch = float(iso_hits.sum().item()) / max(nch.value, 1)
        shade_ms_overlapped = ms3o[1] / max(ncho.value, 1)
        achieved = 512.0 * hits_per_launch / (shade_ms * 1e-3) / 1e9 if shade_ms > 0 else 0.0
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            traffic = next((v for k, v in tj["dram_bytes_per_launch"].items() if k.startswith("ngp_forward")), None)
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 geometry/compositing, f16 table+MLP operands with f32 accumulate", "data": "synthetic",
            "config": {"workload": workload_name(args.config), "rays_per_step_per_gpu": N, "K": sc.K,
                       "triangles": int(sc.faces_np.shape[0]), "hits_per_ray": total_hits.item() / (N * world * args.steps),
                       "l2_policy": f"inputs larger than L2: {n_views} resident ray sets ({n_views * N * 24 / 1e9:.2f} GB) cycled; "
                                    "25 MB table + 2.6 MB BVH are L2-resident by design", "parallelism": f"frames over {world} GPU(s)",
                       "final_image_gather": final_note},
            "ms_per_frame_800x800": ms_total / args.steps if args.config == "c2" else None,
            # headline e2e = the reference's real per-frame input (a host pose; rays are generated on the device as its loader
            # does); the rays-from-host flavour of round 1 stays as `e2e_rays_from_host`
            "e2e": {"value": e2e_pose_value, "unit": UNIT, "h2d_bytes_per_step": 48, "d2h_bytes_per_step": N * 16,
                    "input": "3x4 camera-to-world pose in host memory (48 B, passed as kernel arguments of the ray generator)",
                    "output": "rgb (N,3) + depth (N,1) copied device->pinned host every step"},
            "e2e_png_ready": {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": 48, "d2h_bytes_per_step": N * 4,
                              "output": "uint8 colour (N,3) + uint8 normalised depth (N) — the images the reference's eval loop "
                                        "writes (train_finetune.py:639-646), converted on the device, copied to pinned host every step"},
            "e2e_rays_from_host": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N * 24, "d2h_bytes_per_step": N * 20},
            "gpu_launches": 3 * args.steps,
            "stage_ms_per_step": {"trace": ms3[0] / max(nch.value, 1), "shade": shade_ms, "composite": ms3[2] / max(nch.value, 1),
                                  "note": "each kernel timed alone (single stream)"},
            "stage_ms_per_step_in_timed_region": {"trace": ms3o[0] / max(ncho.value, 1), "shade": shade_ms_overlapped,
                                                  "composite": ms3o[2] / max(ncho.value, 1),
                                                  "note": "frames rotate over the pipeline streams, so kernels of neighbouring frames overlap"},
            "roofline": {"bound": "hbm", "kernel": "ngp_forward_tc_kernel (hash-grid gather + tcgen05 MLPs)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_hit": 512, "hits_per_launch": hits_per_launch, "peak_source": peak_src,
                         "achieved_in_timed_region": (512.0 * hits_per_launch / (shade_ms_overlapped * 1e-3) / 1e9) if shade_ms_overlapped > 0 else 0.0,
                         "note": "kernel timed alone with CUDA events on its stream in the same process (in the timed region it overlaps the "
                                 "next frame's trace kernel); the table is L2-resident, so 'achieved' is an HBM-equivalent gather rate"},
            "clocks": clk,
        }
        line["samples_per_sec"] = total_hits.item() / (ms_total * 1e-3)
        if up2 is not None:
            line["frame_800x800_up_sample2"] = up2
        if dense is not None:
            line["c2_dense"] = dense
        if c4 is not None:
            line["c4"] = c4
        if c5 is not None:
            line["c5"] = c5
        if train is not None:
            line["train"] = train
            line["field_train"] = field_train
            line["field_train_occgrid"] = field_train_occ
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only (torchrun pins OMP_NUM_THREADS=1 anyway)
            threads = os.cpu_count() or 1
            rps, detail, sec = cpu_render_sample(args.cpu_sample, threads, args.config)
            line["cpu_baseline"] = {"value": rps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_sample}x{args.cpu_sample} sub-grid of one frame ({detail.get('rays')} rays, "
                                              f"{detail.get('hits')} hits) in {sec:.1f}s: intersect {detail.get('intersect_s', 0):.1f}s "
                                              f"(OpenMP CPU BVH traversal), field {detail.get('field_s', 0):.2f}s, composite "
                                              f"{detail.get('composite_s', 0):.3f}s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def run_sharded_frame_leg(name, args, dev, rank, world, barrier, baked):
    """BASELINE configs[3] (`c4`: Shelly-shaped 1080p, 1.15 M triangles, K=32, T=2^21, neural field) and configs[4] (`c5`:
    baked spherical-Gaussian textures, 4K): ONE frame per step, its rays sharded over the ranks in 4-row bands dealt
    round-robin (`render_pose(bands=...)`; the reference renders such frames in 160 000-ray splits on one GPU,
    train_finetune.py:590-617, test_baking_texture_images.py:355-371), every rank renders its share from the HOST pose, and
    the (rgb, opacity, depth) shares are gathered to rank 0 with one NCCL gather per frame and put back into image order
    INSIDE the timed region.  Strong
    scaling: the frame is fixed, N grows.  Never fails the main line."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from quadraturefields_b200 import _lib, parallel as P, scene as S
    try:
        lib = _lib.load()
        sc = S.make_scene(name, device=dev, build_field=not baked)
        renderer = sc.baked_renderer if baked else sc.renderer
        N, W, H = sc.n_rays, sc.W, sc.H
        # band-cyclic shares: rank r renders the 4-row bands b with b % world == r (the object sits in the middle of the frame:
        # contiguous bands left the outer ranks idle — N=4, r2i: c4 strong-scaling efficiency 0.61)
        sizes = [P.band_rows(H, r, world) * W for r in range(world)]
        n = sizes[rank]
        out = dict(rgb=torch.empty((n, 3), device=dev), opacity=torch.empty((n, 1), device=dev), depth=torch.empty((n, 1), device=dev))
        band = torch.empty((n, 5), device=dev)
        even = len(set(sizes)) == 1                       # ragged bands (rows not divisible) take parallel.gather_frame's padding path
        recv = [torch.empty((sz, 5), device=dev) for sz in sizes] if (world > 1 and even and rank == 0) else None
        hits = torch.zeros((1,), dtype=torch.int32, device=dev)
        steps, warm = max(3, min(args.steps, 10)), 3
        hit_slots = torch.zeros((warm + steps, 1), dtype=torch.int32, device=dev)
        # N > 1: the frame lives on rank 0 and every rank's composite kernel stores its pixels straight into it over NVLink
        # (parallel.PeerFrame: CUDA IPC peer memory, one 4-byte all-reduce per frame as the fence).  QF_BENCH_PEER_FRAME=0, or
        # a failed mapping / self-check, falls back to one NCCL gather per frame + reassembly.
        peer, peer_note = None, "n/a (one GPU)" if world == 1 else "peer frame disabled (QF_BENCH_PEER_FRAME=0)"
        if world > 1 and os.environ.get("QF_BENCH_PEER_FRAME", "1") != "0":
            ok = torch.ones((1,), device=dev)
            try:
                peer = P.PeerFrame(H, W, dev, dst=0, slots=2)
            except Exception as e:                       # noqa: BLE001 - any failure means "use the collective"
                peer_note = f"unavailable ({type(e).__name__}: {e})"
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) < 1.0:
                peer = None
            else:
                # self-check against the collective path on one frame: same pixels, bit for bit
                renderer.render_pose(sc.poses[0], W, H, sc.focal, sc.cx, sc.cy, hits_out=hits, bands=(rank, world), frame=peer, frame_slot=0)
                peer.sync()
                renderer.render_pose(sc.poses[0], W, H, sc.focal, sc.cx, sc.cy, out=out, hits_out=hits, bands=(rank, world))
                flat = P.gather_frame(torch.cat([out["rgb"], out["opacity"], out["depth"]], dim=1), sizes, dst=0)
                if rank == 0:
                    ref = P.assemble_banded(torch.split(flat, sizes), H, W)
                    got = torch.cat(peer.frame(0), dim=1)
                    if not torch.equal(ref, got):
                        ok.zero_()
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if float(ok.item()) < 1.0:
                    peer_note = "self-check against the NCCL gather FAILED: collective path used"
                    peer.close()
                    peer = None
                else:
                    peer_note = "peer stores over NVLink into rank 0's frame (bit-identical to the NCCL gather on the check frame)"

        def step(i):
            if peer is not None:
                renderer.render_pose(sc.poses[i % len(sc.poses)], W, H, sc.focal, sc.cx, sc.cy, hits_out=hit_slots[i],
                                     bands=(rank, world), frame=peer, frame_slot=i)
                peer.sync()
                return
            renderer.render_pose(sc.poses[i % len(sc.poses)], W, H, sc.focal, sc.cx, sc.cy, out=out, hits_out=hit_slots[i],
                                 bands=(rank, world))
            if world > 1:
                torch.cat([out["rgb"], out["opacity"], out["depth"]], dim=1, out=band)
                if even:
                    dist.gather(band, recv, dst=0)
                    if rank == 0:
                        P.assemble_banded(recv, H, W)                       # the frame, rows back in image order
                else:
                    flat = P.gather_frame(band, sizes, dst=0)
                    if rank == 0:
                        P.assemble_banded(torch.split(flat, sizes), H, W)

        for i in range(warm):
            step(i)
        barrier()
        lib.qf_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(warm + i)
        e1.record()
        barrier()
        ms3 = (C.c_double * 3)()
        nch = C.c_int64()
        _lib.check(lib.qf_profile_read(ms3, C.byref(nch)), "qf_profile_read")
        lib.qf_profile_enable(0)
        ms = P.max_over_ranks(e0.elapsed_time(e1), dev) / steps
        h = hit_slots[warm:].to(torch.int64).sum().reshape(1)
        if world > 1:
            dist.all_reduce(h)
        hits_per_frame = float(h.item()) / steps
        st = [ms3[k] / steps for k in range(3)]            # this rank's band: trace / shade / composite ms per frame
        peak, peak_src = _hbm_peak()
        L = sc.cfg.get("lobes", 0)
        bytes_per_hit = (4 + 6 * L + 36) if baked else 512
        rank_hits = float(hit_slots[warm:].to(torch.int64).sum().item()) / steps
        shade_gbs = bytes_per_hit * rank_hits / (st[1] * 1e-3) / 1e9 if st[1] > 0 else 0.0
        res = {"config": workload_name(name), "n_gpus": world, "scaling": "strong (one frame, 4-row bands dealt round-robin over the ranks)",
               "ms_per_frame": ms, "rays_per_frame": N, "rays_per_sec": N / (ms * 1e-3), "hits_per_ray": hits_per_frame / N,
               "samples_per_sec": hits_per_frame / (ms * 1e-3), "steps": steps, "triangles": int(sc.faces_np.shape[0]), "K": sc.K,
               "bvh_bytes": int(sc.mesh_intersect.rayintersector.info()["device_bytes"]),
               "gather_bytes_per_frame": (N - sizes[0]) * 20 if world > 1 else 0,
               "image_gather": peer_note if (world == 1 or peer is not None) else ("NCCL gather + reassembly; " + peer_note),
               "rank0_stage_ms_per_frame": {"trace": st[0], "shade": st[1], "composite": st[2]},
               "roofline_shade": {"bound": "hbm", "kernel": "baked_shade_kernel" if baked else "ngp_forward_tc_kernel",
                                  "algorithmic_bytes_per_hit": bytes_per_hit, "achieved": shade_gbs, "peak": peak, "unit": "GB/s",
                                  "frac": shade_gbs / peak, "peak_source": peak_src},
               "includes": "pose on the host -> ray generation -> BVH first-K trace -> shade -> composite on every rank's band"
                           + ((" -> pixels stored into rank 0's frame over NVLink by the composite kernel, one 4-byte all-reduce as the fence"
                               if peer is not None else " -> NCCL gather of the bands to rank 0") if world > 1 else "")}
        if peer is not None:
            peer.close()
        del sc, renderer
        torch.cuda.empty_cache()
        return res
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


def run_dense_c2(args, dev):
    """configs[1] with the camera at radius 2.2 instead of 4.03: every ray crosses the shells (h ~ 6 hits / ray instead of
    1.7), the regime SURVEY §8(d) reasons about (512 B and 18.7 kflop per hit sample).  Secondary number, N=1 only."""
    import torch
    from quadraturefields_b200 import scene as S
    from quadraturefields_b200.utils import FramePipeline
    try:
        sc = S.make_scene("c2", device=dev, cam_radius=2.2, views=16)
        N = sc.n_rays
        rays = [sc.rays(v) for v in range(16)]
        NS = 3
        pipe = FramePipeline(sc.renderer, NS)
        outs = [dict(rgb=torch.empty((N, 3), device=dev), opacity=torch.empty((N, 1), device=dev), depth=torch.empty((N, 1), device=dev))
                for _ in range(NS)]
        steps, warm = max(5, min(args.steps, 20)), 5
        hit_slots = torch.zeros((warm + steps, 1), dtype=torch.int32, device=dev)

        def run(n, first):
            pipe.begin()
            for j in range(n):
                o, d = rays[(first + j) % 16]
                pipe.submit(o, d, out=outs[j % NS], hits_out=hit_slots[first + j], image_width=sc.W)
            pipe.join()

        run(warm, 0)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps, warm)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        hits = float(hit_slots[warm:].to(torch.int64).sum().item()) / steps
        return {"ms_per_frame": ms, "rays_per_sec": N / (ms * 1e-3), "hits_per_ray": hits / N, "samples_per_sec": hits / (ms * 1e-3),
                "steps": steps, "camera_radius": 2.2, "l2_policy": "16 resident ray sets (246 MB > L2) cycled"}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


def run_upsample2_frames(args, sc, dev):
    """The reference-faithful variant of configs[1]: the shipped scripts render every "800x800" frame
    at `up_sample=2`, i.e. 1600x1600 = 2.56 M rays, and area-downsample the image by 2 (`cv2.INTER_AREA`,
    train_finetune.py:350,620-627).  Secondary number; `ms_per_step` / `ms_per_frame_800x800` of the main line are the
    plain 800x800 frame BASELINE.json names.  Single stream, no frame overlap.  Never fails the bench."""
    import torch
    try:
        from quadraturefields_b200.datasets import ray_gen
        W = H = 2 * sc.W
        views = [ray_gen.generate_rays(sc.poses[v], W, H, 2.0 * sc.focal, 2.0 * sc.cx, 2.0 * sc.cy, device=dev) for v in range(4)]
        n = W * H
        out = dict(rgb=torch.empty((n, 3), device=dev), opacity=torch.empty((n, 1), device=dev), depth=torch.empty((n, 1), device=dev))
        hits = torch.zeros((1,), dtype=torch.int32, device=dev)

        def frame(i):
            r = views[i % len(views)]
            sc.render(r.origins, r.viewdirs, out=out, hits_out=hits, image_width=W)
            return torch.nn.functional.avg_pool2d(out["rgb"].view(1, H, W, 3).permute(0, 3, 1, 2), 2)     # (1,3,800,800)

        for i in range(3):
            img = frame(i)
        torch.cuda.synchronize(dev)
        steps = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            img = frame(3 + i)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        return {"ms_per_frame": ms, "rays_per_frame": n, "rays_per_sec": n / (ms * 1e-3), "steps": steps, "hits_last_frame": int(hits.item()),
                "image": list(img.shape[-2:]),
                "includes": "1600x1600 trace + shade + composite on one stream, then the 2x2 area downsample to 800x800"}
    except Exception as e:                                             # a secondary number must not cost the main line
        return {"error": f"{type(e).__name__}: {e}"}


def run_field_train_occgrid_steps(args, sc, dev, rank, world, barrier):
    """train_field.py:297-368 with the reference's own sampler: occupancy grid built from the frozen radiance field,
    marcher samples (~2^18 per step, the reference's target_sample_batch_size), weights / reversed weights, quadrature Field
    forward + field_grad, field loss, double backward, Adam.  Reported per SAMPLE (the ray count follows the grid)."""
    import torch
    from quadraturefields_b200 import parallel as P
    from quadraturefields_b200.datasets.utils import Rays
    from quadraturefields_b200.field import Field
    from quadraturefields_b200.occ_grid import OccGridEstimator
    from quadraturefields_b200.utils import train_field_step_occgrid
    rf = sc.radiance_field
    step_size = 5e-3                                                  # train_field.py:195
    est = OccGridEstimator([-1.5] * 3 + [1.5] * 3, resolution=128, levels=1).to(dev)
    est.train()
    torch.manual_seed(7 + rank)
    for it in range(0, 64, 16):                                       # the synthetic field is dense: threshold = mean occupancy
        est.update_every_n_steps(step=it, occ_eval_fn=lambda x: rf.query_density(x) * step_size, occ_thre=1e10)
    net = Field(scale=0.5, precision=16, log2_T=19, L=16, max_res=512, min_res=16, output_dim=1, hidden_size=16,
                num_features=2, back_prop=False, nl="elu").to(dev)
    params = list(net.parameters())
    opt = torch.optim.Adam(params, lr=2e-2, eps=1e-15, fused=True)
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    o_all, d_all = sc.rays(0)
    n_rays = 1 << 16
    steps, warm = max(5, min(args.steps, 10)), 5
    batches = []
    for i in range(steps + warm):
        pi = torch.randint(0, sc.n_rays, (n_rays,), device=dev, generator=g)
        batches.append(Rays(o_all[pi].contiguous(), d_all[pi].contiguous()))
    for p_ in params:
        p_.grad = torch.zeros_like(p_)
    rf.train()
    for i in range(warm):
        train_field_step_occgrid(net, rf, est, batches[i], opt, render_step_size=step_size)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    e0.record()
    samples = 0
    for i in range(steps):
        samples += train_field_step_occgrid(net, rf, est, batches[warm + i], opt, render_step_size=step_size)[1]
    e1.record()
    barrier()
    mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs
    rf.eval()
    ms = P.max_over_ranks(e0.elapsed_time(e1), dev)
    return {"metric": "samples_per_sec_field_train_occgrid", "value": samples * world / (ms * 1e-3), "unit": "samples/s",
            "ms_per_step": ms / steps, "steps": steps, "rays_per_step_per_gpu": n_rays, "samples_per_step": samples / steps,
            "occupied_fraction": float(est.binaries.float().mean()), "cuda_mallocs_in_timed_region": mallocs,
            "includes": "occupancy-grid march + density for visibility culling + frozen field fwd + weights/reversed weights + "
                        "Field fwd with field_grad + loss + double backward + Adam step (no gradient all-reduce in this leg)"}


def run_field_train_steps(args, sc, dev, rank, world, barrier):
    """BASELINE configs[2] with train_field.py semantics (:313-368): frozen radiance field -> weights / reversed weights at
    the samples of a 2^18-ray batch (sampler = the quadrature mesh, the occupancy marcher is out of scope) -> quadrature
    Field forward + field_grad -> field loss -> double backward -> gradient all-reduce -> Adam."""
    import torch
    from quadraturefields_b200 import parallel as P
    from quadraturefields_b200.field import Field
    from quadraturefields_b200.utils import HitTuplePrefetcher, train_field_step
    n = args.train_rays
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    n_views = min(32, len(sc.poses))
    pool = [sc.rays(v) for v in range(n_views)]
    O_all, D_all = torch.stack([p[0] for p in pool]), torch.stack([p[1] for p in pool])
    net = Field(scale=0.5, precision=16, log2_T=19, L=16, max_res=512, min_res=16, output_dim=1, hidden_size=16,
                num_features=2, back_prop=False, nl="elu").to(dev)                  # train_field.py:238-252 (T reduced to 2^19)
    params = list(net.parameters())
    opt = torch.optim.Adam(params, lr=2e-2, eps=1e-15, fused=True)
    steps, warm = max(5, min(args.steps, 20)), 8      # warm-up also fills the side stream's allocator pool
    batches = []
    for i in range(steps + warm + 2):
        vi = torch.randint(0, n_views, (n,), device=dev, generator=g)
        pi = torch.randint(0, sc.n_rays, (n,), device=dev, generator=g)
        batches.append((O_all[vi, pi].contiguous(), D_all[vi, pi].contiguous()))
    for p_ in params:
        p_.grad = torch.zeros_like(p_)
    net.accumulate_grad_in_place = True         # the grid scatter adds into .grad directly
    reduce = (lambda: P.all_reduce_gradients(params, n, n * world)) if world > 1 else None
    # every step trains on a tuple traced two steps earlier and launches the trace of the batch after next on a side stream
    # (the reference's DataLoader workers do the intersection ahead of the step, too): one trace per step, no host wait
    pf = HitTuplePrefetcher(sc.mesh_intersect, ring=4)       # tuples in 4 recycled buffer sets: no allocator calls per step

    def step(i):
        tup = pf.get()
        m = train_field_step(net, sc.radiance_field, sc.mesh_intersect, *batches[i], opt, all_reduce=reduce, tup=tup)[1]
        pf.submit(*batches[i + 2], rays_ready=True)
        return m

    pf.submit(*batches[0], rays_ready=True)
    pf.submit(*batches[1], rays_ready=True)
    for i in range(warm):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    e0.record()
    samples = 0
    for i in range(steps):
        samples += step(warm + i)
    e1.record()
    barrier()
    mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs
    ms = P.max_over_ranks(e0.elapsed_time(e1), dev)
    n_params = sum(p_.numel() for p_ in params)
    return {"metric": "rays_per_sec_field_train_fwd_bwd", "value": n * world * steps / (ms * 1e-3), "unit": "rays/s",
            "ms_per_step": ms / steps, "steps": steps, "rays_per_step_per_gpu": n, "samples_per_ray": samples / (n * steps),
            "cuda_mallocs_in_timed_region": mallocs,
            "params": n_params, "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0,
            "includes": "trace (of the batch after next, side stream) + frozen field fwd + weights/reversed weights + Field fwd with "
                        "field_grad + loss + double backward + grad all-reduce + Adam step"}


def run_train_steps(args, sc, dev, rank, world, barrier):
    """BASELINE configs[2]-shaped step (finetune semantics, train_finetune.py:494-531 without deformation): a batch of
    random rays over all views -> mesh-path render with gradients -> smooth-L1 -> backward (hash table + MLPs) ->
    NCCL all-reduce of the flat gradients -> Adam step.  Returns rays/s (fwd+bwd, all ranks) and ms/step."""
    import torch
    from quadraturefields_b200 import parallel as P
    from quadraturefields_b200.utils import HitTuplePrefetcher, render_train
    n = args.train_rays
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_views = min(32, len(sc.poses))
    pool = [sc.rays(v) for v in range(n_views)]
    O_all, D_all = torch.stack([p[0] for p in pool]), torch.stack([p[1] for p in pool])        # (V, N, 3)
    rf = sc.radiance_field
    params = [rf.mlp_base.params, rf.mlp_head.params]
    # QF_TRAIN_GRAPH=1 replays the step from CUDA graphs (utils.GraphedTrainStep).  Off by default: measured r2 on 2^18 rays,
    # 1.60 ms graphed against 1.52 ms eager — the step is GPU-bound (ngp_backward_kernel 0.63 ms + the next batch's refill
    # trace 0.45 ms sharing the SMs from its side stream), not host-bound, and the fixed capacity adds 8-11 % padding samples
    use_graph = os.environ.get("QF_TRAIN_GRAPH", "0") != "0"
    opt = torch.optim.Adam(params, lr=1e-4, eps=1e-15, fused=True, capturable=use_graph)
    steps, warm = max(5, min(args.steps, 20)), 8      # warm-up also fills the side stream's allocator pool
    extra = 3 if use_graph else 0                    # graph replays between the eager warm-up and the timed region
    batches = []
    for i in range(steps + warm + extra + 2):
        vi = torch.randint(0, n_views, (n,), device=dev, generator=g)
        pi = torch.randint(0, sc.n_rays, (n,), device=dev, generator=g)
        batches.append((O_all[vi, pi].contiguous(), D_all[vi, pi].contiguous(), torch.rand((n, 3), device=dev, generator=g)))
    # the tuple of batch i was traced on a side stream during step i-2 and packed during step i-1 (the reference's
    # DataLoader workers run ahead the same way); every step still performs exactly one trace (of the batch after next)
    pf = HitTuplePrefetcher(sc.mesh_intersect, ring=4)       # tuples in 4 recycled buffer sets: no allocator calls per step

    gs = None            # utils.GraphedTrainStep once captured (after the eager warm-up has shown how many hits a batch has)

    def step(i):
        o, d, target = batches[i]
        tup = pf.get()
        done = False
        if gs is not None:
            try:
                gs.load(tup, d, target)
                gs.step()
                n_hits, done = gs.n_hits, True
            except OverflowError:
                pass                                     # more hits than the captured capacity: this batch runs eagerly
        if not done:
            opt.zero_grad(set_to_none=False)
            rgb, _, _, n_hits = render_train(sc.mesh_intersect, rf, o, d, tup=tup)
            loss = torch.nn.functional.smooth_l1_loss(rgb, target)
            loss.backward()
            P.all_reduce_gradients(params, n, n * world)
            opt.step()
            rf.mark_parameters_changed()                 # (a capturable Adam does not bump the version counters)
        pf.submit(batches[i + 2][0], batches[i + 2][1], rays_ready=True)
        return n_hits

    for p_ in params:
        p_.grad = torch.zeros_like(p_)
    rf.accumulate_grad_in_place = True          # backward adds into .grad directly (no 50 MB zero buffer + autograd add per step)
    pf.submit(batches[0][0], batches[0][1], rays_ready=True)
    pf.submit(batches[1][0], batches[1][1], rays_ready=True)
    warm_hits = [step(i) for i in range(warm)]
    graph_info = None
    if use_graph:
        # the step as two CUDA graphs on fixed-capacity, dummy-padded buffers (utils.GraphedTrainStep): capacity = 8 % above
        # the largest warm-up batch; a batch that still exceeds it falls back to the eager step
        try:
            from quadraturefields_b200.utils import GraphedTrainStep
            cap = (int(max(warm_hits) * 1.08) + 32767) // 32768 * 32768
            reduce = (lambda: P.all_reduce_gradients(params, n, n * world)) if world > 1 else None
            g = GraphedTrainStep(rf, opt, n, cap, sc.mesh_intersect.render_step_size, all_reduce=reduce)
            tup0 = sc.mesh_intersect.sampling_raytrace(batches[0][1], batches[0][0])
            g.load(tup0, batches[0][1], batches[0][2])
            g.capture()
            gs = g
            graph_info = {"capacity": cap, "dummy_rays": g.D}
        except Exception as e:
            graph_info = {"error": f"{type(e).__name__}: {e}"}
        for i in range(extra):                       # replays (or eager steps, had the capture failed) outside the timed region
            step(warm + i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    e0.record()
    hits = 0
    for i in range(steps):
        hits += step(warm + extra + i)
    e1.record()
    barrier()
    mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs
    rf.accumulate_grad_in_place = False
    rf.mark_parameters_changed()
    ms = P.max_over_ranks(e0.elapsed_time(e1), dev)
    n_params = sum(p_.numel() for p_ in params)
    return {"metric": "rays_per_sec_train_fwd_bwd", "value": n * world * steps / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms / steps,
            "cuda_mallocs_in_timed_region": mallocs,
            "steps": steps, "rays_per_step_per_gpu": n, "hits_per_ray": hits / (n * steps), "params": n_params,
            "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0, "cuda_graph": graph_info,
            "includes": "trace (of the batch after next, side stream) + field fwd + composite + loss + field/composite bwd + grad all-reduce "
                        "+ fused Adam step" + ("; fwd/loss/bwd and the Adam step replayed from two CUDA graphs on fixed-capacity, "
                                               "dummy-padded hit buffers" if (graph_info and "error" not in graph_info) else "")}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
