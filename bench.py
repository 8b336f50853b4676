#!/usr/bin/env python
"""bench.py -- Msamples/s of the per-pixel radiance loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- procedural Cornell box (diffuse walls +
2 spheres, one area light), 1920x1080, 64 spp, max depth 5, G19_MODE_PATH. One "step" = one
whole frame = 132 710 400 camera paths. Strong scaling: the frame's 32x32 tiles are interleaved
over the ranks (no data-path collective); the framebuffer is gathered to rank 0 over NCCL and
that gather is inside every timed step.

  value   whole-job Msamples/s, outputs resident in HBM on rank 0 (device timed, CUDA events,
          max over ranks)
  e2e     same metric through the host-buffer C-ABI call g19_render (N=1) / render+gather+
          D2H into pinned host memory (N>1): the reference-facing RayTracer::run equivalent
  roofline  dominant kernel class, algorithmic bytes / CUDA-event time
  cpu_baseline  this repo's FP64 path oracle (same work per sample) on all host cores, on a
          bounded crop of the same frame; `literal_reference` = the unmodified reference
          (oracle/_ref) on the depth-0, 1-spp slice it is able to execute

--impl reference times the CPU implementation (see reference_arm()).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, SPP, DEPTH, SEED = 1920, 1080, 64, 5, 0
WORKLOAD = "cornell_box_1920x1080_64spp_depth5 (BASELINE.json configs[1])"
METRIC, UNIT = "Msamples/s", "Msamples/s"
SAMPLES_PER_STEP = W * H * SPP


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def cpu_path_oracle(threads, budget_s=12.0):
    """The like-for-like CPU path tracer (oracle/path_oracle.c, FP64, brute force over the 14
    primitives) on a centred crop of the SAME frame at the SAME spp/depth; the crop grows until
    the run takes a few seconds. Returns (Msamples/s, description)."""
    g19 = importlib.import_module("2019global_b200")
    from oracle import binding
    orc = binding.CheckerLib("oracle")
    sc, cam, light = g19.Octree.builtin(g19.abi.SCENE_CORNELL, w=W, h=H)
    chk = orc.scene(sc.min, sc.max, sc.entities())
    cw, ch = 96, 54
    while True:
        x0, y0 = (W - cw) // 2, (H - ch) // 2
        t = time.perf_counter()
        binding.path_render(chk, cam, W, H, SPP, DEPTH, seed=SEED, window=(x0, y0, x0 + cw, y0 + ch), threads=threads)
        dt = time.perf_counter() - t
        if dt >= budget_s / 4 or (cw, ch) == (W, H):
            break
        cw, ch = (cw * 2, ch * 2) if cw * 2 <= 1536 else (W, H)  # ... 768x432, 1536x864, then the whole frame
    n = cw * ch * SPP
    return n / dt / 1e6, "centred %dx%d crop of the 1920x1080 frame, %d spp, depth %d (%d paths, %.1f s)" % (
        cw, ch, SPP, DEPTH, n, dt)


def cpu_literal_reference(threads):
    """The UNMODIFIED reference (oracle/_ref) on what it can execute: 1 primary ray per pixel,
    direct shading, on rows of the same Cornell scene. Returns dict or None."""
    from oracle import binding
    if not binding.available("ref"):
        return None
    g19 = importlib.import_module("2019global_b200")
    ref = binding.CheckerLib("ref")
    sc, cam, light = g19.Octree.builtin(g19.abi.SCENE_CORNELL, w=W, h=H)
    chk = ref.scene(sc.min, sc.max, sc.entities())
    rows = 8 * max(1, threads)
    y0 = (H - rows) // 2
    t = time.perf_counter()
    chk.trace(cam, light, W, H, y0=y0, y1=y0 + rows, want=("ids",), threads=threads)
    dt = time.perf_counter() - t
    return {"value": rows * W / dt / 1e6, "unit": UNIT, "cores": threads,
            "sample": "%d rows of the 1080p Cornell frame, depth-0 samples (1 primary ray + direct shade): "
                      "all the reference can execute" % rows}


def reference_arm(args):
    """bench.py --impl reference. The reference's CPU implementation of the path on the box's
    host cores. The compiled reference (oracle/_ref) cannot run this workload -- it has no spp,
    bounces or area light (raytracer.h:32-86) -- so the arm times this repo's CPU port of the
    SAME path-traced workload (kind "port", all host threads) and reports the literal
    reference's depth-0 rate beside it."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    vals, sample = [], ""
    for i in range(args.warmup + args.steps):
        v, sample = cpu_path_oracle(threads, budget_s=8.0 if args.steps > 1 else 16.0)
        if i >= args.warmup:
            vals.append(v)
        if i == 0 and args.warmup + args.steps > 4:
            pass
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * SAMPLES_PER_STEP / (value * 1e6),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "spp": SPP, "max_depth": DEPTH},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "literal_reference": cpu_literal_reference(threads),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "ms_per_step is extrapolated from the bounded crop to the full frame"}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp-per-pass", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather", default="frame", choices=["frame", "nccl"],
                    help="frame: every rank's resolve kernel stores its pixels straight into rank 0's frame over "
                         "NVLink (fused gather, no collective); nccl: compact tile arrays + torch.distributed.gather")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    g19 = importlib.import_module("2019global_b200")
    g19dist = importlib.import_module("2019global_b200.dist")
    abi = g19.abi
    world, rank, local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lib2019global_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    sc, cam, light = g19.Octree.builtin(abi.SCENE_CORNELL, w=W, h=H)
    rt = g19.RayTracer(cam, light, device=local)
    rt.setScene(sc)
    rt.start()
    stream = torch.cuda.current_stream().cuda_stream

    # One payload per rank = its compact tile arrays [float radiance | RGB888], padded to rank 0's
    # length: ONE gather (grouped ncclSend/ncclRecv) per frame carries both outputs.
    pad = g19dist.padded_len(W, H, world)
    rad_bytes = pad * 3 * 4
    payload = torch.zeros(rad_bytes + pad * 3, dtype=torch.uint8, device=dev)
    f_rad = torch.zeros(H * W * 3, dtype=torch.float32, device=dev) if rank == 0 else None
    f_rgb = torch.zeros(H * W * 3, dtype=torch.uint8, device=dev) if rank == 0 else None
    host_rgb = torch.empty(H * W * 3, dtype=torch.uint8).pin_memory() if rank == 0 else None

    def params(profile):
        return rt.params(W, H, mode=abi.MODE_PATH, spp=SPP, max_depth=DEPTH, seed=SEED, rank=rank, world=world,
                         spp_per_pass=args.spp_per_pass, profile=profile)

    def untile_both(r, buf, _frame):
        rt.untile(W, H, r, world, t_rad=buf.data_ptr(), t_rgb=buf.data_ptr() + rad_bytes, d_rad=f_rad.data_ptr(),
                  d_rgb=f_rgb.data_ptr(), stream=stream)

    shared = g19dist.shared_frame(rt, W, H) if args.gather == "frame" else None

    def step(profile=0, read_host=False):
        if shared is not None:
            # fused: resolve stores into rank 0's frame (peer mapping), then a system-scope signal;
            # rank 0 enqueues a wait -- no collective, no staging copy, no host round trip
            shared.render(params(profile), stream=stream)
            if rank == 0:
                shared.wait(world, stream=stream)
                if read_host:
                    shared.read(rgb=host_rgb.data_ptr(), stream=stream)
                shared.release(stream=stream)
            return
        rt.render_tiles(params(profile), t_rad=payload.data_ptr(), t_rgb=payload.data_ptr() + rad_bytes, stream=stream)
        g19dist.gather_frame(payload, W, H, 3, f_rgb, untile_both)
        if read_host and rank == 0:
            host_rgb.copy_(f_rgb, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step()
    # ---- timed region: K steps, device clock, barrier + synchronize on both sides -------------
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        step()
        launches += 0
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.stop()
    st = rt.stats()
    # this library's kernels inside the timed region: every rank's render + rank 0's wait/release
    # (frame) or its untile per payload (nccl)
    launches = (sum_over_ranks(float(st.kernel_launches)) + (2 if shared is not None else world)) * args.steps
    value = SAMPLES_PER_STEP / (ms * 1e-3) / 1e6

    # ---- per-kernel-class CUDA-event times (same steps, event brackets on) --------------------
    barrier()
    cls_ms = [0.0] * 8
    agg = {"extend": 0, "shadow": 0, "shade": 0, "shade_first": 0, "lit": 0, "rad_stores": 0, "samples": 0,
           "launch": [0] * 8}
    for _ in range(args.steps):
        step(profile=1)
        torch.cuda.synchronize()
        s = rt.stats()
        for k in range(8):
            cls_ms[k] += s.class_ms[k]
            agg["launch"][k] += s.class_launches[k]
        agg["extend"] += s.extend_segments
        agg["shadow"] += s.shadow_segments
        agg["shade"] += s.shade_calls
        agg["shade_first"] += s.shade_calls_first
        agg["lit"] += s.lit_samples
        agg["rad_stores"] += s.radiance_stores
        agg["samples"] += s.samples
    barrier()
    # algorithmic HBM bytes per class, this rank. Flat scenes keep DENSE vertex
    # records in the material queues, four float4 planes: (radiance so far | slot) (hit point | primitive)
    # (direction | pixel) (throughput | sample); the camera segment neither writes nor reads the fourth
    E, S0, C, C0, LIT, ST = agg["extend"], agg["samples"], agg["shade"], agg["shade_first"], agg["lit"], agg["rad_stores"]
    npix_local = g19.engine.tile_pixels(W, H, rank, world)
    bytes_cls = {
        "raygen_extend": 48 * C0,                       # record written per shaded camera hit
        "bounce": (48 * C0 + 64 * (C - C0)              # record read per shaded vertex
                   + 64 * (C - C0)                      # record written per continuation hit (= vertices shaded later)
                   + 16 * ST),                          # radiance delivered once per path that ends in this kernel (one float4)
        "accumulate": 16 * S0 + 24 * npix_local * max(1, agg["launch"][abi.K_ACCUM]),  # float4 read per path + accum read-modify-write
    }
    ms_cls = {"raygen_extend": cls_ms[abi.K_EXTEND], "bounce": cls_ms[abi.K_SHADE], "accumulate": cls_ms[abi.K_ACCUM]}
    top = max(ms_cls, key=lambda k: ms_cls[k])
    peak, peak_src = measured_peak()
    n_launch = {"raygen_extend": agg["launch"][abi.K_EXTEND], "bounce": agg["launch"][abi.K_SHADE],
                "accumulate": agg["launch"][abi.K_ACCUM]}
    achieved = bytes_cls[top] / (ms_cls[top] * 1e-3) / 1e9 if ms_cls[top] > 0 else 0.0
    traffic, issue_pct, issue = None, None, None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tj = json.load(f)
            traffic = tj.get(top)
            issue_pct = tj.get("issue_active_pct", {}).get(top)
            # the issue-slot roofline (SURVEY.md 8(d) "issue bound"): warp instructions per launch (ncu, same
            # pass size as this run's) over the live launch time, against SMs x 4 schedulers x SM clock
            winst = tj.get("warp_inst_per_launch", {}).get(top)
            sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
            props = torch.cuda.get_device_properties(local)
            if winst and ms_cls[top] > 0:
                peak_i = props.multi_processor_count * 4 * sm_hz
                ach_i = winst / (ms_cls[top] / max(1, n_launch[top]) * 1e-3)
                issue = {"warp_inst_per_launch": winst, "achieved_Ginst_s": ach_i / 1e9, "peak_Ginst_s": peak_i / 1e9,
                         "frac": ach_i / peak_i, "source": "profiles/roofline_traffic.json (ncu smsp__inst_executed.sum)"}
    except Exception:
        pass
    total_bytes = sum(bytes_cls.values())
    roofline = {
        "bound": "hbm", "kernel": top + "_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "bytes_per_launch": bytes_cls[top] / max(1, n_launch[top]),
        "avg_launch_ms": ms_cls[top] / max(1, n_launch[top]),
        "share_of_step": ms_cls[top] / max(1e-9, sum(cls_ms)),
        "binding_resource": "issue slots (ncu: smsp__issue_active %s%% of peak) -- see DESIGN.md section 5" % issue_pct,
        "issue": issue,
        "note": "per-class times and shares are CUDA-event brackets taken with ONE pass in flight (params.profile); the timed "
                "steps behind `value` keep up to four passes in flight on four streams, so ms_per_step < the sum of the classes",
        "per_class": {k: {"ms_per_step": ms_cls[k] / args.steps, "launches_per_step": n_launch[k] / args.steps,
                          "algorithmic_GB_per_step": bytes_cls[k] / args.steps / 1e9,
                          "GBps": (bytes_cls[k] / (ms_cls[k] * 1e-3) / 1e9) if ms_cls[k] > 0 else None}
                      for k in ms_cls},
        "whole_step": {"algorithmic_GB": total_bytes / args.steps / 1e9, "bytes_per_segment": total_bytes / max(1, E + agg["shadow"]),
                       "segments_per_sample": (E + agg["shadow"]) / max(1, S0),
                       "GBps_over_step": world * total_bytes / args.steps / (ms * 1e-3) / 1e9,
                       "frac_of_peak_per_gpu": total_bytes / args.steps / (ms * 1e-3) / 1e9 / peak},
    }

    # ---- e2e: the reference-facing call with HOST buffers ---------------------------------------
    cam_bytes = 7 * 8 + 3 * 8 + 10 * 4  # g19_camera + light + g19_params
    if world == 1:
        host_np = host_rgb.numpy().reshape(H, W, 3)

        def e2e_step():
            rt.run(W, H, mode=abi.MODE_PATH, want=("rgb",), out={"rgb": host_np}, spp=SPP, max_depth=DEPTH, seed=SEED,
                   spp_per_pass=args.spp_per_pass)
    else:
        def e2e_step():
            step(read_host=True)
            torch.cuda.synchronize()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    e2e = {"value": SAMPLES_PER_STEP / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": cam_bytes, "d2h_bytes_per_step": W * H * 3,
           "api": "g19_render (host RGB888 out)" if world == 1 else (
               "g19_render_to_frame on every rank (peer stores into rank 0's frame over NVLink) + g19_frame_wait + "
               "g19_frame_read into pinned host memory on rank 0" if shared is not None else
               "g19_render_tiles_device + NCCL gather + D2H on rank 0")}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            v, sample = cpu_path_oracle(threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                   "literal_reference": cpu_literal_reference(threads)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "spp": SPP, "max_depth": DEPTH, "seed": SEED,
                       "mode": "PATH", "tiles": "32x32 interleaved, rank = tile % world",
                       "gather": ("shared frame: resolve kernels store into rank 0's HBM over NVLink, no collective"
                                  if shared is not None else "nccl gather of compact tile arrays"),
                       "l2": "per-step wavefront state streams (%.1f GB algorithmic) far exceed the 126 MB L2; "
                             "no flush needed" % (world * total_bytes / args.steps / 1e9),
                       "frame_time_1080p_64spp_ms": ms},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if shared is not None:
        torch.cuda.synchronize()
        if rank == 0 and shared.timeouts():
            sys.stderr.write("bench.py: %d device-side frame spins timed out\n" % shared.timeouts())
        if world > 1:
            dist.barrier()
        shared.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
