#!/usr/bin/env python
"""bench.py — end-to-end Canny throughput on the BASELINE.json workload.

Workload (config.workload): BASELINE configs[2], "batch of 512 synthetic 3840x2160 frames, sigma=1.4",
thresholds 20/60, frame-sharded: every rank owns `--frames` frames (weak scaling, no data-path
collective — the frames are independent units).  A STEP is one pass of the whole hot path
(blur -> Sobel/direction -> NMS -> hysteresis) over the rank's batch.

  value     Mpix/s, whole job, inputs and outputs resident in HBM (u8 gray in, u8 0/255 edge map out).
  e2e       same metric through the public host API (b200_canny_batch_host) with PINNED HOST buffers:
            H2D of every frame and D2H of every edge map are inside the timed region.
  roofline  dominant kernel (the fused front kernel): algorithmic bytes (2 B/px: u8 in + u8 out)
            per launch / its CUDA-event launch duration, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the reference's own CPU path (oracle/_ref, compiled unmodified) — or the C port when the
            prebuilt reference is absent — on a bounded sample of the same frames, on this box's cores.

`--impl reference` times that CPU path instead (all host threads, bounded sample per step).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SIGMA, LO, HI = 1.4, 20, 60
ALG_BYTES_PER_PX = 2.0  # u8 gray read + u8 edge write


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=512, help="frames per GPU (weak scaling)")
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--kind", type=int, default=0, help="0 shapes, 1 uniform noise, 2 constant")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the CPU sample (0 = one per core, <= 32)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="frames", choices=["frames", "bands"],
                    help="frames: batch of frames per GPU (BASELINE configs[2], the default); bands: ONE image of --band-height x "
                         "--band-width split into row bands across the GPUs (configs[4]; strong scaling, NCCL halo + label exchange)")
    ap.add_argument("--band-height", type=int, default=32768)
    ap.add_argument("--band-width", type=int, default=32768)
    ap.add_argument("--sigma", type=float, default=SIGMA)
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation, frame-parallel over host threads
# ----------------------------------------------------------------------------------------------------
def cpu_impl():
    from oracle.bindings import Oracle, Ref
    if Ref.available():
        return Ref(), "reference"
    return Oracle(), "port"


def cpu_run(frames_np, impl, threads, rounds=1):
    """Runs the CPU path on every frame `rounds` times, `threads` at a time (ctypes releases the GIL). Returns wall seconds."""
    n = frames_np.shape[0]
    idx = iter([i % n for i in range(n * rounds)])
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                i = next(idx, None)
            if i is None:
                return
            if hasattr(impl, "lib") and impl.prefix == "ref_":
                impl.lib.ref_canny(C.c_void_p(frames_np[i].ctypes.data), C.c_float(SIGMA), LO, HI, frames_np.shape[1],
                                   frames_np.shape[2], None)
            else:
                impl.canny(frames_np[i], SIGMA, LO, HI)

    ts = [threading.Thread(target=work) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0


def cpu_sample(a):
    import canny_edge_b200 as cb
    cores = os.cpu_count() or 1
    n = a.cpu_frames if a.cpu_frames > 0 else min(cores, 32)
    frames = cb.synth_host(n, a.height, a.width, kind=a.kind, seed=1234, first_frame=0)
    return frames, min(cores, n)


def run_reference(a, rank, world):
    if rank != 0:
        return
    impl, kind = cpu_impl()
    frames, threads = cpu_sample(a)
    t_one = cpu_run(frames, impl, threads)                      # warm-up pass, also calibrates the sample
    # a step = `rounds` passes over the sample frames, sized so that the whole --steps run stays within ~2 minutes
    rounds = max(1, min(8, int(100.0 / max(a.steps, 1) / max(t_one, 1e-3))))
    px = frames.size * rounds
    t = 0.0
    for _ in range(a.steps):
        t += cpu_run(frames, impl, threads, rounds)
    val = px * a.steps / t / 1e6
    sample = (f"{frames.shape[0] * rounds} frames {a.width}x{a.height} per step ({frames.shape[0]} distinct frames of the GPU arm's generator, "
              f"{rounds} passes), one frame per thread")
    print(json.dumps({
        "impl": "reference", "metric": "end-to-end Canny Mpix/s (4K batch)", "value": round(val, 3), "unit": "Mpix/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": round(1e3 * t / a.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i16 (reference CPU arithmetic)",
        "data": "synthetic",
        "config": {"workload": f"batch of {a.width}x{a.height} synthetic frames, sigma={SIGMA}, thresholds {LO}/{HI} (BASELINE configs[2]); bounded CPU sample",
                   "frames_per_step": int(frames.shape[0] * rounds), "height": a.height, "width": a.width},
        "cpu_baseline": {"value": round(val, 3), "unit": "Mpix/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": round(val, 3), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML)
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def run_b200(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import canny_edge_b200 as cb
    from canny_edge_b200._lib import check, load

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = load()
    ctx = cb.Context(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    n, h, w = a.frames, a.height, a.width
    px = n * h * w
    d_in = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_out = torch.empty_like(d_in)
    # frame f of rank r is global frame r*n + f: every rank works on different frames
    check(lib.b200_synth_device(ctx.handle, d_in.data_ptr(), n, h, w, a.kind, 1234, rank * n))
    torch.cuda.synchronize()

    def step():
        cb.canny_batch_device_ptr(ctx, d_in.data_ptr(), n, h, w, SIGMA, LO, HI, d_out.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.kernel_launches - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * px * a.steps / (ms * 1e-3) / 1e6

    edges = C.c_ulonglong()
    check(lib.b200_count_edges_device(ctx.handle, d_out.data_ptr(), px, C.byref(edges)))

    # ---- roofline of the dominant kernel, measured live with CUDA events on the launching stream ----
    ms5, cnt5 = (C.c_float * 5)(), (C.c_int * 5)()
    check(lib.b200_profile_stages_device(ctx.handle, d_in.data_ptr(), n, h, w, C.c_float(SIGMA), LO, HI, d_out.data_ptr(), ms5, cnt5))
    # category 1 = labelling (list-driven link kernel, or tile-local union-find on dense maps), 2 = tile-border merge (dense maps
    # only), 3 = resolve (weak pixels -> 0 / 255)
    names = ["front", "hyst_label", "hyst_merge", "hyst_resolve", "other"]
    total_k = sum(ms5)
    stage = {names[i]: {"ms": round(ms5[i], 3), "launches": cnt5[i], "share": round(ms5[i] / total_k, 4) if total_k else None}
             for i in range(4)}
    peak, peak_src = peaks()
    front_launches = max(cnt5[0], 1)
    bytes_per_launch = ALG_BYTES_PER_PX * px / front_launches
    achieved = bytes_per_launch / (ms5[0] / front_launches * 1e-3) / 1e9 if ms5[0] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "front2_kernel (blur + Sobel + NMS + thresholds, u8 in -> u8 class map)", "achieved": round(achieved, 2), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 5), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_px": ALG_BYTES_PER_PX, "launch_ms": round(ms5[0] / front_launches, 4),
                "note": "issue-bound, not HBM-bound: the reference's rounding order costs ~47 un-fusable FP32 lane-instructions per pixel "
                        "for the 11-tap separable blur alone (DESIGN.md 4.1, profiles/)",
                "whole_pipeline_frac": round(value / world * 1e6 * ALG_BYTES_PER_PX / 1e9 / peak, 5), "stages": stage}
    prof = ROOT / "profiles" / "traffic.json"
    if prof.exists():
        try:
            # measured on a 4-frame launch (ncu --set full); scaled to this run's launch size
            per_px = json.loads(prof.read_text()).get("front_kernel_dram_bytes_per_px")
            roofline["traffic"] = int(per_px * px / front_launches) if per_px else None
        except Exception:
            pass

    out = {
        "metric": "end-to-end Canny Mpix/s (4K batch)", "value": round(value, 1), "unit": "Mpix/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 blur (reference roundings) + int16/int32 gradient/NMS + u8 labels", "data": "synthetic",
        "config": {"workload": f"batch of {n} synthetic {w}x{h} frames per GPU, sigma={SIGMA}, thresholds {LO}/{HI} (BASELINE configs[2])",
                   "frames_per_gpu": n, "frames_total": n * world, "height": h, "width": w, "sigma": SIGMA, "min_val": LO, "max_val": HI,
                   "generator": ["shapes", "noise", "const"][a.kind], "sharding": f"frames x{world} (no data-path collective)",
                   "l2": "inputs (%.2f GB per GPU) far exceed the 126 MB L2; no flush needed" % (px / 1e9)},
        "clocks": clocks, "gpu_launches": int(launches), "edge_fraction": round(edges.value / px, 6), "roofline": roofline,
    }

    # ---- end to end through the host API: pinned host in, pinned host out ----
    if not a.no_e2e:
        # pinned host buffers for the whole per-rank batch (2 x 4.25 GB at the default size); if the host cannot pin that much
        # (N ranks share one host) every rank falls back to the same smaller number of frames
        ne = n
        while True:
            try:
                h_in = torch.empty((ne, h, w), dtype=torch.uint8, pin_memory=True)
                h_out = torch.empty((ne, h, w), dtype=torch.uint8, pin_memory=True)
                ok = 1
            except RuntimeError:
                h_in = h_out = None
                ok = 0
            if world > 1:
                t = torch.tensor([ok], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                ok = int(t.item())
            if ok or ne <= 8:
                break
            h_in = h_out = None
            ne //= 2
        if not ok:
            raise SystemExit("bench.py: cannot pin host memory for the e2e run")
        h_in.copy_(d_in[:ne])  # same frames as the device-resident run
        torch.cuda.synchronize()
        px_e = ne * h * w

        def e2e_step():
            check(lib.b200_canny_batch_host(ctx.handle, h_in.data_ptr(), ne, h, w, C.c_float(SIGMA), LO, HI, h_out.data_ptr()))

        e2e_step()
        barrier()
        b0h, b0d = C.c_ulonglong(), C.c_ulonglong()
        check(lib.b200_ctx_transfer_bytes(ctx.handle, C.byref(b0h), C.byref(b0d)))
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(a.e2e_steps):
            e2e_step()  # blocks until the last edge map is back in host memory
        g1.record()
        barrier()
        wall = time.perf_counter() - t0
        ems = max(g0.elapsed_time(g1), wall * 1e3)  # the call blocks: device timeline and host wall clock must agree
        if world > 1:
            t = torch.tensor([ems], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        b1h, b1d = C.c_ulonglong(), C.c_ulonglong()
        check(lib.b200_ctx_transfer_bytes(ctx.handle, C.byref(b1h), C.byref(b1d)))
        same = bool((h_out.view(-1)[:: 4099] == d_out[:ne].cpu().view(-1)[:: 4099]).all()) if px_e < (1 << 33) else None
        out["e2e"] = {"value": round(world * px_e * a.e2e_steps / (ems * 1e-3) / 1e6, 1), "unit": "Mpix/s", "frames_per_gpu": ne,
                      "h2d_bytes_per_step": (b1h.value - b0h.value) // a.e2e_steps,
                      "d2h_bytes_per_step": (b1d.value - b0d.value) // a.e2e_steps, "steps": a.e2e_steps,
                      "ms_per_step": round(ems / a.e2e_steps, 3),
                      "api": "b200_canny_batch_host (pinned host u8 frames in -> pinned host u8 0/255 edge maps out; the maps cross "
                             "PCIe bit-packed and are expanded by the library's host threads inside the timed call)",
                      "matches_device_run": same}
        del h_in, h_out

    # ---- CPU baseline on this box's cores (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not a.no_cpu:
        impl, kind = cpu_impl()
        frames, threads = cpu_sample(a)
        t_one = cpu_run(frames, impl, threads)                  # warm-up pass, also calibrates the sample to ~15 s of wall time
        rounds = max(1, min(32, int(15.0 / max(t_one, 1e-3))))
        secs = cpu_run(frames, impl, threads, rounds)
        out["cpu_baseline"] = {"value": round(frames.size * rounds / secs / 1e6, 3), "unit": "Mpix/s", "cores": threads, "kind": kind,
                               "sample": f"{frames.shape[0] * rounds} frames {w}x{h} ({frames.shape[0]} distinct frames of the same generator, "
                                         f"{rounds} passes), one per thread, {secs:.1f} s wall"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_bands(a, rank, world, local_rank):
    """configs[4]: one --band-height x --band-width image, row-band sharded; every step = halo exchange (NCCL send/recv) +
    band front kernel + band-local labelling + record all-gather + cross-band union + finalisation."""
    import torch
    import torch.distributed as dist

    import canny_edge_b200 as cb
    from canny_edge_b200 import sharded
    from canny_edge_b200._lib import check, load

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = load()
    ctx = cb.Context(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    H, W = a.band_height, a.band_width
    pipe = sharded.BandPipeline(ctx, H, W, rank, world, a.sigma, LO, HI)
    g = pipe.geo
    band = pipe.band_view()   # the rank's rows live inside the pipeline's persistent halo buffer: no per-step copy
    edges = torch.empty((g.rows, W), dtype=torch.uint8, device="cuda")
    check(lib.b200_synth_rows_device(ctx.handle, band.data_ptr(), g.row0, g.rows, W, a.kind, 1234, 0))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        pipe.run(None, edges)
    barrier()
    pipe.timings = {}
    pipe.run(None, edges)         # one untimed step with per-stage events (reported as config.stage_ms)
    stage_ms = {k: round(v, 3) for k, v in pipe.timings.items()}
    pipe.timings = None
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        pipe.run(None, edges)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.kernel_launches - l0
    ms = e0.elapsed_time(e1)
    cnt = C.c_ulonglong()
    check(lib.b200_count_edges_device(ctx.handle, edges.data_ptr(), edges.numel(), C.byref(cnt)))
    tot = torch.tensor([float(cnt.value)], device="cuda", dtype=torch.float64)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.all_reduce(tot)
    px = H * W
    value = px * a.steps / (ms * 1e-3) / 1e6
    peak, peak_src = peaks()
    halo_bytes = 2 * g.halo * W if world > 1 else 0
    if rank == 0:
        print(json.dumps({
            "metric": "end-to-end Canny Mpix/s (row-band sharded image)", "value": round(value, 1), "unit": "Mpix/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 blur (reference roundings) + int16/int32 gradient/NMS + u8 labels", "data": "synthetic",
            "config": {"workload": f"single {W}x{H} synthetic image, sigma={a.sigma}, thresholds {LO}/{HI}, row-band sharded over {world} GPU(s) "
                                   "(BASELINE configs[4])", "height": H, "width": W, "sigma": a.sigma, "bands": world,
                       "band_rows": g.rows, "halo_rows": g.halo, "halo_bytes_per_interior_rank_per_step": halo_bytes,
                       "record_bytes_all_gathered_per_step": world * pipe.n_records * 8,
                       "generator": ["shapes", "noise", "const"][a.kind], "stage_ms_rank0": stage_ms,
                       "l2": "band (%.2f GB per GPU) exceeds the 126 MB L2" % (g.rows * W / 1e9)},
            "clocks": clocks, "gpu_launches": int(launches), "edge_fraction": round(float(tot.item()) / px, 6),
            "roofline": {"bound": "hbm", "kernel": "whole band pipeline", "achieved": round(value * 1e6 * ALG_BYTES_PER_PX / 1e9 / world, 2),
                         "peak": peak, "unit": "GB/s", "frac": round(value * 1e6 * ALG_BYTES_PER_PX / 1e9 / world / peak, 5),
                         "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_px": ALG_BYTES_PER_PX},
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    # rank 0 prints exactly one JSON line on stdout: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION/INFO) out of it
    # (VERSION and WARN both print it; with the variable unset NCCL prints nothing)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN") and not os.environ.get("B200_KEEP_NCCL_DEBUG"):
        os.environ.pop("NCCL_DEBUG")
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if a.workload == "bands":
        run_bands(a, rank, world, local_rank)
        return
    run_b200(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
