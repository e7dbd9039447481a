#!/usr/bin/env python
"""bench.py -- throughput of the EWA-Jinc resampling hot path on B200, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl b200|reference]

A "step" is one pass of the hot path over one batch of `frames_per_step` distinct synthetic frames of the
chosen BASELINE.json config (default: configs[1], 1080p YUV420P8 -> 2160p Jinc36Resize cplace=MPEG2).
Prints ONE JSON line (rank 0):
  value       output Mpixel/s, kernels only, inputs resident in HBM (CUDA events, max over ranks)
  e2e         the same metric through the C ABI's host-frame call (jinc_filter_submit/wait, what the plugin's
              GetFrame uses) with pinned HOST buffers: H2D + kernels + D2H inside the timed region
  roofline    dominant kernel (luma interior) against the FP32-FMA roofline measured on this box
              (avisynth-jincresize_b200/fma_peak), plus roofline_hbm against MEASURED_PEAKS.json
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, its default SIMD path) on this box's host cores, on a
              bounded sample of the same workload
`--impl reference` times only that CPU reference and prints the same line shape with "impl": "reference".
Under torchrun (N>1) every rank drives its own GPU with the same per-GPU batch (weak scaling; frames are
independent, so there is no collective on the data path -- NCCL is used for the barrier and the max only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "avisynth-jincresize_b200"))

from minihost import avs_host as ah
from oracle import ref as oref  # noqa: E402
from jinc_b200 import paths  # noqa: E402

METRIC = "output Mpixel/s"

CONFIGS = {
    1: dict(name="640x360 YV12 -> 1280x720 Jinc36Resize (tap=3)", fmt=ah.YV12, w=640, h=360, tw=1280, th=720,
            fn="Jinc36Resize", kw=dict(), tap=3, frames=128),
    2: dict(name="1920x1080 YUV420P8 -> 3840x2160 Jinc36Resize cplace=MPEG2", fmt=ah.YUV420P8, w=1920, h=1080, tw=3840,
            th=2160, fn="Jinc36Resize", kw=dict(cplace="MPEG2"), tap=3, frames=24),
    3: dict(name="1920x1080 YUV444P16 -> 3840x2160 Jinc64Resize src_left=10.3 src_top=6.7 quant=256", fmt=ah.YUV444P16,
            w=1920, h=1080, tw=3840, th=2160, fn="Jinc64Resize",
            kw=dict(src_left=10.3, src_top=6.7, quant_x=256, quant_y=256), tap=4, frames=8),
    4: dict(name="3840x2160 RGBPS -> 7680x4320 Jinc256Resize (tap=8)", fmt=ah.RGBPS, w=3840, h=2160, tw=7680, th=4320,
            fn="Jinc256Resize", kw=dict(), tap=8, frames=2),
    5: dict(name="7680x4320 YUV420P10 -> 1920x1080 JincResize tap=6 blur=0.9", fmt=ah.YUV420P10, w=7680, h=4320,
            tw=1920, th=1080, fn="JincResize", kw=dict(tap=6, blur=0.9), tap=6, frames=4),
    # not BASELINE configs: the "next" row of SURVEY.md 8(f), ratios that are not exact 2x / 1/n (general kernel)
    6: dict(name="1280x720 YUV420P8 -> 1920x1080 Jinc36Resize (1.5x, general path)", fmt=ah.YUV420P8, w=1280, h=720, tw=1920,
            th=1080, fn="Jinc36Resize", kw=dict(), tap=3, frames=48),
    7: dict(name="1920x1080 YUV420P8 -> 1280x720 Jinc36Resize (2:3 downscale, periodic path)", fmt=ah.YUV420P8, w=1920, h=1080,
            tw=1280, th=720, fn="Jinc36Resize", kw=dict(), tap=3, frames=48),
    9: dict(name="1920x1080 YUV420P8 -> 2560x1440 Jinc36Resize (4:3 upscale, periodic path)", fmt=ah.YUV420P8, w=1920, h=1080,
            tw=2560, th=1440, fn="Jinc36Resize", kw=dict(), tap=3, frames=32),
    8: dict(name="1920x1080 YUV444P16 -> 2500x1400 Jinc64Resize (irregular ratio, many phases)", fmt=ah.YUV444P16, w=1920, h=1080,
            tw=2500, th=1400, fn="Jinc64Resize", kw=dict(), tap=4, frames=12),
}


def synth_frame(fmt: ah.Format, w: int, h: int, seed: int):
    """Seeded uniform noise over the full sample range (float chroma in [-0.5, 0.5))."""
    rng = np.random.default_rng(0x4A494E43 ^ seed)
    planes = []
    for i in range(len(fmt.planes)):
        ph, pw = fmt.plane_shape(i, w, h)
        if fmt.bits == 32:
            a = rng.random((ph, pw), dtype=np.float32)
            if i in (1, 2) and fmt.family not in ("rgbp", "rgbap"):
                a -= 0.5
        else:
            a = rng.integers(0, fmt.peak + 1, (ph, pw), dtype=np.uint16 if fmt.bits > 8 else np.uint8)
        planes.append(np.ascontiguousarray(a.astype(fmt.dtype)))
    return planes


def algorithmic_work(cfg, fs_luma, fs_chroma):
    """FLOP = 2*fs^2*output samples; bytes = (src + dst samples)*sizeof(T) per plane."""
    fmt = cfg["fmt"]
    sb = np.dtype(fmt.dtype).itemsize
    flop = byts = 0.0
    luma = dict(flop=0.0, bytes=0.0)
    for i in range(len(fmt.planes)):
        sh = fmt.plane_shape(i, cfg["w"], cfg["h"])
        dh = fmt.plane_shape(i, cfg["tw"], cfg["th"])
        chroma_tbl = i in (1, 2) and fmt.subsampling != (0, 0)
        fs = fs_chroma if chroma_tbl else fs_luma
        f = 2.0 * fs * fs * dh[0] * dh[1]
        b = (sh[0] * sh[1] + dh[0] * dh[1]) * sb
        flop += f
        byts += b
        if not chroma_tbl:
            luma["flop"] += f
            luma["bytes"] += b
    return flop, byts, luma


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(config_id: int, frames: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel's launch, from the committed `ncu --set full`
    capture of the same bench command (profiles/ncu_traffic.json); None when no capture exists for this workload."""
    try:
        rec = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic.json"))).get(str(config_id))
        if rec and rec.get("frames_per_launch") == frames:
            return float(rec["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def fma_peak_tflops():
    """FP32 FMA-pipe peak measured on this GPU by the repo's own microbenchmark (best FFMA variant)."""
    tool = paths.fma_peak_tool()
    try:
        out = subprocess.run([tool, "8000"], capture_output=True, text=True, timeout=120).stdout
        best = max(json.loads(l)["tflops"] for l in out.splitlines() if l.startswith("{"))
        return best, "measured (fma_peak: FFMA reg-reg, 148 SMs)"
    except Exception:
        return 148 * 128 * 2 * 1.965e9 / 1e12, "nominal (148 SM x 128 lanes x 2 x 1.965 GHz)"


# ------------------------------------------------------------------------------------------ CPU reference

def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(cfg, steps: int, warmup: int, budget_s: float):
    """Times the reference's own CPU implementation through its plugin API under the mini-host: frame-parallel
    get_frame with one host thread per core (AviSynth Prefetch for this MT_MULTI_INSTANCE filter), default opt
    (AVX2 when the CPU has it, src/JincResize.cpp:898).  Falls back to the scalar oracle port only if the
    prebuilt reference library is absent."""
    fmt, w, h, tw, th = cfg["fmt"], cfg["w"], cfg["h"], cfg["tw"], cfg["th"]
    threads = host_threads()
    mpix = tw * th / 1e6
    if os.path.exists(oref.REF_PLUGIN):
        env = ah.Env()
        env.load_plugin(oref.REF_PLUGIN)
        distinct = min(cfg["frames"], 4)
        src = env.source(fmt, w, h, [synth_frame(fmt, w, h, s) for s in range(distinct)], num_frames=1 << 20)
        clip = env.invoke(cfg["fn"], src, tw, th, **cfg["kw"])
        t1 = clip.pull(0, 1, 1)  # single-thread probe: sizes the bounded sample
        per_step = max(threads, 1)
        est = t1 * per_step / max(1, min(threads, per_step))
        n_steps = steps
        while n_steps > 1 and est * (n_steps + warmup) > budget_s:
            n_steps -= 1
        n_warm = warmup if est * (n_steps + warmup) <= budget_s else 0
        f0 = 0
        for _ in range(n_warm):
            clip.pull(f0, per_step, threads)
            f0 += per_step
        t0 = time.perf_counter()
        for _ in range(n_steps):
            clip.pull(f0, per_step, threads)
            f0 += per_step
        dt = time.perf_counter() - t0
        value = n_steps * per_step * mpix / dt
        res = dict(value=value, unit=METRIC, cores=threads, kind="reference",
                   sample=f"{n_steps} steps x {per_step} frames, frame-parallel on {threads} host threads, reference default opt "
                          f"(AVX2 path), 1-thread frame time {t1 * 1e3:.1f} ms",
                   ms_per_step=dt / n_steps * 1e3, steps=n_steps, warmup=n_warm, frames_per_step=per_step,
                   single_thread_value=mpix / t1)
        clip.release()
        src.release()
        return res
    # scalar port (oracle) -- only when oracle/_ref did not travel
    from oracle import cpu as oc

    planes = synth_frame(fmt, w, h, 0)
    sw, sh = fmt.subsampling
    pp = oc.plane_params(w, h, tw, th, tap=cfg["tap"], sub_w=sw, sub_h=sh, **{k: v for k, v in cfg["kw"].items() if k not in ("tap", "blur")})
    lut = oc.make_lut(cfg["tap"], cfg["kw"].get("blur", 0.0))
    t = oc.Table(pp[0], lut)
    rows = max(8, th // 64)
    t0 = time.perf_counter()
    t.resize(planes[0], float(fmt.peak) if fmt.bits < 32 else 0.0, rows=(th // 2, th // 2 + rows))
    dt = time.perf_counter() - t0
    value = (rows * tw / 1e6) / dt
    return dict(value=value, unit=METRIC, cores=1, kind="port", sample=f"{rows} luma rows of one frame, scalar oracle port",
                ms_per_step=dt * 1e3, steps=1, warmup=0, frames_per_step=rows / th)


def run_plugin_e2e(cfg, threads: int, frames: int, steps: int, device: int):
    """The drop-in path itself: the AviSynth+ C plugin (libjincresize_b200.so) under the mini-host, PAGEABLE host frames
    in and out through get_frame, `threads` concurrent get_frame callers (what Prefetch(threads) does).  Includes the
    plugin's staging copies into its pinned slots, H2D, kernels, D2H and the copy into the host's frame."""
    fmt, w, h, tw, th = cfg["fmt"], cfg["w"], cfg["h"], cfg["tw"], cfg["th"]
    os.environ["JINCRESIZE_B200_DEVICES"] = str(device)
    env = ah.Env()
    env.load_plugin(paths.b200_plugin())
    src = env.source(fmt, w, h, [synth_frame(fmt, w, h, s) for s in range(min(frames, 4))], num_frames=1 << 20)
    clip = env.invoke(cfg["fn"], src, tw, th, **cfg["kw"])
    clip.pull(0, frames, threads)  # warm-up: table build, slot allocation, first-touch of the frame pool
    f0 = frames
    t0 = time.perf_counter()
    for _ in range(steps):
        clip.pull(f0, frames, threads)
        f0 += frames
    dt = time.perf_counter() - t0
    clip.release()
    src.release()
    return {"value": steps * frames * (tw * th / 1e6) / dt, "unit": "Mpixel/s", "threads": threads,
            "api": "avisynth_c_plugin_init / get_frame under the mini-host, pageable frames, %d concurrent callers" % threads,
            "ms_per_step": dt / steps * 1e3}


def bind_to_gpu_numa(device_index: int):
    """One process per GPU: pin this process to the CPUs NVML reports as local to its GPU BEFORE any pinned host buffer is
    allocated, so the frames it DMAs live on the GPU's own NUMA node (with 4+ GPUs on a two-socket box, remote buffers
    cut the end-to-end rate).  Returns the number of CPUs bound to, or None when NVML gives no answer."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------ GPU arm

def run_b200(args, cfg):
    import torch
    import torch.distributed as dist

    from jinc_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    fmt, w, h, tw, th = cfg["fmt"], cfg["w"], cfg["h"], cfg["tw"], cfg["th"]
    F = cfg["frames"]
    sb = np.dtype(fmt.dtype).itemsize
    sw, sh = fmt.subsampling
    kw = cfg["kw"]
    flt = capi.Filter(src_w=w, src_h=h, target_w=tw, target_h=th, n_planes=len(fmt.planes), sample_bytes=sb, bits=fmt.bits,
                      sub_w=sw, sub_h=sh, src_left=kw.get("src_left", 0.0), src_top=kw.get("src_top", 0.0),
                      quant_x=kw.get("quant_x", 256), quant_y=kw.get("quant_y", 256), tap=cfg["tap"],
                      blur=kw.get("blur", 0.0), cplace=kw.get("cplace", "mpeg2"), devices=[local_rank], slots_per_device=args.inflight)
    infos = [flt.table(k).info for k in range(flt.num_tables)]
    fs_l = infos[0].filter_size
    fs_c = infos[-1].filter_size
    tdtype = {1: torch.uint8, 2: torch.uint16, 4: torch.float32}[sb]

    def pitched(shape):  # device plane with a 256-byte aligned pitch, like the pipeline's own buffers
        rows, cols = shape
        pitch_elems = ((cols * sb + 255) // 256 * 256) // sb
        return torch.zeros((rows, pitch_elems), dtype=tdtype, device="cuda")

    shapes = flt.plane_shapes()
    host_src, host_dst, dev_frames, keep = [], [], [], []
    for f in range(F):
        planes = synth_frame(fmt, w, h, seed=1000 * rank + f)
        hs = [torch.from_numpy(p).pin_memory() for p in planes]
        hd = [torch.zeros(dshape, dtype=tdtype).pin_memory() for _, dshape in shapes]
        ds = []
        dd = []
        fr = capi.Frame()
        for i, ((sshape, dshape), p) in enumerate(zip(shapes, hs)):
            s = pitched(sshape)
            s[:, : sshape[1]].copy_(p, non_blocking=True)
            d = pitched(dshape)
            ds.append(s)
            dd.append(d)
            fr.src[i], fr.src_pitch[i] = s.data_ptr(), s.stride(0) * sb
            fr.dst[i], fr.dst_pitch[i] = d.data_ptr(), d.stride(0) * sb
        host_src.append(hs)
        host_dst.append(hd)
        dev_frames.append(fr)
        keep.append((ds, dd))
    torch.cuda.synchronize()

    stream = torch.cuda.Stream()
    sh_ = stream.cuda_stream
    n_tables = flt.num_tables

    frames_arr = (capi.Frame * F)(*dev_frames)

    def kernel_step(pairs=None):
        # one launch per coefficient table covers every frame of the batch: interior tiles + border strips together
        if pairs is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            flt.process_device_batch(frames_arr, 0, 1, args.parts, sh_)
            e1.record(stream)
            pairs.append((e0, e1))
        else:
            flt.process_device_batch(frames_arr, 0, 1, args.parts, sh_)
        if n_tables > 1:
            flt.process_device_batch(frames_arr, 0, 2, args.parts, sh_)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel-only
    for _ in range(args.warmup):
        kernel_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = flt.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pairs = []
    ev0.record(stream)
    for _ in range(args.steps):
        kernel_step(pairs)
    ev1.record(stream)
    barrier()
    gpu_launches = flt.kernel_launches - launches0
    ms_total = ev0.elapsed_time(ev1)
    dom_ms = [a.elapsed_time(b) for a, b in pairs]

    # ---------------- end to end through the host-frame C ABI (pinned host buffers)
    src_np = [[t.numpy() for t in hs] for hs in host_src]
    dst_np = [[t.numpy() for t in hd] for hd in host_dst]
    raw = [flt._frame(s, d) for s, d in zip(src_np, dst_np)]

    def e2e_step():
        tickets = []
        for f in range(F):
            if len(tickets) >= args.inflight:
                flt.wait(tickets.pop(0))
            tickets.append(flt.submit_raw(raw[f]))
        for t in tickets:
            flt.wait(t)

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- reduce over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0]), float(t[1])
        g = torch.tensor([gpu_launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(g)
        gpu_launches = int(g[0])

    if rank == 0:
        mpix = tw * th / 1e6
        frames_total = world * args.steps * F
        value = frames_total * mpix / (ms_total * 1e-3)
        e2e_value = frames_total * mpix / e2e_s
        flop, byts, luma = algorithmic_work(cfg, fs_l, fs_c)
        # dominant kernel = the luma-table launch (all luma-table planes of all F frames, interior tiles + border strips)
        i0 = infos[0]
        share = float(F)
        dom_avg_ms = statistics.mean(dom_ms)
        fma_peak, fma_how = fma_peak_tflops()
        hbm_peak, hbm_how = measured_peaks()
        dom_flop = luma["flop"] * share
        dom_bytes = luma["bytes"] * share
        achieved_tf = dom_flop / (dom_avg_ms * 1e-3) / 1e12
        achieved_gbs = dom_bytes / (dom_avg_ms * 1e-3) / 1e9
        h2d = sum(int(np.prod(s)) for s, _ in shapes) * sb * F
        d2h = sum(int(np.prod(d)) for _, d in shapes) * sb * F
        plugin = None
        if world == 1 and args.plugin_threads > 0:
            try:
                plugin = run_plugin_e2e(cfg, args.plugin_threads, F, max(1, min(args.steps, 5)), local_rank)
            except Exception as ex:  # the C-ABI figures above stand on their own
                plugin = {"error": str(ex)[:200]}
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = run_reference(cfg, steps=3, warmup=1, budget_s=25.0)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "config_id": args.config, "frames_per_step": F,
                       "sample_type": f"{fmt.family}{fmt.bits}", "filter_size": fs_l, **({"parts": args.parts, "INVALID": "diagnostic run, part of the frame skipped"} if args.parts != 3 else {}),
                       "l2": "inputs larger than L2: every step walks %d distinct frames (%.0f MB of planes)" % (F, F * byts / 1e6),
                       "partition": "frame-parallel, one process per GPU, no collective",
                       "host_affinity": ("each rank bound to the %d CPUs local to its GPU" % numa_cpus) if numa_cpus else "unbound"},
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "jinc_filter_submit/jinc_filter_wait (C ABI host-frame call, %d frames in flight per GPU), pinned host planes" % args.inflight,
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "pcie_gbs": (h2d + d2h) * args.steps / e2e_s / 1e9},
            "e2e_plugin": plugin,
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "roofline": {"bound": "fp32_fma", "kernel": {1: "resample_up2x", 2: "resample_down", 3: "resample_down (periodic passes in one launch)"}.get(i0.fast_path, "resample_strips") + " (luma-table planes of all %d frames: interior tiles + border strips)" % F,
                         "achieved": achieved_tf, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved_tf / fma_peak,
                         "peak_source": fma_how, "traffic": ncu_traffic(args.config, F), "launch_ms": dom_avg_ms,
                         "algorithmic_flop_per_launch": dom_flop, "share_of_step": sum(dom_ms) / ms_total},
            "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": achieved_gbs / hbm_peak, "peak_source": hbm_how,
                             "algorithmic_bytes_per_launch": dom_bytes},
            "whole_frame": {"gflop_per_frame": flop / 1e9, "mbytes_per_frame": byts / 1e6,
                            "tflops": flop * frames_total / world / (ms_total * 1e-3) / 1e12},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    flt.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--plugin-threads", type=int, default=8,
                    help="concurrent get_frame callers of the plugin end-to-end leg (N=1 only; 0 skips it)")
    ap.add_argument("--inflight", type=int, default=3, help="frames in flight per GPU in the end-to-end leg (= pipeline slots)")
    ap.add_argument("--parts", type=int, default=3, choices=[1, 2, 3],
                    help="diagnosis only: 1 = interior tiles, 2 = border strips, 3 = both (the only valid bench setting)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_reference(cfg, steps=args.steps, warmup=args.warmup, budget_s=150.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Mpixel/s", "n_gpus": args.gpus,
                "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["name"], "config_id": args.config, "frames_per_step": r["frames_per_step"]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return
    run_b200(args, cfg)


if __name__ == "__main__":
    main()
