#!/usr/bin/env python
"""bench.py -- NV12->I420 frames/s & HBM GB/s on 1..8 B200, next to the reference's CPU loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic decoded surfaces = ONE kernel
launch (default workload: BASELINE.json configs[1], 300 pitched 1920x1080 NV12 surfaces, pitch
2048 -> tight I420).  Per rank the batch is fixed (weak scaling: frames/streams are independent,
ranks never exchange data; torch.distributed only carries the barrier and the max-over-ranks).

  value     whole-job frames/s with the surfaces already resident in HBM (CUDA events on the
            launching stream, max over ranks)
  roofline  algorithmic bytes per launch / launch duration vs the measured HBM copy peak
  e2e       same metric through the C-ABI pipeline with pinned HOST buffers: H2D of every surface,
            convert, D2H of every tight frame inside the timed region
  cpu_baseline / --impl reference
            the UNMODIFIED reference function jm_nvdec_output_frame (nv_dec/nv_dec.cpp:750-828)
            compiled into oracle/_ref, timed on this box's host cores on the same surfaces
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (op, w, h, pitch, frames per launch)
    "nv12_to_i420_1080p_x300_pitch2048": ("i420", 1920, 1080, 2048, 300),
    "nv12_to_i420_1080p_x300_pitch1920": ("i420", 1920, 1080, 1920, 300),      # no padding: shows what the pitch costs
    "nv12_to_i420_4k_x64_pitch4096": ("i420", 3840, 2160, 4096, 64),
    "nv12_to_i420_4k_x64_pitch3840": ("i420", 3840, 2160, 3840, 64),
    "nv12_to_nv12_1080p_x300_pitch2048": ("nv12", 1920, 1080, 2048, 300),
    "i420_to_nv12_1080p_x300_pitch2048": ("pack", 1920, 1080, 2048, 300),
    "i420_to_nv12_4k_x64_pitch4096": ("pack", 3840, 2160, 4096, 64),
    "nv12_to_rgb24_4k_x64_pitch4096": ("rgb", 3840, 2160, 4096, 64),
    "nv12_to_i420_rgb24_4k_x64_pitch4096": ("fused", 3840, 2160, 4096, 64),
    "nv12_to_argb32_4k_x64_pitch4096": ("argb", 3840, 2160, 4096, 64),
    "rgb24_to_nv12_4k_x64_pitch4096": ("rgb2nv12", 3840, 2160, 4096, 64),      # kernel table only
}
DEFAULT_WORKLOAD = "nv12_to_i420_1080p_x300_pitch2048"
N_DISTINCT = 32          # distinct synthetic surfaces, tiled over the batch
FALLBACK_HBM_GBS = 6650.0


# ------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


KERNEL_SOURCES = ("jmc_k_common.cuh", "jmc_k_planes.cuh", "jmc_k_rows.cuh", "jmc_k_rgb.cuh", "jmc_kernels.cu")


def kernel_sources_sha256():
    """Hash of the kernel sources + launch logic: ties a recorded ncu DRAM-traffic figure to the code it was taken from."""
    import hashlib
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "jmcodec_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def recorded_traffic(workload):
    """(dram bytes per launch, provenance dict) from the committed ncu --set full capture (profiles/traffic.json, written by
    tools/record_traffic.py).  A figure captured from other kernel sources than the ones in this tree is STALE: it is
    reported as null, with the reason."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        e = json.load(open(p)).get(workload)
    except Exception:
        e = None
    if not isinstance(e, dict):
        return None, {"status": "no ncu capture recorded for this workload"}
    prov = {k: e.get(k) for k in ("kernel", "capture", "git", "when", "sources_sha256")}
    if e.get("sources_sha256") != kernel_sources_sha256():
        prov["status"] = "stale: kernel sources changed since the capture"
        return None, prov
    prov["status"] = "current"
    return e.get("traffic"), prov


def run_config(name, n_gpus):
    """The `config` object, identical in both arms (b200 / reference) for the same command line."""
    op, w, h, pitch, n = WORKLOADS[name]
    in_b, out_b, out2_b = io_bytes(op, w, h, pitch)
    return {"workload": name, "width": w, "height": h, "pitch": pitch, "frames_per_step_per_gpu": n,
            "distinct_surfaces": N_DISTINCT,
            "l2": f"batch {in_b * n / 1e6:.0f} MB in + {(out_b + out2_b) * n / 1e6:.0f} MB out per step > 126 MB L2 / host LLC, no flush needed",
            "parallelism": f"frames sharded over {n_gpus} GPU(s), no collective"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def aggregate(units: int, ms: float, device):
    """Whole-job units (sum over ranks) and the step time (MAX over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return units, ms
    t = torch.tensor([float(units)], dtype=torch.float64, device=device)
    m = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return int(round(t.item())), m.item()


class ClockSampler:
    """SM clock / throttle reasons sampled while the timed regions run: NVML in a thread (2 ms period),
    or, if NVML is unusable, the nvidia-smi loop of B200_PROFILING.md (200 ms period, whole run)."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._run, self._on = [], set(), None, False, False
        self.index, self.proc, self.source = index, None, "nvml"
        self.period = float(os.environ.get("JMC_BENCH_CLOCK_PERIOD_MS", "2")) * 1e-3
        try:
            if os.environ.get("JMC_BENCH_NO_NVML"):
                raise RuntimeError("NVML disabled by JMC_BENCH_NO_NVML")
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            self.nv, self.source = None, "nvidia-smi"

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                 "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80}
        while self._run:
            if self._on:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(self.period)

    def start(self):
        if self.nv:
            self._run = True
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        else:
            import subprocess
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                              "--format=csv,noheader,nounits", "-lms", "200"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception:
                self.proc = None

    def region(self, on):
        self._on = on

    def stop(self):
        if self.nv and self._run:
            self._run = False
            self.t.join()
        elif self.proc:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                out = ""
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for ln in out.splitlines():
                f = [x.strip() for x in ln.split(",")]
                try:
                    self.samples.append(int(f[0]))
                    self.max_mhz = int(f[1])
                    for nme, v in zip(names, f[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(nme)
                except (ValueError, IndexError):
                    continue
            busy = [x for x in self.samples if self.max_mhz and x > 0.5 * self.max_mhz]
            self.samples = busy or self.samples

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0, "source": self.source}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}


# ------------------------------------------------------------------------------------------------
def make_inputs(op, w, h, pitch, rank):
    from jmcodec_b200 import synth
    if op == "pack":
        return [synth.i420_frame(w, h, rank, f) for f in range(N_DISTINCT)]
    if op == "rgb2nv12":
        return [synth.random_bytes(3 * w * h, synth.frame_key(rank, f) ^ 0x33CC33CC) for f in range(N_DISTINCT)]
    return [synth.nv12_surface(w, h, pitch, rank, f) for f in range(N_DISTINCT)]


def cpu_description():
    """What SURVEY.md 8d wants next to every CPU figure: the CPU model and how the reference code was compiled."""
    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.lower().startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"cpu_model": model,
            "compiler_flags": "g++ -std=gnu++11 -fpermissive -O2 (the reference's MSVC Release /O2 'MaxSpeed'); oracle/Makefile"}


def cpu_reference_fps(w, h, pitch, frames, threads, surfs=None):
    """Time the reference's own CPU function (oracle/_ref, else the C port) on `frames` calls."""
    import oracle
    chk = oracle.best()
    if surfs is None:
        surfs = make_inputs("i420", w, h, pitch, 0)
    S = np.stack(surfs)
    out = np.zeros((max(N_DISTINCT, threads), w * h * 3 // 2), np.uint8)
    chk.nvdec_run(S, out, pitch, w, h, 1, min(frames, 64), threads)          # warm-up: page in, spin up threads
    t = chk.nvdec_run(S, out, pitch, w, h, 1, frames, threads)
    return frames / t, chk.kind


def run_reference_arm(args, name):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    op, w, h, pitch, n = WORKLOADS[name]
    if op != "i420":
        print(json.dumps({"impl": "reference", "unavailable": "the reference has CPU code only for NV12->I420/NV12 on this path"}))
        return 0
    threads = len(os.sched_getaffinity(0))
    surfs = make_inputs("i420", w, h, pitch, 0)
    # one step = one batch of n frames over all host threads (one reference handle per thread).  The K steps run inside
    # ONE call of the driver loop, so the worker threads persist across steps: spawning them per step would cost the CPU
    # arm ~20 % at this batch size, and that is this harness's overhead, not the reference's.
    import oracle
    chk = oracle.best()
    S = np.stack(surfs)
    out = np.zeros((max(N_DISTINCT, threads), w * h * 3 // 2), np.uint8)
    chk.nvdec_run(S, out, pitch, w, h, 1, n * max(args.warmup, 1), threads)
    dt = chk.nvdec_run(S, out, pitch, w, h, 1, n * args.steps, threads)
    fps = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "NV12->I420 frames/s", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": run_config(name, args.gpus),
        "cpu_baseline": dict({"value": fps, "unit": "frames/s", "cores": threads, "kind": chk.kind,
                              "sample": f"{args.steps} steps x {n} frames, jm_nvdec_output_frame out_fmt=1, one handle per thread, "
                                        f"{N_DISTINCT} distinct surfaces"}, **cpu_description()),
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def build_job(ctx, op, w, h, pitch, n, d_in, d_out, d_out2):
    surf_bytes = pitch * h * 3 // 2
    tight_bytes = w * h * 3 // 2
    if op in ("i420", "nv12"):
        j = ctx.job_nvdec(w, h, pitch, 1 if op == "i420" else 0)
        j.surf.base, j.surf.stride, j.tight.base, j.tight.stride = d_in, surf_bytes, d_out, tight_bytes
    elif op == "pack":
        j = ctx.job_nvenc(w, h, pitch, 0x10)
        j.tight.base, j.tight.stride, j.surf.base, j.surf.stride = d_in, tight_bytes, d_out, surf_bytes
    elif op == "argb":
        j = ctx.job_argb(w, h, pitch, 4 * w)
        j.surf.base, j.surf.stride, j.rgb.base, j.rgb.stride = d_in, surf_bytes, d_out, 4 * w * h
    elif op == "rgb2nv12":
        j = ctx.job_rgb_to_nv12(w, h, 3 * w, pitch)
        j.rgb.base, j.rgb.stride, j.surf.base, j.surf.stride = d_in, 3 * w * h, d_out, surf_bytes
    else:
        j = ctx.job_rgb(w, h, pitch, 3 * w, op == "fused")
        j.surf.base, j.surf.stride = d_in, surf_bytes
        if op == "fused":
            j.tight.base, j.tight.stride, j.rgb.base, j.rgb.stride = d_out, tight_bytes, d_out2, 3 * w * h
        else:
            j.rgb.base, j.rgb.stride = d_out, 3 * w * h
    j.n_frames = n
    return j


def io_bytes(op, w, h, pitch):
    surf_bytes, tight_bytes, rgb = pitch * h * 3 // 2, w * h * 3 // 2, 3 * w * h
    if op == "pack":
        return tight_bytes, surf_bytes, 0
    if op == "rgb":
        return surf_bytes, rgb, 0
    if op == "argb":
        return surf_bytes, 4 * w * h, 0
    if op == "rgb2nv12":
        return rgb, surf_bytes, 0
    if op == "fused":
        return surf_bytes, tight_bytes, rgb
    return surf_bytes, tight_bytes, 0


def device_only(ctx, name, rank, iters, warmup):
    """Device-timed launches of one workload with inputs resident in HBM.  Returns a dict."""
    op, w, h, pitch, n = WORKLOADS[name]
    in_b, out_b, out2_b = io_bytes(op, w, h, pitch)
    host = np.concatenate([make_inputs(op, w, h, pitch, rank)[f % N_DISTINCT] for f in range(min(n, N_DISTINCT))])
    d_in = ctx.alloc(in_b * n)
    reps = -(-n // N_DISTINCT)
    for r in range(reps):                        # tile the distinct surfaces over the batch
        cnt = min(N_DISTINCT, n - r * N_DISTINCT)
        ctx.h2d(d_in + r * N_DISTINCT * in_b, host, cnt * in_b)
    d_out = ctx.alloc(out_b * n)
    d_out2 = ctx.alloc(out2_b * n) if out2_b else None
    if op in ("pack", "rgb2nv12"):
        ctx.memset(d_out, 0, out_b * n)
    j = build_job(ctx, op, w, h, pitch, n, d_in, d_out, d_out2)
    for _ in range(warmup):
        ctx.convert(j)
    ctx.sync()
    ms = ctx.convert_timed(j, iters)
    alg = ctx.algorithmic_bytes(j) * n
    res = {"ms_per_launch": ms, "frames_per_s": n / (ms * 1e-3), "algorithmic_bytes_per_launch": alg,
           "gbs": alg / (ms * 1e-3) / 1e9}
    for d in (d_in, d_out, d_out2):
        if d:
            ctx.free(d)
    return res


def ratios_vs_cpu(cpu, value, e2e_value, decode_path):
    """This run's GPU figures over this run's CPU baseline (the reference's loop on all host cores of the same box).
    `e2e` is the honest headline: host buffers both ways on both sides.  The driver's own e2e ratio uses the separate
    `--impl reference` run; this is the same comparison inside one process.  Never raises (it decorates the line)."""
    try:
        base = float(cpu["value"])
        if not base > 0:
            return None
        return {"e2e": e2e_value / base,
                "e2e_device_resident_input": decode_path["value"] / base if decode_path else None,
                "device_only": value / base,
                "basis": "cpu_baseline.value of this run: %s threads, kind %s" % (cpu.get("cores"), cpu.get("kind"))}
    except Exception:                   # noqa: BLE001
        return None


def dropin_api_fps(device, w=1920, h=1080, pitch=2048, frames=400):
    """Frames/s of the per-frame drop-in API, called as test_nv_dec.cpp:215-218 calls it (jm_nvdec_decode_frame
    then jm_nvdec_output_frame, one frame at a time, one handle per thread), measured by tools/jm_dropin -- plain
    C++ on the C-ABI, so that no Python sits in the timed loop.  Keys: see the VARIANTS table of tools/jm_dropin.cpp."""
    import subprocess
    exe = os.path.join(ROOT, "tools", "jm_dropin")
    if not os.path.exists(exe):
        raise RuntimeError("tools/jm_dropin is not built (python -c 'import __graft_entry__ as g; g.build()')")
    p = subprocess.run([exe, "--device", str(device), "--frames", str(frames), "--width", str(w), "--height", str(h),
                        "--pitch", str(pitch)], capture_output=True, text=True, timeout=600)
    if p.returncode != 0:
        raise RuntimeError("tools/jm_dropin failed: " + p.stderr[-300:])
    return json.loads(p.stdout)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(k for k, v in WORKLOADS.items() if v[0] != "rgb2nv12"))
    ap.add_argument("--no-extras", action="store_true", help="skip the per-kernel table and the CPU baseline")
    ap.add_argument("--e2e-sub", type=int, default=0, help="frames per pipeline batch in the e2e leg (default: auto)")
    ap.add_argument("--e2e-depth", type=int, default=3, help="pipeline slots in the e2e leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    name = args.workload
    if args.impl == "reference":
        return run_reference_arm(args, name)

    rank, world, local = dist_env()
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- jmcodec_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import jmcodec_b200 as J

    op, w, h, pitch, n = WORKLOADS[name]
    in_b, out_b, out2_b = io_bytes(op, w, h, pitch)
    # weak scaling: the job is n frames per GPU, and a rank owns one contiguous slice of the job's frames -- no rank ever
    # needs another rank's bytes (the same split the world-2 gloo test checks)
    from jmcodec_b200.shard import frames_for_rank
    my_frames = frames_for_rank(n * world, rank, world)
    if len(my_frames) != n:
        raise SystemExit("bench.py: frame split is not even")
    ctx = J.Ctx(local)
    sampler = ClockSampler(local)
    sampler.start()

    # ---- inputs: pinned host copy of the whole batch (the e2e leg streams from here) ---------------
    distinct = make_inputs(op, w, h, pitch, rank)
    hin = ctx.alloc_host(in_b * n)
    for f in range(n):
        hin.array[f * in_b:(f + 1) * in_b] = distinct[f % N_DISTINCT]
    hout = ctx.alloc_host(out_b * n)
    hout2 = ctx.alloc_host(out2_b * n) if out2_b else None
    d_in = ctx.alloc(in_b * n)
    d_out = ctx.alloc(out_b * n)
    d_out2 = ctx.alloc(out2_b * n) if out2_b else None
    ctx.h2d(d_in, hin.array)
    ctx.memset(d_out, 0, out_b * n)
    job = build_job(ctx, op, w, h, pitch, n, d_in, d_out, d_out2)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident: W warm-up steps, then exactly K timed steps ------------------------------
    for _ in range(args.warmup):
        ctx.convert(job)
    barrier()
    l0 = ctx.launches
    sampler.region(True)
    ms_total = ctx.convert_timed(job, args.steps) * args.steps
    sampler.region(False)
    launches = ctx.launches - l0
    barrier()
    units, ms_max = aggregate(n * args.steps, ms_total, dev)
    value = units / (ms_max * 1e-3)
    ms_per_step = ms_max / args.steps
    alg_launch = ctx.algorithmic_bytes(job) * n
    achieved = alg_launch / ((ms_total / args.steps) * 1e-3) / 1e9          # this rank's kernel, GB/s
    peak, peak_src = peaks()

    # ---- correctness spot-check on rank 0 (bit-exact vs the reference function) ---------------------
    verified = None
    if rank == 0 and op in ("i420", "nv12"):
        import oracle
        chk = oracle.best()
        got = np.empty(out_b, np.uint8)
        want = np.empty(out_b, np.uint8)
        verified = True
        for f in (0, n // 2, n - 1):
            ctx.d2h(got, d_out + f * out_b)
            chk.nvdec_output_frame(distinct[f % N_DISTINCT], pitch, w, h, 1 if op == "i420" else 0, want, out_b)
            verified = verified and bool(np.array_equal(got, want))

    # ---- e2e: pinned host surfaces -> H2D -> convert -> D2H tight frames, overlapped ----------------
    sub = 30 if n % 30 == 0 else (16 if n % 16 == 0 else n)       # frames per pipeline batch
    if args.e2e_sub and n % args.e2e_sub == 0:
        sub = args.e2e_sub
    shape = build_job(ctx, op, w, h, pitch, sub, None, None, None)
    pipe = J.Pipeline(ctx, shape, pitch * h * 3 // 2, depth=args.e2e_depth)
    ev0, ev1 = J.Event(ctx), J.Event(ctx)

    def e2e_step():
        for b in range(n // sub):
            pipe.submit(hin.array[b * sub * in_b:], hout.array[b * sub * out_b:], sub,
                        host_out2=hout2.array[b * sub * out2_b:] if hout2 else None)

    for _ in range(args.warmup):
        e2e_step()
    pipe.drain()
    barrier()
    h0, d0 = pipe.h2d_bytes, pipe.d2h_bytes
    sampler.region(True)
    ev0.record(1)                                  # upload stream: first H2D of the timed region
    for _ in range(args.steps):
        e2e_step()
    ev1.record(2)                                  # delivery stream: after the last D2H
    e2e_ms = ev0.elapsed_ms(ev1)
    pipe.drain()
    sampler.region(False)
    barrier()
    e2e_units, e2e_ms_max = aggregate(n * args.steps, e2e_ms, dev)
    e2e_value = e2e_units / (e2e_ms_max * 1e-3)
    h2d_step = (pipe.h2d_bytes - h0) // args.steps
    d2h_step = (pipe.d2h_bytes - d0) // args.steps
    if rank == 0 and verified is not None:
        import oracle
        want = np.empty(out_b, np.uint8)
        oracle.best().nvdec_output_frame(distinct[(n - 1) % N_DISTINCT], pitch, w, h, 1 if op == "i420" else 0, want, out_b)
        verified = verified and bool(np.array_equal(hout.array[(n - 1) * out_b:n * out_b], want))
    # ---- the reference's real data flow: surfaces already in HBM (NVDEC wrote them), only the tight
    #      frames travel: convert + D2H, no H2D (reported beside e2e, never instead of it) ----------------
    decode_path = None
    if op in ("i420", "nv12", "rgb", "fused", "argb"):
        def resident_step():
            for b in range(n // sub):
                pipe.submit(None, hout.array[b * sub * out_b:], sub, dev_in=d_in + b * sub * in_b,
                            host_out2=hout2.array[b * sub * out2_b:] if hout2 else None)
        for _ in range(args.warmup):
            resident_step()
        pipe.drain()
        barrier()
        d1 = pipe.d2h_bytes
        sampler.region(True)
        ev0.record(0)
        for _ in range(args.steps):
            resident_step()
        ev1.record(2)
        dp_ms = ev0.elapsed_ms(ev1)
        pipe.drain()
        sampler.region(False)
        barrier()
        dp_units, dp_ms_max = aggregate(n * args.steps, dp_ms, dev)
        decode_path = {"value": dp_units / (dp_ms_max * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 0,
                       "d2h_bytes_per_step": int((pipe.d2h_bytes - d1) // args.steps), "ms_per_step": dp_ms_max / args.steps,
                       "note": "surfaces resident in HBM as after NVDEC; kernel + pinned D2H of the tight frames only"}
    pipe.close()
    # ---- the host link's ceiling, measured NOW on every rank at once (barrier-synchronised): both directions
    #      concurrently (what e2e does) and D2H alone (what the device-resident flow does); whole-job = sum over ranks
    ceiling = None
    try:
        barrier()
        up, down = ctx.link_probe(64 << 20, -300, 3)         # every rank copies for the same 300 ms window
        barrier()
        _, down_only = ctx.link_probe(64 << 20, -300, 2)
        barrier()
        sums = []
        for v in (up, down, down_only):
            tot, _ = aggregate(int(v * 1e6), 1.0, dev)
            sums.append(tot / 1e6)
        ceiling = {"bidirectional_h2d_gbs": sums[0], "bidirectional_d2h_gbs": sums[1], "d2h_only_gbs": sums[2],
                   "how": "jmc_link_probe: 64 MiB pinned copies for a 300 ms window on every rank at once, between barriers, device-timed; whole-job = sum over ranks"}
    except Exception as e:      # noqa: BLE001
        ceiling = {"error": repr(e)}
    sampler.stop()

    # ---- extras on rank 0, N=1: per-kernel device table + CPU baseline ------------------------------
    kernels, cpu, batch_sweep, dropin = None, None, None, None
    extras_errors = []
    if rank == 0 and world == 1 and not args.no_extras:
        for d in (d_in, d_out, d_out2):
            if d:
                ctx.free(d)
        d_in = d_out = d_out2 = None
        threads = len(os.sched_getaffinity(0))
        # (1) the reported CPU baseline: the reference function on this box's host cores
        if op == "i420":
            f1, kind = cpu_reference_fps(w, h, pitch, 1500, 1, distinct)
            nfr = 300 * max(1, min(threads, 64))
            fN, _ = cpu_reference_fps(w, h, pitch, nfr, threads, distinct)
            cpu = {"value": fN, "unit": "frames/s", "cores": threads, "kind": kind, "value_1thread": f1,
                   "sample": f"jm_nvdec_output_frame out_fmt=1 on {w}x{h} pitch {pitch}: 1500 frames on 1 thread, "
                             f"{nfr} frames on {threads} threads (one handle per thread), {N_DISTINCT} distinct surfaces"}
            cpu.update(cpu_description())
            try:
                import oracle
                S = np.stack(distinct)
                o3out = np.zeros((max(N_DISTINCT, threads), w * h * 3 // 2), np.uint8)
                oracle.ref_best_effort_run(S, o3out, pitch, w, h, 1, 64, threads)
                t3 = oracle.ref_best_effort_run(S, o3out, pitch, w, h, 1, nfr, threads)
                t31 = oracle.ref_best_effort_run(S, o3out, pitch, w, h, 1, 1500, 1)
                cpu["best_effort_o3_avx2"] = None if not t3 else {
                    "value": nfr / t3, "value_1thread": 1500 / t31 if t31 else None,
                    "note": "same unmodified nv_dec.cpp, gcc -O3 -mavx2 instead of the reference's -O2/MaxSpeed"}
            except Exception as e:      # noqa: BLE001
                extras_errors.append("best_effort_cpu: " + repr(e))
        # (2) optional tables: every kernel device-resident, frames-per-launch sweep, the per-frame drop-in API
        try:
            kernels = {}
            for k in list(WORKLOADS):
                r = device_only(ctx, k, rank, 10, 3)
                kernels[k] = {"frames_per_s": round(r["frames_per_s"], 1), "gbs": round(r["gbs"], 1),
                              "frac_of_peak": round(r["gbs"] / peak, 4), "ms_per_launch": round(r["ms_per_launch"], 4),
                              # the reference has no RGB code: those oracles are builder-defined
                              "parity": "unpinned" if WORKLOADS[k][0] in ("rgb", "fused", "argb", "rgb2nv12") else "pinned"}
        except Exception as e:          # noqa: BLE001
            extras_errors.append("kernels: " + repr(e))
        try:
            batch_sweep = {}            # launch-bound -> bandwidth-bound
            for nb in (1, 8, 64, 300):
                WORKLOADS["_sweep"] = ("i420", 1920, 1080, 2048, nb)
                r = device_only(ctx, "_sweep", rank, 50 if nb < 64 else 10, 5)
                batch_sweep[str(nb)] = {"frames_per_s": round(r["frames_per_s"], 1), "gbs": round(r["gbs"], 1),
                                        "us_per_launch": round(r["ms_per_launch"] * 1e3, 2)}
        except Exception as e:          # noqa: BLE001
            extras_errors.append("frames_per_launch_sweep: " + repr(e))
        WORKLOADS.pop("_sweep", None)
        try:
            dropin = dropin_api_fps(local)
        except Exception as e:          # noqa: BLE001
            extras_errors.append("dropin_api: " + repr(e))

    if rank == 0:
        traffic, traffic_prov = recorded_traffic(name)
        step_s = e2e_ms_max / args.steps * 1e-3
        # whole-job bytes per second each way (h2d_step / d2h_step are this rank's bytes per step; ranks are identical)
        gbs_up, gbs_down = h2d_step * world / step_s / 1e9, d2h_step * world / step_s / 1e9
        e2e_obj = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                   "ms_per_step": e2e_ms_max / args.steps, "frames_per_pipeline_batch": sub, "pipeline_depth": args.e2e_depth,
                   "pcie_gbs_each_way_whole_job": [gbs_up, gbs_down]}
        if ceiling and "error" not in ceiling:
            lim = min(ceiling["bidirectional_h2d_gbs"], ceiling["bidirectional_d2h_gbs"])
            e2e_obj["host_ceiling_gbs"] = lim
            e2e_obj["frac_of_host_ceiling"] = max(gbs_up, gbs_down) / lim if lim > 0 else None
            e2e_obj["host_ceiling"] = ceiling
            if decode_path:
                decode_path["host_ceiling_gbs"] = ceiling["d2h_only_gbs"]
                dp_gbs = decode_path["d2h_bytes_per_step"] * world / (decode_path["ms_per_step"] * 1e-3) / 1e9
                decode_path["frac_of_host_ceiling"] = dp_gbs / ceiling["d2h_only_gbs"] if ceiling["d2h_only_gbs"] > 0 else None
        elif ceiling:
            e2e_obj["host_ceiling"] = ceiling
        line = {
            "metric": "NV12->I420 frames/s" if op == "i420" else f"{name} frames/s",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": run_config(name, world),
            "launches_per_step": 1,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_provenance": traffic_prov, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_launch},
            "e2e": e2e_obj,
            "e2e_device_resident_input": decode_path,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "verified_bit_exact": verified,
            "parity": "unpinned" if op in ("rgb", "fused", "argb", "rgb2nv12") else "pinned",
        }
        if cpu:
            line["cpu_baseline"] = cpu
            line["vs_cpu_baseline_all_cores"] = ratios_vs_cpu(cpu, value, e2e_value, decode_path)
        if kernels:
            line["kernels"] = kernels
        if batch_sweep:
            line["frames_per_launch_sweep_1080p"] = batch_sweep
        if dropin:
            line["dropin_api_1080p"] = dropin
        if extras_errors:
            line["extras_errors"] = extras_errors
        print(json.dumps(line))
    for b in (hin, hout, hout2):
        if b:
            b.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
