#!/usr/bin/env python
"""bench.py - Mode S demodulation throughput on B200 (BASELINE.json metric: Msamples/s).

A "step" is one pass of the hot path over one synthetic capture: configs[1] of BASELINE.json, a
60 s dense-traffic airspy-R2 capture (1.2 G real 16-bit samples = 2.4 GB, ~2000 aircraft with
overlapping frames), decoded to the final frame list (DF18 on, Try/Ok counters on).

  value      whole-job Msamples/s with the capture already resident in HBM (push_device+finish
             through the C ABI: kernels, candidate D2H, host resolve), CUDA-event timed
  e2e        same metric from HOST buffers through adsbdec_b200_push(): pinned host memory ->
             H2D -> kernels -> frames on the host, copies inside the timed region
  roofline   the fused kernel alone: 2 B/sample x samples / its CUDA-event duration vs measured HBM
  cpu_baseline  the reference's own CPU path (oracle/_ref binary built from the unmodified
             sources, or the oracle port) on a bounded sample of the same workload

`--impl reference` times the reference CPU implementation instead (one process per core on
pre-split chunk files) and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import shutil
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FS = 20_000_000
WORKLOAD = "60s dense capture, ~2000 aircraft, DF11/17/18/20/21 mix with overlaps, 20 MS/s real u16 (configs[1])"
SEG_SECONDS = 2.0          # host-rendered dense segment, tiled on the device with fresh noise
AIRCRAFT = 2000
FRAMES_PER_AIRCRAFT_S = 6.0
SIGMA = 25.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=60.0, help="capture length per GPU (default: configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--kernel", type=int, default=0)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# workload

def dense_segment(seed: int, seconds: float = SEG_SECONDS):
    """Noise-free dense traffic (float32, centred on 0) for `seconds`, rendered on the host."""
    from adsbdec_b200.modulator import amplitude_for_snr, envelope, make_frame
    rng = np.random.default_rng(seed)
    n = int(seconds * FS)
    n -= n % 8
    sig = np.zeros(n, dtype=np.float32)
    n_frames = int(AIRCRAFT * FRAMES_PER_AIRCRAFT_S * seconds)
    dfs = np.array([17, 11, 18, 20, 21])
    w = np.array([0.55, 0.2, 0.05, 0.1, 0.1])
    starts = np.sort(rng.integers(0, n - 2600, n_frames))
    snr = rng.uniform(8.0, 30.0, n_frames)
    for s, db in zip(starts, snr):
        df = int(rng.choice(dfs, p=w))
        ov = int(rng.integers(1, 1 << 24)) if df in (20, 21) else 0
        env = envelope(make_frame(df, rng, overlay=ov))
        t = np.arange(s, s + env.size, dtype=np.float64)
        car = np.cos(0.5 * np.pi * t + rng.uniform(0, 2 * np.pi)).astype(np.float32)
        sig[s:s + env.size] += np.float32(amplitude_for_snr(db, SIGMA)) * env * car
    return sig, n_frames


def quantise_host(sig: np.ndarray, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = sig + rng.standard_normal(sig.size, dtype=np.float32) * np.float32(SIGMA)
    return np.clip(np.rint(x + 2048.0), 0, 4095).astype("<u2")


def build_device_capture(torch, dev, seconds: float, seed: int):
    """The capture in HBM: the dense segment tiled, every tile with its own AWGN (torch is only
    used to synthesise the input)."""
    sig, n_frames = dense_segment(seed)
    seg = torch.from_numpy(sig).to(dev)
    tiles = max(1, int(round(seconds / SEG_SECONDS)))
    out = torch.empty(tiles * seg.numel(), dtype=torch.int16, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    for i in range(tiles):
        x = seg + torch.randn(seg.numel(), device=dev, generator=g) * SIGMA
        out[i * seg.numel():(i + 1) * seg.numel()] = torch.clamp(torch.round(x + 2048.0), 0, 4095).to(torch.int16)
    return out, n_frames * tiles, sig


# ---------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    """Samples SM clock and throttle reasons (NVML) while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.004)

    def __enter__(self):
        if self.nv:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._th:
            self._th.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# reference CPU path

def host_cores() -> int:
    return len(os.sched_getaffinity(0))


def ref_binary():
    from oracle import pyoracle
    return pyoracle.ref_binary()


def run_reference_cpu(sample: np.ndarray, n_procs: int, flags=("-a", "-m")):
    """Times `adsbdec -f` (unmodified reference, oracle/_ref) on `sample`: one process over the
    whole sample, then n_procs processes at once, one per core, each on its own pre-split chunk
    file.  Falls back to the oracle port (single thread) if the binary cannot run here.
    Returns dict(single_msps, multi_msps, kind, procs, sample_samples)."""
    binary = ref_binary()
    tmpdir = Path(tempfile.mkdtemp(prefix="adsb_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    try:
        if binary is None:
            from oracle import pyoracle
            t0 = time.perf_counter()
            pyoracle.decode(sample, accept_df18=True)
            dt = time.perf_counter() - t0
            v = sample.size / dt / 1e6
            return {"single_msps": v, "multi_msps": v, "kind": "port", "procs": 1, "sample_samples": int(sample.size)}
        whole = tmpdir / "whole.bin"
        sample.tofile(whole)
        with open(os.devnull, "wb") as null:
            t0 = time.perf_counter()
            subprocess.run([str(binary), *flags, "-f", str(whole)], stdout=null, stderr=null, check=True)
            single = sample.size / (time.perf_counter() - t0) / 1e6
            # pre-split chunks, one per process
            n_files = min(n_procs, 16)
            per = sample.size // n_files
            per -= per % 8
            files = []
            for i in range(n_files):
                f = tmpdir / f"chunk{i}.bin"
                sample[i * per:(i + 1) * per].tofile(f)
                files.append(f)
            t0 = time.perf_counter()
            procs = [subprocess.Popen([str(binary), *flags, "-f", str(files[i % n_files])], stdout=null, stderr=null)
                     for i in range(n_procs)]
            for p in procs:
                p.wait()
            multi = per * n_procs / (time.perf_counter() - t0) / 1e6
        return {"single_msps": single, "multi_msps": multi, "kind": "reference", "procs": n_procs,
                "sample_samples": int(sample.size), "chunk_samples": int(per)}
    finally:
        for f in tmpdir.glob("*"):
            f.unlink()
        tmpdir.rmdir()


def cpu_sample(seed: int, seconds: float) -> np.ndarray:
    sig, _ = dense_segment(seed)
    tiles = max(1, int(round(seconds / SEG_SECONDS)))
    return np.concatenate([quantise_host(sig, seed + 1000 + i) for i in range(tiles)])


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    sample = cpu_sample(1234, 10.0)          # 10 s of the workload = 200 M samples
    times = []
    res = None
    for i in range(args.warmup + args.steps):
        res = run_reference_cpu(sample, cores)
        if i >= args.warmup:
            times.append(res)
        if i == 0 and args.warmup:
            pass
    multi = float(np.mean([r["multi_msps"] for r in times]))
    single = float(np.mean([r["single_msps"] for r in times]))
    samples_per_step = res["chunk_samples"] * res["procs"] if "chunk_samples" in res else res["sample_samples"]
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": multi, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": samples_per_step / multi / 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "10 s of the workload per step, pre-split into per-process chunk files"},
        "realtime_x": multi * 1e6 / FS,
        "cpu_baseline": {"value": multi, "unit": "Msamples/s", "cores": res["procs"], "kind": res["kind"],
                         "sample": f"{res['procs']} processes x {samples_per_step // max(1, res['procs'])} samples (adsbdec -a -m -f), "
                                   f"single process: {single:.1f} Msamples/s", "single_process": single},
        "e2e": {"value": multi, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# B200 arm

def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import adsbdec_b200 as A

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")     # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the demodulation path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    A.lib()   # fail loudly if the extension is missing
    if world > 1 and not os.environ.get("ADSB_B200_NO_BIND"):
        # keep each rank (and the page-locked memory it allocates) on the NUMA node of its GPU
        from adsbdec_b200.parallel import bind_to_gpu_numa_node
        bind_to_gpu_numa_node(local)

    seed = 1234 + rank     # every rank = its own receiver stream (weak scaling, no data-path collective)
    cap, injected, sig = build_device_capture(torch, dev, args.seconds, seed)
    n_samples = cap.numel()
    torch.cuda.synchronize()

    dec = A.Decoder(device=local, accept_df18=True, kernel=args.kernel)   # one handle, reset per step

    def step_resident():
        dec.reset()
        dec.push_device(cap.data_ptr(), n_samples)
        dec.finish()
        return dec

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident: warm-up, then K timed steps ----
    last = None
    for _ in range(args.warmup):
        step_resident()
    barrier()
    dec.fused_time(reset=True)
    dec.kernel_time(reset=True)
    dec.host_time(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fused_ms = fused_launches = fused_samples = 0.0
    launches = 0
    kernel_ms = 0.0
    frames = 0
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            d = step_resident()
            frames = d.num_frames
        ev1.record()
        barrier()
    fused_ms, fused_launches, fused_samples = dec.fused_time()
    kernel_ms, launches = dec.kernel_time()
    host_tel = dec.host_time()
    stats = dec.stats()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev)
    per_rank = [ms_total / args.steps]
    if dist:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [float(x.item()) / args.steps for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * n_samples / (ms_step * 1e-3) / 1e6

    # ---- end to end from pinned host memory through push() ----
    e2e = None
    if not args.no_e2e:
        host = torch.empty(n_samples, dtype=torch.int16, pin_memory=True)
        host.copy_(cap)
        torch.cuda.synchronize()
        host_np = host.numpy().view("<u2")

        def step_e2e():
            dec.reset()
            dec.push(host_np)
            dec.finish()
            fa = dec.frames_array          # the step's result, on the host
            return len(fa), fa.nbytes

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(max(1, args.steps // 2)):
            nf, d2h = step_e2e()
        ev1.record()
        barrier()
        ne = max(1, args.steps // 2)
        te = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if dist:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te.item()) / ne
        e2e = {"value": world * n_samples / (e_ms * 1e-3) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(2 * n_samples), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_ms, "frames": nf, "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / ne}
        assert nf == frames, (nf, frames)
        # the same capture in the airspy's packed 12-bit wire format (8 samples in 12 bytes):
        # a quarter less to move over PCIe, unpacked on the GPU.  Not the reference's file format: reported
        # next to the headline, never instead of it.
        if world == 1:
            n8 = n_samples - n_samples % 8
            packed = torch.empty(n8 // 8 * 3, dtype=torch.int32, pin_memory=True)
            packed.numpy().view("<u4")[:] = A.pack12(host_np[:n8])
            pk_np = packed.numpy().view("<u4")
            best = None
            for _ in range(3):
                dec.reset()
                torch.cuda.synchronize()
                tp = time.perf_counter()
                dec.push_packed12(pk_np, n8)
                dec.push(host_np[n8:])
                dec.finish()
                nfp = len(dec.frames_array)
                dt = time.perf_counter() - tp
                best = dt if best is None else min(best, dt)
            assert nfp == frames, (nfp, frames)
            e2e["packed12"] = {"value": n_samples / best / 1e6, "unit": "Msamples/s", "ms_per_step": best * 1e3,
                               "h2d_bytes_per_step": int(packed.numel() * 4),
                               "what": "pinned host memory, 12-bit packed samples -> adsbdec_b200_push_packed12 -> "
                                       "frames on the host, best of 3, wall clock"}
            del packed
        # the reference's own entry point for this path, fileInput (air.c:219-246): the same capture as a
        # page-cached file through adsbdec_b200_decode_file (rank 0 at N=1 only; needs 2.4 GB of /dev/shm)
        if world == 1 and os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 3 * host_np.nbytes:
            fpath = Path("/dev/shm") / f"adsb_b200_bench_{os.getpid()}.bin"
            try:
                host_np.tofile(fpath)
                best = None
                for _ in range(3):
                    dec.reset()
                    tf = time.perf_counter()
                    dec.decode_file(fpath)
                    nff = dec.num_frames
                    dt = time.perf_counter() - tf
                    best = dt if best is None else min(best, dt)
                assert nff == frames, (nff, frames)
                e2e["file"] = {"value": n_samples / best / 1e6, "unit": "Msamples/s", "ms_per_step": best * 1e3,
                               "what": "page-cached /dev/shm file -> adsbdec_b200_decode_file -> frames on the host, "
                                       "best of 3, wall clock"}
            finally:
                fpath.unlink(missing_ok=True)
        del host

    # ---- roofline of the dominant kernel (fused), CUDA events on its own stream ----
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak = float(json.loads(peaks_path.read_text())["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    roof = None
    if fused_launches:
        per_launch_ms = fused_ms / fused_launches
        achieved = 2.0 * fused_samples / fused_launches / (per_launch_ms * 1e-3) / 1e9
        # DRAM traffic of the kernel from the committed ncu capture (bytes moved per algorithmic byte),
        # scaled to this run's average launch
        traffic, traffic_src = None, None
        tp = ROOT / "profiles" / "r1_traffic.json"
        if tp.exists():
            t = json.loads(tp.read_text())
            ratio = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["algorithmic_bytes"]
            traffic = ratio * 2.0 * fused_samples / fused_launches
            traffic_src = f"profiles/r1_traffic.json: DRAM read+write = {ratio:.3f} x algorithmic bytes"
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "k_fused", "ms_per_launch": per_launch_ms,
                "algorithmic_bytes_per_launch": 2.0 * fused_samples / fused_launches, "peak_source": peak_src,
                "kernel_share_of_step": fused_ms / (ms_step * args.steps)}

    # ---- reference CPU path on this box's cores (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = cpu_sample(1234, 10.0)
        r = run_reference_cpu(sample, host_cores())
        cpu = {"value": r["multi_msps"], "unit": "Msamples/s", "cores": r["procs"], "kind": r["kind"],
               "sample": f"10 s of the workload: {r['procs']} processes x {r.get('chunk_samples', r['sample_samples'])} samples "
                         f"(adsbdec -a -m -f, page-cached chunk files); single process over the 10 s: "
                         f"{r['single_msps']:.1f} Msamples/s",
               "single_process": r["single_msps"]}

    if rank == 0:
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "samples_per_gpu": int(n_samples), "seconds_per_gpu": n_samples / FS,
                       "streams": world, "flags": "-a, Try/Ok stats on", "l2": "input 2.4 GB per step >> 126 MB L2",
                       "injected_frames": int(injected)},
            "realtime_x": value * 1e6 / FS, "frames_per_step": int(frames),
            "frames_per_s": world * frames / (ms_step * 1e-3),
            "stats_try_ok": [list(stats[0]), list(stats[1])],
            "kernel_ms_per_step": kernel_ms / args.steps, "ms_per_step_ranks": per_rank,
            "host_ms_per_step": {k: (v / args.steps if isinstance(v, float) else v) for k, v in host_tel.items()},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
