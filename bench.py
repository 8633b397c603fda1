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
             (+ roofline.path: k_fused + k_gate + k_slice, the whole per-offset path)
  cpu_baseline  the reference's own CPU path (oracle/_ref binary built from the unmodified
             sources, or the oracle port) on the run's own capture: one process over all of it, then
             one process per core on pre-split chunk files
  parity     the run certifies itself: the reference's stdout and Try/Ok on the capture are compared
             byte for byte with what this library decoded from it ("identical": true, sha256 of both)

The capture is rendered on the host (seeded per tile), so `--impl reference` - which times the
reference CPU implementation (one process per core on pre-split chunk files of the same 60 s capture)
and prints the same JSON line with "impl": "reference" - sees the same bytes.
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
SEG_SECONDS = 2.0          # host-rendered dense segment, tiled with fresh noise per tile
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
    ap.add_argument("--mode", default="stream", choices=["stream", "chunked", "streams256"],
                    help="stream: the headline (one 60 s capture per GPU); chunked / streams256: the two halves of "
                         "configs[4] - ONE long capture halo-chunked over the GPUs, 256 receiver streams over the GPUs")
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


PAD_SAMPLES = 1 << 21      # mid-scale tail: lets the reference drain its frame queue before EOF


def quantise_host(sig: np.ndarray, seed: int, tile: int) -> np.ndarray:
    rng = np.random.default_rng([seed, 7, tile])
    x = rng.standard_normal(sig.size, dtype=np.float32)
    x *= np.float32(SIGMA)
    x += sig
    x += np.float32(2048.0)
    np.rint(x, out=x)
    np.clip(x, 0, 4095, out=x)
    return x.astype("<u2")


def build_capture(seconds: float, seed: int, out: np.ndarray | None = None):
    """The workload, rendered on the HOST so that both arms (and the parity check inside the run) see the
    same bytes: the dense 2 s segment tiled, every tile with its own AWGN (seeded per tile), then 2 Mi
    mid-scale samples.  Tiles are rendered by a thread pool (numpy releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    sig, n_frames = dense_segment(seed)
    tiles = max(1, int(round(seconds / SEG_SECONDS)))
    n = tiles * sig.size + PAD_SAMPLES
    cap = out if out is not None else np.empty(n, dtype="<u2")
    assert cap.size == n

    def one(i):
        cap[i * sig.size:(i + 1) * sig.size] = quantise_host(sig, seed, i)

    with ThreadPoolExecutor(max_workers=max(1, min(tiles, host_cores()))) as ex:
        list(ex.map(one, range(tiles)))
    cap[tiles * sig.size:] = 2048
    return cap, n_frames * tiles


def build_device_capture(torch, dev, seconds: float, seed: int):
    """build_capture() uploaded to `dev` (tools/ use this): (int16 tensor, injected frames, None)."""
    cap, injected = build_capture(seconds, seed)
    return torch.from_numpy(cap.view(np.int16)).to(dev), injected, None


def cpu_sample(seed: int, seconds: float) -> np.ndarray:
    """The first `seconds` of the workload as a host array (tools/)."""
    return build_capture(seconds, seed)[0]


def capture_samples(seconds: float) -> int:
    n = int(SEG_SECONDS * FS)
    n -= n % 8
    return max(1, int(round(seconds / SEG_SECONDS))) * n + PAD_SAMPLES


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


class RefFiles:
    """The capture as page-cached files for the reference binary: the whole capture, and the same samples
    pre-split into one chunk per process.  /dev/shm when there is room (2 x the capture), else a shorter
    prefix of the capture for the chunks."""

    def __init__(self, cap: np.ndarray, n_procs: int):
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        self.dir = Path(tempfile.mkdtemp(prefix="adsb_cpu_", dir=base))
        self.cap = cap
        self.n_procs = n_procs
        self.whole = self.dir / "whole.bin"
        cap.tofile(self.whole)
        free = shutil.disk_usage(self.dir).free
        body = cap.size - PAD_SAMPLES
        use = body if free > cap.nbytes + (1 << 28) else min(body, int(free * 0.8) // 2)
        per = use // n_procs
        per -= per % 8
        self.chunk_samples = per
        self.files = []
        for i in range(n_procs):
            f = self.dir / f"chunk{i}.bin"
            cap[i * per:(i + 1) * per].tofile(f)
            self.files.append(f)

    def close(self):
        for f in self.dir.glob("*"):
            f.unlink()
        self.dir.rmdir()


def ref_flavour(binary) -> str:
    return "gcc -O3 -march=native (the reference's own flags)" if str(binary).endswith("_native") else \
        "gcc -O3 -march=x86-64-v3 (same FMA-contracted arithmetic; runs on any AVX2 host)"


def run_reference_single(binary, files: RefFiles, flags=("-a", "-m")):
    """One `adsbdec -a -m -f` process over the whole capture: (Msamples/s, stdout, (try, ok))."""
    from oracle import pyoracle
    t0 = time.perf_counter()
    r = pyoracle.run_ref(files.whole, flags=flags, binary=binary, timeout=900)
    dt = time.perf_counter() - t0
    return files.cap.size / dt / 1e6, r.stdout, r.stats()


def run_reference_multi(binary, files: RefFiles, flags=("-a", "-m")):
    """One process per core, each on its own pre-split chunk file: aggregate Msamples/s."""
    with open(os.devnull, "wb") as null:
        t0 = time.perf_counter()
        procs = [subprocess.Popen([str(binary), *flags, "-f", str(f)], stdout=null, stderr=null) for f in files.files]
        for p in procs:
            p.wait()
        dt = time.perf_counter() - t0
    return files.chunk_samples * len(files.files) / dt / 1e6


def run_port_cpu(cap: np.ndarray):
    """Fallback when the compiled reference cannot run on this host: the oracle port, one thread, on 2 s."""
    from oracle import pyoracle
    n = min(cap.size, 2 * FS)
    t0 = time.perf_counter()
    pyoracle.decode(cap[:n], accept_df18=True)
    return n / (time.perf_counter() - t0) / 1e6, n


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    cap, _ = build_capture(args.seconds, 1234)     # the b200 arm's capture of rank 0, byte for byte
    binary = ref_binary()
    if binary is None:
        v, n = run_port_cpu(cap)
        line = {"impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
                "steps": 1, "warmup": 0, "ms_per_step": n / v / 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "sample": f"first {n} samples, oracle port (oracle/_ref cannot run on this host)"},
                "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": 1, "kind": "port", "sample": f"{n} samples"},
                "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0
    files = RefFiles(cap, cores)
    try:
        single, _, _ = run_reference_single(binary, files)
        vals = []
        for i in range(args.warmup + args.steps):
            v = run_reference_multi(binary, files)
            if i >= args.warmup:
                vals.append(v)
    finally:
        files.close()
    multi = float(np.mean(vals))
    samples_per_step = files.chunk_samples * cores
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": multi, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": samples_per_step / multi / 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "samples_per_gpu": int(cap.size), "seconds_per_gpu": cap.size / FS,
                   "sample": f"the whole capture per step, pre-split into {cores} per-process chunk files of "
                             f"{files.chunk_samples} samples"},
        "realtime_x": multi * 1e6 / FS,
        "cpu_baseline": {"value": multi, "unit": "Msamples/s", "cores": cores, "kind": "reference",
                         "sample": f"{cores} processes x {files.chunk_samples} samples (adsbdec -a -m -f, page-cached chunk "
                                   f"files of the 60 s capture); one process over the whole capture: {single:.1f} "
                                   f"Msamples/s; binary: {Path(binary).name}, {ref_flavour(binary)}",
                         "single_process": single},
        "e2e": {"value": multi, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# configs[4]: one long capture halo-chunked over the GPUs / 256 receiver streams over the GPUs

def _max_over_ranks(torch, dist, dev, x: float) -> float:
    if not dist:
        return x
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def mode_chunked(args, torch, A, dist, world, rank, local):
    """ONE capture (default 240 s = 4.8 G samples, longer than a 2^32-sample fidx epoch) in page-locked host
    memory shared by the ranks; chunks with a 1200-sample halo go round-robin to the GPUs
    (adsbdec_b200_chunk_candidates, upload beside scan), candidate lists to rank 0 (gather_object: results
    only), one sequential resolve there.  Timed: uploads + scans + gather + merge + resolve.  Checked in the
    run against a single decoder fed the same capture."""
    from adsbdec_b200.chunking import assign, merge_candidates, plan_chunks, resolve_merged
    from adsbdec_b200.parallel import pin_host_array, scan_chunks
    dev = torch.device("cuda", local)
    seconds = args.seconds if args.seconds != 60.0 else 240.0
    n = capture_samples(seconds)
    path = Path("/dev/shm") / f"adsb_b200_chunked_{os.environ.get('MASTER_PORT', '0')}.bin"
    if rank == 0:
        cap = np.memmap(path, dtype="<u2", mode="w+", shape=(n,))
        build_capture(seconds, 4242, out=cap)
        cap.flush()
    if dist:
        dist.barrier()
    cap = np.memmap(path, dtype="<u2", mode="r+", shape=(n,))
    pinned = pin_host_array(cap)
    chunks = plan_chunks(n, max(4 * world, -(-n // (256 << 20))))
    mine_chunks = assign(chunks, world, rank)
    dec = A.Decoder(device=local, accept_df18=True, collect_stats=False)
    times, frames = [], None
    for it in range(args.warmup + args.steps):
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        mine = list(scan_chunks(dec, cap, mine_chunks, local))
        t_scan = time.perf_counter() - t0
        if dist:
            # results only: every rank's candidate records as one byte tensor, gathered on rank 0 over NCCL
            # (a pickled gather_object of two million records cost more than the scans)
            counts = torch.tensor([len(c) for _, c in mine] + [0] * (len(chunks) - len(mine)), device=dev)
            all_counts = [torch.zeros_like(counts) for _ in range(world)]
            dist.all_gather(all_counts, counts)
            flat = np.concatenate([c for _, c in mine]) if mine else np.zeros(0, dtype=A.CAND_DTYPE)
            longest = max(int(ac.sum().item()) for ac in all_counts)
            send = torch.zeros(max(1, longest) * A.CAND_DTYPE.itemsize, dtype=torch.uint8, device=dev)
            send[:flat.nbytes] = torch.from_numpy(flat.view(np.uint8)).to(dev)
            recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
            dist.gather(send, recv, dst=0)
            parts = None
            if rank == 0:
                parts = []
                for r in range(world):
                    rc = recv[r].cpu().numpy().view(A.CAND_DTYPE)
                    at = 0
                    for ch, cnt in zip(assign(chunks, world, r), all_counts[r].tolist()):
                        parts.append((ch, rc[at:at + cnt]))
                        at += cnt
        else:
            parts = mine
        t_gather = time.perf_counter() - t0
        if rank == 0:
            frames = resolve_merged(merge_candidates(parts), n, as_array=True)
        dt = _max_over_ranks(torch, dist, dev, time.perf_counter() - t0)
        if it >= args.warmup:
            times.append((dt, _max_over_ranks(torch, dist, dev, t_scan), t_gather - t_scan))
    line = None
    if rank == 0:
        single = A.Decoder(device=local, accept_df18=True, collect_stats=False)
        single.push(cap)
        single.finish()
        ref = single.frames_array
        same = len(ref) == len(frames) and all(np.array_equal(ref[k], frames[k]) for k in ("ts", "pw", "len", "bytes"))
        ms = float(np.mean([t[0] for t in times])) * 1e3
        line = {"mode": "chunked", "metric": "Msamples/s", "value": n / (ms * 1e-3) / 1e6, "unit": "Msamples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "scan_ms_per_step": float(np.mean([t[1] for t in times])) * 1e3,
                "gather_ms_per_step": float(np.mean([t[2] for t in times])) * 1e3,
                "merge_resolve_ms_per_step": float(np.mean([t[0] - t[1] - t[2] for t in times])) * 1e3,
                "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"ONE {seconds:.0f} s capture ({n} samples, {2 * n / 1e9:.1f} GB, crosses a 2^32-sample "
                                       f"fidx wrap) from page-locked host memory, {len(chunks)} chunks with a 1200-sample "
                                       f"halo round-robin over {world} GPU(s), host merge + resolve on rank 0 (configs[4])",
                           "chunks": len(chunks), "pinned": bool(pinned)},
                "realtime_x": n / (ms * 1e-3) / FS, "frames": int(len(frames)),
                "h2d_bytes_per_step": int(sum(2 * (c.sample_end - c.sample_begin) for c in chunks)),
                "parity": {"parity_checked": True, "identical": bool(same),
                           "what": "frames (ts, pw, len, bytes) of the chunked decode vs ONE decoder fed the whole capture"}}
        print(json.dumps(line))
    if dist:
        dist.barrier()
    del cap
    if rank == 0:
        path.unlink(missing_ok=True)
    return 0


def mode_streams256(args, torch, A, dist, world, rank, local):
    """256 receiver streams spread over the GPUs (256 / N decoder handles per GPU), fed the way live
    receivers deliver: 1 Mi-sample buffers (52 ms of signal, the reference's read size) round-robin over the
    streams, from page-locked host memory, frames fetched as they come (flush + poll).  Every stream of a
    rank carries the same 4 s capture (replicated data); all of them must decode to what one plain decode of
    that capture gives."""
    dev = torch.device("cuda", local)
    total_streams = 256
    per = total_streams // world
    seconds = 4.0 if args.seconds == 60.0 else args.seconds
    n = capture_samples(seconds)
    host = torch.empty(n, dtype=torch.int16, pin_memory=True)
    host_np = host.numpy().view("<u2")
    build_capture(seconds, 777 + rank, out=host_np)
    ref = A.Decoder(device=local, accept_df18=True)
    ref.push(host_np)
    ref.finish()
    want, want_stats = ref.frames_array, ref.stats()
    ref.close()
    decs = [A.Decoder(device=local, accept_df18=True, batch_samples=4 << 20) for _ in range(per)]
    piece = 1 << 20
    times = []
    for it in range(args.warmup + args.steps):
        for d in decs:
            d.reset()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for at in range(0, n, piece):
            for d in decs:
                d.push(host_np[at:at + piece])
                d.flush()
            for d in decs:
                d.poll()
        for d in decs:
            d.finish()
        dt = _max_over_ranks(torch, dist, dev, time.perf_counter() - t0)
        if it >= args.warmup:
            times.append(dt)
    ok = True
    for d in decs:
        fa = d.frames_array
        ok = ok and len(fa) == len(want) and all(np.array_equal(fa[k], want[k]) for k in ("ts", "pw", "len", "bytes")) \
            and d.stats() == want_stats
    okt = torch.tensor([1.0 if ok else 0.0], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        ms = float(np.mean(times)) * 1e3
        total = per * world * n
        line = {"mode": "streams256", "metric": "Msamples/s", "value": total / (ms * 1e-3) / 1e6, "unit": "Msamples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic (replicated per stream)",
                "config": {"workload": f"{per * world} receiver streams x {seconds:.0f} s, {per} decoder handles per GPU on "
                                       f"{world} GPU(s), 1 Mi-sample pushes round-robin from page-locked host memory, "
                                       f"flush + poll after every push (configs[4])",
                           "streams": per * world, "samples_per_stream": int(n)},
                "realtime_x_per_stream": n / (ms * 1e-3) / FS, "streams_at_realtime": total / (ms * 1e-3) / FS,
                "h2d_bytes_per_step": int(2 * total),
                "parity": {"parity_checked": True, "identical": bool(okt.item() > 0.5),
                           "what": "every stream's frames, timestamps and Try/Ok vs one plain decode of its capture"}}
        print(json.dumps(line))
    for d in decs:
        d.close()
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
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        # NCCL's version banner goes to stdout when the communicator comes up: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            torch.cuda.set_device(local)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the demodulation path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    A.lib()   # fail loudly if the extension is missing
    if world > 1 and not os.environ.get("ADSB_B200_NO_BIND"):
        # keep each rank (and the page-locked memory it allocates) on the NUMA node of its GPU
        from adsbdec_b200.parallel import bind_to_gpu_numa_node
        bind_to_gpu_numa_node(local)

    if args.mode != "stream":
        rc = (mode_chunked if args.mode == "chunked" else mode_streams256)(args, torch, A, dist, world, rank, local)
        if dist:
            dist.destroy_process_group()
        return rc

    seed = 1234 + rank     # every rank = its own receiver stream (weak scaling, no data-path collective)
    n_samples = capture_samples(args.seconds)
    host = torch.empty(n_samples, dtype=torch.int16, pin_memory=True)   # page-locked: the e2e leg uploads from here
    host_np = host.numpy().view("<u2")
    _, injected = build_capture(args.seconds, seed, out=host_np)
    cap = host.to(dev)
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

    def gather(x: float):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if not dist:
            return [x]
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        return [float(v.item()) for v in allt]

    # ---- device-resident: warm-up, then K timed steps ----
    for _ in range(args.warmup):
        step_resident()
    barrier()
    dec.fused_time(reset=True)
    dec.gate_time(reset=True)
    dec.kernel_time(reset=True)
    dec.host_time(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    frames = 0
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            d = step_resident()
            frames = d.num_frames
        ev1.record()
        barrier()
    fused_ms, fused_launches, fused_samples = dec.fused_time()
    gate_ms = dec.gate_time()
    kernel_ms, launches = dec.kernel_time()
    host_tel = dec.host_time()
    stats = dec.stats()
    frames_arr = dec.frames_array          # the last step's result: compared with the reference below
    per_rank = [v / args.steps for v in gather(ev0.elapsed_time(ev1))]
    ms_step = max(per_rank)
    value = world * n_samples / (ms_step * 1e-3) / 1e6
    host_ranks = gather(sum(v for k, v in host_tel.items() if k.endswith("_ms")) / args.steps)

    # ---- end to end from pinned host memory through push() ----
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            dec.reset()
            dec.push(host_np)
            dec.finish()
            # the step's result is on the host when finish() returns (the library fetches the accepted frames into
            # its page-locked store): read it where it lies - a copy into a fresh numpy array cost 3 ms per step
            fa = dec.frames_view
            assert int(fa["n"][-1]) > int(fa["n"][0])
            return len(fa), fa.nbytes

        step_e2e()
        barrier()
        ne = max(1, args.steps // 2)
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(ne):
            nf, d2h = step_e2e()
        ev1.record()
        barrier()
        e_ranks = [v / ne for v in gather(ev0.elapsed_time(ev1))]
        e_ms = max(e_ranks)
        e2e = {"value": world * n_samples / (e_ms * 1e-3) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(2 * n_samples), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_ms, "frames": nf, "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / ne,
               "ms_per_step_ranks": e_ranks, "h2d_gbs_ranks": [2 * n_samples / (m * 1e-3) / 1e9 for m in e_ranks]}
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
                nfp = len(dec.frames_view)
                dt = time.perf_counter() - tp
                best = dt if best is None else min(best, dt)
            assert nfp == frames, (nfp, frames)
            e2e["packed12"] = {"value": n_samples / best / 1e6, "unit": "Msamples/s", "ms_per_step": best * 1e3,
                               "h2d_bytes_per_step": int(packed.numel() * 4),
                               "what": "pinned host memory, 12-bit packed samples -> adsbdec_b200_push_packed12 -> "
                                       "frames on the host, best of 3, wall clock"}
            del packed

    # ---- roofline of the dominant kernel (k_fused), CUDA events on its own stream ----
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak = float(json.loads(peaks_path.read_text())["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    roof = None
    if fused_launches:
        per_launch_ms = fused_ms / fused_launches
        alg = 2.0 * fused_samples / fused_launches
        achieved = alg / (per_launch_ms * 1e-3) / 1e9
        # DRAM traffic of the kernel from the committed ncu capture of this build (bytes moved per algorithmic
        # byte), scaled to this run's average launch
        traffic, traffic_src = None, None
        tp = ROOT / "profiles" / "r2_traffic.json"
        if tp.exists():
            t = json.loads(tp.read_text())
            ratio = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["algorithmic_bytes"]
            traffic = ratio * alg
            traffic_src = f"profiles/r2_traffic.json: DRAM read+write = {ratio:.3f} x algorithmic bytes"
        path_ms = (fused_ms + gate_ms) / fused_launches
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "k_fused", "ms_per_launch": per_launch_ms,
                "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                "kernel_share_of_step": fused_ms / (ms_step * args.steps),
                # the whole per-offset path (k_fused + k_gate + k_slice: samples -> CRC-valid candidates)
                "path": {"kernels": "k_fused + k_gate + k_slice", "ms_per_launch": path_ms,
                         "achieved": alg / (path_ms * 1e-3) / 1e9, "frac": alg / (path_ms * 1e-3) / 1e9 / peak,
                         "share_of_step": (fused_ms + gate_ms) / (ms_step * args.steps)}}

    # ---- the reference's own CPU path on this box's cores, on the SAME capture, and the parity check
    #      (rank 0, N=1 only): one `adsbdec -a -m -f` process over the whole capture gives the golden output
    #      of this very run; then one process per core on pre-split chunks for the multi-core figure ----
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import hashlib
        binary = ref_binary()
        ours = A.format_frames(1, frames_arr)
        if binary is not None:
            files = RefFiles(host_np, host_cores())
            try:
                single, ref_out, ref_stats = run_reference_single(binary, files)
                multi = run_reference_multi(binary, files)
                # the same capture as a page-cached file through adsbdec_b200_decode_file, the reference's own
                # entry point for this path (fileInput, air.c:219-246)
                best = None
                for _ in range(3):
                    dec.reset()
                    tf = time.perf_counter()
                    dec.decode_file(files.whole)
                    nff = dec.num_frames
                    dt = time.perf_counter() - tf
                    best = dt if best is None else min(best, dt)
                file_out = A.format_frames(1, dec.frames_array)
                if e2e is not None:
                    e2e["file"] = {"value": n_samples / best / 1e6, "unit": "Msamples/s", "ms_per_step": best * 1e3,
                                   "what": "page-cached /dev/shm file -> adsbdec_b200_decode_file -> frames on the "
                                           "host, best of 3, wall clock"}
            finally:
                files.close()
            same = ours == ref_out and file_out == ref_out and ref_stats == (tuple(stats[0]), tuple(stats[1]))
            parity = {"parity_checked": True, "identical": bool(same),
                      "what": "avrmlat byte stream (adsbdec -a -m -f stdout) and Try/Ok counters of the unmodified "
                              "reference on this run's 60 s capture vs push_device and decode_file of this library",
                      "frames": int(len(frames_arr)), "bytes": len(ref_out),
                      "sha256_reference": hashlib.sha256(ref_out).hexdigest(),
                      "sha256_b200": hashlib.sha256(ours).hexdigest(),
                      "try_ok_reference": [list(ref_stats[0] or ()), list(ref_stats[1] or ())]}
            if not same:
                print(f"PARITY FAILURE: {len(ours)} vs {len(ref_out)} bytes, stats {stats} vs {ref_stats}", file=sys.stderr)
            cpu = {"value": multi, "unit": "Msamples/s", "cores": host_cores(), "kind": "reference",
                   "sample": f"the run's own 60 s capture: {host_cores()} processes x {files.chunk_samples} samples "
                             f"(adsbdec -a -m -f, page-cached chunk files); one process over the whole capture: "
                             f"{single:.1f} Msamples/s; binary: {Path(binary).name}, {ref_flavour(binary)}",
                   "single_process": single}
        else:
            v, n = run_port_cpu(host_np)
            cpu = {"value": v, "unit": "Msamples/s", "cores": 1, "kind": "port",
                   "sample": f"first {n} samples of the capture, oracle port (oracle/_ref cannot run on this host)"}
            parity = {"parity_checked": False, "why": "oracle/_ref cannot run on this host"}

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
            "host_ms_per_step_ranks": host_ranks,
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "parity": parity,
        }
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
