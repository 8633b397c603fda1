"""One long capture, time-chunked across the GPUs of a node - torch.distributed plumbing.

One process per GPU.  Every rank scans its chunks with ``adsbdec_b200_chunk_candidates`` (the
chunks carry a halo of one maximum-length frame plus preamble and their absolute start sample, so the
FIR rotation and the fidx wrap line up); the per-chunk candidate lists are gathered on rank 0 -
results only, there is no collective in the data path - merged by offset, and the sequential rules
are applied once by the host resolver.  The outcome is what a single decoder produces for the
whole capture.
"""
from __future__ import annotations

import os

import numpy as np

from . import Decoder
from .chunking import assign, merge_candidates, plan_chunks, resolve_merged


def bind_to_gpu_numa_node(local_rank: int) -> list[int] | None:
    """Pin this process to the CPUs next to its GPU (page-locked buffers are then allocated on that
    NUMA node, which is what the uploads want).  Best effort: returns the CPU list or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        cpus: list[int] = []
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def pin_host_array(a: np.ndarray) -> bool:
    """Page-lock an existing host array in place (cudaHostRegister) so that uploads out of it are real
    asynchronous DMA.  Returns False (and leaves it pageable) if the driver refuses."""
    import torch
    try:
        rc = torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)
        return int(rc) == 0
    except Exception:
        return False


def unpin_host_array(a: np.ndarray) -> None:
    import torch
    try:
        torch.cuda.cudart().cudaHostUnregister(a.ctypes.data)
    except Exception:
        pass


def scan_chunks(dec: Decoder, samples: np.ndarray, chunks, device: int, cand_cap: int = 1 << 20):
    """Upload and scan `chunks` of the host array `samples` (best page-locked) on `device`, the upload of the
    next chunk running on a side stream beside the scan of the current one (two device buffers).  Yields
    (chunk, candidates)."""
    import torch
    if not chunks:
        return
    dev = torch.device("cuda", device)
    longest = max(c.sample_end - c.sample_begin for c in chunks)
    bufs = [torch.empty(longest, dtype=torch.int16, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    host = samples.view(np.int16)

    def upload(i):
        ch = chunks[i]
        n = ch.sample_end - ch.sample_begin
        with torch.cuda.stream(copy_stream):
            bufs[i % 2][:n].copy_(torch.from_numpy(host[ch.sample_begin:ch.sample_end]), non_blocking=True)
            ready[i % 2].record(copy_stream)

    upload(0)
    for i, ch in enumerate(chunks):
        ready[i % 2].synchronize()
        if i + 1 < len(chunks):
            upload(i + 1)          # into the other buffer: its last scan has returned (chunk_candidates is synchronous)
        lo, hi = ch.local_own
        cands, _ = dec.chunk_candidates(bufs[i % 2].data_ptr(), ch.sample_end - ch.sample_begin, ch.sample_begin, lo, hi,
                                        cap=cand_cap)
        yield ch, cands


def decode_chunked(samples: np.ndarray, *, n_chunks: int | None = None, accept_df18: bool = False,
                   emulate_reference_tail: bool = True, device: int | None = None, as_array: bool = False):
    """Decode ``samples`` (host uint16) split over all ranks of the default process group (or on
    one GPU when torch.distributed is not initialised).  Returns the accepted frames on rank 0
    (``None`` elsewhere).  The chunks are uploaded from page-locked memory (the array is registered for the
    duration of the call if it is not already)."""
    import torch
    import torch.distributed as dist

    have_pg = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if have_pg else 0
    world = dist.get_world_size() if have_pg else 1
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if have_pg else 0
    torch.cuda.set_device(device)
    samples = np.ascontiguousarray(samples, dtype="<u2")
    chunks = plan_chunks(samples.size, n_chunks or 4 * world)
    dec = Decoder(device=device, accept_df18=accept_df18, collect_stats=False)
    pinned_here = pin_host_array(samples)
    try:
        mine = list(scan_chunks(dec, samples, assign(chunks, world, rank), device))
    finally:
        if pinned_here:
            unpin_host_array(samples)
    dec.close()
    if have_pg:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)
        if rank != 0:
            return None
        parts = [p for part in gathered for p in part]
    else:
        parts = mine
    return resolve_merged(merge_candidates(parts), samples.size, emulate_reference_tail=emulate_reference_tail,
                          as_array=as_array)
