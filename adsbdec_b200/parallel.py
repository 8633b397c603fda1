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


def decode_chunked(samples: np.ndarray, *, n_chunks: int | None = None, accept_df18: bool = False,
                   emulate_reference_tail: bool = True, device: int | None = None):
    """Decode ``samples`` (host uint16) split over all ranks of the default process group (or on
    one GPU when torch.distributed is not initialised).  Returns the accepted frames on rank 0
    (``None`` elsewhere)."""
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
    mine = []
    for ch in assign(chunks, world, rank):
        seg = torch.from_numpy(samples[ch.sample_begin:ch.sample_end].view(np.int16)).to(f"cuda:{device}")
        lo, hi = ch.local_own
        cands, _ = dec.chunk_candidates(seg.data_ptr(), seg.numel(), ch.sample_begin, lo, hi)
        mine.append((ch, cands))
    dec.close()
    if have_pg:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)
        if rank != 0:
            return None
        parts = [p for part in gathered for p in part]
    else:
        parts = mine
    return resolve_merged(merge_candidates(parts), samples.size, emulate_reference_tail=emulate_reference_tail)
