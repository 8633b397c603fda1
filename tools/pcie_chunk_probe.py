"""Pinned-upload rate as a function of the copy size: one 2.4 GB copy against back-to-back copies of 32 / 64 / 128 /
256 MB on one stream (what adsbdec_b200_push issues per batch), and a push() of the same buffer with different
batch sizes.  usage: python tools/pcie_chunk_probe.py"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import adsbdec_b200 as A

n = 1_202_097_152
host = torch.full((n,), 2048, dtype=torch.int16).pin_memory()
dev = torch.empty(n, dtype=torch.int16, device="cuda")
s = torch.cuda.Stream()
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best
def chunks(mb):
    step = mb * 1024 * 1024 // 2
    def f():
        with torch.cuda.stream(s):
            for o in range(0, n, step):
                dev[o:o + step].copy_(host[o:o + step], non_blocking=True)
    return f
for mb in (0, 32, 64, 128, 256, 512):
    t = timed(chunks(mb) if mb else (lambda: dev.copy_(host, non_blocking=True)))
    print(f"copy in chunks of {mb or 'all'} MB: {2 * n / t / 1e9:.1f} GB/s ({t * 1e3:.2f} ms)")
arr = host.numpy().view("<u2")
for batch in (16 << 20, 32 << 20, 64 << 20, 128 << 20):
    d = A.Decoder(accept_df18=True, batch_samples=batch)
    def f():
        d.reset(); d.push(arr); d.finish()
    f()
    t = timed(f)
    print(f"push() of mid-scale samples, batch {batch >> 20} Mi samples: {n / t / 1e9:.2f} G samples/s = {2 * n / t / 1e9:.1f} GB/s ({t * 1e3:.2f} ms)")
    d.close()
