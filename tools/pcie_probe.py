"""Plain pinned host -> device copy rate of this box, the ceiling of the e2e number (python tools/pcie_probe.py)."""
import torch
n = 1_200_000_000
h = torch.empty(n, dtype=torch.int16, pin_memory=True)
d = torch.empty(n, dtype=torch.int16, device="cuda")
for chunk in (n, 1 << 25):
    for _ in range(2):
        d[:chunk].copy_(h[:chunk], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        for off in range(0, n - chunk + 1, chunk):
            d[off:off + chunk].copy_(h[off:off + chunk], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    moved = (n // chunk) * chunk * 2
    print(f"chunks of {chunk * 2 / 1e6:.0f} MB: {moved / ms / 1e6:.1f} GB/s ({ms:.2f} ms per {moved / 1e9:.2f} GB)")
