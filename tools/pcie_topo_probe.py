"""Host->device copy bandwidth of every GPU alone, of pairs, and of all GPUs at once (one process, one pinned
buffer and one stream per GPU): names what limits the end-to-end number when several ranks upload at the same
time - a shared PCIe uplink shows as a pair that halves, host memory as an aggregate ceiling.
  python tools/pcie_topo_probe.py [MiB per copy]        (needs a multi-GPU box: gpurun --gpus 8)"""
import itertools
import json
import subprocess
import sys
import time

import torch

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = torch.cuda.device_count()
host = [torch.empty(mib << 20, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
dev = [torch.empty(mib << 20, dtype=torch.uint8, device=f"cuda:{g}") for g in range(n)]
streams = [torch.cuda.Stream(device=g) for g in range(n)]


def run(gpus, reps=4):
    for g in gpus:
        torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for _ in range(reps):
        for g in gpus:
            with torch.cuda.stream(streams[g]):
                dev[g].copy_(host[g], non_blocking=True)
    for g in gpus:
        streams[g].synchronize()
    dt = time.perf_counter() - t0
    return reps * (mib << 20) / dt / 1e9   # GB/s per GPU


out = {"gpus": n, "mib_per_copy": mib}
for g in range(n):
    run([g], 1)
out["alone_gbs"] = [round(run([g]), 1) for g in range(n)]
if n > 1:
    out["pairs_gbs_each"] = {f"{a},{b}": round(run([a, b]), 1) for a, b in itertools.combinations(range(n), 2)
                             if a == 0 or b == a + 1}
    for k in (2, 4, 8):
        if k <= n:
            v = run(list(range(k)))
            out[f"first_{k}_gbs_each"] = round(v, 1)
            out[f"first_{k}_gbs_total"] = round(v * k, 1)
try:
    out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
except Exception as e:
    out["topo"] = str(e)
print(json.dumps(out, indent=1))
