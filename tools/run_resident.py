"""One device-resident decode of a synthetic capture (for ncu): python tools/run_resident.py [seconds] [steps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import adsbdec_b200 as A
from bench import build_device_capture

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
cap, injected, _ = build_device_capture(torch, dev, seconds, 1234)
torch.cuda.synchronize()
dec = A.Decoder(accept_df18=True)
import time
for i in range(steps):
    dec.reset()
    dec.host_time(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dec.push_device(cap.data_ptr(), cap.numel())
    dec.finish()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"   wall {wall:.3f} ms, kernels {dec.kernel_time(reset=True)[0]:.3f} ms, host {dec.host_time()}")
    ms, n, s = dec.fused_time(reset=True)
    print(f"step {i}: frames {dec.num_frames}, fused {ms:.3f} ms / {n} launches, {2*s/ms/1e6 if ms else 0:.1f} GB/s, gate+slice {dec.gate_time(reset=True):.3f} ms")
