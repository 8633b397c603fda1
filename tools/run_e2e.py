"""End-to-end timing breakdown from pinned host memory: python tools/run_e2e.py [seconds] [batch_Mi]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import adsbdec_b200 as A
from bench import build_device_capture

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
dev = torch.device("cuda", 0)
cap, injected, _ = build_device_capture(torch, dev, seconds, 1234)
host = torch.empty(cap.numel(), dtype=torch.int16, pin_memory=True)
host.copy_(cap)
torch.cuda.synchronize()
hn = host.numpy().view("<u2")
for batch in [int(x) << 20 for x in (sys.argv[2:] or ["16", "32", "64", "128"])]:
    dec = A.Decoder(accept_df18=True, batch_samples=batch)
    for it in range(3):
        dec.reset(); dec.host_time(reset=True); dec.kernel_time(reset=True); dec.fused_time(reset=True)
        t0 = time.perf_counter(); dec.push(hn); t1 = time.perf_counter(); dec.finish(); t2 = time.perf_counter()
        ht = dec.host_time()
    print(f"batch {batch>>20:4d} Mi: push {1e3*(t1-t0):7.2f} ms finish {1e3*(t2-t1):6.2f} ms total {1e3*(t2-t0):7.2f} ms "
          f"-> {cap.numel()/(t2-t0)/1e9:6.2f} Gsamples/s; kernels {dec.kernel_time()[0]:.2f} ms fused {dec.fused_time(reset=True)[0]:.2f} ms; host " +
          " ".join(f"{k}={v:.2f}" for k, v in ht.items()))
    # resident for comparison
    for it in range(2):
        dec.reset(); dec.host_time(reset=True)
        t0 = time.perf_counter(); dec.push_device(cap.data_ptr(), cap.numel()); dec.finish(); t2 = time.perf_counter()
    print(f"      resident total {1e3*(t2-t0):7.2f} ms host " + " ".join(f"{k}={v:.2f}" for k, v in dec.host_time().items()))
    dec.close()
