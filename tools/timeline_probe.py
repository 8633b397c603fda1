import sys, time
sys.path.insert(0, '/root/repo')
import torch
import adsbdec_b200 as A
from bench import build_device_capture
dev = torch.device("cuda", 0)
cap, injected, _ = build_device_capture(torch, dev, 60.0, 1234)
torch.cuda.synchronize()
dec = A.Decoder(accept_df18=True)
for i in range(5):
    dec.reset(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    dec.push_device(cap.data_ptr(), cap.numel())
    t1 = time.perf_counter()
    dec.finish()
    t2 = time.perf_counter()
    print(f"push_device {1e3*(t1-t0):.3f} ms, finish {1e3*(t2-t1):.3f} ms, kernels {dec.kernel_time(reset=True)[0]:.3f}")
for stats in (False,):
    d2 = A.Decoder(accept_df18=True, collect_stats=False)
    for i in range(4):
        d2.reset(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        d2.push_device(cap.data_ptr(), cap.numel())
        t1 = time.perf_counter()
        d2.finish()
        t2 = time.perf_counter()
        print(f"nostats: push_device {1e3*(t1-t0):.3f} ms, finish {1e3*(t2-t1):.3f} ms, kernels {d2.kernel_time(reset=True)[0]:.3f}")
