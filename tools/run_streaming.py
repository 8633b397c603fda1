"""Live-source pattern (rx_callback -> decodeiq, air.c:173-177): fixed-size pushes from pageable host memory,
flush() + poll() after each.  Reports the time per push and the delay between handing a buffer over and
getting its frames.  usage: python tools/run_streaming.py [seconds] [samples per push]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from bench import cpu_sample
import adsbdec_b200 as A

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
piece = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
cap = cpu_sample(1234, seconds)
cap = np.concatenate([cap, np.full(1 << 21, 2048, dtype="<u2")])
n_frames = [0]
dec = A.Decoder(accept_df18=True, batch_samples=max(piece, 1 << 17))
for rep in range(2):
    dec.reset()
    push_ms, lag_ms, pend = [], [], []
    got = 0
    t_start = time.perf_counter()
    for pos in range(0, cap.size - cap.size % 4, piece):
        t0 = time.perf_counter()
        dec.push(cap[pos:pos + piece])
        k = dec.flush()
        pend.append(t0)
        k += dec.poll()
        t1 = time.perf_counter()
        push_ms.append((t1 - t0) * 1e3)
        if k:
            # frames of the oldest outstanding pushes arrived
            while len(pend) > 2:
                lag_ms.append((t1 - pend.pop(0)) * 1e3)
            got += k
    dec.finish()
    wall = time.perf_counter() - t_start
    p = np.array(push_ms)
    l = np.array(lag_ms) if lag_ms else np.zeros(1)
    print(f"rep {rep}: {cap.size/1e6:.0f} M samples in pushes of {piece} ({piece/20e3:.1f} ms of signal): "
          f"{cap.size/wall/1e6:.0f} Msamples/s = {cap.size/wall/20e6:.0f} x real time; per push median {np.median(p):.3f} ms, "
          f"p99 {np.percentile(p, 99):.3f} ms; push -> frames median {np.median(l):.3f} ms, max {l.max():.3f} ms; "
          f"frames {dec.num_frames} ({got} before finish)")
dec.close()
