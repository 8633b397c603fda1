"""Times adsbdec_b200_decode_file on a page-cached capture (the reference's fileInput path, air.c:219-246).
usage: python tools/run_file.py [seconds] [reps]"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from bench import cpu_sample
import adsbdec_b200 as A

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
f = Path(d) / "adsb_file_cap.bin"
one = cpu_sample(1234, min(seconds, 10.0))
with open(f, "wb") as fh:
    left = int(seconds * 20e6)
    while left > 0:
        part = one[: min(left, one.size)]
        part.tofile(fh)
        left -= part.size
    np.full(1 << 21, 2048, dtype="<u2").tofile(fh)
n = f.stat().st_size // 2
print(f"{n} samples, {os.cpu_count()} host cpus")
ref = None
for thr in (1, 4, 8, 12, 16, 0):
    dec = A.Decoder(accept_df18=True, reader_threads=thr)
    best = 1e9
    for _ in range(reps):
        dec.reset()
        t0 = time.perf_counter()
        dec.decode_file(str(f))
        best = min(best, time.perf_counter() - t0)
    fr = dec.frames_array
    key = (len(fr), hash(fr.tobytes()))
    ref = ref or key
    print(f"reader threads {thr}: {best*1e3:8.1f} ms  {n/best/1e9:6.2f} G samples/s  frames {len(fr)} same={key == ref}  {dec.host_time()}")
    dec.close()
f.unlink()
