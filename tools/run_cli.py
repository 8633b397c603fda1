"""Times the drop-in command line against the unmodified reference binary on one page-cached file.
usage: python tools/run_cli.py [seconds]"""
import subprocess, sys, time, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from bench import cpu_sample
from oracle import pyoracle
import adsbdec_b200 as A

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
f = Path(d) / "adsb_cli_cap.bin"
s = cpu_sample(1234, seconds)
s = np.concatenate([s, np.full(1 << 21, 2048, dtype="<u2")])
s.tofile(f)
n = s.size
outs = {}
for name, cmd in (("adsbdec_b200", [str(A.CLI_PATH), "-a", "-m", "-f", str(f)]),
                  ("reference", [str(pyoracle.ref_binary()), "-a", "-m", "-f", str(f)])):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True)
        best = min(best, time.perf_counter() - t0)
    outs[name] = r.stdout
    print(f"{name:14s} {best*1e3:9.1f} ms  {n/best/1e6:9.1f} Msamples/s  ({len(r.stdout.splitlines())} frames) {r.stderr.decode().splitlines()[1:3]}")
print("stdout identical:", outs["adsbdec_b200"] == outs["reference"],
      "(the reference may drop queued frames at EOF, SURVEY F9)" if outs["adsbdec_b200"] != outs["reference"] else "")
f.unlink()
