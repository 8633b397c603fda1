"""Per-stage / per-source-line instruction counts of k_fused from an ncu report (needs -lineinfo
and --import-source on).  usage: python tools/ncu_lines.py rep.ncu-rep outputs_per_launch"""
import csv, io, subprocess, sys, collections
rep, n_out = sys.argv[1], float(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
files, cur, hdr = collections.OrderedDict(), None, None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1]; files[cur] = collections.OrderedDict(); continue
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}; continue
    if cur and hdr and len(r) > 8 and r[2] == "-":
        try:
            files[cur][int(r[0])] = (float(r[7]), float(r[6]), r[1])
        except ValueError:
            pass
tot = sum(v[0] for f in files.values() for v in f.values())
print(f"total {tot * 32 / n_out:.2f} lane-instructions per envelope sample")
for f, lines in files.items():
    s = sum(v[0] for v in lines.values())
    print(f"\n{f}: {s * 32 / n_out:.2f}")
    if f.endswith("kernel_fused.cu"):
        src = open(f.replace("/root/repo", ".") if not __import__("os").path.exists(f) else f).read().split("\n")
        def find(t):
            return next(i + 1 for i, l in enumerate(src) if t in l)
        marks = [("helpers+setup", 1), ("span fetch", find("fetch a span")), ("stage 1", find("stage 1: envelope")),
                 ("stage 2", find("stage 2: decision bits")), ("stage 3", find("stage 3: preamble AND")),
                 ("stage 4", find("stage 4: slice + CRC")), ("tail", find("q_fir += 32"))]
        marks.sort(key=lambda m: m[1])   # stage 3 sits in front of stage 2 in the source
        for (name, lo), (_, hi) in zip(marks, marks[1:] + [("", 10 ** 9)]):
            a = sum(v[0] for l, v in lines.items() if lo <= l < hi)
            b = sum(v[1] for l, v in lines.items() if lo <= l < hi)
            print(f"  {name:14s} {a * 32 / n_out:6.2f} instr/out   {b:7.0f} samples")
    for l, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"    line {l:4d} {v[0] * 32 / n_out:5.2f}  {v[2].strip()[:100]}")
