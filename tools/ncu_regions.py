"""Per-region dynamic instruction counts, samples and stall reasons of k_fused from an ncu report
(needs -lineinfo and --import-source on).  Regions are the `//@region NAME` markers of the .cu file; every
SASS instruction is attributed through its source line as ncu reports it (inlined intrinsics land in their
header files and are listed separately).
usage: python tools/ncu_regions.py rep.ncu-rep outputs_per_launch [path/to/kernel_fused.cu]"""
import csv, io, subprocess, sys, collections, re
rep, n_out = sys.argv[1], float(sys.argv[2])
cu = sys.argv[3] if len(sys.argv) > 3 else "adsbdec_b200/csrc/kernel_fused.cu"
src = open(cu).read().split("\n")
regions = [(i + 1, re.search(r"//@region\s+(\S+)", l).group(1)) for i, l in enumerate(src) if "//@region" in l]
def region_of(line):
    nm = "head"
    for first, name in regions:
        if line >= first:
            nm = name
    return nm
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = hdr = None
agg = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if not (cur and hdr and len(r) == len(hdr)):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    ix = {h: i for i, h in enumerate(hdr)}
    if r[ix["Address"]] != "-":      # SASS rows repeat the counts of their source line
        continue
    key = region_of(ln) if cur.endswith(".cu") else "inlined:" + cur.split("/")[-1]
    f = lambda x: float(x) if re.fullmatch(r"[0-9.]+", x or "") else 0.0
    a = agg[key]
    a["inst"] += f(r[ix["Instructions Executed"]])
    a["samples"] += f(r[ix["# Samples"]])
    for h in hdr:
        if h.startswith("stall_"):
            a[h] += f(r[ix[h]])
tot_i = sum(a["inst"] for a in agg.values()); tot_s = sum(a["samples"] for a in agg.values())
print(f"total {tot_i * 32 / n_out:.2f} lane-instr/output, {tot_s:.0f} samples")
print("| region | lane-instr/out | % samples | samples per kilo-instr | top stalls |")
print("|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    st = sorted(((h[6:], v) for h, v in a.items() if h.startswith("stall_")), key=lambda kv: -kv[1])[:5]
    s = ", ".join(f"{h} {100 * v / max(1, a['samples']):.0f}%" for h, v in st)
    print(f"| {k} | {a['inst'] * 32 / n_out:.2f} | {100 * a['samples'] / tot_s:.1f} | {1000 * a['samples'] / max(1, a['inst']):.2f} | {s} |")
