"""Turns the scratch outputs of tools/gpu_profile.sh <tag> (gpurun_out/) into the tracked evidence under profiles/:
bench lines, launch list + per-kernel shares, ncu headline metrics of the three hot-path kernels, DRAM traffic
of k_fused (read by bench.py for roofline.traffic), and the SASS evidence of the shipped library.
usage: python tools/make_profiles.py <tag> [<out-prefix, default r2>]"""
import collections, csv, io, json, re, shutil, subprocess, sys, tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else "r2"
G, P = ROOT / "gpurun_out", ROOT / "profiles"

for src, dst in [(f"{tag}_bench.json", f"{out}_bench.json"), (f"{tag}_bench_reference_arm.json", f"{out}_bench_reference_arm.json"),
                 (f"{tag}_launches.csv", f"{out}_launches.csv"), (f"{tag}_tests.log", f"{out}_gpu_tests.log")]:
    if (G / src).exists():
        shutil.copy(G / src, P / dst)

# ---- launch list: share of every kernel in the step (ncu serialises and runs cold: shares, not absolutes) ----
rows = list(csv.reader(open(G / f"{tag}_launches.csv")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
ix = {h: i for i, h in enumerate(rows[hi])}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(rows[hi]) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("adsb::", "").replace("<unnamed>::", "").replace("void ", "")
    tot[name] += float(r[ix["Metric Value"]]) / 1e3
    cnt[name] += 1
allt = sum(tot.values())
lines = [f"# Launch list of `bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline` under ncu ({out}_launches.csv)", "",
         "Serialised, cold-cache per-launch times: the SHARES are what compares with the bench's CUDA-event figures.", "",
         "| kernel | launches | total us | share | us per launch |", "|---|---|---|---|---|"]
for k, v in tot.most_common():
    lines.append(f"| `{k}` | {cnt[k]} | {v:.1f} | {100 * v / allt:.1f} % | {v / cnt[k]:.1f} |")
(P / f"{out}_launches_summary.md").write_text("\n".join(lines) + "\n")

# ---- ncu --set full of the three hot-path kernels ----
rep = G / f"{tag}_kernels.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
txt = [f"# ncu --set full --clock-control none, one launch each of the hot-path kernels ({tag}_kernels.ncu-rep,", 
       "# `python tools/run_resident.py 60 4`: 60 s capture, 400 M samples = 801 MB per launch)", ""]
traffic = None
for r in rr[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "?")
    txt.append(f"## {name}")
    for k in keys:
        if k in d:
            txt.append(f"- `{k}` = {d[k]} {units[hdr.index(k)]}")
    if "k_fused" in name:
        u = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
        rd = float(d["dram__bytes_read.sum"]) * u[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(d["dram__bytes_write.sum"]) * u[units[hdr.index("dram__bytes_write.sum")]]
        traffic = (rd, wr, float(d["smsp__inst_executed.sum"]))
    txt.append("")
(P / f"{out}_kernels_ncu.md").write_text("\n".join(txt))
if traffic:
    bench = json.loads((G / f"{tag}_bench.json").read_text())
    alg = bench["roofline"]["algorithmic_bytes_per_launch"]
    (P / f"{out}_traffic.json").write_text(json.dumps({
        "kernel": "k_fused", "source": f"ncu --set full --clock-control none -k regex:k_fused|k_gate|k_slice -s 9 -c 3 python tools/run_resident.py 60 4 ({tag}_kernels.ncu-rep; summary in profiles/{out}_kernels_ncu.md)",
        "algorithmic_bytes": alg, "dram_bytes_read": traffic[0], "dram_bytes_write": traffic[1],
        "warp_instructions": traffic[2], "lane_instructions_per_envelope_sample": traffic[2] * 32 / (alg / 4)}, indent=1) + "\n")

# ---- SASS evidence of the shipped library ----
so = ROOT / "adsbdec_b200" / "libadsbdec_b200.so"
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(so)], cwd=td, capture_output=True)
    sass = []
    for cb in sorted(Path(td).glob("*.cubin")):
        sass.append(subprocess.run(["nvdisasm", "-c", str(cb)], capture_output=True, text=True).stdout)
sass = "\n".join(sass)
res = [f"# SASS of adsbdec_b200/libadsbdec_b200.so (cuobjdump -xelf all + nvdisasm -c), opcode counts per kernel", ""]
for m in re.finditer(r"\.text\.(\S+):\n(.*?)(?=\n//-{5,}|\Z)", sass, re.S):
    name, body = m.group(1), m.group(2)
    if not re.search(r"k_fusedILb0|k_gate|k_sliceILb0", name):
        continue
    ops = collections.Counter(re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", body))
    res.append(f"## {name}: {sum(ops.values())} instructions")
    res.append(", ".join(f"{o} {n}" for o, n in ops.most_common(28)))
    for pat in ["UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2"]:
        ex = [l.strip() for l in body.split("\n") if re.search(r"\s" + pat + r"[\s.]", l)][:3]
        if ex:
            res.append(f"  first {pat}: " + " | ".join(re.sub(r"\s+", " ", e) for e in ex))
    res.append("")
(P / f"{out}_sass_hot_kernels.md").write_text("\n".join(res))
print("\n".join(txt[:40]))
