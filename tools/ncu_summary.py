"""Summarise an ncu report of k_fused: headline metrics, per-stage instruction counts (split at
barriers) and stall reasons.  usage: python tools/ncu_summary.py rep.ncu-rep outputs_per_launch"""
import csv, io, re, subprocess, sys, collections

rep, n_out = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg']
for k in keys:
    if k in d:
        print(f"- `{k}` = {d[k][0]} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
num = lambda x: float(x) if x.replace('.', '', 1).isdigit() else 0.0
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
seg = 0
segs = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter(), collections.Counter()])
for r in data:
    s = r[ix['Source']]
    n, sm = num(r[ix['Instructions Executed']]), num(r[ix['# Samples']])
    m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', s)
    op = m.group(2) if m else '?'
    segs[seg][0] += n; segs[seg][1] += sm; segs[seg][2][op] += n
    for h in stalls:
        segs[seg][3][h] += num(r[ix[h]])
    if 'BAR.SYNC' in s:
        seg += 1
tot = sum(v[0] for v in segs.values()); ts = sum(v[1] for v in segs.values())
print(f"\ntotal warp instructions {tot:.0f} = {tot * 32 / n_out:.1f} lane-instructions per envelope sample\n")
print("| segment (between barriers) | lane-instr/output | % instr | % samples | top opcodes | top stalls |")
print("|---|---|---|---|---|---|")
for s_, v in segs.items():
    if v[0] / tot < 0.003 and v[1] / ts < 0.01:
        continue
    ops = ", ".join(f"{o} {n * 32 / n_out:.2f}" for o, n in v[2].most_common(7))
    st = ", ".join(f"{o[6:]} {100 * n / max(1, sum(v[3].values())):.0f}%" for o, n in v[3].most_common(4))
    print(f"| {s_} | {v[0] * 32 / n_out:.2f} | {100 * v[0] / tot:.1f} | {100 * v[1] / ts:.1f} | {ops} | {st} |")
all_st = collections.Counter()
for v in segs.values():
    all_st.update(v[3])
st = sum(all_st.values())
print("\nstall samples: " + ", ".join(f"{k[6:]} {100 * n / st:.1f}%" for k, n in all_st.most_common(9)))
