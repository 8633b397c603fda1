#!/usr/bin/env python
"""Static SASS instruction counts of a kernel, bucketed by the source regions of its .cu file.

  python tools/sass_stats.py adsbdec_b200/csrc/build/kernel_fused.o 'k_fusedILb0ELb0' [--ops] [--lines]

Regions are declared in the source with marker comments `//@region NAME` (a region runs to the next
marker).  Every SASS instruction is attributed to the OUTERMOST call site in that file (nvdisasm -gi),
so inlined intrinsics count where they are used.  The unrolled stages of k_fused execute once per
sub-block, so `static count x 32 lanes / outputs per sub-block` is their lane-instructions per output.
Needs no GPU (cuobjdump + nvdisasm on the cross-compiled object).
"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def disassemble(obj: Path) -> str:
    if obj.suffix == ".cubin":
        return subprocess.run(["nvdisasm", "-gi", "-c", str(obj)], capture_output=True, text=True).stdout
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(obj.resolve())], cwd=td, capture_output=True)
        cubins = sorted(Path(td).glob("*.cubin"))
        if not cubins:
            raise SystemExit(f"no cubin in {obj}")
        return subprocess.run(["nvdisasm", "-gi", "-c", str(cubins[0])], capture_output=True, text=True).stdout


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    flags = {a for a in sys.argv[1:] if a.startswith("--")}
    obj, func = Path(args[0]), args[1]
    outputs = int(args[2]) if len(args) > 2 else 896
    text = disassemble(obj)
    src_name = None
    regions = []  # (first line, name)
    cur_line, in_func = None, False
    per_line = collections.Counter()
    per_line_ops = collections.defaultdict(collections.Counter)
    ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
    # Cold code: the divergent-warp fall-backs of collective instructions (targets of BRA.DIV) sit behind
    # the kernel's last EXIT; whole stages are duplicated there.  They are not counted.
    cold_from = None
    if "--all" not in flags:
        labels, pending, div_targets, infn = {}, [], set(), False
        for ln in text.splitlines():
            if ln.startswith("\t.section\t.text."):
                infn = func in ln
                continue
            if not infn:
                continue
            m = re.match(r"^(\.L_x_\d+):", ln)
            if m:
                pending.append(m.group(1))
                continue
            m = ins_re.match(ln)
            if m:
                for lb in pending:
                    labels[lb] = int(m.group(1), 16)
                pending = []
                if "BRA.DIV" in m.group(2):
                    t = re.search(r"`\((\.L_x_\d+)\)", m.group(2))
                    if t:
                        div_targets.add(t.group(1))
        addrs = [labels[t] for t in div_targets if t in labels]
        cold_from = min(addrs) if addrs else None
    file_re = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    for ln in text.splitlines():
        if ln.startswith("\t.section\t.text."):
            in_func = func in ln
            cur_line = None
            continue
        if not in_func:
            continue
        m = file_re.search(ln)
        if m:
            f, l = m.group(1), int(m.group(2))
            if f.endswith(".cu"):
                src_name = f
                cur_line = l
            continue
        m = ins_re.match(ln)
        if m and cur_line is not None:
            if cold_from is not None and int(m.group(1), 16) >= cold_from:
                continue
            toks = m.group(2).split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            op = op.split(".")[0]
            per_line[cur_line] += 1
            per_line_ops[cur_line][op] += 1
    if src_name is None:
        raise SystemExit("function not found / no line info")
    src = Path(src_name).read_text().splitlines()
    for i, s in enumerate(src, 1):
        m = re.search(r"//@region\s+(\S+)", s)
        if m:
            regions.append((i, m.group(1)))
    if not regions:
        regions = [(1, "all")]
    if regions[0][0] > 1:
        regions.insert(0, (1, "head"))

    def region_of(line):
        name = regions[0][1]
        for first, nm in regions:
            if line >= first:
                name = nm
        return name

    tot = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    for l, c in per_line.items():
        r = region_of(l)
        tot[r] += c
        ops[r].update(per_line_ops[l])
    total = sum(tot.values())
    print(f"{func}: {total} SASS instructions; per-output figures assume one execution per {outputs} outputs")
    seen = []
    for _, nm in regions:
        if nm in seen:
            continue
        seen.append(nm)
        c = tot.get(nm, 0)
        top = ", ".join(f"{o} {n}" for o, n in ops[nm].most_common(9))
        print(f"  {nm:14s} {c:5d}  {c * 32 / outputs:6.2f}/out   {top}")
    if "--lines" in flags:
        for l, c in sorted(per_line.items(), key=lambda kv: -kv[1])[:40]:
            print(f"    line {l:4d} {c:4d}  {src[l - 1].strip()[:100]}")
    for fl in flags:
        if fl.startswith("--region="):
            want = fl.split("=", 1)[1]
            for l, c in sorted(per_line.items()):
                if region_of(l) == want:
                    top = ", ".join(f"{o} {n}" for o, n in per_line_ops[l].most_common(6))
                    print(f"    line {l:4d} {c:4d}  {src[l - 1].strip()[:70]:70s} {top}")
    if "--ops" in flags:
        allops = collections.Counter()
        for r in ops.values():
            allops.update(r)
        print("  ops:", ", ".join(f"{o} {n}" for o, n in allops.most_common(30)))


if __name__ == "__main__":
    main()
