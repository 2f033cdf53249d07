"""Attribute ncu per-instruction samples / executed counts to source lines.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep 'env_kernel<(int)16, (int)1>' [top]

ncu's CSV source page has no per-line metrics when the sources are not resolvable, so this joins
`ncu --page source --print-source sass --csv` with `nvdisasm -g` line markers of the same cubin
(both list the function's instructions in address order)."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "mapdn_b200", "libmapdn_b200.so")
m = re.search(r"<\(int\)(\d+), \(int\)(\d+)(?:, \(bool\)(\d))?>", kname)
mangled = f"_ZN5mapdn10env_kernelILi{m.group(1)}ELi{m.group(2)}ELb{m.group(3) or 0}EEEvNS_6ParamsE"
with tempfile.TemporaryDirectory() as d:
    subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
lines, cur, on = [], None, False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        on = ln.startswith(".text." + mangled + ":")
        continue
    if not on:
        continue
    mm = re.search(r'//## File "(.*?)", line (\d+)(.*)', ln)
    if mm:
        fn = os.path.basename(mm.group(1))
        cur = int(mm.group(2)) if fn == "env_kernel.cuh" else -(int(mm.group(2)) + (100000 if "philox" in fn else 200000))
        continue
    mm = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if mm:
        lines.append((int(mm.group(1), 16), cur, mm.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, i = [], 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2; data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) >= len(hdr): data.append(rows[j])
            j += 1
        blocks.append((name, hdr, data)); i = j
    else:
        i += 1
sel = [b for b in blocks if kname in b[0]]
if not sel:
    print("kernels in report:", sorted(set(b[0] for b in blocks))); sys.exit(1)
name, hdr, data = sel[0]
ci = {n: k for k, n in enumerate(hdr)}
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_i = 0
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
for (off, line, txt), r in zip(lines, data):
    s = int(r[ci["# Samples"]] or 0); ins = int(r[ci["Instructions Executed"]] or 0)
    a = agg[line]; a[0] += s; a[1] += ins
    for c in stall_cols:
        v = int(r[ci[c]] or 0)
        if v: a[2][c[6:]] += v
    tot_s += s; tot_i += ins
src = open(os.path.join(ROOT, "mapdn_b200", "csrc", "env_kernel.cuh")).read().splitlines()
print(f"{name}: {tot_s} samples, {tot_i} warp-instructions")
for line, (s, ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if line is not None and line < 0:
        txt = ("philox.cuh:" if -line < 200000 else "other:") + str((-line) % 100000)
    else:
        txt = src[line - 1].strip()[:70] if line and line <= len(src) else "?"
    tops = ",".join(f"{k}:{v}" for k, v in st.most_common(3))
    print(f"{(line or 0):7d} {s:5d} {100*s/tot_s:5.1f}%  inst {ins:9d} {100*ins/tot_i:5.1f}%  [{tops}]  {txt}")
