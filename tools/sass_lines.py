#!/usr/bin/env python3
"""Per-source-line instruction counts (and, with an ncu source-page CSV of the same build, stall samples / executed
instructions) of one kernel of libfxenv.so.  Lines come from `nvdisasm -g` (the library is built with -lineinfo; inlined
code is attributed to the innermost line); the ncu CSV is matched by instruction index.
usage: sass_lines.py <kernel-name-substring> [ncu_source.csv] [top=40]"""
import csv, os, re, subprocess, sys, tempfile, collections

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
key = sys.argv[1]
ncu_csv = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "gym_fx_b200", "libfxenv.so")], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = os.path.join(tmp, "fx_kernels.sm_100a.cubin")
sass = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
instrs = []   # (file, line, text)
infn = False; cur = ("?", 0)
for ln in sass:
    if ln.startswith(".text."):
        infn = key in ln
        continue
    if not infn:
        continue
    if ln.startswith(".section") or ln.startswith("\t.section"):
        infn = False; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?);', ln)
    if m:
        instrs.append((cur[0], cur[1], m.group(1)))
print(f"{len(instrs)} instructions in kernels matching {key!r}")
samples = executed = None
if ncu_csv:
    rows = list(csv.reader(open(ncu_csv)))
    h = rows[1]; data = rows[2:]
    isamp = h.index("Warp Stall Sampling (All Samples)"); iex = h.index("Instructions Executed")
    if len(data) != len(instrs):
        print(f"WARNING: ncu has {len(data)} instructions, the library {len(instrs)}: different builds?")
    n = min(len(data), len(instrs))
    samples = [int(data[i][isamp]) for i in range(n)]; executed = [int(data[i][iex]) for i in range(n)]
agg = collections.defaultdict(lambda: [0, 0, 0])
for i, (f, l, t) in enumerate(instrs):
    a = agg[(f, l)]
    a[0] += 1
    if samples and i < len(samples):
        a[1] += samples[i]; a[2] += executed[i]
src = {}
def line_text(f, l):
    if f not in src:
        p = os.path.join(root, "gym_fx_b200", "csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return src[f][l - 1].strip()[:110] if 0 < l <= len(src[f]) else ""
if samples:
    ts, te = sum(samples), sum(executed)
    print(f"total samples {ts}, executed warp-instructions {te}")
    print("  samples%  exec%  #sass  file:line  source")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {100*a[1]/ts:6.2f}  {100*a[2]/te:6.2f}  {a[0]:5d}  {f}:{l}  {line_text(f, l)}")
else:
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {a[0]:5d}  {f}:{l}  {line_text(f, l)}")
