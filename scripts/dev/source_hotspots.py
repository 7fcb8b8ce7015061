#!/usr/bin/env python
"""Dev tool: attribute the executed instructions and stall samples of one `ncu --set full` capture of cbfnmpc_lane_kernel to
source lines and functions of csrc/cbfnmpc_lane.h.

  python scripts/dev/source_hotspots.py gpurun_out/prof_<tag>.ncu-rep [srb-cbf-nmpc_b200/lib/libcbfnmpc.so [kernel [header]]]

The report's SASS page has no line numbers; they come from `nvdisasm -g` of the cubin inside the library, joined by instruction
offset.  The library must be the build that was profiled (the static instruction counts are compared)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rep = sys.argv[1]
lib = os.path.abspath(sys.argv[2]) if len(sys.argv) > 2 else os.path.join(ROOT, "srb-cbf-nmpc_b200", "lib", "libcbfnmpc.so")
KERNEL = sys.argv[3] if len(sys.argv) > 3 else "cbfnmpc_lane_kernel"
HEADER = sys.argv[4] if len(sys.argv) > 4 else "cbfnmpc_lane.h"

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines = []   # per instruction: innermost (file, line)
chains = []  # per instruction: [(file, line) innermost ... outermost]
inside = False; cur = [("?", 0)]; fresh = False
for l in sass.splitlines():
    if l.startswith("//---------------------"):
        inside = KERNEL in l
        continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if not fresh: cur = [(os.path.basename(m.group(1)), int(m.group(2)))]; fresh = True
        if m.group(3): cur.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l): lines.append(cur[0]); chains.append(cur); fresh = False

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; rows = [r for r in rows[2:] if len(r) > 10]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_")}
print("static instructions: report %d, library %d" % (len(rows), len(lines)))
if len(rows) != len(lines):
    print("WARNING: the library is not the profiled build; attribution would be wrong"); sys.exit(1)

# function ranges of cbfnmpc_lane.h: a definition starts at a line that begins (column 0 or 2) with LDEV / template / struct
fn_at = []
hpath = os.path.join(ROOT, "srb-cbf-nmpc_b200", "csrc", HEADER)
for n, l in enumerate(open(hpath), 1):
    m = re.match(r"\s{0,2}(?:template\s*<[^>]*>\s*)?(?:LDEV|LMEM|XDEV|XMEM|__device__(?:\s+__forceinline__)?)\s+[\w:<>\s\*&]*?(\w+)\s*\(", l)
    if m: fn_at.append((n, m.group(1)))
def fn_of(f, n):
    if f != HEADER: return f
    name = "(top)"
    for a, b in fn_at:
        if a <= n: name = b
        else: break
    return name

ex_line, sm_line = Counter(), Counter(); ex_fn, sm_fn = Counter(), Counter(); st_fn = defaultdict(Counter)
for (f, n), r in zip(lines, rows):
    e, s = int(r[ia]), int(r[isamp])
    ex_line[(f, n)] += e; sm_line[(f, n)] += s
    fn = fn_of(f, n); ex_fn[fn] += e; sm_fn[fn] += s
    for h, i in stall_cols.items():
        if r[i] not in ("", "0"): st_fn[fn][h] += int(r[i])
te, ts = sum(ex_fn.values()), sum(sm_fn.values())
print("\n| function | executed | share | samples share | top stall reasons |\n|---|---|---|---|---|")
for fn, s in sm_fn.most_common(25):
    top = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(s, 1)) for k, v in st_fn[fn].most_common(3))
    print("| %s | %d | %.1f %% | %.1f %% | %s |" % (fn, ex_fn[fn], 100 * ex_fn[fn] / te, 100 * s / ts, top))
src_lines = open(hpath).read().splitlines()
print("\n| line | samples share | executed share | source |\n|---|---|---|---|")
for (f, n), s in sm_line.most_common(40):
    text = src_lines[n - 1].strip()[:110] if f == HEADER and n <= len(src_lines) else ""
    print("| %s:%d | %.2f %% | %.2f %% | `%s` |" % (f, n, 100 * s / ts, 100 * ex_line[(f, n)] / te, text))

# inclusive view: every function on the inline chain of an instruction gets its samples; and the call sites inside the two
# iteration bodies (the line of qp_iter / nlp_iter on the chain)
inc_fn, site = Counter(), Counter(); site_ex = Counter()
for ch, r in zip(chains, rows):
    e, s = int(r[ia]), int(r[isamp])
    seen = set()
    for f, n in ch:
        fn = fn_of(f, n)
        if fn not in seen: inc_fn[fn] += s; seen.add(fn)
        if fn in ("qp_iter", "nlp_iter", "nlp_init", "qp_init", "lane_run_group", "qdense_solve_warp", "qd_factor", "qd_solve") and f == HEADER: site[(fn, n)] += s; site_ex[(fn, n)] += e
print("\n| function (inclusive) | samples share |\n|---|---|")
for fn, s in inc_fn.most_common(30): print("| %s | %.1f %% |" % (fn, 100 * s / ts))
print("\n| call site in the iteration bodies (inclusive) | samples share | executed share | source |\n|---|---|---|---|")
for (fn, n), s in site.most_common(45):
    print("| %s:%d | %.2f %% | %.2f %% | `%s` |" % (fn, n, 100 * s / ts, 100 * site_ex[(fn, n)] / te, src_lines[n - 1].strip()[:120]))
