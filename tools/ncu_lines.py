"""Developer tool: per-source-line instruction counts of one kernel from an ncu report.
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line info of the cubin in libcfx.so.
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep bc7_kernelILi16 [top]
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, pattern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.path.join(root, "cuttlefish_b200", "lib", "libcfx.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
lines = []          # (file, line) per instruction of the kernel
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    cur, active = ("?", 0), False
    for ln in txt.splitlines():
        m = re.match(r"\s*//-+ \.text\.(\S+)", ln)
        if m:
            active = pattern in m.group(1)
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    if lines:
        break
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[hdr_i + 1:] if r and r[0].startswith("0x")]
if len(body) != len(lines):
    print("warning: %d SASS rows in report vs %d in cubin (rebuilt since capture?)" % (len(body), len(lines)))
agg = {}
tot_i = tot_s = 0
for r, key in zip(body, lines):
    a = agg.setdefault(key, [0, 0])
    a[0] += int(r[ci]); a[1] += int(r[cs])
    tot_i += int(r[ci]); tot_s += int(r[cs])
src_cache = {}


def src(f, l):
    if f not in src_cache:
        p = os.path.join(root, "cuttlefish_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    s = src_cache[f]
    return s[l - 1].strip()[:90] if 0 < l <= len(s) else ""


print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * n / tot_i, 100.0 * s / max(tot_s, 1), key[0], key[1], src(*key)))

# ---- the same, aggregated by enclosing function (top-level CFX_HD / __global__ definitions)
import bisect
funcs = {}
for f in {k[0] for k in agg}:
    p = os.path.join(root, "cuttlefish_b200", "csrc", f)
    if not os.path.exists(p):
        continue
    starts = []
    for i, l in enumerate(open(p).read().splitlines(), 1):
        m = re.match(r"(?:CFX_HD(?:_NOINLINE)?|__device__ __forceinline__|__global__|template|static|inline)\b.*?(\w+)\s*\(", l)
        if m and not l.startswith(" "):
            starts.append((i, m.group(1)))
    funcs[f] = starts
byf = {}
for (f, l), (n, s) in agg.items():
    name = "?"
    if f in funcs and funcs[f]:
        idx = bisect.bisect_right([x[0] for x in funcs[f]], l) - 1
        if idx >= 0:
            name = funcs[f][idx][1]
    a = byf.setdefault("%s:%s" % (f, name), [0, 0])
    a[0] += n; a[1] += s
print("---- by function")
for k, (n, s) in sorted(byf.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%5.1f%% inst %5.1f%% smp  %s" % (100.0 * n / tot_i, 100.0 * s / max(tot_s, 1), k))
