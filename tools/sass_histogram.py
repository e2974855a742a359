"""Developer tool: per-kernel SASS opcode histogram of the shipped libcfx.so (cuobjdump -sass), for profiles/.
    python tools/sass_histogram.py > profiles/r02_sass_opcode_histogram.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cuttlefish_b200", "lib", "libcfx.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, {}
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
demangle = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
WATCH = ["HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "LDGSTS", "REDUX", "CREDUX", "LDL", "STL", "ATOMS", "VABSDIFF4", "IDP"]
for (k, h), name in sorted(zip(hist.items(), demangle), key=lambda kv: -sum(kv[0][1].values())):
    n = sum(h.values())
    if n < 50:
        continue
    print("%s\n  %d instructions; %s" % (name[:150], n, ", ".join("%s %d" % (o, c) for o, c in h.most_common(12))))
    print("  watched: " + ", ".join("%s %d" % (o, h[o]) for o in WATCH if h[o]))
