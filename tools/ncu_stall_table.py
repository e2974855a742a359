"""Developer tool: after tools/ncu_stalls.py <report> <kernel pattern> (which joins ncu's per-instruction stall samples with
the cubin's line table into /tmp/astc_prof.pkl), print astc3.cu's source regions with their static SASS size, share of
executed warp instructions, share of stall samples and the samples by stall reason."""
import bisect, collections, os, pickle, re
cols, data = pickle.load(open('/tmp/astc_prof.pkl', 'rb'))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cuttlefish_b200", "csrc", "astc3.cu")).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*// ---- (setup \d+a?|phase \d[abc]?'?)", l)
    if m: marks.append((i, 'kernel: ' + m.group(1)))
    m = re.match(r"(?:template <[^>]*>\s*)?__device__ (?:__forceinline__|__noinline__) \S+ (\w+)\(", l)
    if m: marks.append((i, m.group(1)))
    if l.startswith('__global__'): marks.append((i, 'kernel: head, texel load, address arithmetic'))
    if 'PHASE_SYNC();       // (measured' in l: marks.append((i, 'kernel: pack'))
marks.sort(); starts = [m[0] for m in marks]


def name(f, l):
    if f == 'astc3.cu':
        j = bisect.bisect_right(starts, l) - 1
        return marks[j][1] if j >= 0 else 'astc3 top'
    return f


agg = collections.defaultdict(lambda: [0]*len(cols)); stat = collections.Counter()
for (f, l, _), sass, v in data:
    n = name(f, l); stat[n] += 1
    a = agg[n]
    for i, x in enumerate(v): a[i] += x
tot = [sum(a[i] for a in agg.values()) for i in range(len(cols))]
print("Per source region (own lines; helpers inlined into several phases are listed once): static SASS instructions, share of executed")
print("warp instructions, share of stall samples, and the samples split by stall reason (percent of all samples).")
print("%-44s %6s %6s %6s | %s" % ("where", "static", "inst%", "smp%", " ".join(c.replace('stall_', '')[:7].rjust(7) for c in cols[2:])))
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1]*250 < tot[1]: continue
    print("%-44s %6d %6.1f %6.1f | %s" % (n, stat[n], 100*v[0]/tot[0], 100*v[1]/tot[1], " ".join("%7.1f" % (100*x/tot[1]) for x in v[2:])))
print("%-44s %6d %6.1f %6.1f | %s" % ("total", sum(stat.values()), 100, 100, " ".join("%7.1f" % (100*x/tot[1]) for x in tot[2:])))
