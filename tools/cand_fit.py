"""Developer tool: regress the exact error of the kept candidates on the terms of their estimates (astc_cand_dump output)."""
import re, sys, collections
import numpy as np
path = sys.argv[1]
est = {}; rows = []
blk = None
for line in open(path):
    if line.startswith("=== block"):
        blk = line.split()[2] + ":" + line.split()[3]; continue
    m = re.match(r"EST blk \d+ slot (\d+) mode (\d+) base ([\d.eE+-]+) dec ([\d.eE+-]+) quant ([\d.eE+-]+) colour ([\d.eE+-]+)", line)
    if m:
        est[(blk, int(m.group(1)), int(m.group(2)))] = tuple(float(x) for x in m.groups()[2:]); continue
    m = re.match(r"CAND blk \d+ slot (\d+) mode (\d+) nw (\d+) level (\d+) cl (\d+) est ([\d.eE+-]+) exact ([\d.eE+-]+)", line)
    if m:
        s, mode, nw, lvl, cl = [int(x) for x in m.groups()[:5]]
        e = est.get((blk, s, mode))
        if e: rows.append((blk, s, mode, nw, lvl, cl, float(m.group(6)), float(m.group(7))) + e)
print(len(rows), "candidates with terms")
R = np.array([r[6:] for r in rows]); meta = [r[:6] for r in rows]
estv, exact, base, dec, quant, col = R.T
ok = exact < 1e30
def fit(mask, label):
    A = np.stack([base, dec, quant, col], 1)[mask]; y = exact[mask]
    # relative least squares (weights 1/exact)
    w = 1.0/np.maximum(y, 50.0)
    c, *_ = np.linalg.lstsq(A*w[:, None], y*w, rcond=None)
    pred = A@c
    print("%-28s n=%5d coef base %.2f dec %.2f quant %.2f colour %.2f | median |log2(pred/exact)| %.3f (current est: %.3f)" % (
        label, mask.sum(), c[0], c[1], c[2], c[3], np.median(np.abs(np.log2(np.maximum(pred, 1)/np.maximum(y, 1)))),
        np.median(np.abs(np.log2(np.maximum(estv[mask], 1)/np.maximum(y, 1))))))
fit(ok, "all")
lv = np.array([m[4] for m in meta]); nw = np.array([m[3] for m in meta]); sl = np.array([m[1] for m in meta])
for L in sorted(set(lv)):
    mk = ok & (lv == L)
    if mk.sum() > 30:
        ratio = np.median(exact[mk]/np.maximum(estv[mk], 1))
        print("   level %2d: n=%5d median exact/est %.2f ; share of quant term in est %.2f" % (L, mk.sum(), ratio, np.median(quant[mk]/np.maximum(estv[mk], 1))))
for s in sorted(set(sl)):
    mk = ok & (sl == s)
    if mk.sum() > 30:
        print("   slot %2d: n=%5d median exact/est %.2f" % (s, mk.sum(), np.median(exact[mk]/np.maximum(estv[mk], 1))))
T = max(nw)
for full in (True, False):
    mk = ok & ((nw == T) == full)
    print("   %s grids: n=%5d median exact/est %.2f" % ("full-res" if full else "decimated", mk.sum(), np.median(exact[mk]/np.maximum(estv[mk], 1))))
# how often is the exact winner among the first k by estimate?
byblk = collections.defaultdict(list)
for r in rows: byblk[r[0]].append(r)
for k in (1, 2, 4, 8, 16):
    loss = []
    for b, rs in byblk.items():
        rs = sorted(rs, key=lambda r: r[6])
        best_all = min(r[7] for r in rs); best_k = min(r[7] for r in rs[:k])
        loss.append(best_k/max(best_all, 1))
    print("   top-%2d by estimate: mean exact/best-of-kept %.3f" % (k, np.mean(loss)))
