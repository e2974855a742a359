"""Regenerates cuttlefish_b200/csrc/bc1_tables.cuh: optimal single-colour BC1 end points, by brute
force over all pairs under the ideal interpolation (2*c0 + c1)/3 the decoder uses."""
def e5(v): return (v << 3) | (v >> 2)
def e6(v): return (v << 2) | (v >> 4)
def table(n, ex):
    out = []
    for v in range(256):
        best = min((abs((2*ex(c0) + ex(c1))//3 - v), abs(ex(c0) - ex(c1)), c0, c1) for c0 in range(n) for c1 in range(n))
        out.append(best[2] | (best[3] << 8))
    return out
if __name__ == "__main__":
    print(table(32, e5)[:8], table(64, e6)[:8])
