import re, sys
blk=None; rows={}
for line in open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/cand.txt'):
    if line.startswith('=== block'): blk=int(line.split()[2]); rows[blk]=[]
    m=re.match(r'CAND blk \d+ slot (\d+) nw (\d+) level (\d+) cl (\d+) est ([\d.]+) exact ([\d.]+)',line)
    if m: rows[blk].append(tuple(float(x) for x in m.groups()))
for b,r in rows.items():
    print("block",b,len(r),"candidates; by estimate order (slot nw level cl est exact):")
    for i,x in enumerate(r[:14]): print("   %2d slot %2d nw %2d L %2d cl %2d est %8.1f exact %8.1f"%((i,)+tuple(int(v) if k<4 else v for k,v in enumerate(x))))
    best=sorted(r,key=lambda x:x[5])[:6]
    print("  best by exact:")
    for x in best: print("      slot %2d nw %2d L %2d cl %2d est %8.1f exact %8.1f  (rank by est %d)"%(tuple(int(v) if k<4 else v for k,v in enumerate(x))+(r.index(x),)))
