import csv,io,re,subprocess,sys,collections
rep=sys.argv[1]; pat=sys.argv[2]
cub="/tmp/cub/astc3.sm_100a.cubin"
txt=subprocess.run(["nvdisasm","-g","-c",cub],capture_output=True,text=True).stdout
cur=("?",0);active=False;lines=[];inl=[]
for ln in txt.splitlines():
    m=re.match(r"\s*//-+ \.text\.(\S+)",ln)
    if m: active=pat in m.group(1); continue
    if not active: continue
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)',ln)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2))); 
        # inlined at info
        m2=re.search(r'inlined at "([^"]+)", line (\d+)',m.group(3))
        cur_in = (m2.group(1).split('/')[-1],int(m2.group(2))) if m2 else None
        cur=(cur[0],cur[1],cur_in)
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/",ln): lines.append(cur)
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
hdr=rows[hi]; body=[r for r in rows[hi+1:] if r and r[0].startswith("0x")]
print(len(body),len(lines))
cols=['Instructions Executed','# Samples','stall_long_sb','stall_wait','stall_no_inst','stall_barrier','stall_short_sb','stall_math','stall_branch_resolving','stall_mio','stall_not_selected','stall_selected','stall_dispatch']
idx=[hdr.index(c) for c in cols]
import pickle
data=[]
for r,k in zip(body,lines):
    data.append((k,r[1],[int(r[i] or 0) for i in idx]))
pickle.dump((cols,data),open('/tmp/astc_prof.pkl','wb'))
