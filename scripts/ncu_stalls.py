#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --csv` per-instruction samples: by opcode, by stall reason, hot loop vs rest.

    ncu -i gpurun_out/prof_sdf.ncu-rep --page source --csv > /tmp/src.csv && python scripts/ncu_stalls.py /tmp/src.csv
"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
ix={h:i for i,h in enumerate(hdr)}
data=[]
for r in rows[2:]:
    if len(r)!=len(hdr) or not r[ix['# Samples']].isdigit():
        if data: break
        continue
    data.append(r)
print(len(data), "sass rows (first launch)")
S=lambda r:int(r[ix['# Samples']]); E=lambda r:int(r[ix['Instructions Executed']])
tot=sum(S(r) for r in data); texe=sum(E(r) for r in data)
print("total samples", tot, "instr executed", texe)
byop=collections.Counter(); exe=collections.Counter()
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
stalls=collections.Counter()
def opc(r):
    t=r[ix['Source']].split()
    op=t[1] if t[0].startswith('@') else t[0]
    return op.split('.')[0].rstrip(';')
for r in data:
    op=opc(r); byop[op]+=S(r); exe[op]+=E(r)
    for c in stall_cols: stalls[c]+=int(r[ix[c]])
print("samples by opcode:")
for k,v in byop.most_common(24): print(f"  {k:12s} {v:8d} {100*v/tot:5.1f}%  exec {exe[k]:>12d} {100*exe[k]/texe:5.1f}%")
print("stall totals:")
for k,v in stalls.most_common(12): print(f"  {k:24s} {v:8d} {100*v/tot:5.1f}%")
idx=[i for i,r in enumerate(data) if 'FFMA.SAT' in r[ix['Source']]]
lo,hi=min(idx)-25,max(idx)+45
hot=sum(S(r) for r in data[lo:hi]); hexe=sum(E(r) for r in data[lo:hi])
print(f"hot-loop region rows {lo}-{hi}: samples {100*hot/tot:.1f}%  instr {100*hexe/texe:.1f}%")
hs=collections.Counter()
for r in data[lo:hi]:
    for c in stall_cols: hs[c]+=int(r[ix[c]])
print(" hot stalls:", [(k,round(100*v/hot,1)) for k,v in hs.most_common(8)])
rest=[r for i,r in enumerate(data) if not (lo<=i<hi)]
rs=collections.Counter()
for r in rest:
    for c in stall_cols: rs[c]+=int(r[ix[c]])
rt=sum(S(r) for r in rest)
print(" non-hot stalls:", [(k,round(100*v/rt,1)) for k,v in rs.most_common(8)])
# top non-hot instructions by samples
top=sorted(rest,key=S,reverse=True)[:25]
for r in top: print(f"   {S(r):7d} {E(r):>10d}  {r[ix['Source']].strip()[:90]}")
