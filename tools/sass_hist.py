"""Opcode histogram (executed warp instructions + stall samples) of an `ncu --page source --csv` export."""
import csv, collections, re, sys
rows=list(csv.reader(open(sys.argv[1])))
# find header row
h=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[h]; idx={k:i for i,k in enumerate(hdr)}
ops=collections.Counter(); samples=collections.Counter(); tot=0; stot=0
for r in rows[h+1:]:
    if len(r)<len(hdr) or r[0]=='Address': continue
    sass=r[idx['Source']].strip()
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)',sass)
    op=m.group(2).split('.')[0] if m else sass[:10]
    try: n=int(r[idx['Instructions Executed']] or 0); s=int(r[idx['# Samples']] or 0)
    except ValueError: continue
    ops[op]+=n; tot+=n; samples[op]+=s; stot+=s
print('total warp instr',tot,'samples',stot)
for op,n in ops.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 25):
    print(f"{op:10s} {n:12d} {100*n/tot:5.1f}%  samples {samples[op]:6d} {100*samples[op]/max(stot,1):5.1f}%")
