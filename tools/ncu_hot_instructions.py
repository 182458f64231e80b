"""Rank the SASS instructions of one warp role of a tcgen05 kernel by warp-stall samples (ncu --set full --import-source on
report).  Roles are split at the setmaxnreg instructions.  usage: python tools/ncu_hot_instructions.py report.ncu-rep [loader|wgt+mma|epilogue]"""
import csv,subprocess,sys
rep=sys.argv[1]; role_sel=sys.argv[2] if len(sys.argv)>2 else None
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=rows[1]
ia=h.index('Address'); ie=h.index('Instructions Executed'); ss=h.index('# Samples'); isrc=h.index('Source')
stallcols=[i for i,c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
body=rows[2:]
bidx=[i for i,r in enumerate(body) if 'USETMAXREG' in r[isrc]]
print('rows',len(body),'bounds',bidx)
roles=['prologue','loader','wgt+mma','epilogue']
def role(i): return roles[min(sum(1 for b in bidx if i>=b),3)]
agg={}
for i,r in enumerate(body):
    d=agg.setdefault(role(i),[0,0,{}])
    d[0]+=float(r[ie] or 0); d[1]+=float(r[ss] or 0)
    for c in stallcols:
        v=float(r[c] or 0)
        if v: d[2][h[c]]=d[2].get(h[c],0)+v
for k,d in agg.items():
    print(k,'inst %.0f samples %.0f'%(d[0],d[1]),' '.join('%s=%d'%(a.replace('stall_',''),b) for a,b in sorted(d[2].items(),key=lambda kv:-kv[1])[:8]))
if role_sel:
    sel=[(float(r[ss] or 0),i,r) for i,r in enumerate(body) if role(i)==role_sel]
    sel.sort(key=lambda t:-t[0])
    for s,i,r in sel[:40]:
        st=sorted(((h[c].replace('stall_',''),float(r[c] or 0)) for c in stallcols),key=lambda kv:-kv[1])[:3]
        print('%6d %5.0f ex %9s  %-60s %s'%(i,s,r[ie],r[isrc].strip()[:60],' '.join('%s=%d'%kv for kv in st if kv[1])))
