"""Warp-instructions per 128-voxel tile and the opcode mix of every warp role of conv_tc_kernel, from an ncu source page.
usage: python tools/ncu_instruction_mix.py report.ncu-rep [tiles]"""
import csv,subprocess,sys,collections
rep=sys.argv[1]
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=rows[1]
ie=h.index('Instructions Executed'); isrc=h.index('Source'); ss=h.index('# Samples')
body=rows[2:]
bidx=[i for i,r in enumerate(body) if 'USETMAXREG' in r[isrc]]
roles=['prologue','loader','wgt+mma','epilogue']
def role(i): return roles[min(sum(1 for b in bidx if i>=b),3)]
ntile=float(sys.argv[2]) if len(sys.argv)>2 else 16384.0
for rl in roles[1:]:
    ops=collections.Counter(); tot=0
    for i,r in enumerate(body):
        if role(i)!=rl: continue
        e=float(r[ie] or 0)
        if e<=0: continue
        op=r[isrc].strip().split()
        op=[o for o in op if not o.startswith('@')][0].split('.')[0]
        ops[op]+=e; tot+=e
    print(rl,'warp-instr/tile %.0f'%(tot/ntile),' '.join('%s=%.0f'%(k,v/ntile) for k,v in ops.most_common(22)))
