import csv,sys,subprocess
rep=sys.argv[1]; items=float(sys.argv[2]); thr=float(sys.argv[3]) if len(sys.argv)>3 else 1.5
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=rows[1]; isrc=h.index('Source'); ie=h.index('Instructions Executed'); iss=h.index('# Samples')
R=rows[2:]
print('total/item', sum(int(r[ie]) for r in R)/items, 'samples', sum(int(r[iss]) for r in R))
i=0
while i<len(R):
    e=int(R[i][ie]); j=i; s=0; samp=0; ops={}
    while j<len(R) and abs(int(R[j][ie])-e)<=0.12*max(e,1)+1000:
        s+=int(R[j][ie]); samp+=int(R[j][iss]); t=R[j][isrc].split(); op=t[0] if not t[0].startswith('@') else t[1]
        ops[op]=ops.get(op,0)+1; j+=1
    top=sorted(ops.items(),key=lambda x:-x[1])[:7]
    if s/items>thr: print(f"{i:4d}-{j-1:4d} n={j-i:3d} x{e/items:6.2f} instr/item={s/items:7.1f} samples={samp:6d} {top}")
    i=j
