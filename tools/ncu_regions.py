"""Region-level view of an ncu source page: where do a kernel's warp-stall samples sit?

    ncu -i REPORT.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > page.csv
    python tools/ncu_regions.py page.csv [instructions per region, default 100]

Prints the stall-reason totals and, for every run of N consecutive SASS instructions holding >= 1 % of the samples,
its share, hottest opcodes and top stall reasons (the page lists every inlined copy, so offsets repeat per copy)."""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
body=[r for r in rows[2:] if len(r)>10 and r[ix['# Samples']].isdigit()]
tot=sum(int(r[ix['# Samples']]) for r in body)
print("total samples",tot, "instrs", len(body))
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={h:sum(int(r[ix[h]]) for r in body) for h in stalls}
print([(k,round(100*v/tot,1)) for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:10]])
# region summary: split instruction list in chunks of N and print sample share with a representative opcode mix
N=int(sys.argv[2]) if len(sys.argv)>2 else 100
for c in range(0,len(body),N):
    ch=body[c:c+N]; s=sum(int(r[ix['# Samples']]) for r in ch)
    if s*100/tot<1.0: continue
    ex=max(int(r[ix['Instructions Executed']]) for r in ch)
    ops={}
    for r in ch:
        op=r[ix['Source']].split()[0] if r[ix['Source']].split() else ''
        if op.startswith('@'): op=r[ix['Source']].split()[1]
        ops[op.split('.')[0]]=ops.get(op.split('.')[0],0)+int(r[ix['# Samples']])
    st={h:sum(int(r[ix[h]]) for r in ch) for h in stalls}
    print(f"[{c:5d},{c+N:5d}) {100*s/tot:5.1f}%  maxexec {ex:8d}  ops {sorted(ops.items(), key=lambda kv:-kv[1])[:5]}  stalls {[(k.replace('stall_',''),v) for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:3]]}")
