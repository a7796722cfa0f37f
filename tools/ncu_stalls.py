#!/usr/bin/env python
"""Per-instruction stall summary from `ncu --page source --csv` output (file given as argv[1])."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=[r for r in rows[2:] if len(r)>10 and r[0].startswith('0x')]
ia=hdr.index('Address'); isrc=hdr.index('Source'); ismp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
base=int(data[0][ia],16)
tot=sum(int(r[ismp]) for r in data)
print('total samples',tot,'instrs',len(data))
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
mm=[k for k,r in enumerate(data) if 'UTCHMMA' in r[isrc]]
if mm:
    lo=max(0,mm[0]-60); hi=mm[-1]+10
    reg=data[lo:hi]
    n=sum(int(r[ismp]) for r in reg); print('mma-warp loop samples',n, 'of', tot)
    nm=sum(int(r[ismp]) for r in reg if 'UTCHMMA' in r[isrc] or 'UTCBAR' in r[isrc]); print('  on UTCHMMA/UTCBAR',nm)
    nw=sum(int(r[ismp]) for r in reg if 'SYNCS' in r[isrc] or ('BRA' in r[isrc])); print('  on waits/branches',nw)
    agg={}
    for r in reg:
        for i in stall_cols:
            if r[i] not in ('','0'): agg[hdr[i]]=agg.get(hdr[i],0)+int(r[i])
    print(sorted(agg.items(),key=lambda kv:-kv[1])[:6])
print('--- top instructions')
for r in sorted(data,key=lambda r:-int(r[ismp]))[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    st={hdr[i][6:]:int(r[i]) for i in stall_cols if r[i] not in('','0')}
    top=sorted(st.items(),key=lambda kv:-kv[1])[:2]
    print(f"{int(r[ia],16)-base:6x} {r[ismp]:>6} ex={r[iex]:>8} {r[isrc].strip()[:60]:60s} {top}")
