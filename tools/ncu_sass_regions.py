#!/usr/bin/env python
"""Where does a kernel spend its warp-stall samples?  Reads an `ncu --set full --import-source on` report here (no GPU),
walks the SASS page of one kernel and aggregates samples, executed instructions, hottest opcodes and stall reasons per
region, a region being the code between two BAR.SYNC instructions (the phases of the per-stream kernels).

    python tools/ncu_sass_regions.py gpurun_out/report.ncu-rep stft_features
"""
import csv,collections,subprocess,sys,io
rep=sys.argv[1]; kern=sys.argv[2]
nth=int(sys.argv[3]) if len(sys.argv)>3 else 0   # which of the matching launches
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass","-k","regex:"+kern.split("<")[0]],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
# one block per captured launch of a matching kernel
start=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
start=[i for i in start if kern.replace(" ","") in rows[i][1].replace(" ","").replace("(int)","")] or start
blk=rows[start[nth]: ([j for j in [i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"] if j>start[nth]]+[None])[0]]
print(blk[0][1][:80])
hdr=blk[1]; data=blk[2:]
ix={h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
tot=sum(f(r,"# Samples") for r in data)
print("samples",tot,"instr",sum(f(r,"Instructions Executed") for r in data))
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
region=0
agg=collections.defaultdict(collections.Counter); inst=collections.defaultdict(collections.Counter); st=collections.defaultdict(collections.Counter)
tots=collections.Counter(); itots=collections.Counter()
for r in data:
    src=r[ix["Source"]].strip()
    parts=src.split()
    op=(parts[1] if src.startswith('@') else parts[0]).split('.')[0] if parts else '?'
    agg[region][op]+=f(r,"# Samples"); inst[region][op]+=f(r,"Instructions Executed")
    tots[region]+=f(r,"# Samples"); itots[region]+=f(r,"Instructions Executed")
    for s in stalls: st[region][s]+=f(r,s)
    if "BAR.SYNC" in src: region+=1
for reg in sorted(agg):
    print(f"region {reg}: {int(tots[reg])} samples ({100*tots[reg]/tot:.1f}%), {int(itots[reg]/1e3)}k instr")
    print("   ops:", [(k,int(v)) for k,v in agg[reg].most_common(7)])
    print("   stalls:", [(k[6:],int(v)) for k,v in st[reg].most_common(5)])
