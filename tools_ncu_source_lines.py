import csv, collections, subprocess, sys
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
for i,r in enumerate(rows[:12]):
    if len(r)>5: hi=i; break
hdr=rows[hi]; data=[r for r in rows[hi+1:] if len(r)==len(hdr)]
ismp=[i for i,h in enumerate(hdr) if h=="# Samples"][0]; iex=hdr.index("Instructions Executed")
def num(x):
    try: return int(x)
    except: return 0
agg=collections.OrderedDict(); cur=None
for r in data:
    if r[0] not in ('','-') and r[2] in ('','-'):
        cur=(r[0], r[1].strip()); agg.setdefault(cur,[0,0])
    else:
        if cur is None: continue
        agg[cur][0]+=num(r[ismp]); agg[cur][1]+=num(r[iex])
tot=sum(v[0] for v in agg.values()); tote=sum(v[1] for v in agg.values())
print("samples",tot,"warp-instr",tote)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:topn]:
    print("%5.1f%% smp  %5.1f%% instr  L%s: %s"%(100*v[0]/max(tot,1),100*v[1]/max(tote,1),k[0],k[1][:125]))
