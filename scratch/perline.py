import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
frames=float(sys.argv[2]); thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.4
H=rows[2]
iE=H.index("Instructions Executed"); iN=H.index("# Samples"); iW=H.index("L1 Wavefronts Shared")
def I(x):
    try: return int(x)
    except: return 0
byline=collections.OrderedDict()
for r in rows[3:]:
    if r[0] not in ("","-"):
        try: l=int(r[0])
        except: continue
        s,e,n,w=byline.get(l,(r[1],0,0,0))
        byline[l]=(r[1],e+I(r[iE]),n+I(r[iN]),w+I(r[iW]))
ts=sum(v[2] for v in byline.values())
print("total instr/frame",sum(v[1] for v in byline.values())/frames, "smem wavefronts/frame", sum(v[3] for v in byline.values())/frames)
for l in sorted(byline):
    s,e,n,w=byline[l]
    if e/frames>thr: print(f"{l:4d} {e/frames:7.2f} {n/ts*100:5.1f}% wf {w/frames:6.2f}  {s.strip()[:110]}")
