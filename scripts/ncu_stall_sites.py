import csv,subprocess,sys
rep=sys.argv[1]; thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.02
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; curf=None
hdr=[r for r in rows if len(r)>8 and r[0]=="Line No"][0]
ix={h:i for i,h in enumerate(hdr)}
def ti(x):
    try: return int(x)
    except: return 0
seq={}
for r in rows:
    if len(r)>=2 and r[0]=="File Path": curf=r[1].split('/')[-1]
    elif len(r)>8 and r[0].isdigit(): cur=(curf,int(r[0]))
    elif len(r)>8 and r[0]=="":
        if r[2] not in seq: seq[r[2]]=(r[2], cur, r[3].strip()[:75], ti(r[4]), ti(r[ix['stall_long_sb']]), ti(r[ix['stall_barrier']]), ti(r[ix['stall_wait']]), ti(r[ix['stall_short_sb']]), ti(r[7]))
seq=sorted(seq.values())
tot=sum(x[3] for x in seq)
big=[k for k,x in enumerate(seq) if x[3]>thr*tot]
for k in big:
    print("---- %.1f%%"%(100*seq[k][3]/tot))
    for x in seq[max(0,k-4):k+2]: print(x[0][-5:], x[1][0][5:], x[1][1], x[2], x[3], "long",x[4],"bar",x[5],"wait",x[6],"short",x[7])
