import csv,io,subprocess,sys,re
rep,pat,n=sys.argv[1],sys.argv[2],int(sys.argv[3])
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
pick=next(i for i in starts if re.search(pat,rows[i][1]))
end=next((i for i in starts if i>pick),len(rows))
hdr=rows[pick+1]; ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[pick+2:end] if len(r)>=len(hdr)]
ops={}
for r in data:
    s=r[ix['Source']].strip()
    op=(s.split()[1] if s.startswith('@') else s.split()[0]).split('.')[0]
    ops[op]=ops.get(op,0)+int(r[ix['Instructions Executed']])
tot=sum(ops.values())
print("total/clip",tot/n)
for o,c in sorted(ops.items(),key=lambda x:-x[1])[:40]: print(f"{o:10s} {c/n:8.0f} {100*c/tot:5.1f}%")
