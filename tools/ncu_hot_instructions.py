import csv,io,subprocess,sys,re
rep,pat=sys.argv[1],sys.argv[2]
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
pick=next(i for i in starts if re.search(pat,rows[i][1]))
end=next((i for i in starts if i>pick),len(rows))
hdr=rows[pick+1]; ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[pick+2:end] if len(r)>=len(hdr)]
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot=sum(int(r[ix['# Samples']]) for r in data)
agg={k:0 for k in stalls}
for r in data:
    for k in stalls: agg[k]+=int(r[ix[k]])
print("total samples",tot," ".join(f"{k[6:]}:{100*v/tot:.1f}" for k,v in sorted(agg.items(),key=lambda x:-x[1])[:10]))
top=sorted(range(len(data)),key=lambda i:-int(data[i][ix['# Samples']]))[:int(sys.argv[3]) if len(sys.argv)>3 else 25]
for i in sorted(top):
    r=data[i]; s=int(r[ix['# Samples']])
    st=" ".join(f"{k[6:]}:{int(r[ix[k]])}" for k in sorted(stalls,key=lambda k:-int(r[ix[k]]))[:3])
    print(i, r[ix['Source']].strip()[:64].ljust(64), 'exec',r[ix['Instructions Executed']].rjust(8), 'samples %.2f%%'%(100*s/tot), st)
