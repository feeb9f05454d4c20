import csv,sys,subprocess
rep,kern=sys.argv[1],sys.argv[2]
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name",kern],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[1]
i_src=hdr.index("Source"); i_s=hdr.index("# Samples"); i_ex=hdr.index("Instructions Executed")
stall_cols=[(i,h) for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data=[r for r in rows[2:] if len(r)==len(hdr) and r[i_s].isdigit()]
tot=sum(int(r[i_s]) for r in data)
print("total samples",tot,"n instr",len(data))
n=int(sys.argv[3]) if len(sys.argv)>3 else 20
for r in sorted(data,key=lambda r:-int(r[i_s]))[:n]:
    st=sorted([(int(r[i]),h) for i,h in stall_cols if int(r[i])>0],reverse=True)[:3]
    print("%5d %5.1f%% ex=%8s  %-64s %s"%(int(r[i_s]),100*int(r[i_s])/tot,r[i_ex],r[i_src].strip()[:64],st))
agg={h:sum(int(r[i]) for r in data) for i,h in stall_cols}
print(sorted(agg.items(),key=lambda x:-x[1])[:8])
