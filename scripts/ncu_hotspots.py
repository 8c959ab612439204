import csv,sys,subprocess,io
rep,kid=sys.argv[1],sys.argv[2]
raw=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name',f'regex:{kid}'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=rows[1]; data=[]
for r in rows[2:]:
    if len(r)<len(hdr): break
    data.append(r)
i_src=hdr.index('Source'); i_s=hdr.index('# Samples'); i_ie=hdr.index('Instructions Executed')
tot=sum(int(r[i_s] or 0) for r in data)
print(rows[0][1][:80]); print('total samples',tot,'inst',sum(int(r[i_ie] or 0) for r in data), 'sass lines', len(data))
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={h:sum(int(r[hdr.index(h)] or 0) for r in data) for h in stalls}
print(sorted(agg.items(), key=lambda kv:-kv[1])[:6])
n=int(sys.argv[3]) if len(sys.argv)>3 else 25
top=sorted(range(len(data)), key=lambda i:-int(data[i][i_s] or 0))[:n]
for i in sorted(top):
    r=data[i]; st={h:int(r[hdr.index(h)] or 0) for h in stalls}; m=max(st,key=st.get)
    print(i, r[i_s], r[i_ie], r[i_src][:80], m, st[m])
