import csv, subprocess, sys, io
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur=None; hdr=None; out=[]
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if hdr and r and r[0].isdigit():
        d=dict(zip(hdr,r))
        try: out.append((cur,int(r[0]),int(d['# Samples'] or 0),int(d['Instructions Executed'] or 0),int(d['Thread Instructions Executed'] or 0)))
        except Exception: pass
print(hdr)
groups={}
def grp(f,l):
    if f!='b200rt.cu': return f
    for name,(a,b) in {'setup':(460,526),'regen':(527,574),'flight':(575,713),'tent':(714,748),'event':(749,902),'flush':(903,930),'le':(363,414),'le_generic':(296,362),'abs_tau':(285,295),'find_layer':(275,284),'wrapf':(240,246),'tally':(247,274),'sfc_sample':(419,449)}.items():
        if a<=l<=b: return name
    return 'other'
for f,l,s,i,t in out:
    g=grp(f,l); a=groups.setdefault(g,[0,0,0]); a[0]+=s;a[1]+=i;a[2]+=t
ts=sum(a[0] for a in groups.values()); ti=sum(a[1] for a in groups.values())
for g,a in sorted(groups.items(), key=lambda x:-x[1][1]):
    print('%-22s samples %5.1f%%  warp-inst %5.1f%%  thr/inst %5.1f' % (g,100*a[0]/ts,100*a[1]/ti,a[2]/max(1,a[1])))
