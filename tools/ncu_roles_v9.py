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
ti=sum(o[3] for o in out); ts=sum(o[2] for o in out)
def grp(name, pred):
    s=[o for o in out if pred(o)]
    i=sum(o[3] for o in s); t=sum(o[4] for o in s); sm=sum(o[2] for o in s)
    print('%-34s inst %5.1f%%  samples %5.1f%%  thr/inst %5.1f' % (name, 100*i/ti, 100*sm/ts, t/i if i else 0))
v9=lambda o:o[0]=='transport_v9.cuh'
grp('v9 ring ops (lines 60-150)', lambda o: v9(o) and 60<=o[1]<=150)
grp('v9 flight: poll/wait (254-276)', lambda o: v9(o) and 254<=o[1]<=276)
grp('v9 flight: load+setup (277-292)', lambda o: v9(o) and 277<=o[1]<=292)
grp('v9 flight: loop (293-388)', lambda o: v9(o) and 293<=o[1]<=388)
grp('v9 flight: store+push (389-400)', lambda o: v9(o) and 389<=o[1]<=403)
grp('v9 event: poll/pick (404-442)', lambda o: v9(o) and 404<=o[1]<=442)
grp('v9 event: regen (443-505)', lambda o: v9(o) and 443<=o[1]<=505)
grp('v9 event: tentative (506-612)', lambda o: v9(o) and 506<=o[1]<=612)
grp('v9 event: coll/sfc (613-760)', lambda o: v9(o) and 613<=o[1]<=760)
grp('v9 other', lambda o: v9(o) and (o[1]<60 or 150<o[1]<254 or o[1]>760))
grp('b200rt.cu', lambda o: o[0]=='b200rt.cu')
grp('rt_device.cuh', lambda o: o[0]=='rt_device.cuh')
grp('intrinsics/atomics hdrs', lambda o: o[0] not in ('b200rt.cu','rt_device.cuh','transport_v9.cuh'))
