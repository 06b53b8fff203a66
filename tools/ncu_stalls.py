"""Stall-reason samples of the transport kernel grouped by kernel phase (same grouping as ncu_phase_share.py).
usage: ncu_stalls.py REPORT.ncu-rep"""
import csv, io, os, subprocess, sys
rep = sys.argv[1]
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.argv = [sys.argv[0], rep]
import importlib.util
spec = importlib.util.spec_from_file_location('ps', os.path.join(ROOT, 'tools', 'ncu_phase_share.py'))
src_text = open(os.path.join(ROOT, 'tools', 'ncu_phase_share.py')).read()
# reuse the marker table and group() of ncu_phase_share.py without running its report
head = src_text.split('src = subprocess.run')[0]
ns = {'__file__': os.path.join(ROOT, 'tools', 'ncu_phase_share.py')}
exec(compile(head, 'ncu_phase_share_head', 'exec'), ns)
group = ns['group']
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
REASONS = ['stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_no_inst', 'stall_branch_resolving', 'stall_math', 'stall_mio', 'stall_lg',
           'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_barrier']
agg = {}
cur = None; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and r and r[0].isdigit():
        d = dict(zip(hdr, r))
        g = group(cur, int(r[0]))
        a = agg.setdefault(g, dict.fromkeys(REASONS, 0))
        for k in REASONS:
            try: a[k] += int(d.get(k) or 0)
            except ValueError: pass
tot = sum(sum(a.values()) for a in agg.values()) or 1
print('%-30s %6s | %s' % ('phase', 'all%', ' '.join('%9s' % k.replace('stall_', '')[:9] for k in REASONS)))
for g, a in sorted(agg.items(), key=lambda x: -sum(x[1].values())):
    s = sum(a.values())
    if s * 200 < tot: continue
    print('%-30s %5.1f%% | %s' % (g[:30], 100 * s / tot, ' '.join('%8.1f%%' % (100 * a[k] / tot) for k in REASONS)))
a = {k: sum(x[k] for x in agg.values()) for k in REASONS}
print('%-30s %5.1f%% | %s' % ('TOTAL', 100.0, ' '.join('%8.1f%%' % (100 * a[k] / tot) for k in REASONS)))
