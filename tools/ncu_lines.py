"""Per-source-line warp instructions / samples / active threads of one file region of an .ncu-rep, in line order.
usage: ncu_lines.py REPORT.ncu-rep FILE_SUFFIX FIRST LAST"""
import csv, io, subprocess, sys
rep, suffix, a, b = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; hdr = None; out = []; ti = 0; ts = 0
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and r and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            rec = (cur, int(r[0]), r[1].rstrip()[:110], int(d['# Samples'] or 0), int(d['Instructions Executed'] or 0), int(d['Thread Instructions Executed'] or 0))
        except Exception:
            continue
        ti += rec[4]; ts += rec[3]
        if cur.endswith(suffix) and a <= rec[1] <= b: out.append(rec)
si = sum(o[4] for o in out); ss = sum(o[3] for o in out); st = sum(o[5] for o in out)
print('region: %.1f%% of warp instructions, %.1f%% of samples, %.1f threads/inst' % (100 * si / ti, 100 * ss / ts, st / max(si, 1)))
for o in out:
    if o[4] or o[3]:
        print('%4d smp %5.2f%% inst %5.2f%% thr %5.1f | %s' % (o[1], 100 * o[3] / ts, 100 * o[4] / ti, o[5] / o[4] if o[4] else 0, o[2]))
