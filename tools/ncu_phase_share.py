"""Warp instructions, stall samples and active threads per instruction of the transport kernel, grouped by kernel
phase.  Reads an .ncu-rep captured with --import-source on (no GPU needed).  Phases are located through marker
strings in er3t_b200/csrc/b200rt.cu, so the grouping follows the source as it changes.
usage: ncu_phase_share.py REPORT.ncu-rep [SOURCE.cu]"""
import csv, io, os, subprocess, sys

rep = sys.argv[1]
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
cu = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'er3t_b200', 'csrc', 'b200rt.cu')
lines = open(cu).read().split('\n')

MARKS = [('pool_load/store', 'void pool_load'), ('wrapf', 'float wrapf'), ('tally helpers', 'void tally_add'),
         ('find_layer', 'int find_layer'), ('abs_tau', 'float abs_tau'), ('le_tau_generic', 'float le_tau_generic'),
         ('le_tau', 'float le_tau('), ('le_deposit', 'void le_deposit'), ('inv_dir', 'float3 inv_dir'),
         ('surface_sample', 'float surface_sample'), ('kernel setup', 'transport_kernel(const __grid_constant__'),
         ('queue pick', 'pick the fullest queue'), ('regeneration', '= regeneration'), ('flight', '= flight: geometry only'),
         ('tentative phase', '= tentative collisions (and escapes)'), ('event: load', '= event phase: collisions'),
         ('event: collision/surface', '---- hand-over record of the tentative phase'), ('event: local estimate', '---- local estimates toward every sensor'),
         ('event: new direction+roulette', '---- new direction'), ('flush', '---- flush the block-private tallies'),
         ('(end)', 'typedef void (*transport_fn)')]
starts = []
for name, pat in MARKS:
    for i, l in enumerate(lines):
        if pat in l:
            starts.append((i + 1, name))
            break
starts.sort()


def group(fname, ln):
    if fname != os.path.basename(cu):
        return fname
    g = 'other'
    for s, name in starts:
        if ln >= s:
            g = name
    return g


src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; hdr = None; groups = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr and r and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            v = (int(d['# Samples'] or 0), int(d['Instructions Executed'] or 0), int(d['Thread Instructions Executed'] or 0))
        except Exception:
            continue
        a = groups.setdefault(group(cur, int(r[0])), [0, 0, 0])
        for k in range(3):
            a[k] += v[k]
ts = sum(a[0] for a in groups.values()) or 1
ti = sum(a[1] for a in groups.values()) or 1
tt = sum(a[2] for a in groups.values())
print('total: warp instructions %d, thread instructions %d, active threads per instruction %.2f' % (ti, tt, tt / ti))
for g, a in sorted(groups.items(), key=lambda x: -x[1][1]):
    print('%-32s samples %5.1f%%  warp-inst %5.1f%%  thr/inst %5.1f' % (g, 100 * a[0] / ts, 100 * a[1] / ti, a[2] / max(1, a[1])))
