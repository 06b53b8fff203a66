"""Summarise an .ncu-rep (read here, no GPU): headline counters + per-source-line hot spots of one kernel."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    print('== kernel', d.get('Kernel Name'), 'grid', d.get('Grid Size'), 'block', d.get('Block Size'))
    for k in ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
              'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum',
              'smsp__thread_inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
              'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
              'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
              'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
              'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
              'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
              'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
              'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']:
        if k in d:
            print('  %-85s %s %s' % (k, d[k], u.get(k, '')))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and r and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            out.append((cur, int(r[0]), r[1].strip()[:100], int(d['# Samples'] or 0), int(d['Instructions Executed'] or 0), int(d['Thread Instructions Executed'] or 0)))
        except Exception:
            pass
ts = sum(o[3] for o in out) or 1; ti = sum(o[4] for o in out) or 1; tt = sum(o[5] for o in out)
print('== source: samples %d, warp inst %d, thread inst %d, avg active threads %.2f' % (ts, ti, tt, tt / ti))
out.sort(key=lambda o: -o[4])
for o in out[:top]:
    print('%-16s %4d smp %5.1f%% inst %5.1f%% thr/inst %5.1f | %s' % (o[0], o[1], 100 * o[3] / ts, 100 * o[4] / ti, o[5] / o[4] if o[4] else 0, o[2]))
