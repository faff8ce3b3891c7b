"""Summarise an ncu report: key raw metrics + hottest source lines (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [n_lines]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread','launch__grid_size','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum','sm__cycles_elapsed.max','smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for i, h in enumerate(H):
    if h in want or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
        print('%-90s %-10s %s' % (h, U[i], [r[i] for r in rows[2:]]))
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
cur = None; Hh = None; lines = []; key = None; ops = Counter(); sops = Counter()
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': Hh = r; continue
    if r[0] == 'Function Name': continue
    if Hh and r[0].isdigit():
        ie = Hh.index('Instructions Executed'); ss = Hh.index('# Samples')
        try: lines.append((cur, int(r[0]), r[1].strip(), int(r[ie]), int(r[ss])))
        except Exception: pass
    elif Hh and r[0] == '' and len(r) > 8 and r[7].isdigit():
        t = r[3].split()
        if t:
            op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
            ops[op] += int(r[7]); sops[op] += int(r[6])
tot = sum(l[3] for l in lines) or 1; tots = sum(l[4] for l in lines) or 1
ti = sum(ops.values()) or 1; tsm = sum(sops.values()) or 1
print('--- opcodes (inst%, sample%)')
print(', '.join('%s %.1f/%.1f' % (o, 100*n/ti, 100*sops[o]/tsm) for o, n in ops.most_common(18)))
print('--- top lines by instructions')
for l in sorted(lines, key=lambda l: -l[3])[:N]:
    print('%-20s %4d inst %5.1f%% samp %5.1f%% | %s' % (l[0][:20], l[1], 100*l[3]/tot, 100*l[4]/tots, l[2][:100]))
print('--- top lines by stall samples')
for l in sorted(lines, key=lambda l: -l[4])[:N]:
    print('%-20s %4d inst %5.1f%% samp %5.1f%% | %s' % (l[0][:20], l[1], 100*l[3]/tot, 100*l[4]/tots, l[2][:100]))
