#!/usr/bin/env python
"""Digest an .ncu-rep (read on the CPU box): headline metrics, stall mix, opcode mix, hottest source lines.
usage: python scripts/ncu_digest.py gpurun_out/prof_step.ncu-rep [out_prefix]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__cycles_elapsed.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
lines = []
for k in want:
    if k in hdr:
        i = hdr.index(k)
        lines.append([k, units[i]] + [r[i] for r in rows[2:]])
for l in lines: print(*l, sep=' | ')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
secs = []; cur = None
for r in srows:
    if r and r[0] == 'Kernel Name': cur = {'hdr': None, 'rows': []}; secs.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr'] = r; continue
    cur['rows'].append(r)
s = secs[0]; h = s['hdr']; R = s['rows']; ix = {n: i for i, n in enumerate(h)}
tot = sum(int(r[ix['Instructions Executed']]) for r in R)
print('SASS lines', len(R), 'warp-inst executed', tot)
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
agg = {n: sum(int(r[ix[n]] or 0) for r in R) for n in stalls}; ts = sum(agg.values())
print('stalls:', ', '.join(f'{n[6:]} {100*v/ts:.1f}%' for n, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
ops = collections.Counter()
for r in R:
    m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[ix['Source']])
    ops[m.group(2) if m else '?'] += int(r[ix['Instructions Executed']])
print('opcodes:', ', '.join(f'{o} {100*v/tot:.1f}%' for o, v in ops.most_common(16)))
if out:
    with open(out + '_summary.csv', 'w') as f:
        w = csv.writer(f); w.writerow(['metric', 'unit'] + [f'launch{i}' for i in range(len(rows) - 2)])
        for l in lines: w.writerow(l)
        w.writerow(['warp_inst_executed_launch0', 'inst', tot])
        w.writerow(['stall_mix_launch0', '%'] + [f'{n[6:]}={100*v/ts:.1f}' for n, v in sorted(agg.items(), key=lambda x: -x[1])[:9]])
        w.writerow(['opcode_mix_launch0', '%'] + [f'{o}={100*v/tot:.1f}' for o, v in ops.most_common(16)])
