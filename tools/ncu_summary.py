"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, stall reasons, hottest SASS."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'launch__registers_per_thread', 'launch__block_size',
        'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    print('kernel:', d.get('Kernel Name', '?')[:120])
    for k in keys:
        if k in d: print(f'  {k:75s} {d[k]:>18s} {u[k]}')
    st = [(float(d[h]), h) for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and d[h]]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f'  stall {h.split("stalled_")[1].split("_per_issue")[0]:28s} {v:.3f} per issue')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[h]
iA, iS, iSm, iEx = hdr.index('Address'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
data = []
for r in rows[h + 1:]:
    try: data.append((int(r[iA], 16), r[iS], int(r[iSm]), int(r[iEx])))
    except Exception: pass
base = data[0][0]; tot = sum(d[2] for d in data); totex = sum(d[3] for d in data)
print(f'  source page: {len(data)} SASS instr, samples {tot}, executed {totex}')
ops = {}
for a, s_, sm, ex in data:
    op = s_.split()[0] if not s_.startswith('@') else s_.split()[1]
    op = op.split('.')[0]
    o = ops.setdefault(op, [0, 0]); o[0] += sm; o[1] += ex
print('  by opcode (share of samples / share of executed):')
for op, (sm, ex) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f'    {op:10s} {sm / tot:7.3f} {ex / totex:7.3f}')
print('  hottest instructions:')
for a, s_, sm, ex in sorted(data, key=lambda d: -d[2])[:18]:
    print(f'    +{a - base:05x} {sm / tot:6.3f} ex={ex:>12d}  {s_[:80]}')
