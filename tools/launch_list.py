"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file X cmd`).
usage: python tools/launch_list.py launches.csv "command that was profiled" > profiles/rNN_launches.txt"""
import collections
import csv
import sys

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '?')
rows = [r for r in csv.DictReader(l for l in open(path) if not l.startswith('=='))]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    ms = float(r['Metric Value'].replace(',', '')) / 1e6
    tot[r['Kernel Name']] += ms; cnt[r['Kernel Name']] += 1
total = sum(tot.values())
print(f'# ncu launch list of: {cmd}')
print('# ncu --metrics gpu__time_duration.sum --clock-control none ; per-launch times are cold-cache and serialised: compare SHARES')
print(f'# total GPU time {total:.1f} ms over {len(rows)} launches')
print('share%   total_ms   n   kernel')
for k, v in tot.most_common(12):
    print(f'{100 * v / total:7.3f} {v:11.3f} {cnt[k]:4d}  {k[:110]}')
print('\n# launches of this repo\'s kernels in order (id, kernel, ms)')
for r in rows:
    if 'srb::' in r['Kernel Name'] or 'k_' in r['Kernel Name'].split('(')[0][-16:]:
        print(r['ID'], r['Kernel Name'][:70], f"{float(r['Metric Value'].replace(',', '')) / 1e6:.3f}")
