"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki, vi, mi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Name'), H.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for r in data:
    if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ki]); name = re.sub(r'^void (magb200::)?(\(anonymous namespace\)::)?', '', name)
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f'total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare shares)')
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f'{t:9.1f} us {100 * t / tot:5.1f}%  n={n:3d}  avg={t / n:7.1f}  {k[:100]}')
