"""One solve's launches out of an ncu `--metrics gpu__time_duration.sum --csv` list of a bench.py run.
A solve starts at k_scatter_cutoff; the `--step`-th one (0-based) is printed: per-kernel totals, then the sequence."""
import csv, collections, argparse
ap = argparse.ArgumentParser(); ap.add_argument('csv'); ap.add_argument('--step', type=int, default=4)
ap.add_argument('--header', default='')
a = ap.parse_args()
hdr = None; recs = []
for r in csv.reader(open(a.csv, errors='replace')):
    if len(r) > 5 and r[0] == 'ID':
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        u = d['Metric Unit']
        v = v*1e-3 if u == 'ns' else v*1e3 if u == 'ms' else v
        nm = d['Kernel Name'].split('(')[0].replace('void ', '')
        recs.append((nm, v, d['Grid Size'], d['Block Size']))
starts = [i for i, r in enumerate(recs) if r[0].startswith('k_scatter_cutoff')]
lo = starts[a.step]; hi = starts[a.step + 1] if a.step + 1 < len(starts) else len(recs)
step = recs[lo:hi]
tot = sum(r[1] for r in step)
if a.header:
    print(a.header)
print('One solve (step %d of %d in the run): %d launches, %.3f ms of kernel time (serialised, cold cache).\n' %
      (a.step, len(starts), len(step), tot/1e3))
agg = collections.OrderedDict()
for nm, v, g, b in step:
    k = nm.split('<')[0]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v
print('%-28s %4s %10s %7s' % ('kernel', 'n', 'total us', 'share'))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-28s %4d %10.1f %6.1f%%' % (k, n, v, 100*v/tot))
print('\nfull launch sequence of the step (kernel, us, grid, block):')
for nm, v, g, b in step:
    print('%-42s %7.1f %s %s' % (nm, v, g, b))
