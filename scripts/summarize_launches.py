"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time per kernel name, optionally only
the last `--last-frac` of the launches (to isolate the timed steps from setup and warm-up)."""
import csv, collections, sys, argparse
ap = argparse.ArgumentParser(); ap.add_argument('csv'); ap.add_argument('--tail', type=int, default=0,
                                                                       help='only the last TAIL launches')
a = ap.parse_args()
rows = list(csv.reader(open(a.csv, errors='replace')))
hdr = None; recs = []
for r in rows:
    if len(r) > 5 and r[0] == 'ID':
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        u = d['Metric Unit']
        v = v/1e3 if u == 'us' else v/1e6 if u == 'ns' else v*1e3 if u == 's' else v
        recs.append((d['Kernel Name'].split('(')[0].replace('void ', ''), v, d['Grid Size'], d['Block Size']))
if a.tail:
    recs = recs[-a.tail:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v, g, b in recs:
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print('%d launches, %.3f ms of kernel time' % (len(recs), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-34s n=%5d  %9.3f ms  %5.1f%%  avg %8.1f us' % (k, v[0], v[1], 100*v[1]/tot, 1e3*v[1]/v[0]))
