import csv, collections, sys
for fn in sys.argv[1:]:
    with open(fn) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row['Kernel Name'].split('(')[0][-40:]
        agg.setdefault(name, []).append(float(row['Metric Value'].replace(',', '')))
    print(fn)
    for n, v in agg.items():
        print(f"  {n:42s} n={len(v):3d} mean_us={sum(v)/len(v)/1000:9.1f} min_us={min(v)/1000:9.1f} max_us={max(v)/1000:9.1f}")
