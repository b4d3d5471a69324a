"""Compact an `ncu --csv --metrics ...` launch list (long format) to one short row per launch.
    python profiles/compact_launches.py raw.csv > compact.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
per, metrics = collections.OrderedDict(), []
for r in rows[hi + 1:]:
    if len(r) < len(ix):
        continue
    k = (r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0][:64], r[ix["Grid Size"]], r[ix["Block Size"]], r[ix["Stream"]])
    m = f'{r[ix["Metric Name"]]} [{r[ix["Metric Unit"]]}]'
    if m not in metrics:
        metrics.append(m)
    per.setdefault(k, {})[m] = r[ix["Metric Value"]]
w = csv.writer(sys.stdout)
w.writerow(["ID", "Kernel", "Grid", "Block", "Stream"] + metrics)
for k, d in per.items():
    w.writerow(list(k) + [d.get(m, "") for m in metrics])
