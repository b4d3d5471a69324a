"""Top stall sites per kernel from `ncu -i X.ncu-rep --page source --csv` (one section per profiled launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
for si in range(len(starts) - 1):
    if only is not None and si != only:
        continue
    sec = rows[starts[si]:starts[si + 1]]
    hdr = sec[1]; idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in sec[2:] if len(r) == len(hdr)]
    tot = sum(int(r[idx['# Samples']]) for r in data)
    print('=====', si, sec[0][1][:100]); print('total samples', tot, 'instrs', len(data))
    rank = sorted(range(len(data)), key=lambda i: -int(data[i][idx['# Samples']]))[:n]
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {}
    for r in data:
        for s in stalls:
            agg[s] = agg.get(s, 0) + int(r[idx[s]])
    print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for i in sorted(rank):
        r = data[i]
        st = {s: int(r[idx[s]]) for s in stalls if int(r[idx[s]]) > 0}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(i, r[idx['# Samples']], r[idx['Source']].strip()[:64], top, 'exec', r[idx['Instructions Executed']])
