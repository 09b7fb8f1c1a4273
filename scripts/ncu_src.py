"""Summarise an `ncu --page source --csv` dump: stall-reason totals and hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(x):
    try: return float(x)
    except: return 0.0
tot = sum(f(r[idx['# Samples']]) for r in data)
print('total samples', tot, 'instr lines', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(f(r[idx[s]]) for r in data) for s in stalls}
print([(k, round(v / max(tot, 1) * 100, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
top = sorted(data, key=lambda r: -f(r[idx['# Samples']]))[:n]
for r in top:
    st = sorted(((s[6:], int(f(r[idx[s]]))) for s in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{f(r[idx['# Samples']])/max(tot,1)*100:5.1f}%  exec {r[idx['Instructions Executed']]:>10}  {r[idx['Source']][:100]:100s} {st}")
