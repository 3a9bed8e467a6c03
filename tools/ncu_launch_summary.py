import csv, sys, collections, re
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
# take the last step: find last ~ N launches. Group by (short kernel name, grid, block)
def short(n):
    n = re.sub(r'\(.*', '', n)
    n = n.replace('void ', '').replace('(anonymous namespace)::', '')
    return n[:90]
agg = collections.OrderedDict()
start = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for r in rows[start:]:
    k = (short(r['Kernel Name']), r['Grid Size'], r['Block Size'])
    d = agg.setdefault(k, [0, 0.0])
    d[0] += 1; d[1] += float(r['Metric Value']) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{v[1]:10.1f} us {v[0]:5d}x {v[1]/v[0]:8.1f} us/launch {100*v[1]/tot:5.1f}%  {k[0]}  grid={k[1]} block={k[2]}")
