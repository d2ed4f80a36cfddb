"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (name, grid, block)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        hdr, start = r, i + 1
        break
ik, iv, ig, ib = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'), hdr.index('Block Size')
agg = collections.defaultdict(list)
for r in rows[start:]:
    if len(r) <= iv:
        continue
    name = r[ik].split('(')[0][-56:]
    agg[(name, r[ig], r[ib])].append(float(r[iv].replace(',', '')) / 1000)
tot = sum(sum(v) for v in agg.values())
print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v):9.1f} us {100 * sum(v) / tot:5.1f}%  n={len(v):3d} avg={sum(v) / len(v):7.2f} us  {k[0]} grid={k[1]} block={k[2]}")
