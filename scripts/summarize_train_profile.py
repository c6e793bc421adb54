import csv, collections, sys
rows = list(csv.DictReader(open(sys.argv[1])))
tot = sum(float(r['ms']) for r in rows)
by = collections.OrderedDict()
for r in rows:
    a = by.setdefault(r['kind'], [0, 0.0]); a[0] += 1; a[1] += float(r['ms'])
print(f"total {tot:.2f} ms over {len(rows)} launches")
for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:18s} n={n:3d} {t:8.3f} ms {t / tot * 100:5.1f}%")
for r in sorted(rows, key=lambda r: -float(r['ms']))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"  {r['kind']:16s} {r['layer']:50s} {float(r['ms']):.3f}")
