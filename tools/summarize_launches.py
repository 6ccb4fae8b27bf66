"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list (markdown table)."""
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else v   # -> us
    name = r[ik].split("(")[0][:90]
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print("| kernel | launches | total ms | avg µs | share |\n|---|---|---|---|---|")
for k, v in tot.most_common():
    print(f"| `{k}` | {cnt[k]} | {v / 1e3:.3f} | {v / cnt[k]:.1f} | {100 * v / total:.1f}% |")
print(f"\ntotal {total / 1e3:.3f} ms over {sum(cnt.values())} launches")
