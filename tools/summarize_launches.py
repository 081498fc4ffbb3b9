"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys


def main(path, top=24):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt, rows = collections.defaultdict(float), collections.Counter(), []
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"])
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        t = t / 1e6 if unit == "ns" else t / 1e3 if unit == "us" else t
        name = re.sub(r"vbx::(tc::)?", "", re.sub(r"\(.*", "", row["Kernel Name"]))
        tot[name] += t
        cnt[name] += 1
        rows.append((t, name, row.get("Grid Size", "")))
    total = sum(tot.values())
    print(f"launches {sum(cnt.values())}, serialized kernel time {total:.2f} ms (cold-cache, under ncu: compare shares)")
    print(f"{'ms':>9} {'share':>6} {'n':>5}  kernel")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:top]:
        print(f"{v:9.3f} {100 * v / total:5.1f}% {cnt[k]:5d}  {k[:110]}")
    print("\nlongest single launches")
    for t, s, g in sorted(rows, reverse=True)[:16]:
        print(f"{t:9.3f} ms  {s[:80]}  grid {g}")


if __name__ == "__main__":
    main(sys.argv[1])
