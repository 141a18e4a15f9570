"""Aggregate an ncu launch list (``ncu --metrics gpu__time_duration.sum --csv``) per kernel: launches, total us, share, avg.

    python tools/summarise_launches.py profiles/r1_launches_default_bench_one_step.csv > profiles/r1_launches_default_bench_summary.csv
"""
import csv
import re
import sys


def main(path: str) -> None:
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith(("==", "#"))]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
        name = re.sub(r"^void\s+", "", name)
        rows.append((name, us))
    tot = {}
    for n, us in rows:
        a = tot.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in tot.values())
    w = csv.writer(sys.stdout)
    print(f'"# per-kernel totals of {len(rows)} launches ({total:.1f} us) from {path}"')
    w.writerow(["kernel", "launches", "total_us", "share_pct", "avg_us"])
    for n, (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        w.writerow([n, cnt, round(us, 1), round(100 * us / total, 2), round(us / cnt, 2)])


if __name__ == "__main__":
    main(sys.argv[1])
