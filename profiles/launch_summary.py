"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into launches / total time / share per kernel."""
import csv
import sys
from collections import defaultdict


def main(path, title=""):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        tot[name] += float(r[14].replace(",", "")) / 1e6
        cnt[name] += 1
    total = sum(tot.values())
    if title:
        print(title + "\n")
    print(f"{len(rows)} launches, {total:.1f} ms of kernel time, cold-cache and serialised: compare SHARES, not absolutes.\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for name in sorted(tot, key=lambda n: -tot[n]):
        print(f"| {name[:60]} | {cnt[name]} | {tot[name]:.2f} | {100 * tot[name] / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
