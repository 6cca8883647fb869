"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (per kernel+grid: count, mean us, share)."""
import collections
import csv
import sys


def main(path, top=24):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = row['Kernel Name'].replace('void tcg::gemm_kernel', 'tcg').replace('<unnamed>::', '')[:64]
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row['Metric Unit'], 1e-3)
        agg[(name, row['Grid Size'])].append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
        print("%-66s grid=%-12s n=%4d avg=%7.2f us share=%5.1f%%" % (k[0], k[1], len(v), sum(v) / len(v), 100 * sum(v) / tot))
    print("total us %.1f" % tot)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
