"""Per-(kernel, grid) duration table of an ncu `--metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/ncu_by_grid.py launches.csv [name filter]"""
import collections
import csv
import sys


def main():
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith('==')]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum' or flt not in row['Kernel Name']:
            continue
        t = float(row['Metric Value'].replace(',', '')) / (1000 if row['Metric Unit'] == 'ns' else 1)
        agg.setdefault((row['Kernel Name'][:44], row['Grid Size']), []).append(t)
    for (k, g), v in agg.items():
        print("%-46s %-16s n=%2d avg %6.1f us  min %6.1f max %6.1f" % (k, g, len(v), sum(v) / len(v), min(v), max(v)))


if __name__ == "__main__":
    main()
