"""Summarises an ncu --metrics gpu__time_duration.sum CSV launch list: per-kernel totals and shares."""
import csv, re, sys, collections
def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1000.0 if unit == 'ns' else (v * 1000.0 if unit == 'ms' else v)
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-52s n=%3d total=%9.1f us mean=%8.1f us %5.1f%%" % (k[:52], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
    print("total %.1f us over %d launches" % (tot, sum(len(v) for v in agg.values())))
if __name__ == "__main__":
    main(sys.argv[1])
