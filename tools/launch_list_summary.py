"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel family, launches, total time, share."""
import collections, csv, re, sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    scale = {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rd:
        if len(r) <= vi:
            continue
        name = re.sub(r'[<(].*', '', r[ki].replace('(anonymous namespace)::', '').replace('<unnamed>::', '')).replace('void ', '').replace('la::', '')
        tot[name] += float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
        cnt[name] += 1
    T = sum(tot.values())
    print("%-28s %9s %12s %7s %12s" % ("kernel", "launches", "total ms", "share", "avg ms"))
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print("%-28s %9d %12.3f %6.1f%% %12.4f" % (k, cnt[k], v, 100 * v / T, v / cnt[k]))
    print("%-28s %9d %12.3f" % ("all", sum(cnt.values()), T))


if __name__ == '__main__':
    main(sys.argv[1])
