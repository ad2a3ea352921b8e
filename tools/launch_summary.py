"""Per-kernel totals of ONE bench step from an ncu launch list.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file launches.csv \\
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph
    python tools/launch_summary.py launches.csv > profiles/<name>_summary.txt

A step is delimited by the batched NMS launch that ends it (one per step); the LAST complete step in the list is
summarised.  ncu serialises kernels and runs them with cold caches: compare shares, not absolute times.
"""
import csv
import sys


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        rows.append((int(r[ix['ID']]), r[ix['Kernel Name']], float(r[ix['Metric Value']])))
    ends = [i for i, (_, name, _) in enumerate(rows) if 'nms_small_kernel' in name or 'nms_mask_kernel' in name]
    if len(ends) < 2:
        print('need at least two steps in the list')
        return
    a, b = ends[-2] + 1, ends[-1] + 1
    step = [r for r in rows[a:b] if 'FillFunctor<unsigned char>' not in r[1]]       # bench.py's L2 flush
    # the launches after the NMS up to the first kernel of the next step (finalize/decode of the SAME step) follow
    # the NMS in the head's get_bboxes; include everything up to the next step's first GroupNorm/conv by taking the
    # window between two consecutive NMS launches (same count of every kernel either way)
    tot = sum(t for _, _, t in step)
    agg = {}
    for _, name, t in step:
        key = name.split('(')[0][:70]
        c = agg.setdefault(key, [0.0, 0])
        c[0] += t
        c[1] += 1
    print('one bench step (eager), launches %d..%d: %d launches, %.1f us in total (ncu: serialised, cold caches -- '
          'compare shares, not absolutes)' % (step[0][0], step[-1][0], len(step), tot / 1e3))
    for key, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print('%8.1f us %5.1f%% x%3d %s' % (t / 1e3, 100 * t / tot, n, key))


if __name__ == '__main__':
    main()
