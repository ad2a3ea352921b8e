"""Concurrency view of the replayed training step from a torch.profiler (CUPTI) trace:

    KGDET_TRAIN_TRACE=gpurun_out/train_trace.json python bench.py --mode train --steps 5 --warmup 3
    python tools/train_timeline.py gpurun_out/train_trace.json > profiles/<name>.txt

For the LAST step in the trace (between two L2-flush fills): wall time, summed kernel time, average number of kernels
in flight, time with 1 / 2 / 3+ kernels in flight, and the kernels that run ALONE for the longest total time -- the
serial sections that bound the step.
"""
import json
import sys


def main():
    ev = json.load(open(sys.argv[1]))['traceEvents']
    ks = [e for e in ev if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'dur' in e]
    ks.sort(key=lambda e: e['ts'])
    fl = [i for i, e in enumerate(ks) if 'FillFunctor<unsigned char>' in e['name']]
    if len(fl) < 2:
        print('need two flush fills in the trace, found', len(fl))
        return
    a, b = fl[-2] + 1, fl[-1]
    step = ks[a:b]
    t0 = min(e['ts'] for e in step)
    t1 = max(e['ts'] + e['dur'] for e in step)
    pts = []
    for i, e in enumerate(step):
        pts.append((e['ts'], 1, i))
        pts.append((e['ts'] + e['dur'], -1, i))
    pts.sort()
    active = set()
    last = t0
    hist = {}
    alone = {}
    for t, d, i in pts:
        dt = t - last
        if dt > 0:
            k = len(active)
            hist[k] = hist.get(k, 0.0) + dt
            if k == 1:
                name = step[next(iter(active))]['name'].split('(')[0][:80]
                alone[name] = alone.get(name, 0.0) + dt
        last = t
        if d == 1:
            active.add(i)
        else:
            active.discard(i)
    wall = t1 - t0
    busy = sum(e['dur'] for e in step)
    streams = sorted(set(e.get('args', {}).get('stream', e.get('tid')) for e in step))
    print('last replayed training step: %d kernels on %d streams, wall %.1f us, summed kernel time %.1f us, '
          'average kernels in flight %.2f' % (len(step), len(streams), wall, busy, busy / wall))
    for k in sorted(hist):
        print('  %2d in flight: %8.1f us  %5.1f%%' % (k, hist[k], 100 * hist[k] / wall))
    print('kernels running ALONE (serial sections), by total time:')
    for name, t in sorted(alone.items(), key=lambda kv: -kv[1])[:30]:
        print('  %8.1f us  %s' % (t, name))


if __name__ == '__main__':
    main()
