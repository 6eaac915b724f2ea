"""Markdown summary of bench.py JSON lines (profiles/rNN_bench_*.json): headline numbers and the per-kernel roofline table.
    python tools/summarize_bench.py profiles/r02_bench_C2.json [more.json ...]"""
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:   # noqa: BLE001
        print('%s: unreadable (%s)' % (path, e)); continue
    c = d['config']
    print('### %s  (`%s`)\n' % (c['workload'], path))
    cb = d.get('cpu_baseline') or {}
    fa = d.get('fit_api') or {}
    print('| articles/s (device-resident) | ms/step | e2e (streamed host feeds) | e2e (sync per step) | `fit` API | CPU port | clocks |')
    print('|---|---|---|---|---|---|---|')
    print('| %.3f M | %.4f | %.3f M | %.3f M | %s | %s | %s MHz %s |' % (
        d['value'] / 1e6, d['ms_per_step'], d['e2e']['value'] / 1e6, d['e2e'].get('synchronous_per_step', {}).get('value', float('nan')) / 1e6,
        ('%.3f M' % (fa['value'] / 1e6)) if fa else '-', ('%.1f /s on %d cores' % (cb['value'], cb['cores'])) if cb else '-',
        d['clocks'].get('sm_mhz'), d['clocks'].get('reasons')))
    print('\n| kernel | µs (serialised) | bound | algorithmic work | achieved | peak | frac |')
    print('|---|---|---|---|---|---|---|')
    for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms']):
        if 'bound' in v:
            print('| `%s` | %.1f | %s | %.3g | %.1f %s | %.0f | **%.3f** |' % (k, v['ms'] * 1e3, v['bound'], v['work'], v['achieved'], v['unit'], v['peak'], v['frac']))
        else:
            print('| `%s` | %.1f | latency | | | | |' % (k, v['ms'] * 1e3))
    r = d['roofline']
    print('\n`roofline` (longest kernel): `%s`, %s-bound, frac %.3f; traffic %s B.\n' % (r['kernel'], r['bound'], r['frac'] or 0, r.get('traffic')))
