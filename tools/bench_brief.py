"""One-screen digest of a bench.py JSON line: python tools/bench_brief.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'steps', 'warmup', 'n_gpus', 'scaling') if k in d})
for k in ('e2e', 'cpu_baseline', 'parity', 'clocks'):
    v = d.get(k)
    if isinstance(v, dict):
        v = {kk: (vv if not isinstance(vv, str) else vv[:60]) for kk, vv in v.items()}
    print(k, v)
if d.get('roofline'):
    print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'ms', 'traffic', 'algorithmic_bytes_per_launch')})
    print('roofline_conv', {k: d['roofline_conv'][k] for k in ('achieved', 'frac', 'ms_per_frame')})
rc = d.get('reference_cuda')
if rc:
    print('reference_cuda', {k: (round(rc[k]['value'], 2) if isinstance(rc[k], dict) else '') for k in rc})
for name, o in (d.get('other_configs') or {}).items():
    if 'value' in o:
        print(name, 'value %.1f e2e %.1f ms/step %.1f gn frac %.3f (traffic %s) conv frac %.3f launches %d' % (
            o['value'], o['e2e']['value'], o['ms_per_step'], o['roofline']['frac'], o['roofline']['traffic'], o['roofline_conv']['frac'], o['gpu_launches']))
    else:
        print(name, o)
