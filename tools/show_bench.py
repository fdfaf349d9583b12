"""print the interesting numbers of bench.py JSON lines: python tools/show_bench.py file.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        l = json.load(open(f))
    except Exception as e:
        print(f, 'ERR', e)
        continue
    r = l.get('roofline', {})
    print(f"{f}: N={l['n_gpus']} {l.get('scaling')} value {l['value']:.4g} ms_step {l['ms_per_step']:.5f} "
          f"kernel_ms {r.get('kernel_ms', 0):.5f} (ev {r.get('kernel_ms_event_pairs', 0):.5f}) stream_ms {l.get('ms_per_step_stream_launch', 0):.5f} "
          f"frac {r.get('frac', 0):.4f} traffic {r.get('traffic')}")
    e = l.get('e2e', {})
    if 'frac_of_d2h_ceiling' in e:
        print(f"  e2e {e['value']:.4g} ({e['frac_of_d2h_ceiling']:.3f} of ceil {e['d2h_ceiling_gbs_per_gpu']:.1f} GB/s)  "
              f"resident {l['e2e_resident']['value']:.4g} ms {l['e2e_resident']['ms_per_step']:.3f}")
    for k, v in (l.get('other_workloads') or {}).items():
        if k != 'solvers':
            print('  ', k, {a: (round(b, 5) if isinstance(b, float) else b) for a, b in v.items()
                            if a in ('ms_per_step', 'kernel_ms', 'roofline_frac', 'error')})
        else:
            print('   solvers', {kk: {a: (round(b, 3) if isinstance(b, float) else b) for a, b in vv.items() if a in ('ms', 'iterations_mean', 'converged_frac', 'evals_mean')} for kk, vv in v.items() if isinstance(vv, dict)} if 'error' not in v else v)
    for k, v in (l.get('strong_scaling') or {}).items():
        print('   strong', k, f"P/gpu {v['problems_per_gpu']} ms {v['ms_per_step']:.5f} evals/s {v['evals_per_s']:.4g} frac/gpu {v['roofline_frac_per_gpu']:.4f}")
    if 'cpu_baseline' in l:
        c = l['cpu_baseline']
        print('   cpu', {k: (round(v, 1) if isinstance(v, float) else v) for k, v in c.items() if k != 'sample'})
    if l.get('gather'):
        print('   gather', l['gather'])
    print('   clocks', l.get('clocks'))
