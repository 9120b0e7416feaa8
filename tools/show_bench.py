import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        r = d["roofline"]
        print(f"{f}: {d['value']/1e6:.1f}M trades/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']/1e6:.1f}M  kernels {r['all_kernels_ms']}  frac {r['frac']:.3f} step_frac {r.get('step_frac_physical', r.get('step_frac', float('nan'))):.3f} gate {d['config']['parity_gate_scaled_err']:.1e}")
    except Exception as e:
        print(f, "ERR", e)
