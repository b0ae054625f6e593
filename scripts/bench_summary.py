"""Print the key numbers of bench.py JSON lines (files given on the command line; the JSON line is the last line starting with '{')."""
import json
import sys
for f in sys.argv[1:]:
    line = [l for l in open(f).read().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    r = d.get("roofline", {})
    print(f"{f}: {d['value']:.4g} {d['unit']}  step {d['ms_per_step']:.3f} ms  kernel {r.get('kernel_ms', 0):.3f} ms  frac {r.get('frac', 0):.3f}  "
          f"adam it/s {d.get('adam_steps_per_s', 0):.1f}  e2e {d.get('e2e', {}).get('value', 0):.4g}  path {d['config'].get('kernel_path')}")
