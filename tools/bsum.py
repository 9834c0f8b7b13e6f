"""Runs bench.py (extra args passed through) and prints a compact per-pass summary line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline"] + args, stdout=subprocess.PIPE,
                   stderr=subprocess.PIPE, text=True)
try:
    d = json.loads(r.stdout.strip().splitlines()[-1])
    p = d["roofline"]["pass_ms"]
    print(f"{os.environ.get('LABEL', '')} ms/frame={d['ms_per_step']:.4f} Mpx/s={d['value']:.0f} temporal={p['temporal']*1e3:.1f}us "
          f"variance={p['variance']*1e3:.1f}us levels={[round(x*1e3,1) for x in p['levels']]} e2e={d['e2e']['value']:.0f}")
except Exception as e:  # noqa: BLE001
    print("bench failed:", e, r.stdout[-500:], r.stderr[-1500:])
