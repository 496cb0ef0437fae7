"""Tuning helper: per-kernel timings of bench.py under alternative builds (BFMMM_LIB) / switches."""
import json, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = [("default", {})] + [(f"z_minb{m}", {"BFMMM_LIB": os.path.join(ROOT, "tools", f"lib_mb{m}.so")}) for m in (4, 5, 6)]
for name, extra in variants:
    if "BFMMM_LIB" in extra and not os.path.exists(extra["BFMMM_LIB"]):
        continue
    env = dict(os.environ, **extra)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "60", "--warmup", "5", "--no-cpu-baseline"],
                         capture_output=True, text=True, env=env, cwd=ROOT)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(name, out.stderr[-400:]); continue
    d = json.loads(line[-1])
    print(f"{name}: step {d['ms_per_step']*1e3:.0f}us", {k: round(v['ms']*1e3, 1) for k, v in d['roofline']['kernels'].items()})
