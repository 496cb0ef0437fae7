"""Tuning helper: per-kernel timings of bench.py under alternative builds (tools/lib_<name>.so, see
tools/build_variant.sh) and environment switches.  usage: kbench.py [name[:ENV=VAL,...]] ..."""
import json, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
specs = sys.argv[1:] or ["default"]
for spec in specs:
    name, _, envs = spec.partition(":")
    extra = dict(kv.split("=", 1) for kv in envs.split(",") if kv)
    if name != "default":
        extra["BFMMM_LIB"] = os.path.join(ROOT, "tools", f"lib_{name}.so")
        if not os.path.exists(extra["BFMMM_LIB"]):
            print(spec, "missing"); continue
    env = dict(os.environ, **extra)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "60", "--warmup", "5", "--no-cpu-baseline", "--no-extras"],
                         capture_output=True, text=True, env=env, cwd=ROOT)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(spec, out.stderr[-800:]); continue
    d = json.loads(line[-1])
    print(f"{spec}: step {d['ms_per_step']*1e3:.0f}us e2e {d['e2e']['ms_per_step']*1e3:.0f}us", {k: round(v['ms']*1e3, 1) for k, v in d['roofline']['kernels'].items()}, flush=True)
