import json, os, sys, subprocess
sys.path.insert(0, '/root/repo')
for vz, vc, inj in [(1,2,0),(1,2,1),(2,2,1)]:
    env=dict(os.environ, BFMMM_V_Z=str(vz), BFMMM_V_CHI=str(vc), BFMMM_V_SSR='2')
    if inj: env['BFMMM_Z_INJECT']='1'
    out=subprocess.run([sys.executable,'bench.py','--steps','30','--warmup','5','--no-cpu-baseline'],capture_output=True,text=True,env=env,cwd='/root/repo')
    line=[l for l in out.stdout.splitlines() if l.startswith('{')]
    if not line: print(out.stderr[-500:]); continue
    d=json.loads(line[-1])
    print(f"Vz={vz} Vchi={vc} inject={inj}: step {d['ms_per_step']*1e3:.0f}us", {k:round(v['ms']*1e3,1) for k,v in d['roofline']['kernels'].items()})
