"""Builds a bayesfmmm_b200.Engine for a tests/cases.py case."""
import numpy as np

import bayesfmmm_b200 as bf
from bayesfmmm_b200.engine import FUNCTIONAL, MULTIVARIATE
from tests import cases


def engine_for(name, device_basis=False, **kw):
    kind, _ = cases.CASES[name]
    s, d, st = cases.build(name)
    if kind == "hd":      # BHDFMMM: the functional engine with the tensor-product basis (BFMMM.h:3069-3072)
        eng = bf.Engine(model=FUNCTIONAL, n=s["n"], K=s["K"], P=s["P"], M=s["M"], y=s["y"], B=s["B"], T=s["T"], **kw)
    elif kind == "common":
        if device_basis:
            eng = bf.Engine(model=FUNCTIONAL, n=s["n"], K=s["K"], P=s["P"], M=s["M"], y=s["y"], T=s["T"], t=s["t"],
                            degree=s["degree"], internal_knots=s["internal_knots"], boundary=s["boundary"], X=s["X"], **kw)
        else:
            eng = bf.Engine(model=FUNCTIONAL, n=s["n"], K=s["K"], P=s["P"], M=s["M"], y=s["y"], B=s["B"], T=s["T"],
                            X=s["X"], **kw)
    elif kind == "mv":
        eng = bf.Engine(model=MULTIVARIATE, n=s["n"], K=s["K"], P=s["P"], M=s["M"], y=s["y"], X=s["X"], **kw)
    elif device_basis:
        eng = bf.Engine(model=FUNCTIONAL, n=s["n"], K=s["K"], P=s["P"], M=s["M"], y=s["y"], off=s["off"], t=s["t"],
                        degree=s["degree"], internal_knots=s["internal_knots"], boundary=s["boundary"], X=s["X"],
                        common_grid=False, **kw)
    else:
        eng = bf.Engine(model=FUNCTIONAL, n=s["n"], K=s["K"], P=s["P"], M=s["M"], y=s["y"], B=s["B"], off=s["off"],
                        X=s["X"], common_grid=False, **kw)
    par = s["par"]
    eng.set_state(s["Z"], s["chi"])
    eng.set_globals(par["nu"], par["Phi"], par["sigma_sq"], eta=par["eta"], xi=par["xi"])
    return s, d, st, eng


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
