"""Generates tests/golden/ref_updates.npz: outputs of the REFERENCE's own update functions
(oracle/_ref = Update*.h compiled from /root/reference against oracle/shim) on the seeded
cases of tests/cases.py with the injected draws of cases.draws().  Run in the build container:

    python tests/golden/make_ref_vectors.py

tests/test_oracle_golden.py then checks the oracle restatement against these stored vectors
wherever /root/reference is absent (e.g. on the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from oracle import ref  # noqa: E402
from tests import cases  # noqa: E402


def main():
    assert ref.available(), "oracle/_ref could not be built (needs /root/reference)"
    out = {}
    for name in cases.CASES:
        s, d, st = cases.build(name)
        dr = cases.draws(name, s)
        Pm = None if d.identity_basis else orc.pmat_rw1(d.P)
        for beta in (1.0, 0.6):
            t = beta != 1.0
            key = f"{name}|{beta}|"
            out[key + "Z"] = ref.update_z(d, st, s["pi"], 1.3, cases.A_Z_PM, dr["gam"], dr["u"], beta, t)
            out[key + "chi"] = ref.update_chi(d, st, dr["eps"], beta, t)
            out[key + "sigma"] = ref.update_sigma(d, st, 1.0, 1.0, dr["gsig"], beta, t)
            if beta not in cases.block_betas(name):
                continue
            out[key + "nu"] = ref.update_nu(d, st, dr["tau"], Pm, dr["z_nu"], beta, t)
            out[key + "phi"] = ref.update_phi(d, st, dr["gamma"], dr["tilde_tau"], dr["z_phi"], beta, t)
            if d.D:
                out[key + "eta"] = ref.update_eta(d, st, dr["tau_eta"], Pm, dr["z_eta"], beta, t)
                out[key + "xi"] = ref.update_xi(d, st, dr["gamma_xi"], dr["tilde_tau_xi"], dr["z_xi"], beta, t)
        out[f"{name}|loglik"] = ref.loglik(d, st)
        if not d.identity_basis and "P100" not in name and name not in cases.HEAVY_BLOCK_CASES:      # calcLikelihoodCPO over a short stored chain
            out[f"{name}|cpo"] = ref.cpo(d, cases.stored_iterations(name, st))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_updates.npz"), **out)
    print("wrote", len(out), "reference vectors")


if __name__ == "__main__":
    main()
