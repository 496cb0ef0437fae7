"""Regenerates tests/golden/*.npz from the reference's own test data.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Produces
  bspline_golden.npz   Tensor_BSpline.txt (100x49) and P_mat.txt (49x49): the two stored
                       expected values asserted in src/test-BSplines.cpp:58-82
  sim_inputs.npz       Sim_data.RDS/time.RDS (config 1: 40 functions x 100 points),
                       MVSim_data.RDS (20x10), HDSim_data.RDS/HDtime.RDS (20 x 144, 144x2)
  trace_summaries.npz  posterior summaries of the three stored 150-draw chains
                       inst/test-data/{Functional,Multivariate,HDFunctional}_trace
  (ref_updates.npz is produced by tests/golden/make_ref_vectors.py from oracle/_ref)
"""
import gzip
import os
import struct
import sys

import numpy as np

REF = os.environ.get("BFMMM_REFERENCE", "/root/reference")
TD = os.path.join(REF, "inst", "test-data")
OUT = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------- RDS (gzip + XDR v3)
class _Rds:
    def __init__(self, raw):
        self.b = raw
        self.i = 0

    def int(self):
        v = struct.unpack(">i", self.b[self.i:self.i + 4])[0]
        self.i += 4
        return v

    def item(self):
        flags = self.int()
        typ = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if typ == 254:          # NILVALUE
            return None
        if typ == 255:          # reference to an earlier symbol
            return ("ref", flags >> 8)
        if typ == 1:            # symbol
            return ("sym", self.item())
        if typ == 2:            # pairlist
            out = []
            while True:
                if has_attr:
                    self.item()
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.int()
                t2 = flags & 0xFF
                if t2 == 254:
                    break
                assert t2 == 2, t2
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
            return ("pairlist", out)
        if typ == 9:            # CHARSXP
            ln = self.int()
            if ln == -1:
                return None
            s = self.b[self.i:self.i + ln].decode("latin1")
            self.i += ln
            return s
        if typ == 13:
            ln = self.int()
            v = np.frombuffer(self.b, dtype=">i4", count=ln, offset=self.i).astype(np.int64)
            self.i += 4 * ln
            res = v
        elif typ == 14:
            ln = self.int()
            v = np.frombuffer(self.b, dtype=">f8", count=ln, offset=self.i).astype(np.float64)
            self.i += 8 * ln
            res = v
        elif typ == 16:
            ln = self.int()
            res = [self.item() for _ in range(ln)]
        elif typ == 19:
            ln = self.int()
            res = [self.item() for _ in range(ln)]
        else:
            raise ValueError(f"RDS type {typ} not handled")
        if has_attr:
            attr = self.item()
            dim = None
            if attr and attr[0] == "pairlist":
                for tag, car in attr[1]:
                    name = tag[1] if tag and tag[0] == "sym" else None
                    if name == "dim" or (isinstance(car, np.ndarray) and car.dtype == np.int64 and dim is None
                                         and isinstance(res, np.ndarray) and int(np.prod(car)) == res.size):
                        dim = car
            if dim is not None and isinstance(res, np.ndarray):
                res = res.reshape(tuple(int(x) for x in dim), order="F")
        return res


def read_rds(path):
    raw = gzip.open(path, "rb").read()
    assert raw[:2] == b"X\n", raw[:4]
    r = _Rds(raw)
    r.i = 2
    version = r.int(); r.int(); r.int()
    if version >= 3:
        ln = r.int()
        r.i += ln
    return r.item()


# ---------------------------------------------------------------- Armadillo text formats
def read_arma_txt(path):
    with open(path, "r") as f:
        head = f.readline().strip()
        dims = [int(x) for x in f.readline().split()]
        vals = np.array(f.read().split(), dtype=np.float64)
    if head == "ARMA_MAT_TXT_FN008":
        r, c = dims
        return vals.reshape(r, c)
    if head == "ARMA_CUB_TXT_FN008":
        r, c, s = dims
        return vals.reshape(s, r, c).transpose(1, 2, 0)
    raise ValueError(head)


def read_arma_field_cube_bin(path):
    """ARMA_FLD_BIN field of cubes (Phi0.txt: 150 x 1 field of K x P x M cubes): list of column-major cubes."""
    raw = open(path, "rb").read()
    pos = 0

    def line():
        nonlocal pos
        e = raw.index(b"\n", pos)
        out = raw[pos:e].decode()
        pos = e + 1
        return out
    assert line() == "ARMA_FLD_BIN"
    nr, nc = int(line()), int(line())
    cubes = []
    for _ in range(nr * nc):
        assert line() == "ARMA_CUB_BIN_FN008"
        r, c, sl = (int(x) for x in line().split())
        v = np.frombuffer(raw, dtype="<f8", count=r * c * sl, offset=pos)
        pos += 8 * r * c * sl
        cubes.append(v.reshape(sl, c, r).transpose(2, 1, 0).copy())       # [r][c][slice]
    return cubes


def main():
    if not os.path.isdir(TD):
        sys.exit(f"{TD} not found: run this in the build container")
    np.savez_compressed(os.path.join(OUT, "bspline_golden.npz"),
                        tensor_bspline=read_arma_txt(os.path.join(TD, "Tensor_BSpline.txt")),
                        p_mat=read_arma_txt(os.path.join(TD, "P_mat.txt")))

    sim = read_rds(os.path.join(TD, "Sim_data.RDS"))
    tim = read_rds(os.path.join(TD, "time.RDS"))
    mv = read_rds(os.path.join(TD, "MVSim_data.RDS"))
    hd = read_rds(os.path.join(TD, "HDSim_data.RDS"))
    hdt = read_rds(os.path.join(TD, "HDtime.RDS"))
    np.savez_compressed(os.path.join(OUT, "sim_inputs.npz"),
                        sim_y=np.stack([np.asarray(v).ravel() for v in sim]),
                        sim_t=np.stack([np.asarray(v).ravel() for v in tim]),
                        mv_y=np.asarray(mv),
                        hd_y=np.stack([np.asarray(v).ravel() for v in hd]),
                        hd_t=np.stack([np.asarray(v) for v in hdt]))

    summ = {}
    for fam in ("Functional", "Multivariate", "HDFunctional"):
        d = os.path.join(TD, fam + "_trace")
        sig = read_arma_txt(os.path.join(d, "Sigma0.txt")).ravel()
        a3 = read_arma_txt(os.path.join(d, "alpha_30.txt")).ravel()
        pi = read_arma_txt(os.path.join(d, "Pi0.txt"))
        nu = read_arma_txt(os.path.join(d, "Nu0.txt"))
        Z = read_arma_txt(os.path.join(d, "Z0.txt"))
        half = slice(len(sig) // 2, None)
        summ[fam + "_sigma_q"] = np.quantile(sig[half], [0.05, 0.5, 0.95])
        summ[fam + "_alpha3_med"] = np.median(a3[half])
        summ[fam + "_pi_med"] = np.median(pi[:, half], axis=1)
        summ[fam + "_nu_med"] = np.median(nu[:, :, half], axis=2)
        summ[fam + "_Z_med"] = np.median(Z[:, :, half], axis=2)
        summ[fam + "_sigma_chain"] = sig
        # Posterior summaries of the second half of the stored chain (75 draws).  Z, nu, Phi are only identified up to
        # the mixing transformations Z -> Z A, nu -> A^-1 nu (labels, rotations of the pseudo-eigenfunctions), so the
        # comparable quantities are the function-level ones: the fitted mean coefficients (Z nu)_i (n x P) and the
        # pointwise variances diag(U_i U_i'), U_i = sum_k Z_ik Phi_k (n x P); Z, nu and sum_m phi_km phi_km' are kept
        # with the labels as stored.  Mean, sd and 5 / 95 % quantiles over the draws.
        Phi = read_arma_field_cube_bin(os.path.join(d, "Phi0.txt"))
        idx = range(len(sig) // 2, len(sig))
        fit = np.stack([Z[:, :, i] @ nu[:, :, i] for i in idx])
        cov = np.stack([np.einsum("kpm,kqm->kpq", Phi[i], Phi[i]) for i in idx])
        fvar = np.stack([(np.einsum("nk,kpm->npm", Z[:, :, i], Phi[i]) ** 2).sum(axis=2) for i in idx])
        for name, arr in (("fit", fit), ("fvar", fvar), ("cov", cov), ("Z", np.stack([Z[:, :, i] for i in idx])),
                          ("nu", np.stack([nu[:, :, i] for i in idx]))):
            summ[f"{fam}_{name}_mean"] = arr.mean(axis=0)
            summ[f"{fam}_{name}_q05"] = np.quantile(arr, 0.05, axis=0)
            summ[f"{fam}_{name}_q95"] = np.quantile(arr, 0.95, axis=0)
            summ[f"{fam}_{name}_sd"] = arr.std(axis=0, ddof=1)
    np.savez_compressed(os.path.join(OUT, "trace_summaries.npz"), **summ)
    print("wrote golden fixtures to", OUT)


if __name__ == "__main__":
    main()
