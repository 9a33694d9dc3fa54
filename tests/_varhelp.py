"""Shared helpers for the variational parity tests: oracle state <-> packed state arrays."""
import numpy as np
import torch

from oracle import variational as ov
from stpbenchmark_b200 import varops

LIK = {"VarGP": 0, "VarTGP": 1, "VarLGP": 2}
PRIOR = varops.var_prior_vector()
GH = varops.gauss_hermite_table()


def perturbed_state(Xt, kind, seed=1):
    """An oracle VarState away from every symmetric point (so that no gradient vanishes by accident)."""
    g = torch.Generator().manual_seed(seed)
    n = Xt.shape[0]
    st = ov.init_state(Xt, kind, init_noise=torch.randn(n, dtype=torch.double, generator=g))
    with torch.no_grad():
        st.Z += 0.01 * torch.randn(st.Z.shape, dtype=torch.double, generator=g)
        st.c += 0.1; st.ls *= 1.3; st.os *= 0.9; st.noise *= 2.0
        R = torch.randn(n, n, dtype=torch.double, generator=g) * 0.05
        st.nat2 = -0.5 * torch.eye(n, dtype=torch.double) - R @ R.T
        st.nat1 = torch.randn(n, dtype=torch.double, generator=g) * 0.3
        if st.rawdf is not None:
            st.rawdf += 0.2
    return st


def pack(st, cap):
    """Oracle state -> numpy arrays in the library's layout (one campaign)."""
    n, d = st.Z.shape
    Z = np.zeros((cap, d)); Z[:n] = st.Z.detach().numpy()
    hyp = np.zeros(d + 4)
    hyp[0] = st.c.item(); hyp[1:1 + d] = st.ls.detach().numpy().ravel(); hyp[1 + d] = st.os.item()
    hyp[2 + d] = st.noise.item(); hyp[3 + d] = st.rawdf.item() if st.rawdf is not None else 0.0
    nat1 = np.zeros(cap); nat1[:n] = st.nat1.numpy()
    nat2 = np.zeros((cap, cap)); nat2[:n, :n] = st.nat2.numpy()
    return Z, hyp, nat1, nat2


def split_grads(g, n, d, cap):
    g = np.asarray(g)
    o = cap + cap * cap
    return dict(nat1=g[:n], nat2=g[cap:o].reshape(cap, cap)[:n, :n], Z=g[o:o + n * d].reshape(n, d),
                hyp=g[o + cap * d:o + cap * d + d + 4])


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def check_grads(got, ref, d, tol):
    assert rel(got["nat1"], ref["nat1"].numpy()) <= tol
    assert rel(got["nat2"], ref["nat2"].numpy()) <= tol
    assert rel(got["Z"], ref["Z"].numpy()) <= tol
    h = got["hyp"]
    assert abs(h[0] - ref["c"].item()) <= tol * max(1.0, abs(ref["c"].item()))
    assert rel(h[1:1 + d], ref["ls"].numpy().ravel()) <= tol
    assert abs(h[1 + d] - ref["os"].item()) <= tol * abs(ref["os"].item())
    assert abs(h[2 + d] - ref["noise"].item()) <= tol * abs(ref["noise"].item())
    if "rawdf" in ref:
        assert abs(h[3 + d] - ref["rawdf"].item()) <= tol * abs(ref["rawdf"].item())
