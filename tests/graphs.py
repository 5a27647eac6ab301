"""Seeded test graphs (numpy, int32 CSR) shared by the oracle and GPU parity tests."""
import numpy as np


def random_csr(n, e, seed=0, n_cols=None, hub=None, empty_rows=0.0):
    """Random CSR with multinomial degrees.  hub=(row, degree) plants one heavy row (so the hub
    plan's chunking and in-order combination are exercised); empty_rows = fraction forced to 0."""
    rng = np.random.default_rng(seed)
    n_cols = n if n_cols is None else n_cols
    p = rng.random(n) ** 3 + 1e-3
    if empty_rows > 0:
        p[rng.random(n) < empty_rows] = 0
    p /= p.sum()
    deg = rng.multinomial(e, p) if e > 0 else np.zeros(n, np.int64)
    if hub is not None:
        deg[hub[0]] = hub[1]
    rowptr = np.zeros(n + 1, np.int32)
    rowptr[1:] = np.cumsum(deg)
    colind = rng.integers(0, n_cols, int(rowptr[-1])).astype(np.int32)
    return rowptr, colind


CASES = {
    # name: (n, e, kwargs)
    "tiny": (7, 20, {}),
    "empty_graph": (5, 0, {}),
    "ragged": (300, 2000, {"empty_rows": 0.3}),
    "hub": (400, 3000, {"hub": (17, 1500)}),          # 1500 > chunk 256 -> 6 chunks
    "two_hubs": (1000, 8000, {"hub": (999, 700), "empty_rows": 0.1}),
    "rect": (200, 1500, {"n_cols": 350}),
}


def case(name, seed=0):
    n, e, kw = CASES[name]
    rp, ci = random_csr(n, e, seed=seed, **kw)
    n_cols = kw.get("n_cols", n)
    return rp, ci, n_cols
