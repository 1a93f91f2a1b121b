"""Generates tests/golden/golden_v1.json from the CPU oracle (run once, after test_oracle_pinning passed):
    python tests/golden/make_golden.py
Seeded inputs come from the shared counter-based generator (orc_fill_uniform), so the GPU tests regenerate the
same inputs on the device and compare against these frozen outputs."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def lbfgs_state(orc, n, mem, inverse, npush, scaling=True):
    op = orc.LBFGS(n, mem=mem, scaling=scaling, inverse=inverse)
    for i in range(npush):
        s = orc.uniform(n, 100 + i)
        y = s + 0.1 * orc.uniform(n, 200 + i)
        op.push(s, y)
    return op


def compute_cases(orc):
    out = {}
    n = 1000
    x = orc.uniform(n, 7)
    r0 = orc.uniform(n, 8)
    d = orc.uniform(n, 1)
    v = orc.uniform(n, 2)
    res = r0.copy()
    orc.diag_(res, d, v, 2.0, 2.0)
    out["diag_a2_b2"] = res[:16]
    h = orc.uniform(n, 3)
    h /= np.linalg.norm(h)
    res = np.empty(n)
    orc.householder_(res, h, v, 1.0, 0.0)
    out["householder"] = res[:16]
    # cfg3 chain (H*D + 0.1*I) v
    dd = orc.uniform(n, 4, 0.5, 1.5)
    vv = orc.uniform(n, 5)
    chain = orc.opHouseholder(h) * orc.opDiagonal(dd) + 0.1 * orc.opEye(n)
    out["cfg3_chain"] = chain(vv)[:16]
    for inverse in (False, True):
        op = lbfgs_state(orc, n, 5, inverse, 7)
        out["lbfgs_inv%d_apply" % inverse] = op.apply(x)[:16]
        res = r0.copy()
        op.apply(x, 1.5, -0.25, res=res)
        out["lbfgs_inv%d_apply_ab" % inverse] = res[:16]
    op = lbfgs_state(orc, n, 5, False, 7)
    out["lbfgs_diag"] = op.diag()[:16]
    sr = orc.LSR1(n, mem=5)
    for i in range(7):
        s = orc.uniform(n, 300 + i, -1.0, 1.0)
        y = 2.0 * s + 0.3 * orc.uniform(n, 400 + i, -1.0, 1.0)
        sr.push(s, y)
    out["lsr1_apply"] = sr.apply(x)[:16]
    out["lsr1_diag"] = sr.diag()[:16]
    return {k: np.asarray(v, dtype=np.float64).tolist() for k, v in out.items()}


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
    import oracle
    oracle.build()
    oracle.set_mode(True, 1)
    cases = compute_cases(oracle)
    json.dump({"generator": "tests/golden/make_golden.py", "oracle": "oracle/b2o_oracle.c (long-double reductions)",
               "cases": cases}, open(os.path.join(HERE, "golden_v1.json"), "w"), indent=1)
    print("wrote", len(cases), "cases")
