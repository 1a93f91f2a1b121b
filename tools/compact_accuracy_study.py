"""Accuracy of the two forward L-BFGS representations (CPU, numpy; no GPU):
  (R) the reference's a_k / b_k recursion (src/lbfgs.jl:183-196, 236-250), Float64
  (C) the compact form the library offers with `compact=True` (Byrd-Nocedal-Schnabel 1994, Thm 2.3): Gram matrices in Float64,
      the 2m x 2m middle matrix inverted in long double on the host, coefficients and the combine in Float64
against (T) the reference recursion carried out entirely in 80-bit long double on the same Float64 pairs.
Three families of (s, y) pairs: well conditioned (the bench recipe), y = A s with cond(A) up to 1e12, and nearly dependent steps.
Writes one JSON line per case;  python tools/compact_accuracy_study.py > profiles/r2_compact_accuracy.jsonl"""
import json

import numpy as np

LD = np.longdouble


def recursion(pairs, mem, dt):
    """state after pushing all pairs, reference form; returns apply(x)"""
    n = len(pairs[0][0])
    S, Y, A, B = (np.zeros((mem, n), dt) for _ in range(4))
    ys = np.zeros(mem, dt)
    ins, gamma = 0, dt(1)
    for s, y in pairs:
        s, y = s.astype(dt), y.astype(dt)
        d = s @ y
        if d <= np.finfo(np.float64).eps:
            continue
        S[ins], Y[ins], ys[ins] = s, y, d
        gamma = d / (y @ y)
        B[ins] = y / np.sqrt(d)
        for i in range(1, mem + 1):
            k = (ins + i) % mem
            if ys[k] != 0:
                A[k] = S[k] / gamma
                for j in range(1, i):
                    l = (ins + j) % mem
                    if ys[l] != 0:
                        A[k] = A[k] + (B[l] @ S[k]) * B[l]
                        A[k] = A[k] - (A[l] @ S[k]) * A[l]
                A[k] = A[k] / np.sqrt(S[k] @ A[k])
        ins = (ins + 1) % mem

    def apply(x):
        x = x.astype(dt)
        q = x / gamma
        for i in range(1, mem + 1):
            k = (ins + i - 1) % mem
            if ys[k] != 0:
                q = q + ((B[k] @ x) * B[k] - (A[k] @ x) * A[k])
        return q
    order = [(ins + i) % mem for i in range(mem) if ys[(ins + i) % mem] != 0]
    return apply, S, Y, order, gamma


def invert_ld(M):
    n = M.shape[0]
    M = M.astype(LD).copy()
    I = np.eye(n, dtype=LD)
    for c in range(n):
        p = c + int(np.argmax(np.abs(M[c:, c])))
        M[[c, p]], I[[c, p]] = M[[p, c]], I[[p, c]]
        d = M[c, c]
        M[c], I[c] = M[c] / d, I[c] / d
        for r in range(n):
            if r != c and M[r, c] != 0:
                f = M[r, c]
                M[r], I[r] = M[r] - f * M[c], I[r] - f * I[c]
    return I


def compact_apply(S, Y, order, gamma, x):
    """B x = x/γ + [S Y] W' [Sᵀx; Yᵀx],  W' = -diag(1/γ, 1) [[SᵀS/γ, L], [Lᵀ, -D]]⁻¹ diag(1/γ, 1)   (csrc/b2o_qn.cu build_forward_W)"""
    Sa, Ya = S[order].astype(np.float64), Y[order].astype(np.float64)
    A = len(order)
    g = np.float64(gamma)
    SS, SY = Sa @ Sa.T, Sa @ Ya.T                    # Float64 Gram matrices (the GPU's double accumulators)
    M = np.zeros((2 * A, 2 * A), LD)
    M[:A, :A] = SS.astype(LD) / LD(g)
    L = np.tril(SY, -1).astype(LD)
    M[:A, A:], M[A:, :A] = L, L.T
    M[A:, A:] = -np.diag(np.diag(SY)).astype(LD)
    W = -invert_ld(M)
    W[:A, :] /= LD(g)
    W[:, :A] /= LD(g)
    W = W.astype(np.float64)
    cols = np.concatenate([Sa, Ya])
    dots = cols @ x
    coef = W @ dots
    return x / g + coef @ cols, float(np.linalg.cond(M.astype(np.float64)))


def families(n, npush, rng):
    base = [rng.random(n) for _ in range(npush)]
    yield "well conditioned: y = s + 0.1 u (the bench recipe)", [(s, s + 0.1 * rng.random(n)) for s in base]
    for cond in (1e4, 1e8, 1e12):
        d = np.logspace(0, np.log10(cond), n)
        yield "y = D s, cond(D) = %.0e" % cond, [(s, d * s) for s in base]
    for eps in (1e-3, 1e-6, 1e-9):
        s0 = rng.random(n)
        steps = [s0 + eps * rng.random(n) for _ in range(npush)]
        yield "nearly dependent steps: s_i = s_0 + %.0e u_i, y = s + 0.1 u" % eps, [(s, s + 0.1 * rng.random(n)) for s in steps]
    for decay in (0.5, 0.1, 0.01):
        # a converging iteration: step lengths shrink geometrically, the Hessian y = D s is ill conditioned
        d = np.logspace(0, 6, n)
        yield "shrinking steps: |s_i| ~ %.2g^i, y = D s, cond(D) = 1e6" % decay, [((decay ** i) * s, d * ((decay ** i) * s)) for i, s in enumerate(base)]
    for curv in (1e-4, 1e-8):
        # barely positive curvature: y = curv * s + a component orthogonal to s
        out = []
        for s in base:
            u = rng.random(n)
            u = u - (u @ s) / (s @ s) * s
            out.append((s, curv * s + u))
        yield "small curvature: y = %.0e s + (u orthogonal to s)" % curv, out


def main():
    rng = np.random.default_rng(0)
    n, mem, npush = 400, 10, 25
    rel = lambda a, b: float(np.linalg.norm((a.astype(LD) - b.astype(LD)).astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))
    for name, pairs in families(n, npush, rng):
        x = rng.random(n)
        ap64, S, Y, order, gamma = recursion(pairs, mem, np.float64)
        apT, _, _, _, _ = recursion(pairs, mem, LD)
        truth = apT(x)
        ref = ap64(x)
        comp, condM = compact_apply(S, Y, order, gamma, x)
        print(json.dumps({"family": name, "n": n, "mem": mem, "pushes": npush, "active_pairs": len(order),
                          "rel_err_reference_form": rel(ref, truth), "rel_err_compact_form": rel(comp, truth),
                          "rel_diff_compact_vs_reference": rel(comp, ref), "cond_middle_matrix": condM}), flush=True)


if __name__ == "__main__":
    main()
