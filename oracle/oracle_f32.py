"""TEST INFRASTRUCTURE -- CPU restatement in numpy of the reference's quasi-Newton operators for T = Float32
(LBFGSOperator(Float32, n), InverseLBFGSOperator(Float32, n), LSR1Operator(Float32, n)).  Only tests/ may import this module.

The reference's code is generic in T (src/lbfgs.jl, src/lsr1.jl): with T = Float32 every broadcast statement is Float32
arithmetic (numpy float32 arrays and np.float32 scalars: IEEE single, no contraction, left-to-right inside a statement) and
`dot` / `norm` return a Float32.  Their summation order inside BLAS is unspecified; here they are accumulated in
Float64 and rounded once -- the same convention as the Float64 oracle's long-double reductions.  Parity with the CUDA kernels is
therefore to Float32 rounding of the inner products (tests use 1e-5 norm-wise), not bit for bit.

Pinned on the reference's own predicates for this path (tests/test_oracle_pinning.py::test_f32_*): identity before the first push
and element type Float32 (test/test_lbfgs.jl:162-178, test/test_lsr1.jl:74-86), the secant equation B s = y, H y = s after a push,
H·B = I, agreement with the Float64 oracle to Float32 accuracy."""
import numpy as np

F = np.float32


def _dot(a, b):
    return F(np.dot(a.astype(np.float64), b.astype(np.float64)))


class LBFGS32:
    """src/lbfgs.jl:36-56 (LBFGSData), :117-154 (inverse apply), :173-202 (forward apply), :210-288 (push!), :401-415 (reset!)"""

    def __init__(self, n, mem=5, scaling=True, inverse=False, damped=False, sigma2=0.99, sigma3=10.0):
        self.n, self.mem, self.scaling, self.inverse = n, max(mem, 1), scaling, inverse
        self.damped, self.sigma2, self.sigma3 = damped, F(sigma2), F(sigma3)
        m = self.mem
        self.s = np.zeros((m, n), F)
        self.y = np.zeros((m, n), F)
        self.ys = np.zeros(m, F)
        self.a = np.zeros((m, n), F)
        self.b = np.zeros((m, n), F)
        self.norm_b = np.zeros(m, F)
        self.alpha = np.zeros(m, F)
        self.insert = 1
        self.gamma = F(1)
        self.opnorm_upper_bound = F(1)

    def push(self, s, y):
        s, y = np.asarray(s, F), np.asarray(y, F)
        if self.damped:                                                      # :274-276
            return self.push_damped(s, y)
        ys = _dot(y, s)                                                      # :277
        if ys <= np.finfo(F).eps:                                            # :281
            return False
        return self._push_common(s, y, ys)

    def push_damped(self, s, y, alpha=None, g=None):
        """push!(op, s, y, Bs) :289-321 (forward) / push!(op, s, y, α, g, Bs) :323-357 (inverse: returns the damped y as well)"""
        s, y = np.asarray(s, F), np.asarray(y, F)
        ys = _dot(y, s)
        Bs = self.apply(s) if not self.inverse else (-F(alpha)) * np.asarray(g, F)      # :305 / :341
        sBs = _dot(s, Bs)
        one = F(1)
        theta = None
        if ys < (one - self.sigma2) * sBs:                                   # :308-314
            theta = F(self.sigma2 * sBs / (sBs - ys))
        elif ys > (one + self.sigma3) * sBs:
            theta = F(self.sigma3 * sBs / (ys - sBs))
        if theta is not None:
            y = theta * y + (one - theta) * Bs                               # :316 / :352
            ys = F(theta * ys + (one - theta) * sBs)
        self._push_common(s, y, ys)
        return y

    def diag(self):
        """diag! :379-395"""
        assert not self.inverse
        d = np.ones(self.n, F)
        if self.scaling:
            d = d / self.gamma
        for i in range(1, self.mem + 1):
            k = (self.insert + i - 2) % self.mem
            if self.ys[k] != 0:
                d = d + (self.b[k] * self.b[k] - self.a[k] * self.a[k])
        return d

    def _push_common(self, s, y, ys):
        m, ins = self.mem, self.insert - 1
        self.s[ins] = s                                                      # :220-222
        self.y[ins] = y
        self.ys[ins] = ys
        if self.scaling:                                                     # :223-227
            if self.gamma != 0:
                self.opnorm_upper_bound = F(self.opnorm_upper_bound - F(1) / self.gamma)
            self.gamma = F(ys / _dot(y, y))
            if self.gamma != 0:
                self.opnorm_upper_bound = F(self.opnorm_upper_bound + F(1) / self.gamma)
        if not self.inverse:
            self.opnorm_upper_bound = F(self.opnorm_upper_bound - self.norm_b[ins] * self.norm_b[ins])   # :231
            self.b[ins] = y / np.sqrt(ys)                                    # :232
            self.norm_b[ins] = np.sqrt(_dot(self.b[ins], self.b[ins]))       # :233
            self.opnorm_upper_bound = F(self.opnorm_upper_bound + self.norm_b[ins] * self.norm_b[ins])
            for i in range(1, m + 1):                                        # :236-250
                k = (ins + 1 + i - 1) % m
                if self.ys[k] != 0:
                    self.a[k] = self.s[k] / self.gamma
                    for j in range(1, i):
                        l = (ins + 1 + j - 1) % m
                        if self.ys[l] != 0:
                            self.a[k] = self.a[k] + _dot(self.b[l], self.s[k]) * self.b[l]
                            self.a[k] = self.a[k] - _dot(self.a[l], self.s[k]) * self.a[l]
                    self.a[k] = self.a[k] / np.sqrt(_dot(self.s[k], self.a[k]))
        self.insert = (ins + 1) % m + 1                                      # mod(insert, mem) + 1   :253
        return True

    def apply(self, x, alpha=1.0, beta=0.0, res=None):
        x = np.asarray(x, F)
        alpha, beta = F(alpha), F(beta)
        m, ins = self.mem, self.insert
        q = x.copy()
        if self.inverse:                                                     # :127-153
            for i in range(1, m + 1):
                k = (ins - i - 1) % m
                if self.ys[k] != 0:
                    self.alpha[k] = F(_dot(self.s[k], q) / self.ys[k])
                    q = q - self.alpha[k] * self.y[k]
            if self.scaling:
                q = q * self.gamma
            for i in range(1, m + 1):
                k = (ins + i - 2) % m
                if self.ys[k] != 0:
                    bb = F(self.alpha[k] - F(_dot(self.y[k], q) / self.ys[k]))
                    q = q + bb * self.s[k]
        else:                                                                # :183-196
            if self.scaling:
                q = q / self.gamma
            for i in range(1, m + 1):
                k = (ins + i - 2) % m
                if self.ys[k] != 0:
                    ax, bx = _dot(self.a[k], x), _dot(self.b[k], x)
                    q = q + (bx * self.b[k] - ax * self.a[k])
        out = alpha * q if beta == 0 else alpha * q + beta * np.asarray(res, F)   # :197-201 / :149-153
        return out.astype(F)

    def reset(self):                                                         # :401-415 (Q8: the norm bound stays)
        for arr in (self.s, self.y, self.a, self.b):
            arr[:] = 0
        self.ys[:] = 0
        self.alpha[:] = 0
        self.gamma = F(1)
        self.insert = 1


class LSR1_32:
    """src/lsr1.jl:19-34 (LSR1Data), :89-107 (apply), :119-184 (push!)"""

    def __init__(self, n, mem=5, scaling=True):
        self.n, self.mem, self.scaling = n, max(mem, 1), scaling
        m = self.mem
        self.s = np.zeros((m, n), F)
        self.y = np.zeros((m, n), F)
        self.ys = np.zeros(m, F)
        self.a = np.zeros((m, n), F)
        self.as_ = np.zeros(m, F)
        self.insert = 1
        self.gamma = F(1)
        self.opnorm_upper_bound = F(1)

    def apply(self, x, alpha=1.0, beta=0.0, res=None):
        x = np.asarray(x, F)
        alpha, beta = F(alpha), F(beta)
        m = self.mem
        q = (alpha * x) / self.gamma                                         # :92-96
        if beta != 0:
            q = q + beta * np.asarray(res, F)
        for i in range(1, m + 1):                                            # :98-105
            k = (self.insert + i - 2) % m
            if self.ys[k] != 0:
                ax = F(F(alpha * _dot(self.a[k], x)) / self.as_[k])
                q = q + ax * self.a[k]
        return q.astype(F)

    def diag(self):
        """diag! src/lsr1.jl:196-211"""
        d = np.ones(self.n, F)
        if self.scaling:
            d = d / self.gamma
        for i in range(1, self.mem + 1):
            k = (self.insert + i - 2) % self.mem
            if self.ys[k] != 0:
                d = d + (self.a[k] * self.a[k]) / self.as_[k]
        return d

    def push(self, s, y):
        s, y = np.asarray(s, F), np.asarray(y, F)
        m = self.mem
        eps = np.finfo(F).eps
        ymBs = self.apply(s, -1.0, 1.0, res=y)                               # :124-125
        ys = _dot(y, s)
        sNorm = np.sqrt(_dot(s, s))
        yy = _dot(y, y)
        well_defined = abs(_dot(ymBs, s)) >= eps + eps * np.sqrt(_dot(ymBs, ymBs)) * sNorm       # :131
        sufficient_curvature = scaling_condition = True
        if self.scaling:                                                     # :135-143
            yNorm = np.sqrt(yy)
            sufficient_curvature = abs(ys) >= eps * yNorm * sNorm
            if sufficient_curvature:
                sf = F(ys / yy)
                t = y - s / sf
                scaling_condition = np.sqrt(_dot(t, t)) >= eps * yNorm * sNorm
        if not (well_defined and sufficient_curvature and scaling_condition):
            return False
        ins = self.insert - 1
        self.s[ins] = s
        self.y[ins] = y
        self.ys[ins] = ys
        self.opnorm_upper_bound = F(1)                                       # :156
        if self.scaling:
            self.gamma = F(ys / yy)
            if self.gamma != 0:
                self.opnorm_upper_bound = F(F(1) / abs(self.gamma))
        self.insert = (ins + 1) % m + 1                                      # :163
        for i in range(1, m + 1):                                            # :166-181
            k = (self.insert + i - 2) % m
            if self.ys[k] != 0:
                self.a[k] = self.y[k] - self.s[k] / self.gamma
                for j in range(1, i):
                    l = (self.insert + j - 2) % m
                    if self.ys[l] != 0:
                        as_ = F(_dot(self.a[l], self.s[k]) / self.as_[l])
                        self.a[k] = self.a[k] - as_ * self.a[l]
                self.as_[k] = _dot(self.a[k], self.s[k])
                if self.as_[k] != 0:
                    self.opnorm_upper_bound = F(self.opnorm_upper_bound + _dot(self.a[k], self.a[k]) / abs(self.as_[k]))
        return True
