"""mg_numpy.py — NumPy restatement of the geometric-multigrid preconditioner of csrc/mg.cu.  TEST INFRASTRUCTURE ONLY.

The algorithm is this repository's own (the reference hands `AlgebraicMultigrid.jl` hierarchies to GMRES through `precs`,
docs/src/tutorials/large_systems.md:290-316; on the structured periodic grid a geometric hierarchy plays that role), so there
is no reference arithmetic to follow: this file states the V-cycle independently (whole-array `np.roll` / reshape
operations instead of per-cell kernels) and the tests hold the CUDA kernels to it to rounding, and hold GMRES with the
preconditioner to the C oracle's GMRES on the explicitly preconditioned dense operator.

Hierarchy: N -> N / p for the smallest prime factor p of N, down to one cell; coarse point I sits on fine point p I.
Transfers: multilinear interpolation fine[p I + d] = (1 - d/p) c[I] + (d/p) c[I+1] per direction and its scaled transpose
R = P' / p^dim (p = 2: full weighting 1/4, 1/2, 1/4).
Coarse operators are re-discretisations (a_c = a / p^2, reaction block from the restricted state).  Smoother: damped
block-Jacobi on the 2x2 species blocks, omega = 0.8, nu pre + nu post sweeps with nu = 2 for p = 2 and 2 p for odd p (a
coarsening by p leaves wavelengths up to 2 p h to the smoother), first sweep from the zero guess; the one-cell level is solved
exactly."""
import numpy as np

OMEGA = 0.8


def sweeps(ratio):
    return 2 if ratio == 2 else 2 * ratio


def _factor(n):
    p = 2
    while p * p <= n:
        if n % p == 0:
            return p
        p += 1
    return n


class Level:
    def __init__(self, N, dim, a, A, state):
        self.N, self.dim, self.a, self.A = N, dim, a, A
        shape = (N,) * dim
        self.u = state[:N ** dim].reshape(shape, order="F")
        self.v = state[N ** dim:].reshape(shape, order="F")
        self.ratio = _factor(N) if N > 1 else 0
        if self.ratio > 7 and self.ratio == N and N ** dim > 4096:
            self.ratio = 0

    def split(self, x):
        n = self.N ** self.dim
        shape = (self.N,) * self.dim
        return x[:n].reshape(shape, order="F"), x[n:].reshape(shape, order="F")

    @staticmethod
    def join(a, b):
        return np.concatenate([a.ravel(order="F"), b.ravel(order="F")])

    def lap(self, w):
        s = -2.0 * self.dim * w
        for ax in range(self.dim):
            s = s + np.roll(w, 1, ax) + np.roll(w, -1, ax)
        return s

    def apply(self, x):
        p, q = self.split(x)
        uv2, uu = 2 * self.u * self.v, self.u * self.u
        return self.join(self.a * self.lap(p) + (uv2 - (self.A + 1.0)) * p + uu * q, self.a * self.lap(q) + (self.A - uv2) * p - uu * q)

    def dinv(self, r):
        r0, r1 = self.split(r)
        uv2, uu = 2 * self.u * self.v, self.u * self.u
        ld = -2.0 * self.dim * self.a
        d00, d01, d10, d11 = ld + (uv2 - (self.A + 1.0)), uu, self.A - uv2, ld - uu
        det = d00 * d11 - d01 * d10
        return self.join((d11 * r0 - d01 * r1) / det, (d00 * r1 - d10 * r0) / det)

    def restrict_field(self, w):
        """R = P' / p^dim per direction: coarse[I] = sum_{|d| < p} (1 - |d|/p)/p * fine[p I + d]   (p = 2: 1/4, 1/2, 1/4)."""
        p = self.ratio
        for ax in range(self.dim):
            acc = np.zeros_like(w)
            for d in range(-(p - 1), p):
                acc = acc + ((1.0 - abs(d) / p) / p) * np.roll(w, -d, ax)
            w = np.take(acc, np.arange(0, self.N, p), axis=ax)
        return w

    def prolong_field(self, c):
        """Vertex-centred multilinear interpolation: fine[p I + d] = (1 - d/p) c[I] + (d/p) c[I + 1] per direction."""
        p, dim = self.ratio, self.dim
        for ax in range(dim):
            nxt = np.roll(c, -1, ax)
            fine_shape = list(c.shape)
            fine_shape[ax] *= p
            f = np.empty(fine_shape)
            for d in range(p):
                sl = [slice(None)] * dim
                sl[ax] = slice(d, None, p)
                f[tuple(sl)] = (1.0 - d / p) * c + (d / p) * nxt
            c = f
        return c

    def restrict(self, x):
        a, b = self.split(x)
        return self.join(self.restrict_field(a), self.restrict_field(b))

    def prolong(self, xc):
        n = (self.N // self.ratio) ** self.dim
        shape = (self.N // self.ratio,) * self.dim
        return self.join(self.prolong_field(xc[:n].reshape(shape, order="F")), self.prolong_field(xc[n:].reshape(shape, order="F")))


class Multigrid:
    def __init__(self, N, dim, u, A=3.4, alpha=10.0):
        self.levels = []
        a = alpha * (N - 1) ** 2
        state = np.asarray(u, dtype=float)
        while True:
            L = Level(N, dim, a, A, state)
            self.levels.append(L)
            if L.ratio == 0:
                break
            state = L.restrict(state)
            N //= L.ratio
            a /= L.ratio ** 2

    def sizes(self):
        return [L.N for L in self.levels]

    def vcycle(self, b, l=0):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            x = (1.0 if L.N == 1 else OMEGA) * L.dinv(b)
            if L.N > 1:  # a grid that cannot be coarsened (large prime N): smoothing only
                for _ in range(7):
                    x = x + OMEGA * L.dinv(b - L.apply(x))
            return x
        nu = sweeps(L.ratio)
        x = OMEGA * L.dinv(b)
        for _ in range(nu - 1):
            x = x + OMEGA * L.dinv(b - L.apply(x))
        x = x + L.prolong(self.vcycle(L.restrict(b - L.apply(x)), l + 1))
        for _ in range(nu):
            x = x + OMEGA * L.dinv(b - L.apply(x))
        return x

    def dense_inverse(self):
        n = 2 * self.levels[0].N ** self.levels[0].dim
        M = np.empty((n, n))
        e = np.zeros(n)
        for c in range(n):
            e[c] = 1.0
            M[:, c] = self.vcycle(e)
            e[c] = 0.0
        return M
