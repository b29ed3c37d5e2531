"""ORACLE: time stepping, Newton loop and Krylov solvers (numpy/scipy restatement).

reference: src/solver/04_Time_Domain.jl:1-80                       (GeneralAlpha, update_OneStep!)
           src/solver/linear_solver/02_Preconditioner.jl:32-148    (iterative_Solve!, Pr_Jacobi!)
           src/solver/linear_solver/04_IDRs.jl:1-95                (modify_Omega, idrs!)
           src/solver/linear_solver/03_BiCGstabl.jl:18-96          (bicgstabl_GS!)
Random shadow vectors come from a seeded numpy generator (the reference uses unseeded cuRAND),
so only solver-tolerance agreement of solutions is meaningful (SURVEY.md §7 hard part (f)).
"""
import numpy as np

from . import assembly as asm


def normalized_norm(x):
    return np.linalg.norm(x) / np.sqrt(len(x))


def update_Time(dom):                                   # :10-18
    gf = dom.globalfield
    gf.t += gf.dt
    L = gf.max_time_level
    prod_gamma = [float(np.prod(dom.gamma_params[:i])) for i in range(L + 1)]
    dt_params = [gf.dt ** i for i in range(L + 1)]
    dom.beta_params = [1.0 / (g * d) for g, d in zip(prod_gamma, dt_params)]
    dom.K_params = [a * b for a, b in zip(dom.alpha_params[:L + 1], dom.beta_params)]


def initialize_dx(dom):                                 # :20-30
    gf = dom.globalfield
    n = gf.basicfield_size
    gf.dx[:] = 0.0
    for l in range(gf.max_time_level, 0, -1):
        lo, hi = slice((l - 1) * n, l * n), slice(l * n, (l + 1) * n)
        gf.dx[lo] = gf.dt * (gf.x[hi] + dom.gamma_params[l - 1] * gf.dx[hi])


def update_dx(dom, delta_x):                            # :32-39
    gf = dom.globalfield
    n = gf.basicfield_size
    for l in range(gf.max_time_level + 1):
        gf.dx[l * n:(l + 1) * n] += dom.beta_params[l] * delta_x


def update_x_star(dom):                                 # :41-49
    gf = dom.globalfield
    n = gf.basicfield_size
    gf.x_star[:] = gf.x
    for l in range(gf.max_time_level + 1):
        gf.x_star[l * n:(l + 1) * n] += dom.alpha_params[l] * gf.dx[l * n:(l + 1) * n]


def update_OneStep(dom, max_iter=4, log=None):          # :59-80
    gf = dom.globalfield
    update_Time(dom)
    initialize_dx(dom)
    asm.K_linear_func(dom)
    counter = -1
    history = []
    while True:
        update_x_star(dom)
        asm.K_nonlinear_func(dom)
        res = normalized_norm(gf.residue)
        counter += 1
        history.append(res)
        if log:
            log(f"step {counter} residue = {res}")
        if res < gf.converge_tol or counter > max_iter:
            break
        delta_x = dom.linear_solver(dom)
        update_dx(dom, -delta_x)
    gf.x += gf.dx
    return history


# ---------------------------------------------------------------------------------------------
def Pr_Jacobi(A):
    """Pr_Jacobi! with Jacobi_By_Diagonal + Mat_Div_Jacobi (02_Preconditioner.jl:103-148); A: cpath.CsrOperator, modified in place."""
    n = A.shape[0]
    jac = np.ones(n)
    rows = np.repeat(np.arange(n), np.diff(A.ptr))
    cols = A.col - 1
    dmask = cols == rows
    jac[rows[dmask]] = np.abs(A.data[dmask])
    A.data /= jac[cols]
    return jac


def Pr_Jacobi_column(A):
    """Pr_Jacobi!(normalized_by_column = true): Jacobi2_By_Colomn + sqrt (:110-113,122-129)."""
    jac = np.sqrt(np.bincount(A.col - 1, A.data ** 2, minlength=A.shape[0]))
    A.data /= jac[A.col - 1]
    return jac


class Pl_Jacobi:
    """Pl_Jacobi (:150-166) + _JacobiP (:91-98): b ./= jac_vec in place; by diagonal or by row norm (Jacobi_By_Row :169-176)."""
    def __init__(self, A, normalized_by_row=False):
        n = A.shape[0]
        rows = np.repeat(np.arange(n), np.diff(A.ptr))
        if normalized_by_row:
            self.jac = np.sqrt(np.bincount(rows, A.data ** 2, minlength=n))
        else:
            self.jac = np.ones(n)
            d = (A.col - 1) == rows
            self.jac[rows[d]] = np.abs(A.data[d])

    def __call__(self, b):
        b /= self.jac
        return b


class Pl_ILU:
    """Pl_ILU(A) = ilu02!(copy(A)) + UnitLowerTriangular / UpperTriangular solves (02_Preconditioner.jl:179-194): zero-fill
    incomplete LU of the CSR matrix in ITS row order without pivoting (what cuSPARSE csrilu02 computes), restated with plain loops
    (small systems only)."""

    def __init__(self, A):
        import scipy.sparse as sps
        n = A.shape[0]
        M = sps.csr_matrix((A.data.copy(), A.col - 1, A.ptr - 1), shape=(n, n)) if hasattr(A, "ptr") else sps.csr_matrix(A, copy=True)
        M.sort_indices()
        ptr, col, val = M.indptr, M.indices, M.data
        diag = np.full(n, -1)
        for i in range(n):
            where = {int(col[p]): p for p in range(ptr[i], ptr[i + 1])}
            for p in range(ptr[i], ptr[i + 1]):              # IKJ over the lower entries in ascending column order
                k = int(col[p])
                if k >= i:
                    break
                val[p] /= val[diag[k]]
                for q in range(diag[k] + 1, ptr[k + 1]):
                    t = where.get(int(col[q]))
                    if t is not None:
                        val[t] -= val[p] * val[q]
            diag[i] = where[i]
        self.L = sps.tril(M, -1).tocsr() + sps.identity(n, format="csr")
        self.U = sps.triu(M, 0).tocsr()
        self.M = M

    def __call__(self, b):
        import scipy.sparse.linalg as spl
        y = spl.spsolve_triangular(self.L, b, lower=True, unit_diagonal=True)
        b[:] = spl.spsolve_triangular(self.U, y, lower=False)
        return b


def _mix64(z):
    """splitmix64 finaliser on uint64 arrays (the node hash of the library's elimination order)."""
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return z ^ (z >> np.uint64(31))


class Pl_ILU_block:
    """Restatement of what the LIBRARY does for Pl_ILU (csrc/mfb_ilu.cu) -- not of cuSPARSE: the same zero-fill incomplete
    factorisation, (L U)_ij = A_ij on the pattern, but on the nv x nv node blocks and in an elimination order chosen for
    parallelism. Nodes get the priority hash mix64(id + 0x5bd1e995) of their 0-based id, are coloured greedily in descending
    priority (colour = smallest one no higher-priority neighbour carries) and eliminated colour class by colour class, ties by
    the hash; unit lower factor L_ik = A_ik U_kk^-1, no pivoting. `levels` is the dependency depth of that order (what the
    library reports). Vectors are variable-major ([nv][N]) like GlobalField.x. Plain loops: small systems only."""

    def __init__(self, A, nv, order="color"):
        import scipy.sparse as sps
        n = A.shape[0]
        A = sps.csr_matrix((A.data.copy(), A.col - 1, A.ptr - 1), shape=(n, n)) if hasattr(A, "ptr") else sps.csr_matrix(A)
        N = n // nv
        self.nv, self.N = nv, N
        # node graph and blocks: block (i, j)[v][w] = A[v N + i, w N + j]
        coo = A.tocoo()
        blocks = {}
        for r, c, val in zip(coo.row, coo.col, coo.data):
            i, v, j, w = r % N, r // N, c % N, c // N
            blocks.setdefault((i, j), np.zeros((nv, nv)))[v, w] = val
        nbr = [[] for _ in range(N)]
        for (i, j) in blocks:
            if i != j:
                nbr[i].append(j)
        with np.errstate(over="ignore"):
            key = _mix64(np.arange(N, dtype=np.uint64) + np.uint64(0x5bd1e995))
        if order == "color":
            color = np.full(N, -1)
            for i in np.argsort(key, kind="stable")[::-1]:           # descending priority
                used = {color[j] for j in nbr[i] if color[j] >= 0}
                c = 0
                while c in used:
                    c += 1
                color[i] = c
            key = (color.astype(np.uint64) << np.uint64(56)) | (key >> np.uint64(8))
            self.n_colors = int(color.max()) + 1
        seq = np.argsort(key, kind="stable")                          # elimination order
        pos = np.empty(N, np.int64)
        pos[seq] = np.arange(N)
        self.pos, self.seq = pos, seq
        F = {k: b.copy() for k, b in blocks.items()}
        dinv = [None] * N
        level = np.zeros(N, np.int64)
        for i in seq:
            low = sorted((j for j in nbr[i] if pos[j] < pos[i]), key=lambda j: pos[j])
            for k in low:
                L = F[(i, k)] @ dinv[k]
                F[(i, k)] = L
                for j in nbr[k] + [k]:
                    if j != k and pos[j] > pos[k] and (i, j) in F:
                        F[(i, j)] = F[(i, j)] - L @ F[(k, j)]
                level[i] = max(level[i], level[k] + 1)
            dinv[i] = np.linalg.inv(F[(i, i)])
        self.F, self.dinv, self.nbr, self.blocks = F, dinv, nbr, blocks
        self.levels = int(level.max()) + 1

    def product_defect(self):
        """max |(L U - A)_ij| over the block pattern / max |A_ij|."""
        pos, F, worst, amax = self.pos, self.F, 0.0, 0.0
        I = np.eye(self.nv)
        for (i, j), a in self.blocks.items():
            s = np.zeros_like(a)
            for k in set(self.nbr[i] + [i]) & set(self.nbr[j] + [j]):
                if pos[k] > min(pos[i], pos[j]):
                    continue
                Lik = I if k == i else (F[(i, k)] if pos[k] < pos[i] else None)
                Ukj = F.get((k, j)) if pos[j] >= pos[k] else None
                if Lik is not None and Ukj is not None:
                    s = s + Lik @ Ukj
            worst, amax = max(worst, np.abs(s - a).max()), max(amax, np.abs(a).max())
        return worst / amax

    def __call__(self, b):
        nv, N, pos, F = self.nv, self.N, self.pos, self.F
        v = np.asarray(b, dtype=float).reshape(nv, N).T.copy()        # [N][nv]
        for i in self.seq:                                            # forward, unit L
            for k in self.nbr[i]:
                if pos[k] < pos[i]:
                    v[i] -= F[(i, k)] @ v[k]
        for i in self.seq[::-1]:                                      # backward
            acc = v[i].copy()
            for j in self.nbr[i]:
                if pos[j] > pos[i]:
                    acc -= F[(i, j)] @ v[j]
            v[i] = self.dinv[i] @ acc
        b[:] = v.T.ravel()
        return b


def Identity(b):
    return b


def iterative_Solve(dom, Sv_func, max_pass=4, seed=1234, log=None, Pr_func=Pr_Jacobi, Pl_func=None, **kw):
    """iterative_Solve! (:32-76)."""
    gf = dom.globalfield
    A = asm.csr_operator(gf)          # K_vals = K_total[K_val_ids] (:35): a fresh copy, as in the reference
    jac = Pr_func(A) if Pr_func is not None else np.ones(A.shape[0])
    Pl = Pl_func(A) if Pl_func is not None else Identity
    b = gf.residue
    r = b.copy()
    x = np.zeros_like(b)
    rng = np.random.default_rng(seed)
    pass_number = 1
    tol_factor = 1.0
    iters = []
    while True:
        it = Sv_func(x, A, b, r, tol=tol_factor * gf.converge_tol, rng=rng, Pl=Pl, **kw)
        iters.append(it)
        r[:] = b - A @ x
        res = normalized_norm(r)
        if Pl_func is not None:       # :50-53
            tol_factor = min(normalized_norm(Pl(r)) / res, 1.0)
        if log:
            log(f"pass {pass_number} with res = {res} iter = {it}.")
        if res < gf.converge_tol or pass_number >= max_pass:
            break
        pass_number += 1
    dom.last_solve = dict(passes=pass_number, iters=iters, res=res)
    return x / jac


def modify_Omega(v1, v2):                               # 04_IDRs.jl:1-8
    angle = np.sqrt(2.0) / 2
    n1, n2 = np.linalg.norm(v1), np.linalg.norm(v2)
    d = np.dot(v1, v2)
    rho = abs(d / (n1 * n2))
    omega = d / (n1 * n1)
    return omega * angle / rho if rho < angle else omega


def idrs(x, A, b, r, tol, maxiter, s=4, rng=None, Pl=Identity, **kw):    # 04_IDRs.jl:26-95
    r[:] = Pl(b - A @ x)
    if normalized_norm(r) <= tol:
        return 0
    it = 1
    n = len(b)
    P = [rng.random(n) for _ in range(s)]
    U = [np.zeros(n) for _ in range(s)]
    G = [np.zeros(n) for _ in range(s)]
    M = np.eye(s)
    f = np.zeros(s)
    omega = 1.0
    while True:
        for i in range(s):
            f[i] = np.dot(P[i], r)
        for k in range(s):
            c = np.linalg.solve(np.tril(M[k:, k:]), f[k:])
            V = c[0] * G[k]
            Q = c[0] * U[k]
            for i in range(k + 1, s):
                V += c[i - k] * G[i]
                Q += c[i - k] * U[i]
            V = r - V
            U[k] = Q + omega * V
            G[k] = Pl(A @ U[k])
            for i in range(k):
                alpha = np.dot(P[i], G[k]) / M[i, i]
                G[k] -= alpha * G[i]
                U[k] -= alpha * U[i]
            for i in range(k, s):
                M[i, k] = np.dot(P[i], G[k])
            beta = f[k] / M[k, k]
            x += beta * U[k]
            r -= beta * G[k]
            if normalized_norm(r) <= tol or it >= maxiter:
                return it
            f[k + 1:] -= beta * M[k + 1:, k]
            it += 1
        Ar = Pl(A @ r)
        omega = modify_Omega(Ar, r)
        x += omega * r
        r -= omega * Ar
        if normalized_norm(r) <= tol or it >= maxiter:
            return it
        it += 1


def fem_rand_seeded(n_nodes, nv, seed, stream_id):
    """The library's seeded stand-in for FEM_rand (unseeded cuRAND in the reference, 04_GPU_Utils.jl:22) in REFERENCE layout:
    entry g + v N gets the counter-based value of (seed, stream, g nv + v) -- splitmix64, see k_rand in csrc/mfb_krylov.cu.
    Lets a test run the oracle and the CUDA path with the SAME shadow vectors."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)

    def sm64(z):
        with np.errstate(over="ignore"):
            z = (z + np.uint64(0x9E3779B97F4A7C15)) & M
            z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
            z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
            return z ^ (z >> np.uint64(31))

    with np.errstate(over="ignore"):
        base = sm64(np.uint64(seed) ^ (np.uint64(stream_id) * np.uint64(0xD1B54A32D192ED03) & M))
        g = np.arange(n_nodes, dtype=np.uint64)
        out = np.empty(n_nodes * nv)
        for v in range(nv):
            h = sm64((base + g * np.uint64(nv) + np.uint64(v)) & M)
            out[v * n_nodes:(v + 1) * n_nodes] = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return out


def idrs_original(x, A, b, r, tol, maxiter, s=4, rng=None, Pl=Identity, P=None, **kw):   # 04_IDRs.jl:97-169
    """"not used, not exploiting orthogonality" in the reference; restated as written, including the k == 0 branch that
    overwrites r with Pl(A Q) (:147-148)."""
    r[:] = Pl(b - A @ x)
    if normalized_norm(r) <= tol:
        return 0
    it = 1
    n = len(b)
    P = [rng.random(n) for _ in range(s)] if P is None else P
    U = [np.zeros(n) for _ in range(s)]
    G = [np.zeros(n) for _ in range(s)]
    M = np.zeros((s, s))
    f = np.zeros(s)
    omega = 1.0
    for k in range(s):
        U[k][:] = r
        G[k][:] = Pl(A @ r)
        omega = modify_Omega(G[k], r)
        x += omega * U[k]
        r -= omega * G[k]
        for i in range(s):
            M[i, k] = np.dot(P[i], G[k])
    while True:
        if normalized_norm(r) <= tol or it >= maxiter:
            return it
        it += 1
        for k in range(s + 1):
            for i in range(s):
                f[i] = np.dot(P[i], r)
            c = np.linalg.solve(M, f)
            V = c[0] * G[0]
            Q = c[0] * U[0]
            for i in range(1, s):
                V += c[i] * G[i]
                Q += c[i] * U[i]
            V = r - V
            if k == 0:
                Ar = Pl(A @ V)
                omega = modify_Omega(Ar, V)
                Q += omega * V
                x += Q
                r[:] = Pl(A @ Q)
            else:
                U[k - 1][:] = Q + omega * V
                G[k - 1][:] = Pl(A @ U[k - 1])
                x += U[k - 1]
                r -= G[k - 1]
                for i in range(s):
                    M[i, k - 1] = np.dot(P[i], G[k - 1])


def bicgstabl_GS(x, A, b, r, tol, maxiter, s=2, rng=None, Pl=Identity, **kw):   # 03_BiCGstabl.jl:18-96
    r[:] = Pl(b - A @ x)
    if normalized_norm(r) <= tol:
        return 0
    it = 1
    n = len(b)
    gam, gamp, gampp, sig = np.zeros(s), np.zeros(s), np.zeros(s), np.zeros(s)
    tau = np.zeros((s, s))
    omega = rho0 = 1.0
    alpha = 0.0
    r_shadow = rng.random(n)
    R = [r] + [np.zeros(n) for _ in range(s)]       # R[1] aliases r (:40)
    U = [np.zeros(n) for _ in range(s + 1)]
    while True:
        rho0 *= -omega
        for j in range(s):
            rho1 = np.dot(r_shadow, R[j])
            beta = alpha * rho1 / rho0
            rho0 = rho1
            for i in range(j + 1):
                U[i][:] = R[i] - beta * U[i]
            U[j + 1][:] = Pl(A @ U[j])
            alpha = rho0 / np.dot(r_shadow, U[j + 1])
            for i in range(j + 1):
                R[i] -= alpha * U[i + 1]
            R[j + 1][:] = Pl(A @ R[j])
            x += alpha * U[0]
        for j in range(s):
            for i in range(j):
                tau[i, j] = np.dot(R[i + 1], R[j + 1]) / sig[i]
                R[j + 1] -= tau[i, j] * R[i + 1]
            sig[j] = np.dot(R[j + 1], R[j + 1])
            gamp[j] = np.dot(R[0], R[j + 1]) / sig[j]
        gam[s - 1] = gamp[s - 1]
        omega = gam[s - 1]
        for j in range(s - 2, -1, -1):
            gam[j] = gamp[j] - np.dot(tau[j, j + 1:s], gam[j + 1:s])
        for j in range(s - 1):
            gampp[j] = gam[j + 1] + np.dot(tau[j, j + 1:s - 1], gam[j + 2:s])
        x += gam[0] * R[0]
        R[0] -= gamp[s - 1] * R[s]
        U[0] -= gam[s - 1] * U[s]
        for j in range(s - 1):
            U[0] -= gam[j] * U[j + 1]
            x += gampp[j] * R[j + 1]
            R[0] -= gamp[j] * R[j + 1]
        it += s
        if normalized_norm(R[0]) <= tol or it >= maxiter:
            return it


# ---------------------------------------------------------------------------------------------
# the remaining exported Krylov methods; Pl(.) follows every product as in the reference
def _tmul(A, x):
    """tmul! = A' x (06_LSQR.jl:25,42) through scipy on the oracle's CSR arrays."""
    import scipy.sparse as sps
    n = A.shape[0]
    return sps.csr_matrix((A.data, A.col - 1, A.ptr - 1), shape=(n, n)).T @ x


def bicgstabl(x, A, b, r, tol, maxiter, s=2, rng=None, Pl=Identity, **kw):   # 03_BiCGstabl.jl:98-162
    r[:] = Pl(b - A @ x)
    if np.linalg.norm(r) <= tol:
        return 0
    it = 1
    n = len(b)
    omega = rho0 = 1.0
    alpha = 0.0
    r_shadow = rng.random(n)
    R = [r] + [np.zeros(n) for _ in range(s)]
    U = [np.zeros(n) for _ in range(s + 1)]
    M = np.zeros((s + 1, s + 1))
    while True:
        rho0 *= -omega
        for j in range(s):
            rho1 = np.dot(r_shadow, R[j])
            beta = alpha * rho1 / rho0
            rho0 = rho1
            for i in range(j + 1):
                U[i][:] = R[i] - beta * U[i]
            U[j + 1][:] = Pl(A @ U[j])
            alpha = rho0 / np.dot(r_shadow, U[j + 1])
            for i in range(j + 1):
                R[i] -= alpha * U[i + 1]
            R[j + 1][:] = Pl(A @ R[j])
            x += alpha * U[0]
        for j in range(s + 1):
            for i in range(j + 1):
                M[i, j] = M[j, i] = np.dot(R[i], R[j])
        gam = np.linalg.solve(M[1:, 1:], M[1:, 0])
        for i in range(s):
            U[0] -= gam[i] * U[i + 1]
            x += gam[i] * R[i]
        for i in range(s):
            R[0] -= gam[i] * R[i + 1]
        omega = gam[s - 1]
        it += s
        if normalized_norm(R[0]) <= tol or it >= maxiter:
            return it


def gmres(x, A, b, r, tol, maxiter, s=20, Pl=Identity, **kw):                 # 05_GMRES.jl:46-101
    r[:] = Pl(b - A @ x)
    if normalized_norm(r) <= tol:
        return 0
    it = 1
    n = len(b)
    Q = [np.zeros(n) for _ in range(s + 1)]
    while True:
        H = np.zeros((s + 1, s))
        y = np.zeros(s + 1)
        y[0] = r_norm = np.linalg.norm(r)
        Q[0][:] = r / r_norm
        for i in range(1, s + 1):
            Q[i][:] = Pl(A @ Q[i - 1])
            for j in range(i):
                H[j, i - 1] = np.dot(Q[j], Q[i])
                Q[i] -= H[j, i - 1] * Q[j]
            H[i, i - 1] = np.linalg.norm(Q[i])
            Q[i] /= H[i, i - 1]
        yy = np.linalg.lstsq(H, y, rcond=None)[0]        # Hessenberg(H, y): least squares by Givens rotations (:7-37)
        for i in range(s):
            x += Q[i] * yy[i]
        it += s
        r[:] = Pl(b - A @ x)
        if normalized_norm(r) <= tol or it > maxiter:
            return it


def cgs(x, A, b, r, tol, maxiter, Pl=Identity, **kw):                         # 07_CGS.jl:10-50
    r[:] = Pl(b - A @ x)
    if normalized_norm(r) <= tol:
        return 0
    it = 1
    r0 = r.copy()
    n = len(b)
    rho = 1.0
    u, p, sv, v = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)      # FEM_buffer in the reference (uninitialised)
    while True:
        rhobar = rho
        rho = np.dot(r, r0)
        beta = rho / rhobar
        sv[:] = r + beta * p
        u[:] = sv + beta * (p + beta * u)
        v[:] = Pl(A @ u)
        alpha = rho / np.dot(v, r0)
        p[:] = sv - alpha * v
        x += alpha * (p + sv)
        r[:] = Pl(b - A @ x)
        it += 1
        if normalized_norm(r) <= tol or it > maxiter:
            return it


def cgs2(x, A, b, r, tol, maxiter, rng=None, Pl=Identity, **kw):              # 07_CGS.jl:52-105
    r[:] = Pl(b - A @ x)
    if normalized_norm(r) <= tol:
        return 0
    it = 1
    r0 = r.copy()
    n = len(b)
    s0 = rng.random(n)
    alpha = alphabar = sigma = sigmabar = 1.0
    u, w, sv, v, t, c = (np.zeros(n) for _ in range(6))
    while True:
        rho = np.dot(r, r0)
        beta = 1 / alphabar * rho / sigma
        v[:] = r + beta * u
        rhobar = np.dot(r, s0)
        betabar = 1 / alpha * rhobar / sigmabar
        t[:] = r + betabar * sv
        w[:] = t + beta * (u + betabar * w)
        c[:] = Pl(A @ w)
        sigma = np.dot(c, r0)
        alpha = rho / sigma
        sv[:] = t - alpha * c
        sigmabar = np.dot(c, s0)
        alphabar = rhobar / sigmabar
        u[:] = v - alphabar * c
        x += alpha * v + alphabar * sv
        r[:] = Pl(b - A @ x)
        it += 1
        if normalized_norm(r) <= tol or it > maxiter:
            return it


def tfqmr(x, A, b, r, tol, maxiter, checkiter=200, Pl=Identity, **kw):        # 08_QMR.jl:3-76
    r[:] = Pl(b - A @ x)
    if normalized_norm(r) <= tol:
        return 0
    it = 1
    n = len(b)
    r0, rc, p, u = r.copy(), r.copy(), r.copy(), r.copy()
    q, d = np.zeros(n), np.zeros(n)
    v = Pl(A @ p)
    r_norm = tau = np.linalg.norm(r)
    rho = np.dot(r, r)
    theta = eta = 0.0
    while True:
        alpha = rho / np.dot(v, r0)
        q[:] = u - alpha * v
        v[:] = u + q
        rc -= alpha * Pl(A @ v)
        r_norm_old, r_norm = r_norm, np.linalg.norm(rc)
        d[:] = u + (theta ** 2 * eta / alpha) * d
        theta = r_norm_old / tau
        c = 1 / np.sqrt(1 + theta ** 2)
        tau *= theta * c
        eta = c ** 2 * alpha
        x += eta * d
        d[:] = q + (theta ** 2 * eta / alpha) * d
        theta = np.sqrt(r_norm * r_norm_old) / tau
        c = 1 / np.sqrt(1 + theta ** 2)
        tau *= theta * c
        eta = c ** 2 * alpha
        x += eta * d
        rhobar, rho = rho, np.dot(rc, r0)
        beta = rho / rhobar
        u[:] = rc + beta * q
        p[:] = u + beta * (q + beta * p)
        v[:] = Pl(A @ p)
        it += 1
        if it > maxiter:
            return it
        if it % checkiter == 0:
            r[:] = Pl(b - A @ x)
            if normalized_norm(r) <= tol:
                return it


def lsqr(x, A, b, r, tol, maxiter, Pl=Identity, **kw):                        # 06_LSQR.jl:10-73
    r[:] = Pl(b - A @ x)
    if normalized_norm(r) <= tol:
        return 0
    it = 1
    u = r.copy()
    beta = np.linalg.norm(u)
    u /= beta
    v = Pl(_tmul(A, u))
    alpha = np.linalg.norm(v)
    if alpha != 0:
        v /= alpha
    w = v.copy()
    phibar, rhobar = beta, alpha
    while True:
        u[:] = Pl(A @ v) - alpha * u
        beta = np.linalg.norm(u)
        if beta != 0:
            u /= beta
            v[:] = Pl(_tmul(A, u)) - beta * v
            alpha = np.linalg.norm(v)
            if alpha != 0:
                v /= alpha
        rho = np.sqrt(rhobar ** 2 + beta ** 2)
        c, sn = rhobar / rho, beta / rho
        theta = sn * alpha
        rhobar = -c * alpha
        phi = c * phibar
        phibar = sn * phibar
        x += (phi / rho) * w
        w[:] = v - (theta / rho) * w
        it += 1
        r[:] = Pl(b - A @ x)
        if normalized_norm(r) <= tol or it > maxiter:
            return it
