"""ORACLE (test infrastructure only): numpy restatement of the J2 stress update that the reference's example defines
as its quadrature-point callback.

reference: examples/hypo_elastic_plasticity/J2Plasticity.jl
  MaterialState (:76-101), the callable (:97-101), assemble_strain (:103-112), estimate_stress (:114-126),
  iterate_stress! (:128-188), update_States! (:190-197).
Arrays are [n_el, n_q] (row-major) == the reference's column-major [n_q, n_el]. Voigt slots 1..6 = (1,1) (2,2) (3,3)
(2,3) (1,3) (1,2) (src/symbolics/03_Word.jl:37). The reference accumulates s_dev_2 into an uninitialised FEM_buffer
(:138,153-155); it is taken as zero here.
"""
import numpy as np

_V = {(1, 1): 0, (2, 2): 1, (3, 3): 2, (2, 3): 3, (3, 2): 3, (1, 3): 4, (3, 1): 4, (1, 2): 5, (2, 1): 5}


class MaterialState:
    def __init__(self, shape, Y_initial, lam, mu, Eb, Ep, f_res):
        self.shape, self.Y_initial = shape, float(Y_initial)
        self.lam, self.mu, self.Eb, self.Ep, self.f_res = lam, mu, Eb, Ep, f_res
        self.n_yielded = 0
        self.reset()

    def reset(self):
        z = lambda: [np.zeros(self.shape) for _ in range(6)]
        self.ep_eval, self.b_eval, self.ep, self.b = z(), z(), z(), z()
        self.Y = np.full(self.shape, self.Y_initial)
        self.Y_eval = self.Y.copy()

    def __call__(self, e11, e12, e13, e22, e23, e33):
        et = [None] * 6
        et[_V[1, 1]], et[_V[2, 2]], et[_V[3, 3]] = e11, e22, e33
        et[_V[1, 2]], et[_V[1, 3]], et[_V[2, 3]] = e12, e13, e23
        self.iterate_stress(et)
        return self.ep_eval

    def estimate_stress(self, e):
        s = [(2 * self.mu) * e[k] for k in range(6)]
        tr = e[0] + e[1] + e[2]
        for k in range(3):
            s[k] = s[k] + self.lam * tr
        return s

    def iterate_stress(self, et):
        mu, Eb, Ep = self.mu, self.Eb, self.Ep
        for k in range(6):
            self.ep_eval[k][:] = self.ep[k]
            self.b_eval[k][:] = self.b[k]
        self.Y_eval[:] = self.Y
        sig = self.estimate_stress([et[k] - self.ep_eval[k] for k in range(6)])
        s = [sig[k] - self.b_eval[k] for k in range(6)]
        skk = (s[0] + s[1] + s[2]) / 3
        for k in range(3):
            s[k] = s[k] - skk
        s2 = np.zeros(self.shape)
        for i in (1, 2, 3):
            for j in (1, 2, 3):
                s2 += s[_V[i, j]] * s[_V[i, j]]
        mag = np.sqrt(s2)
        f = np.sqrt(3 / 2) * mag - self.Y_eval
        y = f > self.f_res
        self.n_yielded = int(y.sum())
        if self.n_yielded:
            lp = np.sqrt(3 / 2) * f[y] / (3 * mu + Eb + Ep)
            for k in range(6):
                nd = s[k][y] / mag[y]
                self.ep_eval[k][y] = self.ep[k][y] + nd * lp
                self.b_eval[k][y] = self.b[k][y] + (2 / 3 * Eb) * nd * lp
            self.Y_eval[y] = self.Y[y] + (np.sqrt(2 / 3) * Ep) * lp

    def update_States(self):
        for k in range(6):
            self.ep[k][:] = self.ep_eval[k]
            self.b[k][:] = self.b_eval[k]
        self.Y[:] = self.Y_eval
