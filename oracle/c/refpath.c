/* ORACLE / CPU baseline (test infrastructure only -- never linked into the product path).
 *
 * OpenMP restatement of the reference's three element kernels and of its SpMV, with the
 * reference's loop nests and data layout (element is the slowest index of every table, one
 * "thread" per element, atomics -> omp atomic):
 *   _Var_Basic   reference src/solver/06_FEM_Kernel.jl:1-13
 *   _Kval_Basic  reference src/solver/06_FEM_Kernel.jl:28-45
 *   _Res_Basic   reference src/solver/06_FEM_Kernel.jl:65-79
 *   mul!         reference src/misc/04_GPU_Utils.jl:131 (CSR, Int32 indices, base 1)
 * Table layout: itp_vals[host][slot][a][q] with q fastest == the reference's column-major
 * integral_vals[q, a, sd..., host] restricted to the 4 slots of max_sd_order = 1.
 * Index arrays are 1-based like the reference.
 */
#include <stdint.h>
#include <omp.h>

void ora_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ora_get_threads(void) { return omp_get_max_threads(); }

/* target[t][q] = sum_a itp[host(t)][slot][a][q] * x[cp[a][el(t)] + shift]   (cp: [a][n_cols] 1-based) */
void ora_var_basic(int nq, int na, int64_t nsel, const double *itp, int slot, int64_t shift,
                   const int32_t *cp, int64_t cp_cols, const int64_t *el, const int64_t *host,
                   const double *x, double *target) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < nsel; ++t) {
        const double *iv = itp + ((host[t] * 4 + slot) * na) * (int64_t)nq;
        double *out = target + t * nq;
        for (int q = 0; q < nq; ++q) out[q] = 0.0;
        for (int a = 0; a < na; ++a) {
            double xv = x[cp[a * cp_cols + el[t]] - 1 + shift];
            for (int q = 0; q < nq; ++q) out[q] += iv[a * nq + q] * xv;
        }
    }
}

/* K[sid[a][b][el] + shift] += sum_q itp[..dslot][a][q] * itp[..bslot][b][q] * vals[t][q]   (sid: [a][b][n_cols] 1-based) */
void ora_kval_basic(int nq, int na, int64_t nsel, const double *itp, int dslot, int bslot, const double *vals,
                    const int32_t *sid, int64_t sid_cols, int64_t shift, const int64_t *el, const int64_t *host,
                    double *K) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < nsel; ++t) {
        const double *da = itp + ((host[t] * 4 + dslot) * na) * (int64_t)nq;
        const double *db = itp + ((host[t] * 4 + bslot) * na) * (int64_t)nq;
        const double *v = vals + t * nq;
        for (int a = 0; a < na; ++a)
            for (int b = 0; b < na; ++b) {
                double sum = 0.0;
                for (int q = 0; q < nq; ++q) sum += da[a * nq + q] * db[b * nq + q] * v[q];
                int64_t id = (int64_t)sid[((int64_t)a * na + b) * sid_cols + el[t]] - 1 + shift;
#pragma omp atomic
                K[id] += sum;
            }
    }
}

/* residue[cp[a][el] + shift] += sum_q itp[..slot][a][q] * vals[t][q] */
void ora_res_basic(int nq, int na, int64_t nsel, const double *itp, int slot, const double *vals, int64_t shift,
                   const int32_t *cp, int64_t cp_cols, const int64_t *el, const int64_t *host, double *residue) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < nsel; ++t) {
        const double *iv = itp + ((host[t] * 4 + slot) * na) * (int64_t)nq;
        const double *v = vals + t * nq;
        for (int a = 0; a < na; ++a) {
            double sum = 0.0;
            for (int q = 0; q < nq; ++q) sum += iv[a * nq + q] * v[q];
            int64_t id = (int64_t)cp[a * cp_cols + el[t]] - 1 + shift;
#pragma omp atomic
            residue[id] += sum;
        }
    }
}

/* y = A x, CSR with 1-based ptr/col (the layout GlobalField holds) */
void ora_spmv_csr(int64_t n, const int32_t *ptr, const int32_t *col, const double *val, const double *x, double *y) {
#pragma omp parallel for schedule(static, 256)
    for (int64_t r = 0; r < n; ++r) {
        double s = 0.0;
        for (int64_t k = ptr[r] - 1; k < ptr[r + 1] - 1; ++k) s += val[k] * x[col[k] - 1];
        y[r] = s;
    }
}
