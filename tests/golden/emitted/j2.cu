#include "mfb_skeleton.cuh"

struct F_b0_lin {
  static constexpr int NV = 3, NA = 20, NQ = 27, L1 = 3, BOUNDARY = 0, LINEAR = 1, NW = 0, NCW = 0, NC = 0, HAS_RES = 0, HAS_K = 1, TPB = 64, NSD = 4, KS = 4, ND = 144, NTC = 20, CG = 1, W = 2, LPW = 30, SMEM = 54224, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 4;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {0, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    D[13] += (100000.0) * A.Kp[0];
    D[26] += (50000.0) * A.Kp[0];
    D[29] += (50000.0) * A.Kp[0];
    D[39] += (50000.0) * A.Kp[0];
    D[45] += (50000.0) * A.Kp[0];
    D[62] += (50000.0) * A.Kp[0];
    D[65] += (50000.0) * A.Kp[0];
    D[78] += (100000.0) * A.Kp[0];
    D[91] += (50000.0) * A.Kp[0];
    D[94] += (50000.0) * A.Kp[0];
    D[111] += (50000.0) * A.Kp[0];
    D[117] += (50000.0) * A.Kp[0];
    D[127] += (50000.0) * A.Kp[0];
    D[130] += (50000.0) * A.Kp[0];
    D[143] += (100000.0) * A.Kp[0];
    D[0] += (2000.0) * A.Kp[1];
    D[0] += (1000.0) * A.Kp[2];
    D[52] += (2000.0) * A.Kp[1];
    D[52] += (1000.0) * A.Kp[2];
    D[104] += (2000.0) * A.Kp[1];
    D[104] += (1000.0) * A.Kp[2];
  }
};
extern "C" __global__ void __launch_bounds__(64, 4) mfb_b0_lin(const MfbArgs A) { mfb::assemble<F_b0_lin>(A); }

struct F_b0_nl {
  static constexpr int NV = 3, NA = 20, NQ = 27, L1 = 3, BOUNDARY = 0, LINEAR = 0, NW = 15, NCW = 0, NC = 0, HAS_RES = 1, HAS_K = 0, TPB = 64, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 30256, EVAL = 0, NQPI = 6, NQPO = 0, NGS = 4;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {0, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {1, 2, 3, 0, 0, 1, 2, 3, 0, 0, 1, 2, 3, 0, 0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0, 0, 0, 1, 2, 0, 0, 0, 1, 2, 0, 0, 0, 1, 2}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    const double d1_1 = w[0];
    const double d1_2 = w[1];
    const double d1_3 = w[2];
    const double d1_t1 = w[3];
    const double d1_t2 = w[4];
    const double d2_1 = w[5];
    const double d2_2 = w[6];
    const double d2_3 = w[7];
    const double d2_t1 = w[8];
    const double d2_t2 = w[9];
    const double d3_1 = w[10];
    const double d3_2 = w[11];
    const double d3_3 = w[12];
    const double d3_t1 = w[13];
    const double d3_t2 = w[14];
    const double ep1 = qv[0];
    const double ep2 = qv[1];
    const double ep3 = qv[2];
    const double ep4 = qv[3];
    const double ep5 = qv[4];
    const double ep6 = qv[5];
    const double tmp0 = 50000.0*d1_2 + 50000.0*d2_1 - 100000.0*ep6;
    const double tmp1 = 50000.0*d1_3 + 50000.0*d3_1 - 100000.0*ep5;
    const double tmp2 = 50000.0*d2_3 + 50000.0*d3_2 - 100000.0*ep4;
    R[1] += 100000.0*(d1_1 - ep1);
    R[2] += tmp0;
    R[3] += tmp1;
    R[5] += tmp0;
    R[6] += 100000.0*(d2_2 - ep2);
    R[7] += tmp2;
    R[9] += tmp1;
    R[10] += tmp2;
    R[11] += 100000.0*(d3_3 - ep3);
    R[0] += 2000.0*d1_t1 + 1000.0*d1_t2;
    R[4] += 2000.0*d2_t1 + 1000.0*d2_t2;
    R[8] += 2000.0*d3_t1 + 1000.0*d3_t2;
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b0_nl(const MfbArgs A) { mfb::assemble<F_b0_nl>(A); }

struct F_b0_ev {
  static constexpr int NV = 3, NA = 20, NQ = 27, L1 = 3, BOUNDARY = 0, LINEAR = 0, NW = 15, NCW = 0, NC = 0, HAS_RES = 0, HAS_K = 0, TPB = 64, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 17296, EVAL = 1, NQPI = 0, NQPO = 6, NGS = 0;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {-1, -1, -1, -1}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {1, 2, 3, 0, 0, 1, 2, 3, 0, 0, 1, 2, 3, 0, 0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0, 0, 0, 1, 2, 0, 0, 0, 1, 2, 0, 0, 0, 1, 2}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static __forceinline__ void qp_eval(const double* w, const double* c, const MfbArgs& A, double* out) {
    const double d1_1 = w[0];
    const double d1_2 = w[1];
    const double d1_3 = w[2];
    const double d1_t1 = w[3];
    const double d1_t2 = w[4];
    const double d2_1 = w[5];
    const double d2_2 = w[6];
    const double d2_3 = w[7];
    const double d2_t1 = w[8];
    const double d2_t2 = w[9];
    const double d3_1 = w[10];
    const double d3_2 = w[11];
    const double d3_3 = w[12];
    const double d3_t1 = w[13];
    const double d3_t2 = w[14];
    out[0] = d1_1;
    out[1] = (1.0/2.0)*d1_2 + (1.0/2.0)*d2_1;
    out[2] = (1.0/2.0)*d1_3 + (1.0/2.0)*d3_1;
    out[3] = d2_2;
    out[4] = (1.0/2.0)*d2_3 + (1.0/2.0)*d3_2;
    out[5] = d3_3;
  }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b0_ev(const MfbArgs A) { mfb::assemble<F_b0_ev>(A); }

struct F_b1_lin {
  static constexpr int NV = 3, NA = 20, NQ = 9, L1 = 3, BOUNDARY = 1, LINEAR = 1, NW = 0, NCW = 3, NC = 3, HAS_RES = 0, HAS_K = 1, TPB = 64, NSD = 1, KS = 1, ND = 9, NTC = 20, CG = 1, W = 2, LPW = 30, SMEM = 33744, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {0, -1, -1, -1}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0, 0, 0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0, 1, 2}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    const double dw1 = c[0];
    const double dw2 = c[1];
    const double dw3 = c[2];
    D[0] += (100000000.0) * A.Kp[0];
    D[4] += (100000000.0) * A.Kp[0];
    D[8] += (100000000.0) * A.Kp[0];
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b1_lin(const MfbArgs A) { mfb::assemble<F_b1_lin>(A); }

struct F_b1_nl {
  static constexpr int NV = 3, NA = 20, NQ = 9, L1 = 3, BOUNDARY = 1, LINEAR = 0, NW = 3, NCW = 3, NC = 3, HAS_RES = 1, HAS_K = 0, TPB = 64, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 8400, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {0, -1, -1, -1}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {0, 0, 0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0, 0, 0}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0, 1, 2}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0, 0, 0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0, 1, 2}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    const double d1 = w[0];
    const double d2 = w[1];
    const double d3 = w[2];
    const double dw1 = c[0];
    const double dw2 = c[1];
    const double dw3 = c[2];
    R[0] += 100000000.0*(d1 - dw1);
    R[4] += 100000000.0*(d2 - dw2);
    R[8] += 100000000.0*(d3 - dw3);
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b1_nl(const MfbArgs A) { mfb::assemble<F_b1_nl>(A); }

struct F_b2_nl {
  static constexpr int NV = 3, NA = 20, NQ = 9, L1 = 3, BOUNDARY = 1, LINEAR = 0, NW = 0, NCW = 6, NC = 6, HAS_RES = 1, HAS_K = 0, TPB = 64, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 9744, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {0, -1, -1, -1}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0, 0, 0, 0, 0, 0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0, 1, 2, 3, 4, 5}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    const double sl1 = c[0];
    const double sl2 = c[1];
    const double sl3 = c[2];
    const double sl4 = c[3];
    const double sl5 = c[4];
    const double sl6 = c[5];
    const double n1 = nrm[0];
    const double n2 = nrm[1];
    const double n3 = nrm[2];
    R[0] += -n1*sl1 - n2*sl6 - n3*sl5;
    R[4] += -n1*sl6 - n2*sl2 - n3*sl4;
    R[8] += -n1*sl5 - n2*sl4 - n3*sl3;
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b2_nl(const MfbArgs A) { mfb::assemble<F_b2_nl>(A); }
