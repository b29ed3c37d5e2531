#include "mfb_skeleton.cuh"

struct F_b0_lin {
  static constexpr int NV = 3, NA = 20, NQ = 27, L1 = 1, BOUNDARY = 0, LINEAR = 1, NW = 0, NCW = 0, NC = 0, HAS_RES = 0, HAS_K = 1, TPB = 64, NSD = 3, KS = 3, ND = 81, NTC = 20, CG = 1, W = 2, LPW = 30, SMEM = 35984, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 3;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {-1, 0, 1, 2}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {1, 2, 3}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {1, 2, 3}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {1, 2, 3}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    D[0] += (-1.346153846153846) * A.Kp[0];
    D[4] += (-0.57692307692307687) * A.Kp[0];
    D[8] += (-0.57692307692307687) * A.Kp[0];
    D[10] += (-0.38461538461538458) * A.Kp[0];
    D[12] += (-0.38461538461538458) * A.Kp[0];
    D[20] += (-0.38461538461538458) * A.Kp[0];
    D[24] += (-0.38461538461538458) * A.Kp[0];
    D[28] += (-0.38461538461538458) * A.Kp[0];
    D[30] += (-0.38461538461538458) * A.Kp[0];
    D[36] += (-0.57692307692307687) * A.Kp[0];
    D[40] += (-1.346153846153846) * A.Kp[0];
    D[44] += (-0.57692307692307687) * A.Kp[0];
    D[50] += (-0.38461538461538458) * A.Kp[0];
    D[52] += (-0.38461538461538458) * A.Kp[0];
    D[56] += (-0.38461538461538458) * A.Kp[0];
    D[60] += (-0.38461538461538458) * A.Kp[0];
    D[68] += (-0.38461538461538458) * A.Kp[0];
    D[70] += (-0.38461538461538458) * A.Kp[0];
    D[72] += (-0.57692307692307687) * A.Kp[0];
    D[76] += (-0.57692307692307687) * A.Kp[0];
    D[80] += (-1.346153846153846) * A.Kp[0];
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b0_lin(const MfbArgs A) { mfb::assemble<F_b0_lin>(A); }

struct F_b0_nl {
  static constexpr int NV = 3, NA = 20, NQ = 27, L1 = 1, BOUNDARY = 0, LINEAR = 0, NW = 9, NCW = 0, NC = 0, HAS_RES = 1, HAS_K = 0, TPB = 64, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 19792, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 3;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {-1, 0, 1, 2}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {1, 2, 3}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {1, 2, 3, 1, 2, 3, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0, 0, 0, 1, 1, 1, 2, 2, 2}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    const double d1_1 = w[0];
    const double d1_2 = w[1];
    const double d1_3 = w[2];
    const double d2_1 = w[3];
    const double d2_2 = w[4];
    const double d2_3 = w[5];
    const double d3_1 = w[6];
    const double d3_2 = w[7];
    const double d3_3 = w[8];
    const double tmp0 = 0.57692307692307687*d2_2;
    const double tmp1 = 0.57692307692307687*d3_3;
    const double tmp2 = -0.38461538461538458*(d1_2 + d2_1);
    const double tmp3 = -0.38461538461538458*(d1_3 + d3_1);
    const double tmp4 = 0.57692307692307687*d1_1;
    const double tmp5 = -0.38461538461538458*(d2_3 + d3_2);
    R[1] += -1.346153846153846*d1_1 - tmp0 - tmp1;
    R[2] += tmp2;
    R[3] += tmp3;
    R[5] += tmp2;
    R[6] += -1.346153846153846*d2_2 - tmp1 - tmp4;
    R[7] += tmp5;
    R[9] += tmp3;
    R[10] += tmp5;
    R[11] += -1.346153846153846*d3_3 - tmp0 - tmp4;
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b0_nl(const MfbArgs A) { mfb::assemble<F_b0_nl>(A); }

struct F_b1_lin {
  static constexpr int NV = 3, NA = 20, NQ = 9, L1 = 1, BOUNDARY = 1, LINEAR = 1, NW = 0, NCW = 3, NC = 3, HAS_RES = 0, HAS_K = 1, TPB = 64, NSD = 1, KS = 1, ND = 9, NTC = 20, CG = 1, W = 2, LPW = 30, SMEM = 32784, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
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
    D[0] += (-1000.0) * A.Kp[0];
    D[4] += (-1000.0) * A.Kp[0];
    D[8] += (-1000.0) * A.Kp[0];
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b1_lin(const MfbArgs A) { mfb::assemble<F_b1_lin>(A); }

struct F_b1_nl {
  static constexpr int NV = 3, NA = 20, NQ = 9, L1 = 1, BOUNDARY = 1, LINEAR = 0, NW = 3, NCW = 3, NC = 3, HAS_RES = 1, HAS_K = 0, TPB = 64, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 5712, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
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
    R[0] += 1000.0*(-d1 + dw1);
    R[4] += 1000.0*(-d2 + dw2);
    R[8] += 1000.0*(-d3 + dw3);
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b1_nl(const MfbArgs A) { mfb::assemble<F_b1_nl>(A); }

struct F_b2_nl {
  static constexpr int NV = 3, NA = 20, NQ = 9, L1 = 1, BOUNDARY = 1, LINEAR = 0, NW = 0, NCW = 6, NC = 6, HAS_RES = 1, HAS_K = 0, TPB = 64, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 7056, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
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
    R[0] += n1*sl1 + n2*sl6 + n3*sl5;
    R[4] += n1*sl6 + n2*sl2 + n3*sl4;
    R[8] += n1*sl5 + n2*sl4 + n3*sl3;
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b2_nl(const MfbArgs A) { mfb::assemble<F_b2_nl>(A); }
