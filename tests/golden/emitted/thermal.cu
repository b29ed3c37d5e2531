#include "mfb_skeleton.cuh"

struct F_b0_lin {
  static constexpr int NV = 1, NA = 10, NQ = 14, L1 = 1, BOUNDARY = 0, LINEAR = 1, NW = 0, NCW = 1, NC = 1, HAS_RES = 0, HAS_K = 1, TPB = 64, NSD = 3, KS = 3, ND = 9, NTC = 10, CG = 1, W = 1, LPW = 10, SMEM = 6832, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 3;
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
    const double s = c[0];
    D[0] += (-0.59999999999999998) * A.Kp[0];
    D[4] += (-0.59999999999999998) * A.Kp[0];
    D[8] += (-0.59999999999999998) * A.Kp[0];
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b0_lin(const MfbArgs A) { mfb::assemble<F_b0_lin>(A); }

struct F_b0_nl {
  static constexpr int NV = 1, NA = 10, NQ = 14, L1 = 1, BOUNDARY = 0, LINEAR = 0, NW = 3, NCW = 1, NC = 1, HAS_RES = 1, HAS_K = 0, TPB = 64, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 7392, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 4;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {0, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {1, 2, 3}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0, 0, 0}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0, 0, 0}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    const double T_1 = w[0];
    const double T_2 = w[1];
    const double T_3 = w[2];
    const double s = c[0];
    R[0] += s;
    R[1] += -0.59999999999999998*T_1;
    R[2] += -0.59999999999999998*T_2;
    R[3] += -0.59999999999999998*T_3;
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b0_nl(const MfbArgs A) { mfb::assemble<F_b0_nl>(A); }

struct F_b1_lin {
  static constexpr int NV = 1, NA = 10, NQ = 7, L1 = 1, BOUNDARY = 1, LINEAR = 1, NW = 0, NCW = 0, NC = 0, HAS_RES = 0, HAS_K = 1, TPB = 64, NSD = 1, KS = 1, ND = 1, NTC = 10, CG = 1, W = 1, LPW = 10, SMEM = 2208, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {0, -1, -1, -1}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    D[0] += (-25.0) * A.Kp[0];
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b1_lin(const MfbArgs A) { mfb::assemble<F_b1_lin>(A); }

struct F_b1_nl {
  static constexpr int NV = 1, NA = 10, NQ = 7, L1 = 1, BOUNDARY = 1, LINEAR = 0, NW = 1, NCW = 0, NC = 0, HAS_RES = 1, HAS_K = 0, TPB = 64, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 2032, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {0, -1, -1, -1}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    const double T = w[0];
    R[0] += 7328.7499999999991 - 25.0*T;
  }
};
extern "C" __global__ void __launch_bounds__(64, 6) mfb_b1_nl(const MfbArgs A) { mfb::assemble<F_b1_nl>(A); }
