#include "mfb_skeleton.cuh"

struct F_b0_lin {
  static constexpr int NV = 4, NA = 20, NQ = 27, L1 = 2, BOUNDARY = 0, LINEAR = 1, NW = 0, NCW = 0, NC = 0, HAS_RES = 0, HAS_K = 1, TPB = 160, NSD = 4, KS = 4, ND = 256, NTC = 10, CG = 2, W = 5, LPW = 32, SMEM = 78256, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 4;
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
    D[0] += (1000.0) * A.Kp[1];
    D[17] += (100.0) * A.Kp[0];
    D[34] += (100.0) * A.Kp[0];
    D[51] += (100.0) * A.Kp[0];
    D[80] += (-10.5) * A.Kp[0];
    D[85] += (210000.0) * A.Kp[0];
    D[102] += (105000.0) * A.Kp[0];
    D[105] += (105000.0) * A.Kp[0];
    D[119] += (105000.0) * A.Kp[0];
    D[125] += (105000.0) * A.Kp[0];
    D[150] += (105000.0) * A.Kp[0];
    D[153] += (105000.0) * A.Kp[0];
    D[160] += (-10.5) * A.Kp[0];
    D[170] += (210000.0) * A.Kp[0];
    D[187] += (105000.0) * A.Kp[0];
    D[190] += (105000.0) * A.Kp[0];
    D[215] += (105000.0) * A.Kp[0];
    D[221] += (105000.0) * A.Kp[0];
    D[235] += (105000.0) * A.Kp[0];
    D[238] += (105000.0) * A.Kp[0];
    D[240] += (-10.5) * A.Kp[0];
    D[255] += (210000.0) * A.Kp[0];
    D[0] += (0.001575) * A.Kp[0];
    D[5] += (-10.5) * A.Kp[0];
    D[10] += (-10.5) * A.Kp[0];
    D[15] += (-10.5) * A.Kp[0];
    D[68] += (10.0) * A.Kp[1];
    D[136] += (10.0) * A.Kp[1];
    D[204] += (10.0) * A.Kp[1];
  }
};
extern "C" __global__ void __launch_bounds__(160, 2) mfb_b0_lin(const MfbArgs A) { mfb::assemble<F_b0_lin>(A); }

struct F_b0_nl {
  static constexpr int NV = 4, NA = 20, NQ = 27, L1 = 2, BOUNDARY = 0, LINEAR = 0, NW = 17, NCW = 0, NC = 0, HAS_RES = 1, HAS_K = 0, TPB = 160, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 29232, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 4;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {0, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0, 1, 2, 3}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    const double T = w[0];
    const double T_1 = w[1];
    const double T_2 = w[2];
    const double T_3 = w[3];
    const double T_t1 = w[4];
    const double d1_1 = w[5];
    const double d1_2 = w[6];
    const double d1_3 = w[7];
    const double d1_t1 = w[8];
    const double d2_1 = w[9];
    const double d2_2 = w[10];
    const double d2_3 = w[11];
    const double d2_t1 = w[12];
    const double d3_1 = w[13];
    const double d3_2 = w[14];
    const double d3_3 = w[15];
    const double d3_t1 = w[16];
    const double tmp0 = 10.5*T;
    const double tmp1 = 105000.0*(d1_2 + d2_1);
    const double tmp2 = 105000.0*(d1_3 + d3_1);
    const double tmp3 = 105000.0*(d2_3 + d3_2);
    R[0] += 0.001575*T + 1000.0*T_t1 - 10.5*d1_1 - 10.5*d2_2 - 10.5*d3_3;
    R[1] += 100.0*T_1;
    R[2] += 100.0*T_2;
    R[3] += 100.0*T_3;
    R[5] += 210000.0*d1_1 - tmp0;
    R[6] += tmp1;
    R[7] += tmp2;
    R[9] += tmp1;
    R[10] += 210000.0*d2_2 - tmp0;
    R[11] += tmp3;
    R[13] += tmp2;
    R[14] += tmp3;
    R[15] += 210000.0*d3_3 - tmp0;
    R[4] += 10.0*d1_t1;
    R[8] += 10.0*d2_t1;
    R[12] += 10.0*d3_t1;
  }
};
extern "C" __global__ void __launch_bounds__(160, 2) mfb_b0_nl(const MfbArgs A) { mfb::assemble<F_b0_nl>(A); }

struct F_b1_lin {
  static constexpr int NV = 4, NA = 20, NQ = 9, L1 = 2, BOUNDARY = 1, LINEAR = 1, NW = 0, NCW = 0, NC = 0, HAS_RES = 0, HAS_K = 1, TPB = 160, NSD = 1, KS = 1, ND = 16, NTC = 10, CG = 2, W = 5, LPW = 32, SMEM = 55088, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
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
    D[5] += (210000000.0) * A.Kp[0];
    D[10] += (210000000.0) * A.Kp[0];
    D[15] += (210000000.0) * A.Kp[0];
  }
};
extern "C" __global__ void __launch_bounds__(160, 2) mfb_b1_lin(const MfbArgs A) { mfb::assemble<F_b1_lin>(A); }

struct F_b1_nl {
  static constexpr int NV = 4, NA = 20, NQ = 9, L1 = 2, BOUNDARY = 1, LINEAR = 0, NW = 3, NCW = 0, NC = 0, HAS_RES = 1, HAS_K = 0, TPB = 160, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 6768, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
  __device__ static constexpr int gslot(int i) { constexpr int t[] = {0, -1, -1, -1}; return t[i]; }
  __device__ static constexpr int gslot_id(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int dslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int bslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int wslot(int i) { constexpr int t[] = {0, 0, 0}; return t[i]; }
  __device__ static constexpr int wlev(int i) { constexpr int t[] = {0, 0, 0}; return t[i]; }
  __device__ static constexpr int wpos(int i) { constexpr int t[] = {1, 2, 3}; return t[i]; }
  __device__ static constexpr int cslot(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static constexpr int cfield(int i) { constexpr int t[] = {0}; return t[i]; }
  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {
    const double d1 = w[0];
    const double d2 = w[1];
    const double d3 = w[2];
    R[4] += 210000000.0*d1;
    R[8] += 210000000.0*d2;
    R[12] += 210000000.0*d3;
  }
};
extern "C" __global__ void __launch_bounds__(160, 2) mfb_b1_nl(const MfbArgs A) { mfb::assemble<F_b1_nl>(A); }

struct F_b2_lin {
  static constexpr int NV = 4, NA = 20, NQ = 9, L1 = 2, BOUNDARY = 1, LINEAR = 1, NW = 0, NCW = 1, NC = 1, HAS_RES = 0, HAS_K = 1, TPB = 160, NSD = 1, KS = 1, ND = 16, NTC = 10, CG = 2, W = 5, LPW = 32, SMEM = 55088, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
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
    const double Te = c[0];
    D[0] += (100.0) * A.Kp[0];
  }
};
extern "C" __global__ void __launch_bounds__(160, 2) mfb_b2_lin(const MfbArgs A) { mfb::assemble<F_b2_lin>(A); }

struct F_b2_nl {
  static constexpr int NV = 4, NA = 20, NQ = 9, L1 = 2, BOUNDARY = 1, LINEAR = 0, NW = 1, NCW = 1, NC = 1, HAS_RES = 1, HAS_K = 0, TPB = 160, NSD = 0, KS = 0, ND = 0, NTC = 2, CG = 1, W = 1, LPW = 1, SMEM = 7056, EVAL = 0, NQPI = 0, NQPO = 0, NGS = 1;
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
    const double Te = c[0];
    R[0] += 100.0*(T - Te);
  }
};
extern "C" __global__ void __launch_bounds__(160, 2) mfb_b2_nl(const MfbArgs A) { mfb::assemble<F_b2_nl>(A); }
