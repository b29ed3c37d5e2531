// mfb_qp.cu -- INTEGRATION_POINT_VAR arrays and the built-in J2 return map.
//
// The reference evaluates a user callback on whole [n_q, n_el] device arrays between the interpolation of its
// arguments and the residual terms that read its outputs (src/symbolics/08_Tensor.jl:175-183,210). The only callback
// in the reference's examples is the J2 stress update of examples/hypo_elastic_plasticity/J2Plasticity.jl:76-198
// (MaterialState callable -> iterate_stress!; update_States!), ~40 broadcast kernels and a findall per call there.
// Here it is ONE element-wise, HBM-bound kernel: 6 + 13 doubles read, 13 written per quadrature point.
#include <cstring>

#include "mfb_internal.h"

namespace {
constexpr int TPB = 256;

struct J2Ptrs {
    const double* e[6];        // callback argument order: e11 e12 e13 e22 e23 e33
    const double* ep[6];       // committed state, Voigt order 11 22 33 23 13 12
    const double* b[6];
    const double* Y;
    double* ep_eval[6];
    double* b_eval[6];
    double* Y_eval;
};

// iterate_stress! (J2Plasticity.jl:118-188), one thread per quadrature point
__global__ void k_j2_iterate(J2Ptrs P, mfb_j2_params m, int64_t n, unsigned long long* n_yielded) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ea[6], ep0[6], b0[6], ep[6], b[6], Yn;
#pragma unroll
    for (int k = 0; k < 6; ++k) { ea[k] = P.e[k][i]; ep0[k] = P.ep[k][i]; b0[k] = P.b[k][i]; }
    if (mfb::j2_return_map(ea, ep0, b0, P.Y[i], m.lambda, m.mu, m.Eb, m.Ep, m.f_res, ep, b, Yn)) atomicAdd(n_yielded, 1ULL);
#pragma unroll
    for (int k = 0; k < 6; ++k) { P.ep_eval[k][i] = ep[k]; P.b_eval[k][i] = b[k]; }
    P.Y_eval[i] = Yn;
}

__global__ void k_fill(double* p, double v, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
}  // namespace

int mfb_qp_lookup(mfb_ctx* ctx, const std::string& name, double** p) {
    MFB_REQUIRE(ctx->n_el > 0 && ctx->n_q > 0, MFB_ERR_STATE, "integration-point arrays need mfb_mesh_set first");
    DevBuf<double>& a = ctx->qp[name];
    const size_t n = (size_t)ctx->n_q * ctx->n_el;
    if (a.n != n || !a.p) {
        MFB_CUDA(a.alloc(n));
        MFB_CUDA(cudaMemsetAsync(a.p, 0, n * sizeof(double), ctx->stream));
    }
    *p = a.p;
    return MFB_OK;
}

extern "C" int mfb_qp_array(mfb_ctx* ctx, const char* name, double** device_ptr, int64_t* n) {
    if (!ctx || !name) return MFB_ERR_ARG;
    MFB_CUDA(cudaSetDevice(ctx->device));
    double* p = nullptr;
    MFB_TRY(mfb_qp_lookup(ctx, name, &p));
    if (device_ptr) *device_ptr = p;
    if (n) *n = (int64_t)ctx->n_q * ctx->n_el;
    return MFB_OK;
}

extern "C" int mfb_qp_set(mfb_ctx* ctx, const char* name, const double* values, int64_t n) {
    if (!ctx || !name) return MFB_ERR_ARG;
    MFB_CUDA(cudaSetDevice(ctx->device));
    double* p = nullptr;
    MFB_TRY(mfb_qp_lookup(ctx, name, &p));
    MFB_REQUIRE(n == (int64_t)ctx->n_q * ctx->n_el, MFB_ERR_ARG, "mfb_qp_set: wrong length");
    MFB_TRY(mfb_stage_in(ctx, values, n * sizeof(double), p));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}

extern "C" int mfb_qp_get(mfb_ctx* ctx, const char* name, double* values, int64_t n) {
    if (!ctx || !name) return MFB_ERR_ARG;
    MFB_CUDA(cudaSetDevice(ctx->device));
    double* p = nullptr;
    MFB_TRY(mfb_qp_lookup(ctx, name, &p));
    MFB_REQUIRE(n == (int64_t)ctx->n_q * ctx->n_el, MFB_ERR_ARG, "mfb_qp_get: wrong length");
    MFB_TRY(mfb_stage_out(ctx, p, n * sizeof(double), values));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}

static std::string j2name(const char* prefix, const char* what, int k) {
    return std::string(prefix) + "." + what + (k >= 0 ? std::to_string(k + 1) : std::string());
}

// MaterialState(wp, Y_initial, ...) (J2Plasticity.jl:76-101): zero ep, b; Y = Y_initial
extern "C" int mfb_j2_init(mfb_ctx* ctx, const char* prefix, double Y_initial, const char* const* e_names,
                           const char* const* ep_names) {
    if (!ctx || !prefix || !e_names || !ep_names) return MFB_ERR_ARG;
    MFB_CUDA(cudaSetDevice(ctx->device));
    mfb_ctx::J2State& st = ctx->j2[prefix];
    const int64_t n = (int64_t)ctx->n_q * ctx->n_el;
    double* p = nullptr;
    for (int k = 0; k < 6; ++k) {
        st.e[k] = e_names[k];
        st.ep[k] = ep_names[k];
        for (const char* w : {"ep", "b", "b_eval"}) {
            MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, w, k), &p));
            MFB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double), ctx->stream));
        }
        MFB_TRY(mfb_qp_lookup(ctx, st.e[k], &p));
        MFB_TRY(mfb_qp_lookup(ctx, st.ep[k], &p));
        MFB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double), ctx->stream));
    }
    for (const char* w : {"Y", "Y_eval"}) {
        MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, w, -1), &p));
        k_fill<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(p, Y_initial, n);
        ctx->launches++;
    }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}

extern "C" int mfb_j2_iterate_stress(mfb_ctx* ctx, const char* prefix, const mfb_j2_params* params, int64_t* n_yielded) {
    if (!ctx || !prefix || !params) return MFB_ERR_ARG;
    auto it = ctx->j2.find(prefix);
    MFB_REQUIRE(it != ctx->j2.end(), MFB_ERR_STATE, "mfb_j2_iterate_stress: call mfb_j2_init first");
    MFB_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = (int64_t)ctx->n_q * ctx->n_el;
    J2Ptrs P;
    double* p = nullptr;
    for (int k = 0; k < 6; ++k) {
        MFB_TRY(mfb_qp_lookup(ctx, it->second.e[k], &p)); P.e[k] = p;
        MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, "ep", k), &p)); P.ep[k] = p;
        MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, "b", k), &p)); P.b[k] = p;
        MFB_TRY(mfb_qp_lookup(ctx, it->second.ep[k], &P.ep_eval[k]));
        MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, "b_eval", k), &P.b_eval[k]));
    }
    MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, "Y", -1), &p)); P.Y = p;
    MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, "Y_eval", -1), &P.Y_eval));
    MFB_CUDA(ctx->qp_counter.alloc(1));
    unsigned long long* cnt = ctx->qp_counter.p;
    MFB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), ctx->stream));
    k_j2_iterate<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(P, *params, n, cnt);
    ctx->launches++;
    MFB_CUDA(cudaGetLastError());
    if (n_yielded) {
        unsigned long long h = 0;
        MFB_CUDA(cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        *n_yielded = (int64_t)h;
    }
    return MFB_OK;
}

// update_States! (J2Plasticity.jl:190-197)
extern "C" int mfb_j2_update_states(mfb_ctx* ctx, const char* prefix) {
    if (!ctx || !prefix) return MFB_ERR_ARG;
    auto it = ctx->j2.find(prefix);
    MFB_REQUIRE(it != ctx->j2.end(), MFB_ERR_STATE, "mfb_j2_update_states: call mfb_j2_init first");
    MFB_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)ctx->n_q * ctx->n_el * sizeof(double);
    double *src = nullptr, *dst = nullptr;
    for (int k = 0; k < 6; ++k) {
        MFB_TRY(mfb_qp_lookup(ctx, it->second.ep[k], &src));
        MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, "ep", k), &dst));
        MFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, "b_eval", k), &src));
        MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, "b", k), &dst));
        MFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, "Y_eval", -1), &src));
    MFB_TRY(mfb_qp_lookup(ctx, j2name(prefix, "Y", -1), &dst));
    MFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return MFB_OK;
}

// Fused path (the return map is inlined in the element kernel): number of points that yielded in the last
// mfb_assemble_nonlinear; the counter is the first word of the integration-point array "<prefix>.count".
extern "C" int mfb_j2_yield_count(mfb_ctx* ctx, const char* prefix, int64_t* n_yielded) {
    if (!ctx || !prefix || !n_yielded) return MFB_ERR_ARG;
    MFB_CUDA(cudaSetDevice(ctx->device));
    double* p = nullptr;
    MFB_TRY(mfb_qp_lookup(ctx, std::string(prefix) + ".count", &p));
    unsigned long long h = 0;
    MFB_CUDA(cudaMemcpyAsync(&h, p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_yielded = (int64_t)h;
    return MFB_OK;
}
