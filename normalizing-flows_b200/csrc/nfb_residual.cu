// nfb_residual.cu -- element-wise kernels of the invertible residual block (flows/residual.py:12-251 `Residual` /
// `iResBlock`) with a Lipschitz MLP (nets/lipschitz.py:14-67: [Swish, InducedNormLinear] x n).
//
// The Linear layers of g(x), of its Jacobian-vector products (forward mode: the exact 2 x 2 Jacobian of the 2-D eval
// path, residual.py:148-161) and of its vector-Jacobian products (reverse mode: the Hutchinson power series,
// :355-379) are tensor-core GEMMs (csrc/nfb_gemm_tc.cu); what is left is element-wise and HBM-bound:
//   swish        a = x sigmoid(b x) / 1.1,  da = d a / d x          (nets/lipschitz.py:642-648, b = softplus(beta))
//   mul_rows     T[t, i] *= m[i]                                     (tangents / cotangents through the activation)
//   logdet2      log |det(I + J)| for [B, 2, 2] Jacobians given as two tangent outputs
//   rowdot       out[r] (+)= c * sum_j a[r, j] b[r, j]              (trace estimate v^T J^k eps per sample)
#include "nfb_kernels.h"

namespace nfb {

__global__ void swish_kernel(const float* __restrict__ x, float b, long long n, float* __restrict__ a,
                             float* __restrict__ da) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    const float s = 1.f / (1.f + __expf(-b * v));
    a[i] = v * s * (1.f / 1.1f);
    if (da) da[i] = (s + b * v * s * (1.f - s)) * (1.f / 1.1f);
}
int launch_swish(const float* x, float b, long long n, float* a, float* da, cudaStream_t st) {
    if (n == 0) return NFB_OK;
    swish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, b, n, a, da);
    NFB_LAUNCH_CHECK();
    return NFB_OK;
}

// T[t * n + i] = S[t * n + i] * m[i]  for t < nt  (in place allowed)
__global__ void mul_rows_kernel(const float* __restrict__ S, const float* __restrict__ m, long long n, int nt,
                                float* __restrict__ T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float w = m[i];
    for (int t = 0; t < nt; ++t) T[(long long)t * n + i] = S[(long long)t * n + i] * w;
}
int launch_mul_rows(const float* S, const float* m, long long n, int nt, float* T, cudaStream_t st) {
    if (n == 0 || nt == 0) return NFB_OK;
    mul_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S, m, n, nt, T);
    NFB_LAUNCH_CHECK();
    return NFB_OK;
}

// jt: [2, B, 2] -- jt[t, r, :] = J(x_r) e_t (column t of the Jacobian of g at sample r).  out[r] = log |det(I + J)|.
__global__ void logdet2_kernel(const float* __restrict__ jt, long long B, float* __restrict__ out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const float2 c0 = reinterpret_cast<const float2*>(jt)[r];        // (J00, J10)
    const float2 c1 = reinterpret_cast<const float2*>(jt)[B + r];    // (J01, J11)
    out[r] = logf(fabsf((c0.x + 1.f) * (c1.y + 1.f) - c1.x * c0.y));
}
int launch_logdet2(const float* jt, long long B, float* out, cudaStream_t st) {
    if (B == 0) return NFB_OK;
    logdet2_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(jt, B, out);
    NFB_LAUNCH_CHECK();
    return NFB_OK;
}

// out = h + t * sigmoid(c): the GLU gate of a context-conditioned residual block (nets/resnet.py:48-50,
// nets/made.py:212-214: F.glu(cat(temps, context_layer(context))) + inputs)
__global__ void glu_residual_kernel(const float* __restrict__ h, const float* __restrict__ t, const float* __restrict__ c,
                                    long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fmaf(t[i], 1.f / (1.f + __expf(-c[i])), h[i]);
}
int launch_glu_residual(const float* h, const float* t, const float* c, long long n, float* out, cudaStream_t st) {
    if (n == 0) return NFB_OK;
    glu_residual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h, t, c, n, out);
    NFB_LAUNCH_CHECK();
    return NFB_OK;
}

// out[r] = (accumulate ? out[r] : 0) + c * sum_j a[r, j] * b[r, j]   (one warp per row)
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long rows,
                                                     int d, float c, int accumulate, float* __restrict__ out) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s = 0.f;
    for (int j = lane; j < d; j += 32) s = fmaf(a[row * d + j], b[row * d + j], s);
    s = warp_sum(s);
    if (lane == 0) out[row] = (accumulate ? out[row] : 0.f) + c * s;
}
int launch_rowdot(const float* a, const float* b, long long rows, int d, float c, int accumulate, float* out,
                  cudaStream_t st) {
    if (rows == 0) return NFB_OK;
    rowdot_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(a, b, rows, d, c, accumulate, out);
    NFB_LAUNCH_CHECK();
    return NFB_OK;
}

}  // namespace nfb
