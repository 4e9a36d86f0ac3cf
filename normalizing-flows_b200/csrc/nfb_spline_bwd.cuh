// nfb_spline_bwd.cuh -- analytic backward of the monotone rational-quadratic spline element
// (utils/splines.py:100-219, forward branch :200-219), the arithmetic core of SURVEY 8f-1 ("backward of the
// fused blocks").  NOT yet wired into a kernel: the function is host/device and templated on the scalar type so
// that tests/native can check it in double precision against finite differences and against gradients minted
// from the reference's autograd (tests/test_spline_host.py) before a backward kernel is built around it.
//
// Same formulation as rqs_core (nfb_spline.cuh): logits in the log2 domain, knots on the unit interval
// (knot j = a * prefix_{j-1} + 1e-3 j, a = (1 - 1e-3 K) / sum), bin by search on the width knots, softplus on the
// two selected derivative logits only.  Given the upstream gradients (gy, glad) of one element it returns the
// gradients w.r.t. x, the K + K log2-domain logits and the K - 1 derivative logits.  Only the two knots of the
// selected bin and the two selected derivatives receive gradient directly; the softmax couples all K logits.
#pragma once
#include "nfb_common.cuh"
#include "nfb_spline.cuh"

namespace nfb {

template <typename T> __host__ __device__ __forceinline__ T t_exp2(T x);
template <> __host__ __device__ __forceinline__ float t_exp2<float>(float x) { return fast_ex2(x); }
template <> __host__ __device__ __forceinline__ double t_exp2<double>(double x) { return exp2(x); }
template <typename T> __host__ __device__ __forceinline__ T t_log(T x);
template <> __host__ __device__ __forceinline__ float t_log<float>(float x) { return kLn2 * fast_lg2(x); }
template <> __host__ __device__ __forceinline__ double t_log<double>(double x) { return log(x); }
template <typename T> __host__ __device__ __forceinline__ T t_exp(T x);
template <> __host__ __device__ __forceinline__ float t_exp<float>(float x) { return fast_ex2(x * kLog2e); }
template <> __host__ __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }

// Forward (y, lad) and backward in one pass (the backward needs every forward intermediate).
template <int K, typename T>
__host__ __device__ inline void rqs_fwd_bwd(T x, const T (&lw)[K], const T (&lh)[K], const T (&ud)[K - 1], T tail,
                                            T gy, T glad, T& y, T& lad, T& gx, T (&glw)[K], T (&glh)[K],
                                            T (&gud)[K - 1]) {
    const T m = (T)1e-3, md = (T)1e-3, cfac = (T)1 - m * (T)K, ln2 = (T)0.6931471805599453;
#pragma unroll
    for (int i = 0; i < K; ++i) { glw[i] = (T)0; glh[i] = (T)0; }
#pragma unroll
    for (int i = 0; i < K - 1; ++i) gud[i] = (T)0;
    if (!(x >= -tail && x <= tail)) {  // linear tails (and NaN): identity, lad 0  (:28,:40-41)
        y = x; lad = (T)0; gx = gy;
        return;
    }
    // ---- forward ----
    T mw = lw[0], mh = lh[0];
#pragma unroll
    for (int i = 1; i < K; ++i) { mw = lw[i] > mw ? lw[i] : mw; mh = lh[i] > mh ? lh[i] : mh; }
    T ew[K], eh[K], cw[K], ch[K];
    T sw = (T)0, sh = (T)0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        ew[i] = t_exp2<T>(lw[i] - mw); eh[i] = t_exp2<T>(lh[i] - mh);
        sw += ew[i]; sh += eh[i];
        cw[i] = sw; ch[i] = sh;
    }
    const T aw = cfac / sw, ah = cfac / sh;
    T kw[K + 1], kh[K + 1];
    kw[0] = (T)0; kh[0] = (T)0; kw[K] = (T)1; kh[K] = (T)1;
#pragma unroll
    for (int i = 0; i < K - 1; ++i) { kw[i + 1] = aw * cw[i] + m * (T)(i + 1); kh[i + 1] = ah * ch[i] + m * (T)(i + 1); }
    const T two_b = (T)2 * tail;
    const T xu = x / two_b + (T)0.5;
    int b = 0;
#pragma unroll
    for (int i = 1; i < K; ++i) b += (xu >= kw[i]) ? 1 : 0;  // knots increase: count = bin index
    const T l_w = kw[b], r_w = kw[b + 1], l_h = kh[b], r_h = kh[b + 1];
    const T u0 = b == 0 ? (T)NFB_BOUNDARY_UD : ud[b - 1], u1 = b == K - 1 ? (T)NFB_BOUNDARY_UD : ud[b];
    auto softplus = [](T u) { return u > (T)20 ? u : (u < (T)-30 ? t_exp<T>(u) : t_log<T>((T)1 + t_exp<T>(u))); };
    auto sigmoid = [](T u) { return (T)1 / ((T)1 + t_exp<T>(-u)); };
    const T d0 = md + softplus(u0), d1 = md + softplus(u1);
    const T w = r_w - l_w, h = r_h - l_h;
    const T delta = h / w, theta = (xu - l_w) / w, omt = (T)1 - theta;
    const T A = theta * theta, Bq = theta * omt, C = omt * omt;
    const T s = d0 + d1 - (T)2 * delta;
    const T den = delta + s * Bq;
    const T P = delta * A + d0 * Bq;
    const T num = h * P;
    const T outu = l_h + num / den;
    const T Q = d1 * A + (T)2 * delta * Bq + d0 * C;
    const T dnum = delta * delta * Q;
    y = outu * two_b - tail;
    lad = t_log<T>(dnum) - (T)2 * t_log<T>(den);
    // ---- backward of the local rational function ----
    const T g_out = gy * two_b;
    const T g_num = g_out / den;
    const T g_den = -g_out * num / (den * den) - (T)2 * glad / den;
    const T g_dnum = glad / dnum;
    T g_delta = g_dnum * ((T)2 * delta * Q + delta * delta * (T)2 * Bq);
    const T g_Q = g_dnum * delta * delta;
    T g_d1 = g_Q * A, g_d0 = g_Q * C;
    T g_A = g_Q * d1, g_B = g_Q * (T)2 * delta;
    const T g_C = g_Q * d0;
    T g_h = g_num * P;
    const T g_P = g_num * h;
    g_delta += g_P * A; g_A += g_P * delta; g_d0 += g_P * Bq; g_B += g_P * d0;
    g_delta += g_den;
    const T g_s = g_den * Bq;
    g_B += g_den * s;
    g_d0 += g_s; g_d1 += g_s; g_delta -= (T)2 * g_s;
    const T g_theta = (T)2 * theta * g_A + ((T)1 - (T)2 * theta) * g_B - (T)2 * omt * g_C;
    const T g_xu = g_theta / w;
    T g_lw = -g_theta / w;
    T g_w = -g_theta * theta / w;
    g_h += g_delta / w;
    g_w -= g_delta * delta / w;
    const T g_rw = g_w;
    g_lw -= g_w;
    const T g_rh = g_h;
    const T g_lh = g_out - g_h;
    gx = g_xu / two_b;
    if (b > 0) gud[b - 1] += g_d0 * (u0 > (T)20 ? (T)1 : sigmoid(u0));
    if (b < K - 1) gud[b] += g_d1 * (u1 > (T)20 ? (T)1 : sigmoid(u1));
    // ---- knots -> softmax logits.  knot j = cfac * prefix_{j-1} / sum + m j  (1 <= j <= K-1) ----
    //   d knot_j / d e_t = cfac * ([t <= j-1] - prefix_{j-1} / sum) / sum ;  d e_t / d logit_t = ln2 * e_t
    auto knots_bwd = [&](const T (&e)[K], const T (&c)[K], T sum, T g_left, T g_right, T (&g)[K]) {
        const T gk[2] = {b >= 1 ? g_left : (T)0, b + 1 <= K - 1 ? g_right : (T)0};
        const int jj[2] = {b, b + 1};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (gk[q] == (T)0) continue;
            const int j = jj[q];
            const T frac = c[j - 1] / sum, base = gk[q] * cfac / sum;
#pragma unroll
            for (int t = 0; t < K; ++t) g[t] += base * ((t <= j - 1 ? (T)1 : (T)0) - frac) * ln2 * e[t];
        }
    };
    knots_bwd(ew, cw, sw, g_lw, g_rw, glw);
    knots_bwd(eh, ch, sh, g_lh, g_rh, glh);
}

}  // namespace nfb
