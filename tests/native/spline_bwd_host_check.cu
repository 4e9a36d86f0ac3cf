// Compiles csrc/nfb_spline_bwd.cuh for the HOST (double and float) so that the `not gpu` suite can check the
// analytic spline backward against finite differences and against gradients minted from the reference's
// autograd.  Test-only object; the product library never contains or calls this.
#include "../../normalizing-flows_b200/csrc/nfb_spline_bwd.cuh"
void nfb_set_error(const char*, ...) {}

template <typename T>
static void run(const double* x, const double* lw, const double* lh, const double* ud, int n, double tail,
                const double* gy, const double* gl, double* y, double* lad, double* gx, double* glw, double* glh,
                double* gud) {
    constexpr int K = 8;
    for (int i = 0; i < n; ++i) {
        T a[K], b[K], d[K - 1], ga[K], gb[K], gd[K - 1], yy, ll, gxx;
        for (int k = 0; k < K; ++k) { a[k] = (T)lw[i * K + k]; b[k] = (T)lh[i * K + k]; }
        for (int k = 0; k < K - 1; ++k) d[k] = (T)ud[i * (K - 1) + k];
        nfb::rqs_fwd_bwd<K, T>((T)x[i], a, b, d, (T)tail, (T)gy[i], (T)gl[i], yy, ll, gxx, ga, gb, gd);
        y[i] = yy; lad[i] = ll; gx[i] = gxx;
        for (int k = 0; k < K; ++k) { glw[i * K + k] = ga[k]; glh[i * K + k] = gb[k]; }
        for (int k = 0; k < K - 1; ++k) gud[i * (K - 1) + k] = gd[k];
    }
}

extern "C" __attribute__((visibility("default")))
void spline_bwd_check(const double* x, const double* lw, const double* lh, const double* ud, int n, double tail,
                      const double* gy, const double* gl, int use_float, double* y, double* lad, double* gx,
                      double* glw, double* glh, double* gud) {
    if (use_float) run<float>(x, lw, lh, ud, n, tail, gy, gl, y, lad, gx, glw, glh, gud);
    else run<double>(x, lw, lh, ud, n, tail, gy, gl, y, lad, gx, glw, glh, gud);
}
