// Compiles the product's spline arithmetic (csrc/nfb_spline.cuh) for the HOST so that the `not gpu`
// test-suite can check it against the golden vectors without a device.  Test-only object; the
// product library never contains or calls this.
#include "../../normalizing-flows_b200/csrc/nfb_spline.cuh"
void nfb_set_error(const char*, ...) {}
extern "C" __attribute__((visibility("default")))
void spline_host_check(const float* x, const float* params, int n, int K, float tail, float wh_scale,
                       int inverse, int templated, float* y, float* lad) {
    const int P = 3 * K - 1;
    for (int i = 0; i < n; ++i) {
        const float* p = params + (size_t)i * P;
        auto acc = [p](int k) { return p[k]; };
        if (templated && K == 8) {
            if (inverse) nfb::rqs_eval<8, true>(x[i], acc, tail, wh_scale, y[i], lad[i]);
            else nfb::rqs_eval<8, false>(x[i], acc, tail, wh_scale, y[i], lad[i]);
        } else {
            if (inverse) nfb::rqs_eval_dyn<true>(K, x[i], acc, tail, wh_scale, y[i], lad[i]);
            else nfb::rqs_eval_dyn<false>(K, x[i], acc, tail, wh_scale, y[i], lad[i]);
        }
    }
}

// per-feature tails (circular NSF layers): nd derivative parameters per element, circular flag and tail bound per feature
extern "C" __attribute__((visibility("default")))
void spline_host_check_tails(const float* x, const float* params, int rows, int feats, int K, int nd, const float* tail,
                             const int* circ, float wh_scale, int inverse, float* y, float* lad) {
    const int P = 2 * K + nd;
    for (int r = 0; r < rows; ++r)
        for (int f = 0; f < feats; ++f) {
            const size_t e = (size_t)r * feats + f;
            const float* p = params + e * P;
            auto acc = [p](int k) { return p[k]; };
            if (inverse) nfb::rqs_eval_dyn<true>(K, x[e], acc, tail[f], wh_scale, y[e], lad[e], nd, circ[f] != 0);
            else nfb::rqs_eval_dyn<false>(K, x[e], acc, tail[f], wh_scale, y[e], lad[e], nd, circ[f] != 0);
        }
}
