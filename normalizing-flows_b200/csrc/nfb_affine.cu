// nfb_affine.cu -- the affine family for low-dimensional flows (D <= 16), whole stack in ONE kernel.
//
// One thread per sample keeps z[D] and the running log-det in registers and walks the layer list:
//   MaskedAffineFlow      flows/affine/coupling.py:208-229   (s,t = MLPs on b*z; non-finite -> NaN)
//   AffineCouplingBlock   flows/affine/coupling.py:253-267 -> AffineCoupling :113-171
//   AffineConstFlow/ActNorm (after init)  flows/affine/coupling.py:38-54
//   Permute               flows/mixing.py:31-54
// HBM traffic per sample is D*4 bytes in, D*4 out (+4 for log_q) for the entire stack -- the
// reference launches ~26 ATen ops per layer.  MLP weights are read through the
// read-only path with warp-uniform addresses (broadcast).
#include "nfb_kernels.h"

namespace nfb {

__device__ __forceinline__ void mlp_eval(const AffMlp& m, const float* in, float* out, float slope) {
    float a[kAffMaxW], b[kAffMaxW];
    const int n0 = m.sizes[0];
    for (int i = 0; i < n0; ++i) a[i] = in[i];
    float* cur = a;
    float* nxt = b;
    for (int l = 0; l < m.n_layers; ++l) {
        const int ni = m.sizes[l], no = m.sizes[l + 1];
        const float* w = m.w[l];
        const float* bias = m.b[l];
        const bool last = (l + 1 == m.n_layers);
        for (int o = 0; o < no; ++o) {
            float acc = __ldg(bias + o);
            for (int i = 0; i < ni; ++i) acc = fmaf(cur[i], __ldg(w + o * ni + i), acc);
            nxt[o] = last ? acc : (acc >= 0.f ? acc : acc * slope);
        }
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    const int no = m.sizes[m.n_layers];
    for (int o = 0; o < no; ++o) out[o] = cur[o];
}

// direction: 0 = "inverse" (density pass: ops applied last-to-first), 1 = "forward" (sampling)
__global__ void __launch_bounds__(128)
affine_stack_kernel(const AffineOp* __restrict__ ops, int n_ops, const float* __restrict__ zin,
                    float* __restrict__ zout, float* __restrict__ logq, long long rows, int d,
                    int accumulate, int direction) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    float z[kAffMaxD];
    for (int j = 0; j < d; ++j) z[j] = zin[row * d + j];
    float ld = 0.f;
    for (int k = 0; k < n_ops; ++k) {
        const AffineOp& op = ops[direction ? k : (n_ops - 1 - k)];
        if (op.type == kOpMasked) {
            float zm[kAffMaxD], s[kAffMaxD], t[kAffMaxD];
            for (int j = 0; j < d; ++j) zm[j] = __ldg(op.p0 + j) * z[j];
            if (op.s.n_layers) mlp_eval(op.s, zm, s, op.slope); else for (int j = 0; j < d; ++j) s[j] = 0.f;
            if (op.t.n_layers) mlp_eval(op.t, zm, t, op.slope); else for (int j = 0; j < d; ++j) t[j] = 0.f;
            for (int j = 0; j < d; ++j) {
                const float b = __ldg(op.p0 + j);
                const float sj = isfinite(s[j]) ? s[j] : __int_as_float(0x7fc00000);
                const float tj = isfinite(t[j]) ? t[j] : __int_as_float(0x7fc00000);
                if (direction) {
                    z[j] = zm[j] + (1.f - b) * (z[j] * expf(sj) + tj);
                    ld += (1.f - b) * sj;
                } else {
                    z[j] = zm[j] + (1.f - b) * (z[j] - tj) * expf(-sj);
                    ld -= (1.f - b) * sj;
                }
            }
        } else if (op.type == kOpConst) {
            float ssum = 0.f;
            for (int j = 0; j < d; ++j) {
                const float s = __ldg(op.p0 + j), t = __ldg(op.p1 + j);
                z[j] = direction ? z[j] * expf(s) + t : (z[j] - t) * expf(-s);
                ssum += s;
            }
            ld += direction ? ssum : -ssum;
        } else if (op.type == kOpCoupling) {
            const int h = (d + 1) / 2;                 // torch.chunk(2): first chunk ceil(d/2)
            const bool inv_split = (op.flags >> 3) & 1;  // channel_inv: z1 is the SECOND chunk
            const int o1 = inv_split ? h : 0, n1 = inv_split ? d - h : h;
            const int o2 = inv_split ? 0 : h, n2 = d - n1;
            float param[2 * kAffMaxD];
            mlp_eval(op.s, z + o1, param, op.slope);
            if (!(op.flags & 1)) {
                for (int j = 0; j < n2; ++j) z[o2 + j] += direction ? param[j] : -param[j];
            } else {
                const int smap = (op.flags >> 1) & 3;
                for (int j = 0; j < n2; ++j) {
                    const float shift = param[2 * j], sc = param[2 * j + 1];
                    float& v = z[o2 + j];
                    if (smap == 0) {
                        if (direction) { v = v * expf(sc) + shift; ld += sc; }
                        else { v = (v - shift) * expf(-sc); ld -= sc; }
                    } else {
                        const float sg = 1.f / (1.f + expf(-(sc + 2.f)));
                        const float lsg = logf(sg);
                        const bool div = (smap == 1) == (direction != 0);
                        if (direction) v = div ? v / sg + shift : v * sg + shift;
                        else v = div ? (v - shift) / sg : (v - shift) * sg;
                        ld += div ? -lsg : lsg;
                    }
                }
            }
        } else {  // permute
            const int* idx = direction ? op.fwd_idx : op.inv_idx;
            float tmp[kAffMaxD];
            for (int j = 0; j < d; ++j) tmp[j] = z[__ldg(idx + j)];
            for (int j = 0; j < d; ++j) z[j] = tmp[j];
        }
    }
    for (int j = 0; j < d; ++j) zout[row * d + j] = z[j];
    if (logq) logq[row] = accumulate ? logq[row] + ld : ld;
}

int launch_affine_stack(const void* ops_dev, int n_ops, const float* zin, float* zout, float* logq,
                        long long rows, int d, int accumulate, int direction, cudaStream_t st) {
    NFB_CHECK(d >= 1 && d <= kAffMaxD, NFB_ERR_UNSUPPORTED, "affine stack: dim %d > %d", d, kAffMaxD);
    if (rows == 0) return NFB_OK;
    affine_stack_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(
        static_cast<const AffineOp*>(ops_dev), n_ops, zin, zout, logq, rows, d, accumulate, direction);
    NFB_LAUNCH_CHECK();
    return NFB_OK;
}

size_t affine_op_size() { return sizeof(AffineOp); }

}  // namespace nfb
