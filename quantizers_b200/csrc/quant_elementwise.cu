// quant_elementwise.cu -- CT quantize / fake_quantize / dequantize with caller-supplied qparams for ANY
// strategy (TENSOR, CHANNEL, GROUP, BLOCK incl. ragged 128x128 edges): the general, un-fused path
// (CT:quantization/lifecycle/forward.py:185-241 -> forward_helpers.py:57-268).  One thread = 8 contiguous
// elements of a row; the qparam index is recomputed per element so any group/block geometry works.
#include "common.cuh"
#include "kernels.cuh"

namespace b200q {

__device__ __forceinline__ int64_t qindex(const ElemParams& p, int64_t r, int64_t c, int64_t pc) {
    switch (p.strategy) {
    case ST_TENSOR: return 0;
    case ST_CHANNEL: return r;
    case ST_GROUP: return r * pc + c / p.group;
    default: return (r / p.bh) * pc + c / p.bw;
    }
}

template <int DT, int QT, int OP>
__global__ void __launch_bounds__(256) elementwise_kernel(const ElemParams p) {
    const int64_t cpr = (p.cols + 7) / 8;  // chunks per row
    const int64_t pc = p.strategy == ST_GROUP ? (p.cols + p.group - 1) / p.group
                     : (p.strategy == ST_BLOCK ? (p.cols + p.bw - 1) / p.bw : 1);
    const float lo = (QT == QT_INT) ? -(float)(1 << (p.nbits - 1)) : (QT == QT_FP8 ? -448.0f : -6.0f);
    const float hi = (QT == QT_INT) ? (float)((1 << (p.nbits - 1)) - 1) : (QT == QT_FP8 ? 448.0f : 6.0f);
    const bool use_zp = (QT == QT_INT) ? (p.zp != nullptr) : (p.has_zp != 0);
    const float gs = p.gs ? p.gs[0] : 1.0f;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < p.rows * cpr; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / cpr, c0 = (t % cpr) * 8;
        const int n = (int)min((int64_t)8, p.cols - c0);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i >= n) break;
            const int64_t c = c0 + i, e = r * p.cols + c, k = qindex(p, r, c, pc);
            float s = load_T<DT>(p.scale, k);
            const float z = (QT == QT_INT && p.zp) ? (float)p.zp[k] : 0.0f;
            if (OP == EW_DEQUANT) {
                float q;
                if (QT == QT_INT) q = (float)((const int8_t*)p.x)[e];
                else if (QT == QT_FP8) q = e4m3_decode(((const uint8_t*)p.x)[e]);
                else q = load_T<DT>(p.x, e);
                float d;
                if (p.gs) {  // scale / global_scale promotes to fp32 (forward_helpers.py:255-256)
                    const float se = fdiv(s, gs);
                    d = use_zp ? fadd(q, -z) : q;
                    d = fmul(d, se);
                } else d = dequant_val<DT>(q, s, z, use_zp);
                store_T<DT>(p.out, e, d);
                continue;
            }
            const float x = load_T<DT>(p.x, e);
            if (QT == QT_FP4 || p.gs) s = fdiv(s, gs);
            if (OP == EW_QUANT) {
                if (QT == QT_INT) ((int8_t*)p.out)[e] = (int8_t)quant_int<DT>(x, s, z, use_zp, lo, hi);
                else if (QT == QT_FP8) ((uint8_t*)p.out)[e] = quant_fp8<DT>(x, s, use_zp);
                else {
                    const uint32_t nb = quant_fp4(x, s);
                    const float v = e2m1_value(nb & 7u);
                    store_T<DT>(p.out, e, (nb & 8u) ? -v : v);
                }
            } else {
                float y;
                if (QT == QT_INT) y = fq_int<DT>(x, s, z, use_zp, lo, hi);
                else if (QT == QT_FP8) y = fq_fp8<DT>(x, s, use_zp);
                else y = fq_fp4<DT>(x, s);
                store_T<DT>(p.out, e, y);
            }
        }
    }
}

template <int DT, int QT>
static int launch_op(int op, const ElemParams& p, cudaStream_t st) {
    const int64_t work = p.rows * ((p.cols + 7) / 8);
    if (work == 0) return B200Q_OK;
    const int blocks = (int)min((int64_t)kNumSMs * 16, (work + 255) / 256);
    switch (op) {
    case EW_QUANT: elementwise_kernel<DT, QT, EW_QUANT><<<blocks, 256, 0, st>>>(p); break;
    case EW_FQ: elementwise_kernel<DT, QT, EW_FQ><<<blocks, 256, 0, st>>>(p); break;
    case EW_DEQUANT: elementwise_kernel<DT, QT, EW_DEQUANT><<<blocks, 256, 0, st>>>(p); break;
    default: set_error("bad op %d", op); return B200Q_EINVAL;
    }
    B200Q_CHECK_LAUNCH();
    return B200Q_OK;
}

int launch_elementwise(int op, int dt, int qt, const ElemParams& p, cudaStream_t st) {
    if (p.strategy == ST_GROUP) B200Q_REQUIRE(p.group > 0 && (p.cols < p.group || p.cols % p.group == 0),
        "tensor column shape must be divisible by the given group_size %d but got %lld", p.group, (long long)p.cols);
    if (p.strategy == ST_BLOCK) B200Q_REQUIRE(p.bh > 0 && p.bw > 0, "block_structure must be positive");
    B200Q_DISPATCH_DT(dt, {
        switch (qt) {
        case QT_INT: return launch_op<DT, QT_INT>(op, p, st);
        case QT_FP8: return launch_op<DT, QT_FP8>(op, p, st);
        case QT_FP4: return launch_op<DT, QT_FP4>(op, p, st);
        default: set_error("bad qtype %d", qt); return B200Q_EINVAL;
        }
    });
    return B200Q_EINVAL;
}

}  // namespace b200q
