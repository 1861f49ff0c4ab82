// MSE range estimator kernels.
//
// Reference: MSE_Estimator.loss_fx (range_estimators.py:248-256) is evaluated once per candidate
// -- 100 (1-D grid), 12,800 (2-D grid) or ~30-900 times (golden section) -- and each evaluation
// deep-copies a quantizer, runs ~14 ATen kernels over the whole tensor and syncs the host.
//
// Here the tensor is read from HBM ONCE: every CTA stages its slice of x in shared memory (up to
// 160 KB, B200 has 227 KB per CTA) and sweeps the whole candidate table over it.  The work is
// FP32-CUDA-core bound (IEEE division + rint + clamp + dequant + squared error per (element,
// candidate)); 512 threads x 2 independent float4 per iteration keep the FMA/ALU pipes busy.
// Summation: squared errors fp32, per-thread partial fp32, everything above in fp64 and in a
// fixed order (per-CTA partial rows, then a column-sum kernel) -> bitwise run-to-run reproducible.
#include "tq_common.cuh"

namespace tq {

constexpr int kMThreads = 512;
constexpr int kSliceMax = 40960;          // floats staged per CTA iteration (160 KB)
constexpr int kCandChunk = 128;           // candidates per block-reduction round

template <bool FAST>
__device__ __forceinline__ float2 sqerr2(float2 x, const QP2& p, float2 acc) {
    const float2 y = qdq2_t<FAST>(x, p);
    const float2 d = __fadd2_rn(x, make_float2(-y.x, -y.y));
    return __fadd2_rn(acc, __fmul2_rn(d, d));
}

template <bool FAST>
__device__ __forceinline__ void sse_slice(const float4* __restrict__ xs4, int nv, int tid, const QP& p, float& acc0,
                                          float& acc1) {
    const QP2 p2 = pair_of(p);
    float2 a0 = make_float2(0.0f, 0.0f), a1 = a0, a2 = a0, a3 = a0;
    int i = tid;
    for (; i + kMThreads < nv; i += 2 * kMThreads) {
        const float4 a = xs4[i], b = xs4[i + kMThreads];
        a0 = sqerr2<FAST>(make_float2(a.x, a.y), p2, a0);
        a1 = sqerr2<FAST>(make_float2(a.z, a.w), p2, a1);
        a2 = sqerr2<FAST>(make_float2(b.x, b.y), p2, a2);
        a3 = sqerr2<FAST>(make_float2(b.z, b.w), p2, a3);
    }
    if (i < nv) {
        const float4 a = xs4[i];
        a0 = sqerr2<FAST>(make_float2(a.x, a.y), p2, a0);
        a1 = sqerr2<FAST>(make_float2(a.z, a.w), p2, a1);
    }
    acc0 = (a0.x + a0.y) + (a1.x + a1.y);
    acc1 = (a2.x + a2.y) + (a3.x + a3.y);
}

__global__ void __launch_bounds__(kMThreads, 1)
mse_sse_kernel(const float* __restrict__ x, int64_t n, int64_t per_cta, int vec_ok,
               const float* __restrict__ cand, int32_t n_cand, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);                                   // [kSliceMax]
    QP* ctab = reinterpret_cast<QP*>(smem_raw + (size_t)kSliceMax * 4);                   // [kCandChunk]
    double* wpart = reinterpret_cast<double*>(smem_raw + (size_t)kSliceMax * 4 + kCandChunk * sizeof(QP));  // [16][kCandChunk]

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t begin = (int64_t)blockIdx.x * per_cta;
    int64_t end = begin + per_cta;
    if (end > n) end = n;
    double* row = partial + (size_t)blockIdx.x * n_cand;
    bool first_slice = true;

    for (int64_t s0 = begin; s0 < end || first_slice; s0 += kSliceMax) {
        int64_t s1 = s0 + kSliceMax;
        if (s1 > end) s1 = end;
        const int len = s1 > s0 ? (int)(s1 - s0) : 0;
        const int len4 = (len + 3) & ~3;                       // zero padded: QDQ(0) == 0 exactly
        // ---- stage the slice (the only HBM read of x) ----
        if (vec_ok) {
            const float4* xv = reinterpret_cast<const float4*>(x + s0);
            for (int i = tid; i < (len >> 2); i += kMThreads)
                reinterpret_cast<float4*>(xs)[i] = ld_stream(xv + i);
            for (int i = (len & ~3) + tid; i < len4; i += kMThreads) xs[i] = i < len ? x[s0 + i] : 0.0f;
        } else {
            for (int i = tid; i < len4; i += kMThreads) xs[i] = i < len ? x[s0 + i] : 0.0f;
        }
        __syncthreads();
        const int nv = len4 >> 2;
        const float4* xs4 = reinterpret_cast<const float4*>(xs);

        for (int c0 = 0; c0 < n_cand; c0 += kCandChunk) {
            const int cn = (n_cand - c0) < kCandChunk ? (n_cand - c0) : kCandChunk;
            if (tid < cn)
                ctab[tid] = make_qp(cand[c0 + tid], cand[n_cand + c0 + tid], cand[2 * n_cand + c0 + tid],
                                    cand[3 * n_cand + c0 + tid]);
            __syncthreads();
            for (int c = 0; c < cn; ++c) {
                const QP p = ctab[c];
                float acc0 = 0.0f, acc1 = 0.0f;
                if (p.exact) sse_slice<false>(xs4, nv, tid, p, acc0, acc1);      // uniform per candidate
                else sse_slice<true>(xs4, nv, tid, p, acc0, acc1);
                const double w = warp_sum((double)acc0 + (double)acc1);
                if (lane == 0) wpart[wid * kCandChunk + c] = w;
            }
            __syncthreads();
            if (tid < cn) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < kMThreads / 32; ++w) tot += wpart[w * kCandChunk + tid];
                if (first_slice) row[c0 + tid] = tot;
                else row[c0 + tid] += tot;
            }
            __syncthreads();
        }
        first_slice = false;
    }
}

__global__ void mse_colsum_kernel(const double* __restrict__ partial, int rows, int32_t n_cand,
                                  double* __restrict__ loss_accum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cand) return;
    double tot = 0.0;
    for (int r = 0; r < rows; ++r) tot += partial[(size_t)r * n_cand + c];
    loss_accum[c] += tot;          // range_estimators.py:366 / :401 accumulate over batches
}

struct Best {
    double v;
    int idx;
};
__device__ __forceinline__ bool better(const Best& a, const Best& b) {   // np.argmin: first min, NaN wins
    const bool an = a.v != a.v, bn = b.v != b.v;
    if (an || bn) return an && (!bn || a.idx < b.idx);
    return a.v < b.v || (a.v == b.v && a.idx < b.idx);
}

__global__ void __launch_bounds__(1024, 1)
mse_argmin_kernel(const double* __restrict__ loss, int32_t n, const float* __restrict__ cxmin,
                  const float* __restrict__ cxmax, float* __restrict__ xmin_out,
                  float* __restrict__ xmax_out, int32_t* __restrict__ idx_out) {
    __shared__ Best sb[32];
    Best b{__longlong_as_double(0x7ff0000000000000LL), 0x7fffffff};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const Best c{loss[i], i};
        if (better(c, b)) b = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best c;
        c.v = __shfl_xor_sync(0xffffffffu, b.v, o);
        c.idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
        if (better(c, b)) b = c;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sb[wid] = b;
    __syncthreads();
    if (wid == 0) {
        b = sb[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Best c;
            c.v = __shfl_xor_sync(0xffffffffu, b.v, o);
            c.idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
            if (better(c, b)) b = c;
        }
        if (lane == 0) {
            const int idx = b.idx == 0x7fffffff ? 0 : b.idx;
            if (idx_out != nullptr) *idx_out = idx;
            if (xmin_out != nullptr) *xmin_out = cxmin[idx];
            if (xmax_out != nullptr) *xmax_out = cxmax[idx];
        }
    }
}

static size_t mse_smem_bytes() {
    return (size_t)kSliceMax * 4 + (size_t)kCandChunk * sizeof(QP) + (size_t)(kMThreads / 32) * kCandChunk * 8;
}

}  // namespace tq

extern "C" {

size_t tq_mse_workspace_bytes(int32_t n_cand) {
    if (n_cand < 1) n_cand = 1;
    return (size_t)tq::sm_count() * (size_t)n_cand * sizeof(double);
}

int tq_mse_sse_f32(const float* x, int64_t n, const float* cand, int32_t n_cand, double* loss_accum,
                   void* ws, size_t ws_bytes, void* stream) {
    if (x == nullptr || cand == nullptr || loss_accum == nullptr || ws == nullptr) return TQ_EINVAL;
    if (n < 1 || n_cand < 1) return TQ_EINVAL;
    if (ws_bytes < tq_mse_workspace_bytes(n_cand)) return TQ_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = tq::mse_smem_bytes();
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tq::mse_sse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    int64_t grid = (n + 2047) / 2048;
    if (grid > tq::sm_count()) grid = tq::sm_count();
    int64_t per_cta = (n + grid - 1) / grid;
    per_cta = (per_cta + 3) & ~(int64_t)3;
    grid = (n + per_cta - 1) / per_cta;
    const int vec_ok = tq::aligned16(x) ? 1 : 0;
    tq::mse_sse_kernel<<<(int)grid, tq::kMThreads, smem, st>>>(x, n, per_cta, vec_ok, cand, n_cand,
                                                             (double*)ws);
    if (int e = tq::launch_status()) return e;
    tq::mse_colsum_kernel<<<(n_cand + 255) / 256, 256, 0, st>>>((const double*)ws, (int)grid, n_cand,
                                                              loss_accum);
    return tq::launch_status();
}

int tq_mse_argmin_f64(const double* loss, int32_t n_cand, const float* cand_xmin, const float* cand_xmax,
                      float* xmin_out, float* xmax_out, int32_t* idx_out, void* stream) {
    if (loss == nullptr || n_cand < 1) return TQ_EINVAL;
    if ((xmin_out != nullptr && cand_xmin == nullptr) || (xmax_out != nullptr && cand_xmax == nullptr))
        return TQ_EINVAL;
    tq::mse_argmin_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(loss, n_cand, cand_xmin, cand_xmax, xmin_out,
                                                              xmax_out, idx_out);
    return tq::launch_status();
}

}  // extern "C"
