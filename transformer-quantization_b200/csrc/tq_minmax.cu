// Range-estimation reductions: one-pass min/max, per tensor,
// per axis (no transpose copy) and per embedding group, plus the estimator state updates and the
// device-side set_quant_range (rows a4/a5) -- everything the calibration loop needs without a
// single host synchronisation.
//
// Reduction scheme: thread-local fminf/fmaxf over 128-bit loads -> warp shuffles -> one
// order-preserving-integer atomicMax per (CTA, channel) into the workspace -> the last CTA (ticket)
// decodes to fp32, writes the result and re-zeroes the workspace.  min/max are exactly associative
// and commutative, so the atomics are deterministic.  NaN propagates like torch.min/max.
#include "tq_common.cuh"

namespace tq {

constexpr int kRThreads = 256;
constexpr int kRUnroll = 4;

// workspace: word 0 = ticket; words 4.. : [C] ~ord(min) | [C] ord(max) | [C] nan flag  (all zero idle)
struct MMWs {
    uint32_t* ticket;
    uint32_t* emin;
    uint32_t* emax;
    uint32_t* nanf;
};
__host__ __device__ inline MMWs mm_ws(void* ws, int64_t C) {
    uint32_t* w = reinterpret_cast<uint32_t*>(ws);
    MMWs r;
    r.ticket = w;
    r.emin = w + 4;
    r.emax = w + 4 + C;
    r.nanf = w + 4 + 2 * C;
    return r;
}

__device__ __forceinline__ void publish(const MMWs& w, int64_t c, float mn, float mx, bool nan) {
    atomicMax(w.emin + c, ~f2ord(mn));
    atomicMax(w.emax + c, f2ord(mx));
    if (nan) atomicOr(w.nanf + c, 1u);
}

// last CTA: decode all channels, reset workspace
__device__ void finalize(const MMWs& w, int64_t C, float* mn, float* mx, int64_t mx_stride,
                         uint32_t total_ctas) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const uint32_t t = atomicAdd(w.ticket, 1u);
        is_last = (t == total_ctas - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y;
    for (int64_t c = tid; c < C; c += nthr) {
        const uint32_t a = atomicExch(w.emin + c, 0u);
        const uint32_t b = atomicExch(w.emax + c, 0u);
        const uint32_t f = atomicExch(w.nanf + c, 0u);
        float vmin = ord2f(~a), vmax = ord2f(b);
        if (f) vmin = vmax = __int_as_float(0x7fc00000);
        mn[c] = vmin;
        mx[c * mx_stride] = vmax;
    }
    if (tid == 0) *w.ticket = 0u;
}

// Column kernels: ONE TICKET PER COLUMN BLOCK instead of one per launch -- the last CTA of a block's slabs decodes that
// block's columns.  (A launch-wide ticket made 768 CTAs queue on one address: ~27 cycles per same-address atomic =
// 10 us of a 25 us launch.)  The ticket lives in bits 8.. of the block's first NaN-flag word (bit 0 stays the flag), so
// the workspace layout and size are unchanged; the decoding pass re-zeroes it with the flags.
__device__ void finalize_colblock(const MMWs& w, int64_t c0, int64_t c1, float* mn, float* mx, uint32_t slabs) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) {
        const uint32_t t = atomicAdd(w.nanf + c0, 256u) >> 8;
        is_last = (t == slabs - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int nthr = blockDim.x * blockDim.y;
    for (int64_t c = c0 + tid; c < c1; c += nthr) {
        const uint32_t a = atomicExch(w.emin + c, 0u);
        const uint32_t b = atomicExch(w.emax + c, 0u);
        const uint32_t f = atomicExch(w.nanf + c, 0u) & 1u;
        float vmin = ord2f(~a), vmax = ord2f(b);
        if (f) vmin = vmax = __int_as_float(0x7fc00000);
        mn[c] = vmin;
        mx[c] = vmax;
    }
}

__device__ __forceinline__ void acc4(const float4& v, float& mn, float& mx, bool& nan) {
    mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
    mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
    nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
}

__device__ __forceinline__ void block_reduce_publish(float mn, float mx, bool nan, const MMWs& w,
                                                     int64_t c) {
    __shared__ float smn[32], smx[32];
    __shared__ int snan;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    if (tid == 0) snan = 0;
    __syncthreads();
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (nan) atomicOr(&snan, 1);
    if (lane == 0) {
        smn[wid] = mn;
        smx[wid] = mx;
    }
    __syncthreads();
    if (wid == 0) {
        mn = lane < nw ? smn[lane] : __int_as_float(0x7f800000);
        mx = lane < nw ? smx[lane] : __int_as_float(0xff800000);
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) publish(w, c, mn, mx, snan != 0);
    }
}

// ---- per-tensor -----------------------------------------------------------------------------
__global__ void __launch_bounds__(kRThreads, 4)
minmax_tensor_kernel(const float* __restrict__ x, int64_t n, int vec_ok, float* __restrict__ out,
                     void* ws) {
    const MMWs w = mm_ws(ws, 1);
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    bool nan = false;
    if (vec_ok) {
        const int64_t nvec = n >> 2;
        const float4* xv = reinterpret_cast<const float4*>(x);
        const int64_t stride = (int64_t)gridDim.x * kRThreads * kRUnroll;
        for (int64_t base = (int64_t)blockIdx.x * kRThreads * kRUnroll + threadIdx.x; base < nvec;
             base += stride) {
            float4 v[kRUnroll];
#pragma unroll
            for (int u = 0; u < kRUnroll; ++u) {
                const int64_t idx = base + (int64_t)u * kRThreads;
                if (idx < nvec) v[u] = ld_stream(xv + idx);
            }
#pragma unroll
            for (int u = 0; u < kRUnroll; ++u) {
                const int64_t idx = base + (int64_t)u * kRThreads;
                if (idx < nvec) acc4(v[u], mn, mx, nan);
            }
        }
        if (blockIdx.x == 0) {
            const int64_t i = (nvec << 2) + threadIdx.x;
            if (i < n) {
                const float v = x[i];
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
                nan |= (v != v);
            }
        }
    } else {
        const int64_t stride = (int64_t)gridDim.x * kRThreads;
        for (int64_t i = (int64_t)blockIdx.x * kRThreads + threadIdx.x; i < n; i += stride) {
            const float v = x[i];
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
            nan |= (v != v);
        }
    }
    block_reduce_publish(mn, mx, nan, w, 0);
    finalize(w, 1, out, out + 1, 1, gridDim.x);
}

// ---- per-axis, inner == 1: x viewed [rows, C]; block = 64 vector columns x 4 rows --------------
__global__ void __launch_bounds__(256, 4)
minmax_cols_vec_kernel(const float* __restrict__ x, int64_t rows, int32_t C, int64_t rows_per_slab,
                       float* __restrict__ mn_out, float* __restrict__ mx_out, void* ws) {
    const MMWs w = mm_ws(ws, C);
    const int32_t CV = C >> 2;
    const int32_t vc = blockIdx.x * 64 + threadIdx.x;
    const float inf = __int_as_float(0x7f800000);
    float4 mn = make_float4(inf, inf, inf, inf), mx = make_float4(-inf, -inf, -inf, -inf);
    bool nan = false;
    if (vc < CV) {
        const float4* xv = reinterpret_cast<const float4*>(x);
        const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
        int64_t r1 = r0 + rows_per_slab;
        if (r1 > rows) r1 = rows;
        int64_t r = r0 + threadIdx.y;
        for (; r + 28 < r1; r += 32) {                 // eight 16-byte loads in flight per thread (~35 KB per SM needed)
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = ld_stream(xv + (r + 4 * u) * CV + vc);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                mn.x = fminf(mn.x, v[u].x); mn.y = fminf(mn.y, v[u].y);
                mn.z = fminf(mn.z, v[u].z); mn.w = fminf(mn.w, v[u].w);
                mx.x = fmaxf(mx.x, v[u].x); mx.y = fmaxf(mx.y, v[u].y);
                mx.z = fmaxf(mx.z, v[u].z); mx.w = fmaxf(mx.w, v[u].w);
                nan |= (v[u].x != v[u].x) | (v[u].y != v[u].y) | (v[u].z != v[u].z) | (v[u].w != v[u].w);
            }
        }
        for (; r + 12 < r1; r += 16) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld_stream(xv + (r + 4 * u) * CV + vc);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                mn.x = fminf(mn.x, v[u].x); mn.y = fminf(mn.y, v[u].y);
                mn.z = fminf(mn.z, v[u].z); mn.w = fminf(mn.w, v[u].w);
                mx.x = fmaxf(mx.x, v[u].x); mx.y = fmaxf(mx.y, v[u].y);
                mx.z = fmaxf(mx.z, v[u].z); mx.w = fmaxf(mx.w, v[u].w);
                nan |= (v[u].x != v[u].x) | (v[u].y != v[u].y) | (v[u].z != v[u].z) | (v[u].w != v[u].w);
            }
        }
        for (; r < r1; r += 4) {
            const float4 v = ld_stream(xv + r * CV + vc);
            mn.x = fminf(mn.x, v.x); mn.y = fminf(mn.y, v.y);
            mn.z = fminf(mn.z, v.z); mn.w = fminf(mn.w, v.w);
            mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y);
            mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
            nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
        }
    }
    __shared__ float4 smn[4][64], smx[4][64];
    __shared__ int snan[64];
    if (threadIdx.y == 0) snan[threadIdx.x] = 0;
    smn[threadIdx.y][threadIdx.x] = mn;
    smx[threadIdx.y][threadIdx.x] = mx;
    __syncthreads();
    // a NaN anywhere in a column poisons only that column (torch semantics); conservatively flag
    // the 4 columns of the vector lane that saw it, then refine per component below.
    if (nan) atomicOr(&snan[threadIdx.x], 1);
    __syncthreads();
    if (threadIdx.y == 0 && vc < CV) {
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            const float4 a = smn[j][threadIdx.x], b = smx[j][threadIdx.x];
            mn.x = fminf(mn.x, a.x); mn.y = fminf(mn.y, a.y); mn.z = fminf(mn.z, a.z); mn.w = fminf(mn.w, a.w);
            mx.x = fmaxf(mx.x, b.x); mx.y = fmaxf(mx.y, b.y); mx.z = fmaxf(mx.z, b.z); mx.w = fmaxf(mx.w, b.w);
        }
        const int64_t c = (int64_t)vc * 4;
        const bool nf = snan[threadIdx.x] != 0;
        publish(w, c + 0, mn.x, mx.x, false);
        publish(w, c + 1, mn.y, mx.y, false);
        publish(w, c + 2, mn.z, mx.z, false);
        publish(w, c + 3, mn.w, mx.w, false);
        if (nf) {  // rare: re-scan this CTA's slab per component to attribute the NaN exactly
            const float4* xv = reinterpret_cast<const float4*>(x);
            const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
            int64_t r1 = r0 + rows_per_slab;
            if (r1 > rows) r1 = rows;
            bool n0 = false, n1 = false, n2 = false, n3 = false;
            for (int64_t r = r0; r < r1; ++r) {
                const float4 v = xv[r * CV + vc];
                n0 |= (v.x != v.x); n1 |= (v.y != v.y); n2 |= (v.z != v.z); n3 |= (v.w != v.w);
            }
            if (n0) atomicOr(w.nanf + c + 0, 1u);
            if (n1) atomicOr(w.nanf + c + 1, 1u);
            if (n2) atomicOr(w.nanf + c + 2, 1u);
            if (n3) atomicOr(w.nanf + c + 3, 1u);
        }
    }
    const int64_t c0 = (int64_t)blockIdx.x * 256;
    finalize_colblock(w, c0, c0 + 256 < C ? c0 + 256 : C, mn_out, mx_out, gridDim.y);
}

// ---- per-axis, general [outer, C, inner]: one CTA per (channel, outer-split) --------------------
__global__ void __launch_bounds__(kRThreads, 4)
minmax_rows_kernel(const float* __restrict__ x, int64_t outer, int64_t C, int64_t inner, int vec_ok,
                   float* __restrict__ mn_out, float* __restrict__ mx_out, void* ws) {
    const MMWs w = mm_ws(ws, C);
    for (int64_t c = blockIdx.x; c < C; c += gridDim.x) {
        float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
        bool nan = false;
        for (int64_t o = blockIdx.y; o < outer; o += gridDim.y) {
            const float* row = x + (o * C + c) * inner;
            if (vec_ok) {
                const float4* xv = reinterpret_cast<const float4*>(row);
                const int64_t iv = inner >> 2;
                for (int64_t i = threadIdx.x; i < iv; i += kRThreads) acc4(ld_stream(xv + i), mn, mx, nan);
            } else {
                for (int64_t i = threadIdx.x; i < inner; i += kRThreads) {
                    const float v = row[i];
                    mn = fminf(mn, v);
                    mx = fmaxf(mx, v);
                    nan |= (v != v);
                }
            }
        }
        block_reduce_publish(mn, mx, nan, w, c);
        __syncthreads();
    }
    finalize(w, C, mn_out, mx_out, 1, gridDim.x * gridDim.y);
}

// scalar column kernel (inner == 1, C not a multiple of 4 or misaligned): thread per column
__global__ void __launch_bounds__(kRThreads, 4)
minmax_cols_scalar_kernel(const float* __restrict__ x, int64_t rows, int64_t C, int64_t rows_per_slab,
                          float* __restrict__ mn_out, float* __restrict__ mx_out, void* ws) {
    const MMWs w = mm_ws(ws, C);
    const int64_t c = (int64_t)blockIdx.x * kRThreads + threadIdx.x;
    if (c < C) {
        float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
        bool nan = false;
        const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
        int64_t r1 = r0 + rows_per_slab;
        if (r1 > rows) r1 = rows;
        for (int64_t r = r0; r < r1; ++r) {
            const float v = x[r * C + c];
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
            nan |= (v != v);
        }
        if (r1 > r0) publish(w, c, mn, mx, nan);
    }
    finalize(w, C, mn_out, mx_out, 1, gridDim.x * gridDim.y);
}

// ---- per-embedding-group statistics on the [C] vectors ------------------------------------------
__global__ void __launch_bounds__(1024, 1)
group_minmax_kernel(const float* __restrict__ mn, const float* __restrict__ mx, int32_t C,
                    int32_t n_groups, const float* __restrict__ ranges, float* __restrict__ mn_out,
                    float* __restrict__ mx_out) {
    extern __shared__ uint32_t gsm[];   // [G] ~ord(min) | [G] ord(max) | [G] nan | [C] group id
    uint32_t* gmin = gsm;
    uint32_t* gmax = gsm + n_groups;
    uint32_t* gnan = gsm + 2 * n_groups;
    int32_t* gid = reinterpret_cast<int32_t*>(gsm + 3 * n_groups);
    const int32_t gs = C / n_groups;
    for (int g = threadIdx.x; g < 3 * n_groups; g += blockDim.x) gsm[g] = 0u;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        int32_t pos = c;
        if (ranges != nullptr) {           // stable rank of dim c in ascending range order
            const float rc = ranges[c];
            int32_t rank = 0;
            for (int j = 0; j < C; ++j) {
                const float rj = ranges[j];
                rank += (rj < rc) || (rj == rc && j < c);
            }
            pos = rank;
        }
        const int32_t g = pos / gs;
        gid[c] = g;
        const float a = mn[c], b = mx[c];
        atomicMax(gmin + g, ~f2ord(a));
        atomicMax(gmax + g, f2ord(b));
        if (a != a || b != b) atomicOr(gnan + g, 1u);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int32_t g = gid[c];
        float a = ord2f(~gmin[g]), b = ord2f(gmax[g]);
        if (gnan[g]) a = b = __int_as_float(0x7fc00000);
        mn_out[c] = a;
        mx_out[c] = b;
    }
}

__global__ void dim_ranges_kernel(const float* __restrict__ mn, const float* __restrict__ mx, int64_t C,
                                  int first, float* __restrict__ ranges) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float r = __fsub_rn(mx[c], mn[c]);
    if (!first) r = __fadd_rn(__fmul_rn(0.1f, r), __fmul_rn((float)(1 - 0.1), r));  // range_estimators.py:78-79
    ranges[c] = r;
}

__global__ void range_update_kernel(const float* __restrict__ nmin, const float* __restrict__ nmax,
                                    float* __restrict__ cmin, float* __restrict__ cmax, int64_t k,
                                    int mode, float m_new, float m_old, int first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const float a = nmin[i], b = nmax[i];
    if (mode == 0 || first) {
        cmin[i] = a;
        cmax[i] = b;
    } else if (mode == 1) {   // (1 - momentum) * new + momentum * cur, range_estimators.py:213-214
        cmin[i] = __fadd_rn(__fmul_rn(m_new, a), __fmul_rn(m_old, cmin[i]));
        cmax[i] = __fadd_rn(__fmul_rn(m_new, b), __fmul_rn(m_old, cmax[i]));
    } else {                  // torch.min/max(cur, new), range_estimators.py:166-167 (NaN propagates)
        const float c0 = cmin[i], c1 = cmax[i];
        cmin[i] = (a != a || c0 != c0) ? __int_as_float(0x7fc00000) : fminf(c0, a);
        cmax[i] = (b != b || c1 != c1) ? __int_as_float(0x7fc00000) : fmaxf(c1, b);
    }
}

// torch.min(x_min, 0) / torch.max(x_max, eps), quantizers.py:258-259 (NaN propagates)
__device__ __forceinline__ float tmin0(float v) { return (v != v) ? v : fminf(v, 0.0f); }
__device__ __forceinline__ float tmaxe(float v, float eps) { return (v != v) ? v : fmaxf(v, eps); }

__global__ void set_range_asym_kernel(const float* __restrict__ xmin, const float* __restrict__ xmax,
                                      int64_t k, float int_max, float eps, int log_domain,
                                      float* __restrict__ delta, float* __restrict__ zero_float) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const float a = tmin0(xmin[i]), b = tmaxe(xmax[i], eps);
    const float d = __fdiv_rn(__fsub_rn(b, a), int_max);      // quantizers.py:276
    zero_float[i] = __fdiv_rn(-a, d);                         // quantizers.py:277
    delta[i] = log_domain ? logf(d) : d;                      // quantizers.py:279-280
}

__global__ void __launch_bounds__(1024, 1)
set_range_sym_kernel(const float* __restrict__ xmin, const float* __restrict__ xmax, int64_t k,
                     int n_bits, float eps, int log_domain, float* __restrict__ delta,
                     uint8_t* __restrict__ is_signed) {
    int neg = 0;
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) neg |= (tmin0(xmin[i]) < 0.0f);
    const int sg = __syncthreads_or(neg);                     // quantizers.py:336
    const float int_max = (float)(1u << (n_bits - (sg ? 1 : 0))) - 1.0f;
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        const float a = tmin0(xmin[i]), b = tmaxe(xmax[i], eps);
        const float fa = fabsf(a);
        const float am = (fa != fa || b != b) ? __int_as_float(0x7fc00000) : fmaxf(fa, b);  // :338
        const float d = __fdiv_rn(am, int_max);               // quantizers.py:339
        delta[i] = log_domain ? logf(d) : d;
    }
    if (threadIdx.x == 0) *is_signed = sg ? 1 : 0;
}

// Rows per CTA of the column kernels.  Every CTA ends with one atomic on the shared ticket and two on each of its
// columns' words, and same-address L2 atomics serialise (~27 cycles each): 768 CTAs on a 12.6 MB activation spent
// ~10 us queueing on the ticket alone (22.5 us per launch).  So the CTA count follows the tensor size -- >= 128 KB of
// input per CTA, between one CTA per SM and eight (and the vector kernel keeps one ticket per column block).
static int64_t slab_rows(int64_t rows, int64_t col_blocks, int unit, int64_t total_bytes) {
    int64_t target_ctas = total_bytes / 131072;
    const int64_t lo = sm_count(), hi = (int64_t)sm_count() * 8;
    target_ctas = target_ctas < lo ? lo : (target_ctas > hi ? hi : target_ctas);
    int64_t slabs = target_ctas / (col_blocks > 0 ? col_blocks : 1);
    if (slabs < 1) slabs = 1;
    int64_t rps = (rows + slabs - 1) / slabs;
    rps = ((rps + unit - 1) / unit) * unit;
    if (rps < unit) rps = unit;
    return rps;
}


// Calibration-time fused GEMM (SURVEY.md section 7 "hard parts", reference quantization_manager.py:99-106 after
// hijacker.py:98-116): the GEMM epilogue reduced min / max of its own output into two ordered-int words
// (tile_minmax); this ONE single-thread launch decodes them, applies the estimator update (current / running EMA /
// all-time min-max, range_estimators.py:142-143, 166-167, 205-214), and sets the per-tensor quantizer range
// (quantizers.py:263-282 asymmetric, 334-344 symmetric) -- the min/max pass over the tensor and two scalar launches
// are gone.  The two words are reset for the next calibration batch.
__global__ void calib_finalize_kernel(uint32_t* __restrict__ tile_mm, float* __restrict__ cmin, float* __restrict__ cmax,
                                      int mode, float m_new, float m_old, int first, int symmetric, int n_bits, float eps,
                                      int log_domain, float* __restrict__ delta, float* __restrict__ zero_float,
                                      uint8_t* __restrict__ is_signed) {
    const float a0 = ord2f(~tile_mm[0]), b0 = ord2f(tile_mm[1]);
    tile_mm[0] = 0u;
    tile_mm[1] = 0u;
    float mn, mx;
    if (mode == 0 || first) {
        mn = a0;
        mx = b0;
    } else if (mode == 1) {
        mn = __fadd_rn(__fmul_rn(m_new, a0), __fmul_rn(m_old, cmin[0]));
        mx = __fadd_rn(__fmul_rn(m_new, b0), __fmul_rn(m_old, cmax[0]));
    } else {
        const float c0 = cmin[0], c1 = cmax[0];
        mn = (a0 != a0 || c0 != c0) ? __int_as_float(0x7fc00000) : fminf(c0, a0);
        mx = (b0 != b0 || c1 != c1) ? __int_as_float(0x7fc00000) : fmaxf(c1, b0);
    }
    cmin[0] = mn;
    cmax[0] = mx;
    const float a = tmin0(mn), b = tmaxe(mx, eps);
    if (!symmetric) {
        const float int_max = (float)(1u << n_bits) - 1.0f;
        const float d = __fdiv_rn(__fsub_rn(b, a), int_max);
        zero_float[0] = __fdiv_rn(-a, d);
        delta[0] = log_domain ? logf(d) : d;
    } else {
        const int sg = a < 0.0f;
        const float int_max = (float)(1u << (n_bits - (sg ? 1 : 0))) - 1.0f;
        const float fa = fabsf(a);
        const float am = (fa != fa || b != b) ? __int_as_float(0x7fc00000) : fmaxf(fa, b);
        const float d = __fdiv_rn(am, int_max);
        delta[0] = log_domain ? logf(d) : d;
        *is_signed = sg ? 1 : 0;
    }
}

}  // namespace tq

extern "C" {

size_t tq_minmax_workspace_bytes(int64_t C) {
    if (C < 1) C = 1;
    return (size_t)(4 + 3 * C) * sizeof(uint32_t);
}

int tq_minmax_f32(const float* x, int64_t n, float* out, void* ws, size_t ws_bytes, void* stream) {
    if (x == nullptr || out == nullptr || ws == nullptr || n < 1) return TQ_EINVAL;
    if (ws_bytes < tq_minmax_workspace_bytes(1)) return TQ_EWORKSPACE;
    const int vec_ok = tq::aligned16(x) ? 1 : 0;
    const int64_t per_block = (int64_t)tq::kRThreads * (vec_ok ? 4 * tq::kRUnroll : 1);
    int64_t blocks = (n + per_block - 1) / per_block;
    // CTA count by tensor size (>= 64 KB per CTA, one to four CTAs per SM): every CTA queues on the shared ticket and on
    // the two result words with same-address atomics (see slab_rows)
    int64_t cap = n * 4 / 65536;
    const int64_t lo = tq::sm_count(), hi = (int64_t)tq::sm_count() * 4;
    cap = cap < lo ? lo : (cap > hi ? hi : cap);
    if (blocks > cap) blocks = cap;
    tq::minmax_tensor_kernel<<<(int)blocks, tq::kRThreads, 0, (cudaStream_t)stream>>>(x, n, vec_ok, out, ws);
    return tq::launch_status();
}

int tq_minmax_axis_f32(const float* x, int64_t outer, int64_t C, int64_t inner, float* mn, float* mx,
                       void* ws, size_t ws_bytes, void* stream) {
    if (x == nullptr || mn == nullptr || mx == nullptr || ws == nullptr) return TQ_EINVAL;
    if (outer < 1 || C < 1 || inner < 1) return TQ_EINVAL;
    if (ws_bytes < tq_minmax_workspace_bytes(C)) return TQ_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (inner == 1) {
        const int64_t rows = outer;
        if (tq::aligned16(x) && (C & 3) == 0) {
            const int64_t colb = ((C >> 2) + 63) / 64;
            const int64_t rps = tq::slab_rows(rows, colb, 32, rows * C * 4);
            dim3 grid((unsigned)colb, (unsigned)((rows + rps - 1) / rps));
            tq::minmax_cols_vec_kernel<<<grid, dim3(64, 4), 0, st>>>(x, rows, (int32_t)C, rps, mn, mx, ws);
        } else {
            const int64_t colb = (C + tq::kRThreads - 1) / tq::kRThreads;
            const int64_t rps = tq::slab_rows(rows, colb, 8, rows * C * 4);
            dim3 grid((unsigned)colb, (unsigned)((rows + rps - 1) / rps));
            tq::minmax_cols_scalar_kernel<<<grid, tq::kRThreads, 0, st>>>(x, rows, C, rps, mn, mx, ws);
        }
        return tq::launch_status();
    }
    const int vec_ok = (tq::aligned16(x) && (inner & 3) == 0) ? 1 : 0;
    int64_t gx = C < 65535 ? C : 65535;
    int64_t gy = 1;
    const int64_t target = (int64_t)tq::sm_count() * 8;
    if (gx < target && outer > 1) {
        gy = target / gx;
        if (gy > outer) gy = outer;
        if (gy > 65535) gy = 65535;
        if (gy < 1) gy = 1;
    }
    dim3 grid((unsigned)gx, (unsigned)gy);
    tq::minmax_rows_kernel<<<grid, tq::kRThreads, 0, st>>>(x, outer, C, inner, vec_ok, mn, mx, ws);
    return tq::launch_status();
}

int tq_group_minmax_f32(const float* mn, const float* mx, int64_t C, int32_t n_groups,
                        const float* ranges, float* mn_out, float* mx_out, void* stream) {
    if (mn == nullptr || mx == nullptr || mn_out == nullptr || mx_out == nullptr) return TQ_EINVAL;
    if (C < 1 || n_groups < 1 || C % n_groups != 0) return TQ_EINVAL;
    const size_t smem = (size_t)(3 * n_groups + C) * 4;
    if (smem > 200 * 1024) return TQ_EUNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(tq::group_minmax_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    tq::group_minmax_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(mn, mx, (int32_t)C, n_groups, ranges,
                                                                    mn_out, mx_out);
    return tq::launch_status();
}

int tq_dim_ranges_f32(const float* mn, const float* mx, int64_t C, int32_t first, float* ranges,
                      void* stream) {
    if (mn == nullptr || mx == nullptr || ranges == nullptr || C < 1) return TQ_EINVAL;
    tq::dim_ranges_kernel<<<(unsigned)((C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mn, mx, C, first, ranges);
    return tq::launch_status();
}

int tq_range_update_f32(const float* new_min, const float* new_max, float* cur_min, float* cur_max,
                        int64_t k, int32_t mode, double momentum, int32_t first, void* stream) {
    if (new_min == nullptr || new_max == nullptr || cur_min == nullptr || cur_max == nullptr || k < 1)
        return TQ_EINVAL;
    if (mode < 0 || mode > 2) return TQ_EINVAL;
    // python: (1 - momentum) and momentum are doubles that torch casts to fp32 (range_estimators.py:213)
    const float m_new = (float)(1.0 - momentum), m_old = (float)momentum;
    tq::range_update_kernel<<<(unsigned)((k + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        new_min, new_max, cur_min, cur_max, k, mode, m_new, m_old, first);
    return tq::launch_status();
}

int tq_set_range_asym_f32(const float* x_min, const float* x_max, int64_t k, int32_t n_bits, float eps,
                          int32_t log_domain, float* delta, float* zero_float, void* stream) {
    if (x_min == nullptr || x_max == nullptr || delta == nullptr || zero_float == nullptr || k < 1)
        return TQ_EINVAL;
    if (n_bits < 1 || n_bits > 16) return TQ_EINVAL;
    const float int_max = (float)(1u << n_bits) - 1.0f;
    tq::set_range_asym_kernel<<<(unsigned)((k + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x_min, x_max, k, int_max, eps, log_domain, delta, zero_float);
    return tq::launch_status();
}

int tq_set_range_sym_f32(const float* x_min, const float* x_max, int64_t k, int32_t n_bits, float eps,
                         int32_t log_domain, float* delta, uint8_t* is_signed, void* stream) {
    if (x_min == nullptr || x_max == nullptr || delta == nullptr || is_signed == nullptr || k < 1)
        return TQ_EINVAL;
    if (n_bits < 1 || n_bits > 16) return TQ_EINVAL;
    tq::set_range_sym_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x_min, x_max, k, n_bits, eps, log_domain,
                                                                  delta, is_signed);
    return tq::launch_status();
}


int tq_calib_finalize_f32(void* tile_minmax, float* cur_min, float* cur_max, int32_t mode, double momentum, int32_t first,
                          int32_t symmetric, int32_t n_bits, float eps, int32_t log_domain, float* delta, float* zero_float,
                          void* is_signed, void* stream) {
    if (tile_minmax == nullptr || cur_min == nullptr || cur_max == nullptr || delta == nullptr) return TQ_EINVAL;
    if (mode < 0 || mode > 2 || n_bits < 1 || n_bits > 16) return TQ_EINVAL;
    if (symmetric ? is_signed == nullptr : zero_float == nullptr) return TQ_EINVAL;
    const float m_new = (float)(1.0 - momentum), m_old = (float)momentum;
    tq::calib_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint32_t*>(tile_minmax), cur_min, cur_max, mode,
                                                                m_new, m_old, first, symmetric, n_bits, eps, log_domain, delta,
                                                                zero_float, reinterpret_cast<uint8_t*>(is_signed));
    return tq::launch_status();
}

}  // extern "C"
