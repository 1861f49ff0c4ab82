// Training-time quantizer kernels.
//
//  * tq_qdq_bwd_f32: autograd of AsymmetricUniformQuantizer.forward / SymmetricUniformQuantizer.forward
//    (reference quantization/quantizers.py:142-153, 172-211 with the straight-through round of :12-20)
//    in ONE pass: reads x and the upstream gradient once, writes grad_x, and reduces the gradients of
//    the learnable range parameters (`_delta`, `_zero_float`; make_range_trainable, :284-288, 346-349).
//    The reference's autograd graph runs ~14 elementwise / reduction kernels over the tensor.
//    HBM bound: 12 B / element (x 4 R + g 4 R + grad_x 4 W).
//  * tq_adaround_*: AdaRoundQuantizer.to_integer_forward in the relaxation modes
//    (quantization/adaround/quantizer.py:46-92): alpha initialisation, soft / hard targets, gradient
//    of alpha.  Elementwise over weight tensors.
//
// Reductions are deterministic: fp32 per thread, fp64 across threads and CTAs, CTA partials in the
// caller's workspace, summed in a fixed order by the last CTA (ticket).
#include "tq_common.cuh"

namespace tq {

constexpr int kBThreads = 256;
constexpr int kBUnroll = 4;

// ---- per-element backward ------------------------------------------------------------------------
//   t = x / s; u = rint(t) + zp; in = lo <= u <= hi; w = clamp(u) - zp          (forward, :184-185)
//   h = g * s                      MulBackward of  s * (x_int - zp)              (:209)
//   hm = in ? h : 0                ClampBackward
//   gx = hm / s                    DivBackward (self)                            (:184)
//   ds = g * w - hm * ((x / s) / s)   both uses of `scale`                       (:184, :209)
//   dz = hm - h                    both uses of `zero_point`                     (:184, :209)
template <bool FAST>
__device__ __forceinline__ void bwd_elem(float x, float g, const QP& p, float& gx, float& ds, float& dz) {
    const float t = div_rn_t<FAST>(x, p);
    const float u = __fadd_rn(rint_even(t), p.zp);
    const bool in = (u >= p.lo) && (u <= p.hi);          // false for NaN, like torch's clamp mask
    float xi = u < p.lo ? p.lo : u;
    xi = xi > p.hi ? p.hi : xi;
    const float w = __fsub_rn(xi, p.zp);
    const float h = __fmul_rn(g, p.scale);
    const float hm = in ? h : 0.0f;
    gx = div_rn_t<FAST>(hm, p);
    const float t2 = div_rn_t<FAST>(t, p);
    ds += __fsub_rn(__fmul_rn(g, w), __fmul_rn(hm, t2));
    dz += __fsub_rn(hm, h);
}

// d scale / d delta (quantizers.py:142-147) and d zero_point / d zero_float (:149-153)
__device__ __forceinline__ float delta_grad(const tq_qspec& q, int64_t c, float gs) {
    const float d = q.delta[c];
    if (q.log_domain) return __fmul_rn(gs, expf(d));
    return d >= q.eps ? gs : 0.0f;
}
__device__ __forceinline__ float zf_grad(const tq_qspec& q, int64_t c, float gz, float lo, float hi) {
    const float r = rintf(q.zero_float[c]);
    return (r >= lo && r <= hi) ? gz : 0.0f;
}

// sum over the CTA, result valid in thread 0 (fixed tree: deterministic)
__device__ __forceinline__ double block_sum(double v, double* sm) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                       // sm may still be read from a previous call
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < nw ? sm[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// ticket: true in every thread of the LAST CTA to arrive (its reads see all other CTAs' partials)
__device__ __forceinline__ bool last_cta(uint32_t* ticket, uint32_t total) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) is_last = (atomicAdd(ticket, 1u) == total - 1);
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

// Workspace layout shared by all variants (the same caller buffer serves every call on a stream, so the
// ticket words must never be used as partial-sum storage): bytes 0..1023 = 256 ticket words, zero when idle
// (word 0: per-tensor kernel; word 4 + b: column block b of the column kernel), partial sums from byte 1024.
constexpr size_t kBwdHeader = 1024;
constexpr int64_t kBwdMaxColBlocks = 252;
__host__ __device__ inline double* bwd_partials(void* ws) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + kBwdHeader);
}

// ---- per-tensor ------------------------------------------------------------------------------------
template <bool FAST>
__device__ __forceinline__ void bwd_tensor_vec(const float4* __restrict__ xv, const float4* __restrict__ gv,
                                               float4* __restrict__ gxv, int64_t nvec, const QP& p, float& ds,
                                               float& dz) {
    const int64_t stride = (int64_t)gridDim.x * kBThreads * kBUnroll;
    for (int64_t base = (int64_t)blockIdx.x * kBThreads * kBUnroll + threadIdx.x; base < nvec; base += stride) {
        float4 a[kBUnroll], b[kBUnroll];
#pragma unroll
        for (int u = 0; u < kBUnroll; ++u) {
            const int64_t idx = base + (int64_t)u * kBThreads;
            if (idx < nvec) {
                a[u] = ld_stream(xv + idx);
                b[u] = ld_stream(gv + idx);
            }
        }
#pragma unroll
        for (int u = 0; u < kBUnroll; ++u) {
            const int64_t idx = base + (int64_t)u * kBThreads;
            if (idx < nvec) {
                float4 o;
                bwd_elem<FAST>(a[u].x, b[u].x, p, o.x, ds, dz);
                bwd_elem<FAST>(a[u].y, b[u].y, p, o.y, ds, dz);
                bwd_elem<FAST>(a[u].z, b[u].z, p, o.z, ds, dz);
                bwd_elem<FAST>(a[u].w, b[u].w, p, o.w, ds, dz);
                if (gxv != nullptr) st_stream(gxv + idx, o);
            }
        }
    }
}

template <bool FAST>
__device__ __forceinline__ void bwd_scalar_range(const float* __restrict__ x, const float* __restrict__ g,
                                                 float* __restrict__ gx, int64_t i0, int64_t i1, int64_t step,
                                                 const QP& p, float& ds, float& dz) {
    for (int64_t i = i0; i < i1; i += step) {
        float o;
        bwd_elem<FAST>(x[i], g[i], p, o, ds, dz);
        if (gx != nullptr) gx[i] = o;
    }
}

__global__ void __launch_bounds__(kBThreads, 3)
qdq_bwd_tensor_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gx, int64_t n,
                      int vec_ok, tq_qspec q, float* __restrict__ grad_delta, float* __restrict__ grad_zf, void* ws) {
    __shared__ double sm[32];
    float lo, hi;
    grid_of(q, lo, hi);
    const QP p = resolve(q, 0, lo, hi);
    float ds = 0.0f, dz = 0.0f;
    if (vec_ok) {
        const int64_t nvec = n >> 2;
        const float4* xv = reinterpret_cast<const float4*>(x);
        const float4* gv = reinterpret_cast<const float4*>(g);
        float4* gxv = reinterpret_cast<float4*>(gx);
        if (p.exact) bwd_tensor_vec<false>(xv, gv, gxv, nvec, p, ds, dz);
        else bwd_tensor_vec<true>(xv, gv, gxv, nvec, p, ds, dz);
        if (blockIdx.x == 0) {                                    // ragged tail (n % 4 elements)
            const int64_t i = (nvec << 2) + threadIdx.x;
            if (p.exact) bwd_scalar_range<false>(x, g, gx, i, n, kBThreads, p, ds, dz);
            else bwd_scalar_range<true>(x, g, gx, i, n, kBThreads, p, ds, dz);
        }
    } else {
        const int64_t i0 = (int64_t)blockIdx.x * kBThreads + threadIdx.x, step = (int64_t)gridDim.x * kBThreads;
        if (p.exact) bwd_scalar_range<false>(x, g, gx, i0, n, step, p, ds, dz);
        else bwd_scalar_range<true>(x, g, gx, i0, n, step, p, ds, dz);
    }
    double* part = bwd_partials(ws);
    const double bs = block_sum((double)ds, sm);
    const double bz = block_sum((double)dz, sm);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = bs;
        part[2 * blockIdx.x + 1] = bz;
    }
    if (!last_cta(reinterpret_cast<uint32_t*>(ws), gridDim.x)) return;
    double as = 0.0, az = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += kBThreads) {
        as += __ldcg(part + 2 * i);
        az += __ldcg(part + 2 * i + 1);
    }
    as = block_sum(as, sm);
    az = block_sum(az, sm);
    if (threadIdx.x == 0) {
        if (grad_delta != nullptr) grad_delta[0] = delta_grad(q, 0, (float)as);
        if (grad_zf != nullptr && q.zero_float != nullptr) grad_zf[0] = zf_grad(q, 0, (float)az, lo, hi);
        *reinterpret_cast<uint32_t*>(ws) = 0u;
    }
}

// Non-persistent variant for tensors far larger than L2: every CTA streams kBwdSpan consecutive 16 KB chunks of
// x and grad_y and retires; CTAs that retire and get replaced one by one keep reads and writes mixed, whereas
// the CTAs of a persistent grid fall into lock step (arithmetic-free probe tq_probe_copy_f32: 6.8 vs 6.0 TB/s).
// One chunk per CTA was tried and lost (4.6 TB/s): a ticket atomic + two block reductions per 16 KB cost more
// than the access pattern gains.  One partial pair per CTA; the last
// CTA sums them in a fixed order.
constexpr int kBwdSpan = 8;
constexpr int64_t kBwdSpanMinDefault = 6144;            // CTAs: measured 6.17 vs 5.54 TB/s at 8192 CTAs (256 Mi elements),
                                                        // 5.41 vs 5.78 at 2048 (64 Mi): only very large tensors gain
__global__ void __launch_bounds__(kBThreads, 3)
qdq_bwd_span_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gx, int64_t n,
                    tq_qspec q, float* __restrict__ grad_delta, float* __restrict__ grad_zf, void* ws) {
    __shared__ double sm[32];
    const int64_t nvec = n >> 2;
    const float4* xv = reinterpret_cast<const float4*>(x);
    const float4* gv = reinterpret_cast<const float4*>(g);
    float4* gxv = reinterpret_cast<float4*>(gx);
    float lo, hi;
    grid_of(q, lo, hi);
    const QP p = resolve(q, 0, lo, hi);
    float ds = 0.0f, dz = 0.0f;
    const int64_t first = (int64_t)blockIdx.x * kBwdSpan * kBThreads * kBUnroll + threadIdx.x;
    for (int c = 0; c < kBwdSpan; ++c) {
        const int64_t base = first + (int64_t)c * kBThreads * kBUnroll;
        if (base - threadIdx.x >= nvec) break;
        float4 a[kBUnroll], b[kBUnroll];
#pragma unroll
        for (int u = 0; u < kBUnroll; ++u) {
            const int64_t idx = base + (int64_t)u * kBThreads;
            if (idx < nvec) {
                a[u] = ld_stream(xv + idx);
                b[u] = ld_stream(gv + idx);
            }
        }
#pragma unroll
        for (int u = 0; u < kBUnroll; ++u) {
            const int64_t idx = base + (int64_t)u * kBThreads;
            if (idx < nvec) {
                float4 o;
                if (p.exact) {
                    bwd_elem<false>(a[u].x, b[u].x, p, o.x, ds, dz);
                    bwd_elem<false>(a[u].y, b[u].y, p, o.y, ds, dz);
                    bwd_elem<false>(a[u].z, b[u].z, p, o.z, ds, dz);
                    bwd_elem<false>(a[u].w, b[u].w, p, o.w, ds, dz);
                } else {
                    bwd_elem<true>(a[u].x, b[u].x, p, o.x, ds, dz);
                    bwd_elem<true>(a[u].y, b[u].y, p, o.y, ds, dz);
                    bwd_elem<true>(a[u].z, b[u].z, p, o.z, ds, dz);
                    bwd_elem<true>(a[u].w, b[u].w, p, o.w, ds, dz);
                }
                if (gxv != nullptr) st_stream(gxv + idx, o);
            }
        }
    }
    if (blockIdx.x == 0) {                                        // ragged tail (n % 4 elements)
        const int64_t i = (nvec << 2) + threadIdx.x;
        if (p.exact) bwd_scalar_range<false>(x, g, gx, i, n, kBThreads, p, ds, dz);
        else bwd_scalar_range<true>(x, g, gx, i, n, kBThreads, p, ds, dz);
    }
    double2* part = reinterpret_cast<double2*>(bwd_partials(ws));
    const double bs = block_sum((double)ds, sm);
    const double bz = block_sum((double)dz, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = make_double2(bs, bz);
    if (!last_cta(reinterpret_cast<uint32_t*>(ws), gridDim.x)) return;
    double as = 0.0, az = 0.0;
    for (unsigned i0 = threadIdx.x; i0 < gridDim.x; i0 += kBThreads * 8) {      // 8 independent L2 loads in flight
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned i = i0 + u * kBThreads;
            v[u] = i < gridDim.x ? __ldcg(part + i) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            as += v[u].x;
            az += v[u].y;
        }
    }
    as = block_sum(as, sm);
    az = block_sum(az, sm);
    if (threadIdx.x == 0) {
        if (grad_delta != nullptr) grad_delta[0] = delta_grad(q, 0, (float)as);
        if (grad_zf != nullptr && q.zero_float != nullptr) grad_zf[0] = zf_grad(q, 0, (float)az, lo, hi);
        *reinterpret_cast<uint32_t*>(ws) = 0u;
    }
}

// ---- per-embedding / per-embedding-group: x viewed [rows, C]; CTA = 64 vector columns x 4 row lanes ---
template <bool FAST>
__device__ __forceinline__ void bwd_vec4(const float4& a, const float4& b, float4& o, const QP (&p)[4], float (&ds)[4],
                                         float (&dz)[4]) {
    bwd_elem<FAST>(a.x, b.x, p[0], o.x, ds[0], dz[0]);
    bwd_elem<FAST>(a.y, b.y, p[1], o.y, ds[1], dz[1]);
    bwd_elem<FAST>(a.z, b.z, p[2], o.z, ds[2], dz[2]);
    bwd_elem<FAST>(a.w, b.w, p[3], o.w, ds[3], dz[3]);
}

constexpr int kColsInFlight = 4;      // rows in flight per thread: 8 x 128-bit loads (2 CTAs/SM -> 64 KB per SM)

template <bool FAST>
__device__ __forceinline__ void bwd_cols_body(const float4* __restrict__ xv, const float4* __restrict__ gv,
                                              float4* __restrict__ gxv, int64_t r0, int64_t r1, int32_t CV, int32_t vc,
                                              const QP (&p)[4], float (&ds)[4], float (&dz)[4]) {
    int64_t r = r0 + threadIdx.y;
    for (; r + 4 * (kColsInFlight - 1) < r1; r += 4 * kColsInFlight) {
        float4 a[kColsInFlight], b[kColsInFlight];
#pragma unroll
        for (int u = 0; u < kColsInFlight; ++u) {
            a[u] = ld_stream(xv + (r + 4 * u) * CV + vc);
            b[u] = ld_stream(gv + (r + 4 * u) * CV + vc);
        }
#pragma unroll
        for (int u = 0; u < kColsInFlight; ++u) {
            float4 o;
            bwd_vec4<FAST>(a[u], b[u], o, p, ds, dz);
            if (gxv != nullptr) st_stream(gxv + (r + 4 * u) * CV + vc, o);
        }
    }
    for (; r < r1; r += 4) {
        const float4 a0 = ld_stream(xv + r * CV + vc), b0 = ld_stream(gv + r * CV + vc);
        float4 o0;
        bwd_vec4<FAST>(a0, b0, o0, p, ds, dz);
        if (gxv != nullptr) st_stream(gxv + r * CV + vc, o0);
    }
}

__global__ void __launch_bounds__(256, 2)
qdq_bwd_cols_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gx, int64_t rows,
                    int32_t C, int64_t rows_per_slab, tq_qspec q, float* __restrict__ grad_delta,
                    float* __restrict__ grad_zf, void* ws) {
    __shared__ float sds[4][64][4], sdz[4][64][4];
    float lo, hi;
    grid_of(q, lo, hi);
    const int32_t CV = C >> 2;
    const int32_t vc = blockIdx.x * 64 + threadIdx.x;
    const bool active = vc < CV;
    QP p[4];
    int need_exact = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        p[j] = resolve(q, active ? (int64_t)vc * 4 + j : 0, lo, hi);
        need_exact |= p[j].exact;
    }
    const int exact = __syncthreads_or(need_exact);
    float ds[4] = {0.f, 0.f, 0.f, 0.f}, dz[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
        const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
        int64_t r1 = r0 + rows_per_slab;
        if (r1 > rows) r1 = rows;
        const float4* xv = reinterpret_cast<const float4*>(x);
        const float4* gv = reinterpret_cast<const float4*>(g);
        float4* gxv = reinterpret_cast<float4*>(gx);
        if (exact) bwd_cols_body<false>(xv, gv, gxv, r0, r1, CV, vc, p, ds, dz);
        else bwd_cols_body<true>(xv, gv, gxv, r0, r1, CV, vc, p, ds, dz);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sds[threadIdx.y][threadIdx.x][j] = ds[j];
        sdz[threadIdx.y][threadIdx.x][j] = dz[j];
    }
    __syncthreads();
    // partial sums of this CTA: [slab][C] pairs {d scale, d zero_point}
    double2* part = reinterpret_cast<double2*>(bwd_partials(ws));
    if (threadIdx.y == 0 && active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                a += (double)sds[y][threadIdx.x][j];
                b += (double)sdz[y][threadIdx.x][j];
            }
            part[(int64_t)blockIdx.y * C + (int64_t)vc * 4 + j] = make_double2(a, b);
        }
    }
    // the last CTA of this COLUMN BLOCK (one ticket per block of 256 columns) sums the slabs of its columns:
    // the four row lanes take every fourth slab (independent 16-byte loads), then a fixed-order combine
    if (!last_cta(reinterpret_cast<uint32_t*>(ws) + 4 + blockIdx.x, gridDim.y)) return;
    __shared__ double2 fin[4][64][4];
    if (active) {
        double2 acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = make_double2(0.0, 0.0);
        for (unsigned sl = threadIdx.y; sl < gridDim.y; sl += 4) {
            const double2* row = part + (int64_t)sl * C + (int64_t)vc * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 v = __ldcg(row + j);
                acc[j].x += v.x;
                acc[j].y += v.y;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) fin[threadIdx.y][threadIdx.x][j] = acc[j];
    }
    __syncthreads();
    if (active) {                                   // thread (x, y) finishes column 4 * vc + y
        const int j = threadIdx.y;
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            a += fin[y][threadIdx.x][j].x;
            b += fin[y][threadIdx.x][j].y;
        }
        const int64_t c = (int64_t)vc * 4 + j;
        if (grad_delta != nullptr) grad_delta[c] = delta_grad(q, c, (float)a);
        if (grad_zf != nullptr && q.zero_float != nullptr) grad_zf[c] = zf_grad(q, c, (float)b, lo, hi);
    }
    if (threadIdx.x == 0 && threadIdx.y == 0) reinterpret_cast<uint32_t*>(ws)[4 + blockIdx.x] = 0u;
}

// ---- general [outer, C, inner] (per-channel weights, odd shapes): one CTA per channel ---------------
__global__ void __launch_bounds__(kBThreads, 4)
qdq_bwd_rows_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gx, int64_t outer,
                    int64_t C, int64_t inner, int vec_ok, tq_qspec q, float* __restrict__ grad_delta,
                    float* __restrict__ grad_zf) {
    __shared__ double sm[32];
    float lo, hi;
    grid_of(q, lo, hi);
    for (int64_t c = blockIdx.x; c < C; c += gridDim.x) {
        const QP p = resolve(q, c, lo, hi);
        float ds = 0.0f, dz = 0.0f;
        if (vec_ok) {
            const int64_t iv = inner >> 2, tot = outer * iv;
            for (int64_t k = threadIdx.x; k < tot; k += kBThreads) {
                const int64_t o = k / iv, j = k - o * iv;
                const int64_t idx = ((o * C + c) * inner >> 2) + j;
                const float4 a = ld_stream(reinterpret_cast<const float4*>(x) + idx);
                const float4 b = ld_stream(reinterpret_cast<const float4*>(g) + idx);
                float4 o4;
                if (p.exact) {
                    bwd_elem<false>(a.x, b.x, p, o4.x, ds, dz);
                    bwd_elem<false>(a.y, b.y, p, o4.y, ds, dz);
                    bwd_elem<false>(a.z, b.z, p, o4.z, ds, dz);
                    bwd_elem<false>(a.w, b.w, p, o4.w, ds, dz);
                } else {
                    bwd_elem<true>(a.x, b.x, p, o4.x, ds, dz);
                    bwd_elem<true>(a.y, b.y, p, o4.y, ds, dz);
                    bwd_elem<true>(a.z, b.z, p, o4.z, ds, dz);
                    bwd_elem<true>(a.w, b.w, p, o4.w, ds, dz);
                }
                if (gx != nullptr) st_stream(reinterpret_cast<float4*>(gx) + idx, o4);
            }
        } else {
            const int64_t tot = outer * inner;
            for (int64_t k = threadIdx.x; k < tot; k += kBThreads) {
                const int64_t o = k / inner, j = k - o * inner;
                const int64_t idx = (o * C + c) * inner + j;
                float o1;
                if (p.exact) bwd_elem<false>(x[idx], g[idx], p, o1, ds, dz);
                else bwd_elem<true>(x[idx], g[idx], p, o1, ds, dz);
                if (gx != nullptr) gx[idx] = o1;
            }
        }
        const double a = block_sum((double)ds, sm);
        const double b = block_sum((double)dz, sm);
        if (threadIdx.x == 0) {
            if (grad_delta != nullptr) grad_delta[c] = delta_grad(q, c, (float)a);
            if (grad_zf != nullptr && q.zero_float != nullptr) grad_zf[c] = zf_grad(q, c, (float)b, lo, hi);
        }
    }
}

// ---- per-channel weights [C, inner] with many channels: one WARP per channel row -------------------------
// A BERT weight row is 768 - 3072 floats: a CTA per row leaves most threads idle and pays two block
// reductions per 3 - 12 KB.  Here a warp streams its row with 4 x 2 independent 128-bit loads per lane and
// reduces with shuffles only.
template <bool FAST>
__device__ __forceinline__ void bwd_row_warp(const float4* __restrict__ xv, const float4* __restrict__ gv,
                                             float4* __restrict__ gxv, int iv, int lane, const QP& p, float& ds,
                                             float& dz) {
    for (int j0 = lane; j0 < iv; j0 += 32 * kBUnroll) {
        float4 a[kBUnroll], b[kBUnroll];
#pragma unroll
        for (int u = 0; u < kBUnroll; ++u) {
            const int j = j0 + 32 * u;
            if (j < iv) {
                a[u] = ld_stream(xv + j);
                b[u] = ld_stream(gv + j);
            }
        }
#pragma unroll
        for (int u = 0; u < kBUnroll; ++u) {
            const int j = j0 + 32 * u;
            if (j < iv) {
                float4 o;
                bwd_elem<FAST>(a[u].x, b[u].x, p, o.x, ds, dz);
                bwd_elem<FAST>(a[u].y, b[u].y, p, o.y, ds, dz);
                bwd_elem<FAST>(a[u].z, b[u].z, p, o.z, ds, dz);
                bwd_elem<FAST>(a[u].w, b[u].w, p, o.w, ds, dz);
                if (gxv != nullptr) st_stream(gxv + j, o);
            }
        }
    }
}

__global__ void __launch_bounds__(kBThreads, 2)
qdq_bwd_rowwarp_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gx, int64_t C,
                       int64_t inner, tq_qspec q, float* __restrict__ grad_delta, float* __restrict__ grad_zf) {
    float lo, hi;
    grid_of(q, lo, hi);
    const int lane = threadIdx.x & 31;
    const int iv = (int)(inner >> 2);                 // inner <= 65536 (host dispatch)
    const int64_t warps = (int64_t)gridDim.x * (kBThreads / 32);
    for (int64_t c = (int64_t)blockIdx.x * (kBThreads / 32) + (threadIdx.x >> 5); c < C; c += warps) {
        const QP p = resolve(q, c, lo, hi);
        const float4* xv = reinterpret_cast<const float4*>(x + c * inner);
        const float4* gv = reinterpret_cast<const float4*>(g + c * inner);
        float4* gxv = gx != nullptr ? reinterpret_cast<float4*>(gx + c * inner) : nullptr;
        float ds = 0.0f, dz = 0.0f;
        if (p.exact) bwd_row_warp<false>(xv, gv, gxv, iv, lane, p, ds, dz);      // warp-uniform
        else bwd_row_warp<true>(xv, gv, gxv, iv, lane, p, ds, dz);
        const double a = warp_sum((double)ds), b = warp_sum((double)dz);
        if (lane == 0) {
            if (grad_delta != nullptr) grad_delta[c] = delta_grad(q, c, (float)a);
            if (grad_zf != nullptr && q.zero_float != nullptr) grad_zf[c] = zf_grad(q, c, (float)b, lo, hi);
        }
    }
}

static int64_t bwd_cols_slabs(int64_t rows, int64_t C) {
    // one wave: column blocks x slabs <= 2 resident CTAs per SM (a 297th CTA would run alone in a second wave)
    const int64_t col_blocks = ((C >> 2) + 63) / 64;
    int64_t slabs = ((int64_t)sm_count() * 2) / col_blocks;
    const int64_t max_slabs = (rows + 7) / 8;                                     // >= 8 rows per slab
    if (slabs > max_slabs) slabs = max_slabs;
    return slabs < 1 ? 1 : slabs;
}
// the non-persistent span kernel takes over at this many CTAs (TQ_BWD_SPAN_MIN overrides; 0 = always)
static int64_t bwd_span_min() {
    static int64_t v = -1;
    if (v < 0) {
        const char* e = getenv("TQ_BWD_SPAN_MIN");
        v = e != nullptr ? atoll(e) : kBwdSpanMinDefault;
    }
    return v;
}
static int64_t bwd_spans(int64_t n) {
    const int64_t per_cta = (int64_t)kBwdSpan * kBThreads * kBUnroll;
    return ((n >> 2) + per_cta - 1) / per_cta;
}
static bool bwd_use_spans(int64_t n) {
    const int64_t c = bwd_spans(n);
    return c >= bwd_span_min() && c > 0 && c < 0x7fffffff;
}

static int bwd_tensor_grid(int64_t n) {
    int64_t blocks = ((n >> 2) + kBThreads * kBUnroll) / (kBThreads * kBUnroll);
    const int64_t cap = (int64_t)sm_count() * 3;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

// ---- AdaRound ----------------------------------------------------------------------------------------
enum { ADA_SIGMOID = 0, ADA_HARD_SIGMOID = 1, ADA_TEMP_DECAY = 2 };
constexpr float kZeta = 1.1f, kGamma = -0.1f;      // adaround/quantizer.py:29,34 defaults

__device__ __forceinline__ float sigmoidf_(float a) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-a))); }

// soft target h(alpha) and dh / dalpha (adaround/quantizer.py:29-31, 84-92)
__device__ __forceinline__ float ada_rest(float alpha, int mode, float temperature, float* deriv) {
    if (mode == ADA_HARD_SIGMOID) {
        const float s = sigmoidf_(alpha);
        const float v = __fadd_rn(__fmul_rn(s, kZeta - kGamma), kGamma);
        if (deriv != nullptr)
            *deriv = (v >= 0.0f && v <= 1.0f) ? __fmul_rn(__fmul_rn(s, __fsub_rn(1.0f, s)), kZeta - kGamma) : 0.0f;
        return fminf(fmaxf(v, 0.0f), 1.0f);
    }
    if (mode == ADA_TEMP_DECAY) {
        const float s = sigmoidf_(__fdiv_rn(alpha, temperature));
        if (deriv != nullptr) *deriv = __fdiv_rn(__fmul_rn(s, __fsub_rn(1.0f, s)), temperature);
        return s;
    }
    const float s = sigmoidf_(alpha);
    if (deriv != nullptr) *deriv = __fmul_rn(s, __fsub_rn(1.0f, s));
    return s;
}

// what: 0 = alpha initialisation (out0 = alpha), 1 = forward (out0 = y and / or out1 = x_int),
//       2 = backward (out0 = grad_alpha for upstream gradient `aux`)
__global__ void __launch_bounds__(kBThreads, 4)
adaround_kernel(int what, const float* __restrict__ w, const float* __restrict__ alpha, const float* __restrict__ aux,
                float* __restrict__ out0, float* __restrict__ out1, int64_t n, int64_t C, int64_t inner, tq_qspec q,
                int mode, int soft, float temperature) {
    float lo, hi;
    grid_of(q, lo, hi);
    const int64_t stride = (int64_t)gridDim.x * kBThreads;
    for (int64_t i = (int64_t)blockIdx.x * kBThreads + threadIdx.x; i < n; i += stride) {
        const int64_t c = (C == 1) ? 0 : (i / inner) % C;
        const QP p = resolve(q, c, lo, hi);
        const float t = __fdiv_rn(w[i], p.scale);
        const float fl = floorf(t);
        if (what == 0) {
            const float rest = __fsub_rn(t, fl);
            float a;
            if (mode == ADA_HARD_SIGMOID) {
                a = -logf(__fdiv_rn(__fsub_rn(kZeta, rest), __fsub_rn(rest, kGamma)));
            } else {
                const float pr = fminf(fmaxf(rest, 1e-16f), (float)(1 - 1e-16));
                a = -logf(__fsub_rn(__fdiv_rn(1.0f, pr), 1.0f));
                if (mode == ADA_TEMP_DECAY) a = __fmul_rn(temperature, a);
            }
            out0[i] = a;
            continue;
        }
        const float al = alpha[i];
        float deriv = 0.0f;
        const float up = (soft || what == 2) ? ada_rest(al, mode, temperature, what == 2 ? &deriv : nullptr)
                                             : (al >= 0.0f ? 1.0f : 0.0f);
        const float u = __fadd_rn(__fadd_rn(fl, up), p.zp);
        if (what == 2) {
            const bool in = (u >= lo) && (u <= hi);
            out0[i] = in ? __fmul_rn(__fmul_rn(aux[i], p.scale), deriv) : 0.0f;
            continue;
        }
        float xi = u < lo ? lo : u;
        xi = xi > hi ? hi : xi;
        if (out1 != nullptr) out1[i] = xi;
        if (out0 != nullptr) out0[i] = dequant(xi, p);
    }
}

static int ada_launch(int what, const float* w, const float* alpha, const float* aux, float* out0, float* out1,
                      int64_t outer, int64_t C, int64_t inner, tq_qspec q, int mode, int soft, float temperature,
                      cudaStream_t st) {
    if (outer < 0 || C < 1 || inner < 0) return TQ_EINVAL;
    if (mode < 0 || mode > 2) return TQ_EINVAL;
    if (mode == ADA_TEMP_DECAY && !(temperature > 0.0f)) return TQ_EINVAL;
    if (int e = check_qspec(q)) return e;
    const int64_t n = outer * C * inner;
    if (n == 0) return TQ_OK;
    if (w == nullptr) return TQ_EINVAL;
    int64_t blocks = (n + kBThreads - 1) / kBThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    adaround_kernel<<<(int)blocks, kBThreads, 0, st>>>(what, w, alpha, aux, out0, out1, n, C, inner, q, mode, soft,
                                                       temperature);
    return launch_status();
}

}  // namespace tq

extern "C" {

size_t tq_qdq_bwd_workspace_bytes(int64_t outer, int64_t C, int64_t inner) {
    if (outer < 0 || C < 1 || inner < 0) return 0;
    if (C == 1) {
        const int64_t n = outer * inner;
        const size_t ctas = tq::bwd_use_spans(n) ? (size_t)tq::bwd_spans(n) : (size_t)tq::sm_count() * 3;
        return tq::kBwdHeader + ctas * 2 * sizeof(double);
    }
    if (inner == 1 && (C & 3) == 0 && ((C >> 2) + 63) / 64 <= tq::kBwdMaxColBlocks)
        return tq::kBwdHeader + (size_t)tq::bwd_cols_slabs(outer, C) * (size_t)C * 2 * sizeof(double);
    return tq::kBwdHeader;
}

int tq_qdq_bwd_f32(const float* x, const float* grad_y, float* grad_x, float* grad_delta, float* grad_zero_float,
                   int64_t outer, int64_t C, int64_t inner, tq_qspec q, void* ws, size_t ws_bytes, void* stream) {
    if (outer < 0 || C < 1 || inner < 0) return TQ_EINVAL;
    if (int e = tq::check_qspec(q)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = outer * C * inner;
    if (n == 0) {                                        // empty sums
        if (grad_delta != nullptr) cudaMemsetAsync(grad_delta, 0, (size_t)C * sizeof(float), st);
        if (grad_zero_float != nullptr) cudaMemsetAsync(grad_zero_float, 0, (size_t)C * sizeof(float), st);
        return tq::launch_status();
    }
    if (x == nullptr || grad_y == nullptr) return TQ_EINVAL;
    if (ws == nullptr || ws_bytes < tq_qdq_bwd_workspace_bytes(outer, C, inner)) return TQ_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(ws) & 15u) != 0) return TQ_EALIGN;
    const bool al = tq::aligned16(x) && tq::aligned16(grad_y) && (grad_x == nullptr || tq::aligned16(grad_x));
    if (C == 1) {
        if (al && tq::bwd_use_spans(n)) {
            tq::qdq_bwd_span_kernel<<<(int)tq::bwd_spans(n), tq::kBThreads, 0, st>>>(x, grad_y, grad_x, n, q, grad_delta,
                                                                                grad_zero_float, ws);
            return tq::launch_status();
        }
        tq::qdq_bwd_tensor_kernel<<<tq::bwd_tensor_grid(n), tq::kBThreads, 0, st>>>(
            x, grad_y, grad_x, n, al ? 1 : 0, q, grad_delta, grad_zero_float, ws);
        return tq::launch_status();
    }
    if (inner == 1 && (C & 3) == 0 && al && ((C >> 2) + 63) / 64 <= tq::kBwdMaxColBlocks) {
        const int64_t slabs = tq::bwd_cols_slabs(outer, C);
        const int64_t rps = (outer + slabs - 1) / slabs;
        const dim3 grid((unsigned)(((C >> 2) + 63) / 64), (unsigned)((outer + rps - 1) / rps));
        tq::qdq_bwd_cols_kernel<<<grid, dim3(64, 4), 0, st>>>(x, grad_y, grad_x, outer, (int32_t)C, rps, q, grad_delta,
                                                              grad_zero_float, ws);
        return tq::launch_status();
    }
    const int vec_ok = (al && (inner & 3) == 0) ? 1 : 0;
    const int64_t cap = (int64_t)tq::sm_count() * 8;
    if (outer == 1 && vec_ok && C >= (int64_t)tq::sm_count() * 2 && inner <= 65536) {
        const int64_t blocks = (C + 7) / 8, resident = (int64_t)tq::sm_count() * 2;      // persistent: one wave
        tq::qdq_bwd_rowwarp_kernel<<<(int)(blocks < resident ? blocks : resident), tq::kBThreads, 0, st>>>(
            x, grad_y, grad_x, C, inner, q, grad_delta, grad_zero_float);
        return tq::launch_status();
    }
    tq::qdq_bwd_rows_kernel<<<(int)(C < cap ? C : cap), tq::kBThreads, 0, st>>>(x, grad_y, grad_x, outer, C, inner,
                                                                               vec_ok, q, grad_delta, grad_zero_float);
    return tq::launch_status();
}

int tq_adaround_init_alpha_f32(const float* w, float* alpha, int64_t outer, int64_t C, int64_t inner, tq_qspec q,
                               int32_t mode, float temperature, void* stream) {
    if (alpha == nullptr && outer * C * inner > 0) return TQ_EINVAL;
    return tq::ada_launch(0, w, nullptr, nullptr, alpha, nullptr, outer, C, inner, q, mode, 1, temperature,
                          (cudaStream_t)stream);
}

int tq_adaround_fwd_f32(const float* w, const float* alpha, float* y, float* x_int, int64_t outer, int64_t C,
                        int64_t inner, tq_qspec q, int32_t mode, int32_t soft_targets, float temperature, void* stream) {
    if (outer * C * inner > 0 && (alpha == nullptr || (y == nullptr && x_int == nullptr))) return TQ_EINVAL;
    return tq::ada_launch(1, w, alpha, nullptr, y, x_int, outer, C, inner, q, mode, soft_targets ? 1 : 0, temperature,
                          (cudaStream_t)stream);
}

int tq_adaround_bwd_f32(const float* w, const float* alpha, const float* grad_y, float* grad_alpha, int64_t outer,
                        int64_t C, int64_t inner, tq_qspec q, int32_t mode, float temperature, void* stream) {
    if (outer * C * inner > 0 && (alpha == nullptr || grad_y == nullptr || grad_alpha == nullptr)) return TQ_EINVAL;
    return tq::ada_launch(2, w, alpha, grad_y, grad_alpha, nullptr, outer, C, inner, q, mode, 1, temperature,
                          (cudaStream_t)stream);
}

}  // extern "C"
