// Quantize -> round -> clamp -> dequantize kernels.
// Replaces the 6-pass ATen chain of AsymmetricUniformQuantizer.forward / to_integer_forward
// (reference quantization/quantizers.py:172-211) by ONE pass: 4 B read + 4 B written per element.
//
// HBM-bound elementwise work.  Layout / mapping (B200: 148 SMs, 4 CTAs x 256 threads resident per
// SM, 4 independent 128-bit loads in flight per thread = 64 KB in flight per SM):
//   * per-tensor QDQ: one 16 KB chunk per CTA, as many CTAs as chunks (qdq_tensor_chunk_kernel; measured
//     6.86 TB/s vs 5.95 for every persistent variant); integer outputs: grid-stride over float4 vectors.
//     Quantizer parameters are resolved once per thread from the device-resident
//     `_delta/_zero_float/_signed` buffers (no host sync).
//   * per-embedding / per-embedding-group (x viewed [rows, C], inner == 1): the resolved per-dim
//     {scale, zero_point} table is staged in shared memory once per CTA and indexed by hidden dim;
//     the column of each vector is tracked incrementally (no 64-bit modulo in the loop).
//   * per-channel weights (inner > 1): one row (channel) per CTA iteration, scalar parameters.
#include "tq_common.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

namespace tq {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

enum OutMode { OUT_QDQ = 0, OUT_INT = 1 };

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int MODE, bool FAST>
__device__ __forceinline__ void emit_vec(const float4& v, const QP2& pa, const QP2& pb, float4* y, float4* yint,
                                         uint2* yctr, int64_t idx) {
    const float2 qa = quant_int2_t<FAST>(make_float2(v.x, v.y), pa);
    const float2 qb = quant_int2_t<FAST>(make_float2(v.z, v.w), pb);
    if (MODE == OUT_QDQ) {
        const float2 oa = dequant2(qa, pa), ob = dequant2(qb, pb);
        st_stream(y + idx, make_float4(oa.x, oa.y, ob.x, ob.y));
    } else {
        if (yint != nullptr) st_stream(yint + idx, make_float4(qa.x, qa.y, qb.x, qb.y));
        if (yctr != nullptr) {
            const float2 ca = centre2(qa, pa), cb = centre2(qb, pb);
            uint2 c;
            c.x = pack_bf16x2(ca.x, ca.y);
            c.y = pack_bf16x2(cb.x, cb.y);
            yctr[idx] = c;
        }
    }
}

template <int MODE>
__device__ __forceinline__ void emit_scalar(float v, const QP& p, float* y, float* yint,
                                            __nv_bfloat16* yctr, int64_t i) {
    const float qi = quant_int(v, p);
    if (MODE == OUT_QDQ) {
        y[i] = dequant(qi, p);
    } else {
        if (yint != nullptr) yint[i] = qi;
        if (yctr != nullptr) yctr[i] = __float2bfloat16_rn(__fsub_rn(qi, p.zp));
    }
}

// ---- per-tensor ---------------------------------------------------------------------------------
template <int MODE, bool FAST, int U = kUnroll>
__device__ __forceinline__ void tensor_body(const float4* __restrict__ xv, float4* __restrict__ yv,
                                            float4* __restrict__ yiv, uint2* __restrict__ ycv, int64_t nvec,
                                            const QP& p) {
    const QP2 p2 = pair_of(p);
    const int64_t stride = (int64_t)gridDim.x * kThreads * U;
    for (int64_t base = (int64_t)blockIdx.x * kThreads * U + threadIdx.x; base < nvec; base += stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t idx = base + (int64_t)u * kThreads;
            if (idx < nvec) v[u] = ld_stream(xv + idx);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t idx = base + (int64_t)u * kThreads;
            if (idx < nvec) emit_vec<MODE, FAST>(v[u], p2, p2, yv, yiv, ycv, idx);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 4)
qdq_tensor_vec_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ yint,
                      __nv_bfloat16* __restrict__ yctr, int64_t n, tq_qspec q) {
    float lo, hi;
    grid_of(q, lo, hi);
    const QP p = resolve(q, 0, lo, hi);
    const int64_t nvec = n >> 2;
    const float4* xv = reinterpret_cast<const float4*>(x);
    float4* yv = reinterpret_cast<float4*>(y);
    float4* yiv = reinterpret_cast<float4*>(yint);
    uint2* ycv = reinterpret_cast<uint2*>(yctr);
    if (p.exact) tensor_body<MODE, false>(xv, yv, yiv, ycv, nvec, p);     // uniform; see tq::div_rn_t
    else tensor_body<MODE, true>(xv, yv, yiv, ycv, nvec, p);
    // ragged tail (n % 4 elements)
    if (blockIdx.x == 0) {
        const int64_t i = (nvec << 2) + threadIdx.x;
        if (i < n) emit_scalar<MODE>(x[i], p, y, yint, yctr, i);
    }
}

// One 16 KB chunk per CTA, as many CTAs as chunks (no grid-stride loop).  Measured with the arithmetic-free
// probe (tq_probe_copy_f32, profiles/r1_copy_probe.json): the same 128-bit accesses reach 6.8 TB/s this way
// and 6.0 TB/s from a persistent grid-stride grid, whose CTAs fall into lock step (all reading, then all
// writing); CTAs that retire and get replaced one by one keep reads and writes mixed.  The data loads are
// issued before the quantizer parameters are resolved (two dependent L2 round trips per CTA otherwise).
constexpr int kChunkVec = kThreads * kUnroll;        // float4 per CTA
__global__ void __launch_bounds__(kThreads, 4)
qdq_tensor_chunk_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, tq_qspec q) {
    const int64_t nvec = n >> 2;
    const float4* xv = reinterpret_cast<const float4*>(x);
    float4* yv = reinterpret_cast<float4*>(y);
    const int64_t base = (int64_t)blockIdx.x * kChunkVec + threadIdx.x;
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const int64_t idx = base + (int64_t)u * kThreads;
        if (idx < nvec) v[u] = ld_stream(xv + idx);
    }
    float lo, hi;
    grid_of(q, lo, hi);
    const QP p = resolve(q, 0, lo, hi);
    const QP2 p2 = pair_of(p);
    if (p.exact) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t idx = base + (int64_t)u * kThreads;
            if (idx < nvec) emit_vec<OUT_QDQ, false>(v[u], p2, p2, yv, nullptr, nullptr, idx);
        }
    } else {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t idx = base + (int64_t)u * kThreads;
            if (idx < nvec) emit_vec<OUT_QDQ, true>(v[u], p2, p2, yv, nullptr, nullptr, idx);
        }
    }
    if (blockIdx.x == 0) {                               // ragged tail (n % 4 elements)
        const int64_t i = (nvec << 2) + threadIdx.x;
        if (i < n) emit_scalar<OUT_QDQ>(x[i], p, y, nullptr, nullptr, i);
    }
}

// ---- per-tensor, bulk-copy staged (TMA engine, shared-memory ring) --------------------------------
// For tensors much larger than L2 the memory-level parallelism of the LDG kernel is capped by its
// register budget (4 x 16 B per thread).  Here the copy engine moves 16 KB tiles global -> shared
// memory (cp.async.bulk + mbarrier complete_tx) and back (cp.async.bulk shared -> global, bulk
// groups); threads only touch shared memory.  One elected thread issues both directions, a ring of
// kBulkStages tiles keeps (stages - 1) loads and the stores in flight per CTA, two CTAs per SM.
constexpr int kBulkThreads = 256;
constexpr int kBulkTileVec = 1024;                 // float4 per tile (16 KB)
constexpr int kBulkStages = 4;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool FAST>
__device__ __forceinline__ void bulk_compute(float4* tile, int nvec_tile, const QP& p) {
    const QP2 p2 = pair_of(p);
#pragma unroll
    for (int u = 0; u < kBulkTileVec / kBulkThreads; ++u) {
        const int i = u * kBulkThreads + threadIdx.x;
        if (i < nvec_tile) {
            const float4 v = tile[i];
            const float2 a = qdq2_t<FAST>(make_float2(v.x, v.y), p2), b = qdq2_t<FAST>(make_float2(v.z, v.w), p2);
            tile[i] = make_float4(a.x, a.y, b.x, b.y);
        }
    }
}

__global__ void __launch_bounds__(kBulkThreads, 2)
qdq_tensor_bulk_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, tq_qspec q) {
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    float4* ring = reinterpret_cast<float4*>(bulk_smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(bulk_smem + (size_t)kBulkStages * kBulkTileVec * 16);
    float lo, hi;
    grid_of(q, lo, hi);
    const QP p = resolve(q, 0, lo, hi);
    const int64_t nvec = n >> 2;
    const int64_t tiles = (nvec + kBulkTileVec - 1) / kBulkTileVec;
    const int64_t my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kBulkStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto tile_vecs = [&](int64_t i) -> int {      // i-th tile of this CTA
        const int64_t t = blockIdx.x + i * gridDim.x;
        const int64_t left = nvec - t * kBulkTileVec;
        return (int)(left < kBulkTileVec ? left : kBulkTileVec);
    };
    auto issue_load = [&](int64_t i) {
        const int s = (int)(i % kBulkStages);
        const int64_t t = blockIdx.x + i * gridDim.x;
        const uint32_t bytes = (uint32_t)tile_vecs(i) * 16u;
        const uint32_t bar = smem_addr(bars + s);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_addr(ring + (size_t)s * kBulkTileVec)), "l"(x + t * kBulkTileVec * 4), "r"(bytes), "r"(bar)
                     : "memory");
    };
    if (threadIdx.x == 0)
        for (int64_t i = 0; i < kBulkStages - 1 && i < my_tiles; ++i) issue_load(i);

    for (int64_t i = 0; i < my_tiles; ++i) {
        const int s = (int)(i % kBulkStages);
        const uint32_t parity = (uint32_t)((i / kBulkStages) & 1);
        const uint32_t bar = smem_addr(bars + s);
        uint32_t done = 0;
        while (!done)
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
        float4* tile = ring + (size_t)s * kBulkTileVec;
        const int nv = tile_vecs(i);
        if (p.exact) bulk_compute<false>(tile, nv, p);
        else bulk_compute<true>(tile, nv, p);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // smem writes -> visible to the copy engine
        __syncthreads();
        if (threadIdx.x == 0) {
            const int64_t t = blockIdx.x + i * gridDim.x;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(y + t * kBulkTileVec * 4), "r"(smem_addr(tile)), "r"((uint32_t)nv * 16u)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the buffer of tile i-1 is re-used by tile i + stages - 1: its store must have been read
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (i + kBulkStages - 1 < my_tiles) issue_load(i + kBulkStages - 1);
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    // ragged tail (n % 4 elements)
    if (blockIdx.x == 0) {
        const int64_t i = (nvec << 2) + threadIdx.x;
        if (i < n) emit_scalar<OUT_QDQ>(x[i], p, y, nullptr, nullptr, i);
    }
}

// generic scalar kernel: any alignment, any [outer, C, inner] view (C == 1: per-tensor)
template <int MODE>
__global__ void __launch_bounds__(kThreads, 4)
qdq_generic_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ yint,
                   __nv_bfloat16* __restrict__ yctr, int64_t n, int64_t C, int64_t inner,
                   tq_qspec q) {
    float lo, hi;
    grid_of(q, lo, hi);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        const int64_t c = (C == 1) ? 0 : (i / inner) % C;
        const QP p = resolve(q, c, lo, hi);
        emit_scalar<MODE>(x[i], p, y, yint, yctr, i);
    }
}

// ---- per-embedding / per-embedding-group: x viewed [rows, C], parameters indexed by column ------
template <int MODE, bool FAST>
__device__ __forceinline__ void cols_body(const float4* __restrict__ xv, float4* __restrict__ yv,
                                          float4* __restrict__ yiv, uint2* __restrict__ ycv, int64_t nvec, int32_t CV,
                                          const float4* sv, const float4* zv, const float4* rv, float lo, float hi) {
    const int64_t stride = (int64_t)gridDim.x * kThreads * kUnroll;
    int64_t base = (int64_t)blockIdx.x * kThreads * kUnroll + threadIdx.x;
    int32_t col0 = (int32_t)(base % CV);                 // vector column of the first access
    const int32_t step = (int32_t)(stride % CV);
    int32_t off[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) off[u] = (u * kThreads) % CV;

    for (; base < nvec; base += stride) {
        float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t idx = base + (int64_t)u * kThreads;
            if (idx < nvec) v[u] = ld_stream(xv + idx);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t idx = base + (int64_t)u * kThreads;
            if (idx < nvec) {
                int32_t cv = col0 + off[u];
                cv -= (cv >= CV) ? CV : 0;
                const float4 s = sv[cv];
                const float4 z = zv[cv];
                const float4 r = rv[cv];
                QP2 pa, pb;
                pa.scale = make_float2(s.x, s.y); pa.nscale = make_float2(-s.x, -s.y); pa.rcp = make_float2(r.x, r.y);
                pa.zp = make_float2(z.x, z.y); pa.nzp = make_float2(-z.x, -z.y); pa.lo = lo; pa.hi = hi;
                pb.scale = make_float2(s.z, s.w); pb.nscale = make_float2(-s.z, -s.w); pb.rcp = make_float2(r.z, r.w);
                pb.zp = make_float2(z.z, z.w); pb.nzp = make_float2(-z.z, -z.w); pb.lo = lo; pb.hi = hi;
                emit_vec<MODE, FAST>(v[u], pa, pb, yv, yiv, ycv, idx);
            }
        }
        col0 += step;
        col0 -= (col0 >= CV) ? CV : 0;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 4)
qdq_cols_vec_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ yint,
                    __nv_bfloat16* __restrict__ yctr, int64_t nvec, int32_t C, tq_qspec q) {
    extern __shared__ __align__(16) float tab[];   // [C] scale | [C] zero_point | [C] 1/scale, by hidden dim
    float lo, hi;
    grid_of(q, lo, hi);
    int need_exact = 0;
    for (int c = threadIdx.x; c < C; c += kThreads) {
        const QP p = resolve(q, c, lo, hi);
        tab[c] = p.scale;
        tab[C + c] = p.zp;
        tab[2 * C + c] = p.rcp;
        need_exact |= p.exact;
    }
    const int exact = __syncthreads_or(need_exact);   // rare: any column needing the IEEE divide
    const int32_t CV = C >> 2;
    const float4* xv = reinterpret_cast<const float4*>(x);
    float4* yv = reinterpret_cast<float4*>(y);
    float4* yiv = reinterpret_cast<float4*>(yint);
    uint2* ycv = reinterpret_cast<uint2*>(yctr);
    const float4* sv = reinterpret_cast<const float4*>(tab);
    const float4* zv = reinterpret_cast<const float4*>(tab + C);
    const float4* rv = reinterpret_cast<const float4*>(tab + 2 * C);

    if (exact) cols_body<MODE, false>(xv, yv, yiv, ycv, nvec, CV, sv, zv, rv, lo, hi);
    else cols_body<MODE, true>(xv, yv, yiv, ycv, nvec, CV, sv, zv, rv, lo, hi);
}

// ---- per-channel rows: x viewed [rows = outer*C, inner], one parameter per row ------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads, 4)
qdq_rows_vec_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ yint,
                    __nv_bfloat16* __restrict__ yctr, int64_t rows, int64_t C, int64_t inner,
                    tq_qspec q) {
    float lo, hi;
    grid_of(q, lo, hi);
    const int64_t ivec = inner >> 2;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const QP p = resolve(q, r % C, lo, hi);
        const float4* xv = reinterpret_cast<const float4*>(x + r * inner);
        float4* yv = reinterpret_cast<float4*>(y + r * inner);
        float4* yiv = yint ? reinterpret_cast<float4*>(yint + r * inner) : nullptr;
        uint2* ycv = yctr ? reinterpret_cast<uint2*>(yctr + r * inner) : nullptr;
        const QP2 p2 = pair_of(p);
        if (p.exact) {
            for (int64_t i = threadIdx.x; i < ivec; i += kThreads)
                emit_vec<MODE, false>(ld_stream(xv + i), p2, p2, yv, yiv, ycv, i);
        } else {
            for (int64_t i = threadIdx.x; i < ivec; i += kThreads)
                emit_vec<MODE, true>(ld_stream(xv + i), p2, p2, yv, yiv, ycv, i);
        }
    }
}

// per-tensor QDQ switches from the LDG kernel to the bulk-copy staged kernel at this many elements
// (TQ_QDQ_BULK_MIN overrides; 0 disables the LDG kernel, a huge value disables the bulk kernel)
static int64_t qdq_bulk_threshold() {
    static int64_t thr = -1;
    if (thr < 0) {
        const char* e = getenv("TQ_QDQ_BULK_MIN");
        thr = e != nullptr ? atoll(e) : (int64_t)8 * 1024 * 1024;
    }
    return thr;
}

// TQ_QDQ_VARIANT = ldg | bulk | chunk forces one per-tensor QDQ kernel for every size (tuning / profiling);
// unset: the size policy in launch_any
static int qdq_forced_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TQ_QDQ_VARIANT");
        v = 0;
        if (e != nullptr) v = e[0] == 'l' ? 1 : (e[0] == 'b' ? 3 : (e[0] == 'c' ? 4 : 0));
    }
    return v;
}

// the one-chunk-per-CTA kernel is the default at every size (measured faster from 3 M to 256 M elements,
// profiles/r1_qdq_variants.json); TQ_QDQ_CHUNK_MIN = minimum number of chunks for it
static int64_t qdq_chunk_min_ctas() {
    static int64_t v = -1;
    if (v < 0) {
        const char* e = getenv("TQ_QDQ_CHUNK_MIN");
        v = e != nullptr ? atoll(e) : 1;
    }
    return v;
}

static int grid_for(int64_t work_items, int per_block, int ctas_per_sm) {
    int64_t blocks = (work_items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <int MODE>
static int launch_any(const float* x, float* y, float* yint, __nv_bfloat16* yctr, int64_t outer,
                      int64_t C, int64_t inner, tq_qspec q, cudaStream_t st) {
    const int64_t n = outer * C * inner;
    if (n == 0) return TQ_OK;
    const bool al = aligned16(x) && (MODE == OUT_QDQ ? aligned16(y)
                                                     : ((yint == nullptr || aligned16(yint)) &&
                                                        (yctr == nullptr ||
                                                         (reinterpret_cast<uintptr_t>(yctr) & 7u) == 0)));
    if (C == 1) {
        const int forced = MODE == OUT_QDQ && al ? qdq_forced_variant() : 0;
        // one-chunk CTAs beat every persistent variant (LDG grid-stride, bulk-copy ring) at all measured sizes
        const int64_t chunks = ((n >> 2) + kChunkVec - 1) / kChunkVec;
        if (MODE == OUT_QDQ && al && (forced == 4 || (forced == 0 && chunks >= qdq_chunk_min_ctas())) && chunks > 0 &&
            chunks < 0x7fffffff) {
            qdq_tensor_chunk_kernel<<<(int)chunks, kThreads, 0, st>>>(x, y, n, q);
            return launch_status();
        }
        if (MODE == OUT_QDQ && al && x != y && forced != 1 && (forced == 3 || n >= qdq_bulk_threshold())) {
            const size_t smem = (size_t)kBulkStages * kBulkTileVec * 16 + kBulkStages * 8;
            static bool attr_set = false;
            if (!attr_set) {
                cudaError_t e = cudaFuncSetAttribute(qdq_tensor_bulk_kernel,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return (int)e;
                attr_set = true;
            }
            const int64_t tiles = ((n >> 2) + kBulkTileVec - 1) / kBulkTileVec;
            const int64_t cap = (int64_t)sm_count() * 2;
            qdq_tensor_bulk_kernel<<<(int)(tiles < cap ? tiles : cap), kBulkThreads, smem, st>>>(x, y, n, q);
            return launch_status();
        }
        if (al) {
            const int grid = grid_for((n >> 2) + 1, kThreads * kUnroll, 4);
            qdq_tensor_vec_kernel<MODE><<<grid, kThreads, 0, st>>>(x, y, yint, yctr, n, q);
        } else {
            qdq_generic_kernel<MODE><<<grid_for(n, kThreads, 8), kThreads, 0, st>>>(x, y, yint, yctr, n,
                                                                                   1, 1, q);
        }
        return launch_status();
    }
    if (inner == 1 && al && (C & 3) == 0 && C * 12 <= 200 * 1024) {
        const size_t smem = (size_t)C * 12;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(qdq_cols_vec_kernel<MODE>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        const int grid = grid_for(n >> 2, kThreads * kUnroll, 4);
        qdq_cols_vec_kernel<MODE><<<grid, kThreads, smem, st>>>(x, y, yint, yctr, n >> 2, (int32_t)C, q);
        return launch_status();
    }
    if (inner > 1 && al && (inner & 3) == 0) {
        const int64_t rows = outer * C;
        const int grid = (int)(rows < (int64_t)sm_count() * 16 ? rows : (int64_t)sm_count() * 16);
        qdq_rows_vec_kernel<MODE><<<grid, kThreads, 0, st>>>(x, y, yint, yctr, rows, C, inner, q);
        return launch_status();
    }
    qdq_generic_kernel<MODE><<<grid_for(n, kThreads, 8), kThreads, 0, st>>>(x, y, yint, yctr, n, C, inner,
                                                                           q);
    return launch_status();
}

}  // namespace tq

extern "C" {

int tq_qdq_f32(const float* x, float* y, int64_t n, tq_qspec q, void* stream) {
    if (n < 0 || (n > 0 && (x == nullptr || y == nullptr))) return TQ_EINVAL;
    if (int e = tq::check_qspec(q)) return e;
    return tq::launch_any<tq::OUT_QDQ>(x, y, nullptr, nullptr, 1, 1, n, q, (cudaStream_t)stream);
}

int tq_qdq_axis_f32(const float* x, float* y, int64_t outer, int64_t C, int64_t inner, tq_qspec q,
                    void* stream) {
    if (outer < 0 || C < 1 || inner < 0) return TQ_EINVAL;
    if (outer * C * inner > 0 && (x == nullptr || y == nullptr)) return TQ_EINVAL;
    if (int e = tq::check_qspec(q)) return e;
    return tq::launch_any<tq::OUT_QDQ>(x, y, nullptr, nullptr, outer, C, inner, q,
                                       (cudaStream_t)stream);
}

int tq_quant_int_f32(const float* x, float* x_int_f32, void* x_ctr_bf16, int64_t outer, int64_t C,
                     int64_t inner, tq_qspec q, void* stream) {
    if (outer < 0 || C < 1 || inner < 0) return TQ_EINVAL;
    if (outer * C * inner > 0 && x == nullptr) return TQ_EINVAL;
    if (x_int_f32 == nullptr && x_ctr_bf16 == nullptr) return TQ_EINVAL;
    if (int e = tq::check_qspec(q)) return e;
    if (C == 1) {  // per-tensor: fold everything into one flat run
        return tq::launch_any<tq::OUT_INT>(x, nullptr, x_int_f32, (__nv_bfloat16*)x_ctr_bf16, 1, 1,
                                           outer * inner, q, (cudaStream_t)stream);
    }
    return tq::launch_any<tq::OUT_INT>(x, nullptr, x_int_f32, (__nv_bfloat16*)x_ctr_bf16, outer, C,
                                       inner, q, (cudaStream_t)stream);
}

}  // extern "C"
