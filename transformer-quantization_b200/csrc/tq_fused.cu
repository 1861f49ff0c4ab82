// Fused encoder blocks around the quantizer sites (pulled forward
// because 24 + 36 of the 161 activation-quantizer sites of a BERT-base forward sit in them):
//
//   tq_attention_qdq_bf16 : QK^T -> QDQ(scores) -> /sqrt(d) + mask -> softmax -> QDQ(probs) -> PV ->
//                           QDQ(context)      (reference models/quantized_bert.py:153-213)
//                           one CTA per (batch, head); both GEMMs on tcgen05 with exact integer
//                           operands (bf16 carriers), scores/probs never leave the SM.
//   tq_ln_qdq_bf16        : LayerNorm with fake-quantized gamma over a quantized input + output QDQ
//                           (reference autoquant_utils.py:55-66 after the residual quantizer)
//   tq_embed_ln_qdq_bf16  : word + token-type -> QDQ -> + position -> QDQ -> LayerNorm -> QDQ
//                           (reference models/quantized_bert.py:59-88)
//
// All three read / write the CENTRED INTEGER GRID of the fake-quantized tensors in bf16
// (x_int - zero_point, exact for n_bits <= 8): 2 B per element instead of the 4 B fp32 tensor the
// reference materialises at every site, and exactly the operand format of tq_linear_qdq_bf16.
// The dequantized fp32 value the reference would have produced is scale * ctr, recomputed on the fly.
#include "tq_common.cuh"
#include "tq_attn.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>

namespace tq {
namespace fused {

// ---------------------------------------------------------------------------------------------------
// shared small helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf16_lo(uint32_t pair) { return __uint_as_float(pair << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t pair) { return __uint_as_float(pair & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

struct ColQ {              // quantizer with 1 or C parameter slots
    tq_qspec q;
    int64_t params;
    float lo, hi;
    QP p0;                 // resolved once when per-tensor
    __device__ __forceinline__ void init() {
        grid_of(q, lo, hi);
        p0 = resolve(q, 0, lo, hi);
    }
    __device__ __forceinline__ QP at(int64_t c) const { return params > 1 ? resolve(q, c, lo, hi) : p0; }
    // PT (compile time): the caller has established params == 1 -> no per-element branch
    template <bool PT>
    __device__ __forceinline__ QP get(int64_t c) const { return PT ? p0 : at(c); }
};

// ---------------------------------------------------------------------------------------------------
// LayerNorm (+ optional embedding prologue) : one warp per row, D % 256 == 0, D <= 1024
// ---------------------------------------------------------------------------------------------------
constexpr int kLnWarps = 8;
constexpr int kLnMaxIter = 4;    // D / 256

struct LnArgs {
    // input: either a quantized tensor (x_ctr + in_q) or the embedding prologue (ids != null)
    const __nv_bfloat16* x_ctr;
    ColQ in_q;
    const int64_t* ids;          // [M] token ids            -- embedding prologue
    const int64_t* type_ids;     // [M] or null (all zero)
    const int64_t* pos_ids;      // [M] or null (position = row % T)
    int64_t T;
    const float* word;           // [V, D] fake-quantized tables
    const float* type_tab;       // [2, D]
    const float* pos_tab;        // [P, D]
    ColQ e_tok, e_pos;           // quantizers of the two embedding sums
    // LayerNorm
    const float* gamma_q;        // [D] fake-quantized weight
    const float* beta;           // [D]
    float eps;
    ColQ out_q;
    __nv_bfloat16* out_ctr;      // [M, D] centred grid (optional when out_u8 is given)
    unsigned char* out_u8;       // optional [M, D] x_int bytes (8-bit operand mode of the GEMMs)
    float* out_f32;              // optional [M, D]
    int64_t M;
    int32_t D;
};

template <bool EMBED, bool FAST>
__device__ __forceinline__ void ln_row(const LnArgs& a, int64_t row, int lane, ColQ& in_q, ColQ& e_tok, ColQ& e_pos,
                                       ColQ& out_q) {
    const int iters = a.D >> 8;                       // 8 elements per lane per iteration
    float v[kLnMaxIter][8];
    if (EMBED) {
        const int64_t id = a.ids[row];
        const int64_t tt = a.type_ids != nullptr ? a.type_ids[row] : 0;
        const int64_t pp = a.pos_ids != nullptr ? a.pos_ids[row] : (row % a.T);
#pragma unroll
        for (int it = 0; it < kLnMaxIter; ++it) {
            if (it < iters) {
                const int c = (it * 32 + lane) * 8;
                const float4* w = reinterpret_cast<const float4*>(a.word + id * a.D + c);
                const float4* t = reinterpret_cast<const float4*>(a.type_tab + tt * a.D + c);
                const float4* p = reinterpret_cast<const float4*>(a.pos_tab + pp * a.D + c);
                const float4 w0 = w[0], w1 = w[1], t0 = t[0], t1 = t[1], p0 = p[0], p1 = p[1];
                const float ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                const float ts[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                const float ps[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float e = qdq_t<FAST>(__fadd_rn(ws[j], ts[j]), e_tok.get<FAST>(c + j));      // quantized_bert.py:78-79
                    e = qdq_t<FAST>(__fadd_rn(e, ps[j]), e_pos.get<FAST>(c + j));                // :83-84
                    v[it][j] = e;
                }
            }
        }
    } else {
#pragma unroll
        for (int it = 0; it < kLnMaxIter; ++it) {
            if (it < iters) {
                const int c = (it * 32 + lane) * 8;
                const uint4 raw = *reinterpret_cast<const uint4*>(a.x_ctr + row * a.D + c);
                const uint32_t pr[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[it][2 * j] = __fmul_rn(in_q.get<FAST>(c + 2 * j).scale, bf16_lo(pr[j]));          // scale * ctr
                    v[it][2 * j + 1] = __fmul_rn(in_q.get<FAST>(c + 2 * j + 1).scale, bf16_hi(pr[j]));
                }
            }
        }
    }
    float mean, rstd;
    if (!EMBED && FAST) {
        // per-tensor quantized input x = s * k: statistics from the exact integer sums of k (see ln_stats_from_sums),
        // bit-identical to the fused GEMM + LayerNorm epilogues whatever their tiling
        const float sc = in_q.p0.scale;
        float f1 = 0.0f, f2 = 0.0f;                    // <= 32 values per lane: exact in fp32
#pragma unroll
        for (int it = 0; it < kLnMaxIter; ++it) {
            if (it < iters) {
                const int c = (it * 32 + lane) * 8;
                const uint4 raw = *reinterpret_cast<const uint4*>(a.x_ctr + row * a.D + c);
                const uint32_t pr[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float k0 = bf16_lo(pr[j]), k1 = bf16_hi(pr[j]);
                    f1 = __fadd_rn(f1, __fadd_rn(k0, k1));
                    f2 = __fmaf_rn(k0, k0, f2);
                    f2 = __fmaf_rn(k1, k1, f2);
                }
            }
        }
        int i1 = __float2int_rn(f1), i2 = __float2int_rn(f2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            i1 += __shfl_xor_sync(0xffffffffu, i1, o);
            i2 += __shfl_xor_sync(0xffffffffu, i2, o);
        }
        ln_stats_from_sums((long long)i1, (long long)i2, (int64_t)a.D, sc, a.eps, mean, rstd);
    } else {
        // mean / variance over the row (fp32, two passes: the values are in registers)
        float s = 0.0f;
#pragma unroll
        for (int it = 0; it < kLnMaxIter; ++it)
            if (it < iters)
#pragma unroll
                for (int j = 0; j < 8; ++j) s += v[it][j];
        s = warp_sum(s);
        mean = s / (float)a.D;
        float ss = 0.0f;
#pragma unroll
        for (int it = 0; it < kLnMaxIter; ++it)
            if (it < iters)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = v[it][j] - mean;
                    ss += d * d;
                }
        ss = warp_sum(ss);
        rstd = 1.0f / sqrtf(ss / (float)a.D + a.eps);
    }
#pragma unroll
    for (int it = 0; it < kLnMaxIter; ++it) {
        if (it < iters) {
            const int c = (it * 32 + lane) * 8;
            const float4* g = reinterpret_cast<const float4*>(a.gamma_q + c);
            const float4* b = reinterpret_cast<const float4*>(a.beta + c);
            const float4 g0 = g[0], g1 = g[1], b0 = b[0], b1 = b[1];
            const float gs[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float ctr[8], deq[8];
            if (FAST) {
                const QP2 p2 = pair_of(out_q.p0);
                const float2 nmean = make_float2(-mean, -mean), rs2 = make_float2(rstd, rstd);
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    float2 y = __fmul2_rn(__fadd2_rn(make_float2(v[it][j], v[it][j + 1]), nmean), rs2);
                    y = __ffma2_rn(y, make_float2(gs[j], gs[j + 1]), make_float2(bs[j], bs[j + 1]));    // as the fused epilogues
                    const float2 ci = centre2(quant_int2_t<true>(y, p2), p2);
                    const float2 dq = __fmul2_rn(p2.scale, ci);
                    ctr[j] = ci.x; ctr[j + 1] = ci.y;
                    deq[j] = dq.x; deq[j + 1] = dq.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float y = (v[it][j] - mean) * rstd * gs[j] + bs[j];
                    const QP p = out_q.get<FAST>(c + j);
                    ctr[j] = __fsub_rn(quant_int_t<FAST>(y, p), p.zp);
                    deq[j] = __fmul_rn(p.scale, ctr[j]);
                }
            }
            uint4 o;
            o.x = pack2(ctr[0], ctr[1]);
            o.y = pack2(ctr[2], ctr[3]);
            o.z = pack2(ctr[4], ctr[5]);
            o.w = pack2(ctr[6], ctr[7]);
            if (a.out_ctr != nullptr) *reinterpret_cast<uint4*>(a.out_ctr + row * a.D + c) = o;
            if (a.out_u8 != nullptr) {
                uint32_t xi[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    xi[j] = __float_as_uint(__fadd_rn(ctr[j], __fadd_rn(out_q.get<FAST>(c + j).zp, 12582912.0f)));
                uint2 o8;
                o8.x = __byte_perm(__byte_perm(xi[0], xi[1], 0x0040), __byte_perm(xi[2], xi[3], 0x0040), 0x5410);
                o8.y = __byte_perm(__byte_perm(xi[4], xi[5], 0x0040), __byte_perm(xi[6], xi[7], 0x0040), 0x5410);
                *reinterpret_cast<uint2*>(a.out_u8 + row * a.D + c) = o8;
            }
            if (a.out_f32 != nullptr) {
                float4* of = reinterpret_cast<float4*>(a.out_f32 + row * a.D + c);
                of[0] = make_float4(deq[0], deq[1], deq[2], deq[3]);
                of[1] = make_float4(deq[4], deq[5], deq[6], deq[7]);
            }
        }
    }
}

template <bool EMBED>
__global__ void __launch_bounds__(kLnWarps * 32, 2) ln_qdq_kernel(LnArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kLnWarps + warp;
    pdl_trigger();
    pdl_wait();
    if (row >= a.M) return;
    ColQ in_q = a.in_q, e_tok = a.e_tok, e_pos = a.e_pos, out_q = a.out_q;
    out_q.init();
    // FAST (division-free quotient) only when every quantizer involved is per-tensor and inside
    // div_rn's proven domain; otherwise the IEEE-divide instantiation (always valid)
    bool fast = out_q.params == 1 && !out_q.p0.exact;
    if (EMBED) {
        e_tok.init();
        e_pos.init();
        fast = fast && e_tok.params == 1 && !e_tok.p0.exact && e_pos.params == 1 && !e_pos.p0.exact;
    } else {
        in_q.init();
        fast = fast && in_q.params == 1;
    }
    if (fast) ln_row<EMBED, true>(a, row, lane, in_q, e_tok, e_pos, out_q);
    else ln_row<EMBED, false>(a, row, lane, in_q, e_tok, e_pos, out_q);
}

// ---------------------------------------------------------------------------------------------------
// attention: T = 128 keys/queries per (batch, head), head_dim = 64
// ---------------------------------------------------------------------------------------------------
using attn::AT;
using attn::AD;
// 8 warps, two threads per query row; thread 0 also issues the TMA loads and the MMAs, warp 0 owns the
// TMEM allocation.  Resources are sized for THREE CTAs per SM (444 slots >= the 384 (batch, head)
// pairs of BERT-base at batch 32: one wave): 128 TMEM columns (O reuses the first 64 columns of S once
// the probabilities have left TMEM), ~52 KB shared memory (the P tile reuses the Q | K buffers once the
// score MMA has retired), <= 85 registers.
constexpr int kAttnThreads = 256;
constexpr int kAttnTmemCols = 128;      // S: [0,128)  then O: [0,64)
constexpr int kAttnSmem = 16384 * 3 + 512 + 2048 + 64 + 192 + 1024;   // Q K (later P) V | mask | max/sum exchange | barriers | quantizers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) {
            printf("tq_attention: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
// K-major SWIZZLE_128B operand (rows of 128 B, 8-row atoms 1024 B apart)
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B operand: 64 MN-elements (128 B) contiguous per K row, 8-row atoms 1024 B
// apart along K (stride byte offset); a single 64-wide MN block (leading byte offset unused)
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

struct AttnArgs {
    int32_t B, H;                 // batch, heads; rows of qkv = B * AT, columns = 3 * H * AD
    tq_qspec q_q, k_q, v_q;       // per-tensor quantizers of the Q / K / V projections (scales only)
    tq_qspec s_q, p_q, c_q;       // scores / probs / context quantizers (per-tensor)
    const float* mask;            // [B, AT] additive mask (0 / -10000) or null
    __nv_bfloat16* c_ctr;         // [B * AT, H * AD] centred context grid (or null)
    unsigned char* c_u8;          // [B * AT, H * AD] context x_int, one byte each (8-bit operand mode; or null)
    float inv_sqrt_d;             // 1 / sqrt(head_dim)
    float sqrt_d;                 // != 0: scores are DIVIDED by this value (true head_dim not a power of 4: 1 / sqrt(d) is not exact)
    int32_t qkv_params, c_params; // parameter slots of q_q / k_q / v_q and of c_q: 1, or a divisor of H (per-embedding-group
                                  // quantizers whose groups hold whole heads: head h uses slot h / (H / params))
};

__device__ __forceinline__ float scale_of(const tq_qspec& q) {
    float lo, hi;
    grid_of(q, lo, hi);
    return resolve(q, 0, lo, hi).scale;
}

__device__ __forceinline__ QP load_qp(const float* o) {
    QP p;
    p.scale = o[0]; p.zp = o[1]; p.lo = o[2]; p.hi = o[3]; p.rcp = o[4]; p.exact = __float_as_int(o[5]);
    return p;
}

// per-row work of the softmax / epilogue warps (tq_attn.cuh) around this kernel's barriers: thread 0 issues the PV MMA
template <bool FAST, bool DIVD>
__device__ __forceinline__ void attn_rows(const AttnArgs& a, const QP& qs, const QP& qp, const QP& qc, float sqk,
                                          float spv, uint32_t trow, int row, int quarter, int lane, int warp, int b,
                                          int h, int32_t dmodel, const float* smask, unsigned char* pP, float* xchg,
                                          uint32_t bar_s, uint32_t bar_p, uint32_t bar_o, uint32_t bar_v, uint32_t sP,
                                          uint32_t sV, uint32_t tmem) {
    const int hs = warp >> 2;                       // which half of the keys / of the context columns
    attn::RowArgs ra;
    ra.inv_sqrt_d = a.inv_sqrt_d; ra.sqrt_d = a.sqrt_d; ra.c_ctr = a.c_ctr; ra.c_u8 = a.c_u8;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    attn::softmax_rows<FAST, DIVD>(ra, qs, qp, sqk, trow, row, hs, smask, pP, xchg);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tensor core reads
    tc_fence_before();
    mbar_arrive(bar_p);
    if (threadIdx.x == 0) {
        // O[q, d] = sum_k P[q, k] * V[k, d]      (A = P K-major over keys; B = V MN-major); O overwrites
        // the first 64 score columns: every thread has read its scores before arriving on bar_p
        mbar_wait(bar_p, 0);
        mbar_wait(bar_v, 0);
        tc_fence_after();
        constexpr uint32_t id2 = idesc_bf16(128, 64, 1);
#pragma unroll
        for (int k = 0; k < AT / 16; ++k) {
            const uint64_t ad = desc_k_sw128(sP + (k >> 2) * 16384) + (uint64_t)(2 * (k & 3));
            const uint64_t bd = desc_mn_sw128(sV + k * 2048);
            tc_mma(tmem, ad, bd, id2, k != 0);
        }
        tc_commit(bar_o);
    }
    __syncwarp();
    // context: O * (s_p * s_v) -> QDQ -> centred bf16 / bytes; this thread owns 32 of the 64 head dims
    mbar_wait(bar_o, 0);
    tc_fence_after();
    attn::context_rows<FAST>(ra, qc, spv, trow, hs, ((int64_t)b * AT + row) * dmodel + h * AD + hs * 32, dmodel);
}

__global__ void __launch_bounds__(kAttnThreads, 3)
attention_kernel(const __grid_constant__ CUtensorMap map_qkv, AttnArgs a) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* bp = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t sQ = base, sK = base + 16384, sV = base + 32768;
    const uint32_t sP = base;                          // P (2 x 16 KB K-major halves) reuses Q | K after the score MMA
    unsigned char* pP = bp;
    float* smask = reinterpret_cast<float*>(bp + 49152);
    float* xchg = reinterpret_cast<float*>(bp + 49152 + 512);
    float* qsm = reinterpret_cast<float*>(bp + 49152 + 512 + 2048 + 64);      // [6][8] resolved quantizers
    const uint32_t bar0 = base + 49152 + 512 + 2048;
    const uint32_t bar_qk = bar0, bar_v = bar0 + 8, bar_s = bar0 + 16, bar_p = bar0 + 24, bar_o = bar0 + 32;
    const uint32_t tmem_slot = bar0 + 40;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(bp + 49152 + 512 + 2048 + 40);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const int32_t dmodel = a.H * AD;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
        mbar_init(bar_qk, 1);
        mbar_init(bar_v, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, kAttnThreads);
        mbar_init(bar_o, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // the loads start before the TMEM allocation / block barrier below
        pdl_wait();                                   // qkv is produced by the previous kernel
        mbar_expect_tx(bar_qk, 32768);
        tma_load_2d(sQ, &map_qkv, h * AD, b * AT, bar_qk);
        tma_load_2d(sK, &map_qkv, dmodel + h * AD, b * AT, bar_qk);
        mbar_expect_tx(bar_v, 16384);
        tma_load_2d(sV, &map_qkv, 2 * dmodel + h * AD, b * AT, bar_v);
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)kAttnTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 1 && lane < 6) {
        // quantizer parameters are calibration constants, not outputs of the previous kernel: no dependency wait
        const tq_qspec& q = lane == 0 ? a.s_q : lane == 1 ? a.p_q : lane == 2 ? a.c_q : lane == 3 ? a.q_q : lane == 4 ? a.k_q : a.v_q;
        float lo, hi;
        grid_of(q, lo, hi);
        const int np = lane == 2 ? a.c_params : (lane >= 3 ? a.qkv_params : 1);
        const QP p = resolve(q, np > 1 ? h / (a.H / np) : 0, lo, hi);
        float* o = qsm + lane * 8;
        o[0] = p.scale; o[1] = p.zp; o[2] = p.lo; o[3] = p.hi; o[4] = p.rcp; o[5] = __int_as_float(p.exact);
    }
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x < AT) smask[threadIdx.x] = a.mask != nullptr ? a.mask[(int64_t)b * AT + threadIdx.x] : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;

    if (threadIdx.x == 0) {
        // S[q, k] = sum_d Q[q, d] * K[k, d]      (both operands K-major, K = head_dim)
        mbar_wait(bar_qk, 0);
        tc_fence_after();
        constexpr uint32_t id1 = idesc_bf16(128, 128, 0);
#pragma unroll
        for (int k = 0; k < AD / 16; ++k)
            tc_mma(tmem, desc_k_sw128(sQ) + (uint64_t)(2 * k), desc_k_sw128(sK) + (uint64_t)(2 * k), id1, k != 0);
        tc_commit(bar_s);
    }
    __syncwarp();
    {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;                      // query index == TMEM lane
        const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
        // the six quantizers were resolved by six lanes of warp 1 before the block barrier (qsm): one global round
        // trip for the whole CTA instead of a chain of dependent loads in every thread
        const QP qs = load_qp(qsm + 0 * 8), qp = load_qp(qsm + 1 * 8), qc = load_qp(qsm + 2 * 8);
        const float sqk = qsm[3 * 8] * qsm[4 * 8];
        const float spv = qp.scale * qsm[5 * 8];

        if (a.sqrt_d != 0.0f) {
            if (qs.exact | qp.exact | qc.exact)
                attn_rows<false, true>(a, qs, qp, qc, sqk, spv, trow, row, quarter, lane, warp, b, h, dmodel, smask, pP, xchg, bar_s, bar_p, bar_o, bar_v, sP, sV, tmem);
            else
                attn_rows<true, true>(a, qs, qp, qc, sqk, spv, trow, row, quarter, lane, warp, b, h, dmodel, smask, pP, xchg, bar_s, bar_p, bar_o, bar_v, sP, sV, tmem);
        } else if (qs.exact | qp.exact | qc.exact)
            attn_rows<false, false>(a, qs, qp, qc, sqk, spv, trow, row, quarter, lane, warp, b, h, dmodel, smask, pP, xchg, bar_s, bar_p, bar_o, bar_v, sP, sV, tmem);
        else
            attn_rows<true, false>(a, qs, qp, qc, sqk, spv, trow, row, quarter, lane, warp, b, h, dmodel, smask, pP, xchg, bar_s, bar_p, bar_o, bar_v, sP, sV, tmem);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kAttnTmemCols)
                     : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------
// classification head: first token -> pooler (dense + tanh + QDQ) -> classifier (dense + QDQ)   (reference
// models/quantized_bert.py:525-560: QuantizedBertPooler, then the classifier QuantLinear).  B x hidden x hidden
// int8 products: two launches of the persistent GEMM (TMEM allocation, TMA ring, 128-row tiles for 32 rows) and a
// gather kernel took ~27 us of a 1 ms step; this is one CTA per sequence on the CUDA cores (dp4a), a warp per output
// column with coalesced weight rows.  Same arithmetic as tq_linear_qdq_i8's generic epilogue: exact integer
// accumulation, fma(acc - zp * rowsum, s_a * s_w, bias), tanhf, clamp(rint(RN(. / s))).
// ---------------------------------------------------------------------------------------------------
constexpr int kHeadThreads = 1024;
constexpr int kHeadIlp = 4;              // output columns per warp iteration: their weight rows are requested together (L2 latency)
constexpr int kHeadMaxWords = 8;            // hidden <= 1024: <= 8 packed words of the input row per lane

struct HeadArgs {
    const unsigned char* x;       // x_int bytes; row b of the head input = x + b * row_stride
    int64_t row_stride;
    int32_t D, L;                 // hidden size, classifier outputs
    const int8_t* wp; const int32_t* wp_rowsum; const float* bp;     // pooler [D, D]
    tq_qspec a_q, wp_q, pool_q;
    const int8_t* wc; const int32_t* wc_rowsum; const float* bc;     // classifier [>= L, D]
    tq_qspec wc_q, cls_q;
    float* logits;                // [B, ldl] dequantized classifier outputs
    int64_t ldl;
};

__device__ __forceinline__ int dp4a_us(uint32_t a_u8, uint32_t b_s8, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8), "r"(b_s8), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_uu(uint32_t a_u8, uint32_t b_u8, int c) {
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8), "r"(b_u8), "r"(c));
    return d;
}

__global__ void __launch_bounds__(kHeadThreads, 1) head_kernel(HeadArgs a) {
    __shared__ __align__(16) uint32_t xs[kHeadMaxWords * 32], ps[kHeadMaxWords * 32];
    __shared__ float qsm[5][8];
    __shared__ int wsigned[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kHeadThreads / 32;
    const int words = a.D >> 2, per_lane = words >> 5;
    if (warp == 0 && lane < 5) {                      // quantizers are calibration constants: resolved before the dependency wait
        const tq_qspec& q = lane == 0 ? a.a_q : lane == 1 ? a.wp_q : lane == 2 ? a.pool_q : lane == 3 ? a.wc_q : a.cls_q;
        float lo, hi;
        grid_of(q, lo, hi);
        const QP p = resolve(q, 0, lo, hi);
        float* o = qsm[lane];
        o[0] = p.scale; o[1] = p.zp; o[2] = p.lo; o[3] = p.hi; o[4] = p.rcp; o[5] = __int_as_float(p.exact);
        if (lane == 1 || lane == 3) wsigned[lane >> 1] = (q.zero_float == nullptr && q.is_signed != nullptr && *q.is_signed) ? 1 : 0;
    }
    pdl_trigger();
    pdl_wait();                                       // x is produced by the previous kernel
    const uint32_t* xrow = reinterpret_cast<const uint32_t*>(a.x + (int64_t)blockIdx.x * a.row_stride);
    for (int i = threadIdx.x; i < words; i += kHeadThreads) xs[i] = xrow[i];
    __syncthreads();
    const QP qa = load_qp(qsm[0]), qwp = load_qp(qsm[1]), qpool = load_qp(qsm[2]), qwc = load_qp(qsm[3]), qcls = load_qp(qsm[4]);
    uint32_t xw[kHeadMaxWords];
#pragma unroll
    for (int i = 0; i < kHeadMaxWords; ++i) xw[i] = i < per_lane ? xs[lane + 32 * i] : 0u;
    {   // pooler: tanh(dense(x)) -> QDQ -> x_int bytes in shared memory
        const float cs = __fmul_rn(qa.scale, qwp.scale);
        const int zp = (int)qa.zp;
        const bool sg = wsigned[0] != 0;
        unsigned char* pb = reinterpret_cast<unsigned char*>(ps);
        for (int n0 = warp * kHeadIlp; n0 < a.D; n0 += nwarps * kHeadIlp) {
            uint32_t w[kHeadIlp][kHeadMaxWords];
#pragma unroll
            for (int u = 0; u < kHeadIlp; ++u) {
                const int n = n0 + u < a.D ? n0 + u : a.D - 1;
                const uint32_t* wrow = reinterpret_cast<const uint32_t*>(a.wp + (int64_t)n * a.D);
#pragma unroll
                for (int i = 0; i < kHeadMaxWords; ++i)
                    if (i < per_lane) w[u][i] = __ldg(wrow + lane + 32 * i);
            }
            int acc[kHeadIlp];
#pragma unroll
            for (int u = 0; u < kHeadIlp; ++u) {
                acc[u] = 0;
#pragma unroll
                for (int i = 0; i < kHeadMaxWords; ++i)
                    if (i < per_lane) acc[u] = sg ? dp4a_us(xw[i], w[u][i], acc[u]) : dp4a_uu(xw[i], w[u][i], acc[u]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int u = 0; u < kHeadIlp; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
            if (lane < kHeadIlp && n0 + lane < a.D) {         // lane u finishes output n0 + u
                const int n = n0 + lane;
                int mine = acc[0];
#pragma unroll
                for (int u = 1; u < kHeadIlp; ++u) mine = lane == u ? acc[u] : mine;
                const float pre = __fmaf_rn(__int2float_rn(mine - zp * __ldg(a.wp_rowsum + n)), cs, a.bp != nullptr ? __ldg(a.bp + n) : 0.0f);
                pb[n] = (unsigned char)(int)quant_int(tanhf(pre), qpool);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kHeadMaxWords; ++i) xw[i] = i < per_lane ? ps[lane + 32 * i] : 0u;
    {   // classifier: dense(pooled) -> QDQ -> dequantized logits
        const float cs = __fmul_rn(qpool.scale, qwc.scale);
        const int zp = (int)qpool.zp;
        const bool sg = wsigned[1] != 0;
        for (int n = warp; n < a.L; n += nwarps) {
            const uint32_t* wrow = reinterpret_cast<const uint32_t*>(a.wc + (int64_t)n * a.D);
            int acc = 0;
#pragma unroll
            for (int i = 0; i < kHeadMaxWords; ++i)
                if (i < per_lane) {
                    const uint32_t w = __ldg(wrow + lane + 32 * i);
                    acc = sg ? dp4a_us(xw[i], w, acc) : dp4a_uu(xw[i], w, acc);
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                const float pre = __fmaf_rn(__int2float_rn(acc - zp * __ldg(a.wc_rowsum + n)), cs, a.bc != nullptr ? __ldg(a.bc + n) : 0.0f);
                a.logits[(int64_t)blockIdx.x * a.ldl + n] = __fmul_rn(qcls.scale, __fsub_rn(quant_int(pre, qcls), qcls.zp));
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace fused
}  // namespace tq

extern "C" {

static int attention_impl(const void* qkv_ctr_bf16, void* c_ctr_bf16, void* c_u8, int32_t B, int32_t T, int32_t H,
                          int32_t head_dim, tq_qspec q_q, tq_qspec k_q, tq_qspec v_q, tq_qspec s_q, tq_qspec p_q,
                          tq_qspec c_q, const float* mask, void* stream, int32_t qkv_params = 1, int32_t c_params = 1,
                          int32_t true_head_dim = 0) {
    using namespace tq::fused;
    if (qkv_ctr_bf16 == nullptr || (c_ctr_bf16 == nullptr && c_u8 == nullptr) || B < 1 || H < 1) return TQ_EINVAL;
    if (T != AT || head_dim != AD) return TQ_EUNSUPPORTED;
    if (!tq::aligned16(qkv_ctr_bf16) || !tq::aligned16(c_ctr_bf16) || !tq::aligned16(c_u8)) return TQ_EALIGN;
    if (c_u8 != nullptr && c_q.n_bits > 8) return TQ_EUNSUPPORTED;
    const tq_qspec* all[6] = {&q_q, &k_q, &v_q, &s_q, &p_q, &c_q};
    for (int i = 0; i < 6; ++i)
        if (int e = tq::check_qspec(*all[i])) return e;
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return TQ_EUNSUPPORTED;
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)(3 * H * AD), (cuuint64_t)B * AT};
    const cuuint64_t strides[1] = {(cuuint64_t)(3 * H * AD) * 2};
    const cuuint32_t box[2] = {(cuuint32_t)AD, (cuuint32_t)AT};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv_ctr_bf16), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return TQ_EINVAL;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
        if (e != cudaSuccess) return (int)e;
        cudaFuncSetAttribute(attention_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    AttnArgs a;
    a.B = B;
    a.H = H;
    a.q_q = q_q; a.k_q = k_q; a.v_q = v_q; a.s_q = s_q; a.p_q = p_q; a.c_q = c_q;
    a.mask = mask;
    a.c_ctr = reinterpret_cast<__nv_bfloat16*>(c_ctr_bf16);
    a.c_u8 = reinterpret_cast<unsigned char*>(c_u8);
    if (true_head_dim < 0 || true_head_dim > head_dim) return TQ_EINVAL;
    const int32_t dd = true_head_dim > 0 ? true_head_dim : head_dim;
    a.inv_sqrt_d = 1.0f / sqrtf((float)dd);
    const float sq = sqrtf((float)dd);                 // math.sqrt(d) in the reference (float64 there; the tensor is fp32)
    a.sqrt_d = (sq * sq == (float)dd && (dd & (dd - 1)) == 0) ? 0.0f : sq;      // power of 4: multiply by the exact reciprocal
    if (qkv_params < 1 || c_params < 1 || H % qkv_params != 0 || H % c_params != 0) return TQ_EINVAL;
    a.qkv_params = qkv_params;
    a.c_params = c_params;
    return tq::launch_pdl(attention_kernel, dim3(B * H), dim3(kAttnThreads), kAttnSmem, (cudaStream_t)stream, 1, map, a);
}

int tq_attention_qdq_bf16(const void* qkv_ctr_bf16, void* c_ctr_bf16, int32_t B, int32_t T, int32_t H,
                          int32_t head_dim, tq_qspec q_q, tq_qspec k_q, tq_qspec v_q, tq_qspec s_q, tq_qspec p_q,
                          tq_qspec c_q, const float* mask, void* stream) {
    if (c_ctr_bf16 == nullptr) return TQ_EINVAL;
    return attention_impl(qkv_ctr_bf16, c_ctr_bf16, nullptr, B, T, H, head_dim, q_q, k_q, v_q, s_q, p_q, c_q, mask, stream);
}

int tq_attention_qdq_i8(const void* qkv_ctr_bf16, void* c_i8, int32_t B, int32_t T, int32_t H, int32_t head_dim,
                        tq_qspec q_q, tq_qspec k_q, tq_qspec v_q, tq_qspec s_q, tq_qspec p_q, tq_qspec c_q,
                        const float* mask, void* stream) {
    if (c_i8 == nullptr) return TQ_EINVAL;
    return attention_impl(qkv_ctr_bf16, nullptr, c_i8, B, T, H, head_dim, q_q, k_q, v_q, s_q, p_q, c_q, mask, stream);
}

int tq_attention_peg_qdq_i8(const void* qkv_ctr_bf16, void* c_i8, int32_t B, int32_t T, int32_t H, int32_t head_dim,
                            tq_qspec q_q, tq_qspec k_q, tq_qspec v_q, int32_t qkv_params, tq_qspec s_q, tq_qspec p_q,
                            tq_qspec c_q, int32_t c_params, const float* mask, void* stream) {
    if (c_i8 == nullptr) return TQ_EINVAL;
    return attention_impl(qkv_ctr_bf16, nullptr, c_i8, B, T, H, head_dim, q_q, k_q, v_q, s_q, p_q, c_q, mask, stream, qkv_params,
                          c_params);
}

int tq_attention_pad_qdq_i8(const void* qkv_ctr_bf16, void* c_i8, int32_t B, int32_t T, int32_t H, int32_t head_dim,
                            int32_t true_head_dim, tq_qspec q_q, tq_qspec k_q, tq_qspec v_q, tq_qspec s_q, tq_qspec p_q,
                            tq_qspec c_q, const float* mask, void* stream) {
    if (c_i8 == nullptr) return TQ_EINVAL;
    return attention_impl(qkv_ctr_bf16, nullptr, c_i8, B, T, H, head_dim, q_q, k_q, v_q, s_q, p_q, c_q, mask, stream, 1, 1,
                          true_head_dim);
}

int tq_head_qdq_i8(const void* x_i8, int64_t row_stride, int32_t B, int32_t D, int32_t L, const void* wp_i8, const int32_t* wp_rowsum,
                   const float* bp, tq_qspec a_q, tq_qspec wp_q, tq_qspec pool_q, const void* wc_i8, const int32_t* wc_rowsum,
                   const float* bc, tq_qspec wc_q, tq_qspec cls_q, float* logits, int64_t ldl, void* stream) {
    using namespace tq::fused;
    if (x_i8 == nullptr || wp_i8 == nullptr || wp_rowsum == nullptr || wc_i8 == nullptr || wc_rowsum == nullptr || logits == nullptr) return TQ_EINVAL;
    if (B < 1 || L < 1 || ldl < L || row_stride < D) return TQ_EINVAL;
    if (D < 128 || (D & 127) != 0 || D > 128 * kHeadMaxWords) return TQ_EUNSUPPORTED;
    const tq_qspec* all[5] = {&a_q, &wp_q, &pool_q, &wc_q, &cls_q};
    for (int i = 0; i < 5; ++i) {
        if (int e = tq::check_qspec(*all[i])) return e;
        if (all[i]->n_bits > 8) return TQ_EUNSUPPORTED;
    }
    if (a_q.zero_float == nullptr || pool_q.zero_float == nullptr) return TQ_EUNSUPPORTED;      // x_int bytes: unsigned activation grids
    if ((reinterpret_cast<uintptr_t>(x_i8) & 3u) != 0 || (row_stride & 3) != 0 || (reinterpret_cast<uintptr_t>(wp_i8) & 3u) != 0 ||
        (reinterpret_cast<uintptr_t>(wc_i8) & 3u) != 0)
        return TQ_EALIGN;
    HeadArgs a;
    a.x = static_cast<const unsigned char*>(x_i8); a.row_stride = row_stride; a.D = D; a.L = L;
    a.wp = static_cast<const int8_t*>(wp_i8); a.wp_rowsum = wp_rowsum; a.bp = bp;
    a.a_q = a_q; a.wp_q = wp_q; a.pool_q = pool_q;
    a.wc = static_cast<const int8_t*>(wc_i8); a.wc_rowsum = wc_rowsum; a.bc = bc;
    a.wc_q = wc_q; a.cls_q = cls_q; a.logits = logits; a.ldl = ldl;
    return tq::launch_pdl(head_kernel, dim3((unsigned)B), dim3(kHeadThreads), 0, (cudaStream_t)stream, 1, a);
}

static int ln_common(tq::fused::LnArgs& a, bool embed, void* stream) {
    using namespace tq::fused;
    if (a.M < 1 || a.D < 256 || (a.D & 255) != 0 || a.D > 256 * kLnMaxIter) return TQ_EUNSUPPORTED;
    if (a.gamma_q == nullptr || a.beta == nullptr || (a.out_ctr == nullptr && a.out_u8 == nullptr)) return TQ_EINVAL;
    if (int e = tq::check_qspec(a.out_q.q)) return e;
    const unsigned grid = (unsigned)((a.M + kLnWarps - 1) / kLnWarps);
    static bool carve_set = false;
    if (!carve_set) {   // same shared-memory carveout as the GEMMs around it: no SM reconfiguration between kernels
        cudaFuncSetAttribute(ln_qdq_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(ln_qdq_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carve_set = true;
    }
    if (embed) return tq::launch_pdl(ln_qdq_kernel<true>, dim3(grid), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, 1, a);
    return tq::launch_pdl(ln_qdq_kernel<false>, dim3(grid), dim3(kLnWarps * 32), 0, (cudaStream_t)stream, 1, a);
}

int tq_ln_qdq_bf16(const void* x_ctr_bf16, tq_qspec in_q, int64_t in_q_params, const float* gamma_q,
                   const float* beta, float eps, tq_qspec out_q, int64_t out_q_params, void* out_ctr_bf16,
                   float* out_f32, int64_t M, int32_t D, void* stream) {
    tq::fused::LnArgs a = {};
    if (x_ctr_bf16 == nullptr) return TQ_EINVAL;
    if (int e = tq::check_qspec(in_q)) return e;
    a.x_ctr = reinterpret_cast<const __nv_bfloat16*>(x_ctr_bf16);
    a.in_q.q = in_q;
    a.in_q.params = in_q_params;
    a.gamma_q = gamma_q;
    a.beta = beta;
    a.eps = eps;
    a.out_q.q = out_q;
    a.out_q.params = out_q_params;
    a.out_ctr = reinterpret_cast<__nv_bfloat16*>(out_ctr_bf16);
    a.out_f32 = out_f32;
    a.M = M;
    a.D = D;
    return ln_common(a, false, stream);
}

static int embed_impl(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int64_t T,
                      const float* word_q, const float* type_q, const float* pos_q, tq_qspec e_tok,
                      int64_t e_tok_params, tq_qspec e_pos, int64_t e_pos_params, const float* gamma_q,
                      const float* beta, float eps, tq_qspec out_q, int64_t out_q_params, void* out_ctr_bf16,
                      void* out_u8, float* out_f32, int64_t M, int32_t D, void* stream) {
    tq::fused::LnArgs a = {};
    a.out_u8 = reinterpret_cast<unsigned char*>(out_u8);
    if (out_u8 != nullptr && out_q.n_bits > 8) return TQ_EUNSUPPORTED;
    if (ids == nullptr || word_q == nullptr || type_q == nullptr || pos_q == nullptr || T < 1) return TQ_EINVAL;
    if (int e = tq::check_qspec(e_tok)) return e;
    if (int e = tq::check_qspec(e_pos)) return e;
    a.ids = ids;
    a.type_ids = type_ids;
    a.pos_ids = pos_ids;
    a.T = T;
    a.word = word_q;
    a.type_tab = type_q;
    a.pos_tab = pos_q;
    a.e_tok.q = e_tok;
    a.e_tok.params = e_tok_params;
    a.e_pos.q = e_pos;
    a.e_pos.params = e_pos_params;
    a.gamma_q = gamma_q;
    a.beta = beta;
    a.eps = eps;
    a.out_q.q = out_q;
    a.out_q.params = out_q_params;
    a.out_ctr = reinterpret_cast<__nv_bfloat16*>(out_ctr_bf16);
    a.out_f32 = out_f32;
    a.M = M;
    a.D = D;
    return ln_common(a, true, stream);
}

int tq_embed_ln_qdq_bf16(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int64_t T,
                         const float* word_q, const float* type_q, const float* pos_q, tq_qspec e_tok,
                         int64_t e_tok_params, tq_qspec e_pos, int64_t e_pos_params, const float* gamma_q,
                         const float* beta, float eps, tq_qspec out_q, int64_t out_q_params, void* out_ctr_bf16,
                         float* out_f32, int64_t M, int32_t D, void* stream) {
    if (out_ctr_bf16 == nullptr) return TQ_EINVAL;
    return embed_impl(ids, type_ids, pos_ids, T, word_q, type_q, pos_q, e_tok, e_tok_params, e_pos, e_pos_params, gamma_q,
                      beta, eps, out_q, out_q_params, out_ctr_bf16, nullptr, out_f32, M, D, stream);
}

int tq_embed_ln_qdq_i8(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int64_t T,
                       const float* word_q, const float* type_q, const float* pos_q, tq_qspec e_tok,
                       int64_t e_tok_params, tq_qspec e_pos, int64_t e_pos_params, const float* gamma_q,
                       const float* beta, float eps, tq_qspec out_q, int64_t out_q_params, void* out_i8, int64_t M,
                       int32_t D, void* stream) {
    if (out_i8 == nullptr) return TQ_EINVAL;
    return embed_impl(ids, type_ids, pos_ids, T, word_q, type_q, pos_q, e_tok, e_tok_params, e_pos, e_pos_params, gamma_q,
                      beta, eps, out_q, out_q_params, nullptr, out_i8, nullptr, M, D, stream);
}

}  // extern "C"
