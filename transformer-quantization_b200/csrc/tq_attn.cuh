// Per-row work of the fake-quantized attention block (reference models/quantized_bert.py:153-213), shared by the
// stand-alone attention kernel (tq_fused.cu: one CTA per (batch, head)) and the attention stage of the encoder chain
// kernel (tq_linear.cu namespace chain: a cluster per sequence).  T = 128 keys / queries, head_dim = 64; the caller owns
// the MMAs and the barriers around these two pieces:
//
//   softmax_rows   S (fp32, TMEM) -> QDQ(scores) -> / sqrt(d) + mask -> softmax -> QDQ(probs) -> centred bf16 integers in
//                  the K-major SWIZZLE_128B A tile of the PV product (shared memory)
//   context_rows   O (fp32, TMEM) * (s_p * s_v) -> QDQ(context) -> centred bf16 grid and / or x_int bytes
//
// TWO threads per query row: warps w and w + 4 of the eight share a TMEM lane quarter, each owns one 64-key half of the
// row (= one swizzle span of P) and 32 of the 64 head dimensions of the context.  Named barrier 1 (256 threads) is used
// for the two row exchanges.
#pragma once
#include "tq_common.cuh"
#include <cuda_bf16.h>

namespace tq {
namespace attn {

constexpr int AT = 128, AD = 64;

struct RowArgs {
    float inv_sqrt_d;             // 1 / sqrt(head_dim)
    float sqrt_d;                 // != 0: scores are DIVIDED by this value (true head_dim not a power of 4: 1 / sqrt(d) is not exact)
    __nv_bfloat16* c_ctr;         // [B * AT, dmodel] centred context grid (or null)
    unsigned char* c_u8;          // [B * AT, dmodel] context x_int, one byte each (8-bit operand mode; or null)
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st16_nowait(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// hs: which half of the keys (0 / 1) this thread owns; trow: TMEM address of this thread's lane quarter, column 0 of S;
// smask: [AT] additive mask of the sequence; pP: the P tile (2 x 16 KB K-major halves); xchg: [4][AT] floats
template <bool FAST, bool DIVD>
__device__ __forceinline__ void softmax_rows(const RowArgs& a, const QP& qs, const QP& qp, float sqk, uint32_t trow, int row, int hs,
                                             const float* smask, unsigned char* pP, float* xchg) {
    const int k0 = hs * 64;
    const QP2 qs2 = pair_of(qs), qp2 = pair_of(qp);
    const float2 sqk2 = make_float2(sqk, sqk), inv2 = make_float2(a.inv_sqrt_d, a.inv_sqrt_d);
    const float2 sinv2 = make_float2(__fmul_rn(qs.scale, a.inv_sqrt_d), __fmul_rn(qs.scale, a.inv_sqrt_d));
    // pass 1: scores -> QDQ -> / sqrt(d) + mask, written back to TMEM; row max
    float vmax = __int_as_float(0xff800000);
#pragma unroll 1
    for (int c0 = k0; c0 < k0 + 64; c0 += 16) {
        uint32_t v[16];
        ld16(trow + c0, v);
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float2 sc = __fmul2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sqk2);
            float2 t;                                                               // quantized_bert.py:153-154
            if (FAST && !DIVD) {
                // centred integers, then ONE multiply by scale / sqrt(d): 1 / sqrt(d) is a power of two here, so
                // fl(fl(scale * c) / sqrt(d)) == fl((scale / sqrt(d)) * c) bit for bit -- three packed FP32 ops fewer
                t = __fmul2_rn(quant_ctr2_finite(sc, qs2), sinv2);
            } else {
                if (FAST) t = dequant2(quant_int2_finite(sc, qs2), qs2);
                else t = make_float2(qdq_t<false>(sc.x, qs), qdq_t<false>(sc.y, qs));
                if (DIVD) t = make_float2(__fdiv_rn(t.x, a.sqrt_d), __fdiv_rn(t.y, a.sqrt_d));       // scores / math.sqrt(d)
                else t = __fmul2_rn(t, inv2);                                                         // (exact for d = 4^n)
            }
            t = __fadd2_rn(t, *reinterpret_cast<const float2*>(smask + c0 + j));                      // :190-194
            vmax = fmaxf(vmax, fmaxf(t.x, t.y));
            v[j] = __float_as_uint(t.x);
            v[j + 1] = __float_as_uint(t.y);
        }
        st16_nowait(trow + c0, v);
    }
    xchg[hs * AT + row] = vmax;
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    vmax = fmaxf(vmax, xchg[(hs ^ 1) * AT + row]);
    // pass 2: exp(t - max) = 2^(t * log2 e - max * log2 e): one packed FMA and one ex2 per score.  (The reference's exp
    // is a libm / CUDA expf within 1-2 ulp; ex2.approx adds <= 2 ulp and the single rounding of the exponent
    // |t - max| * 2^-24 <~ 5e-6 relative -- far inside the probability quantizer's step; the flip budget of the
    // attention block, tests/test_gpu_fullsize_parity.py, covers it.)  Row sum: two interleaved partial sums.
    const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
    const float nm = -__fmul_rn(vmax, 1.4426950408889634f);
    const float2 nm2 = make_float2(nm, nm);
    float2 vs2 = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int c0 = k0; c0 < k0 + 64; c0 += 16) {
        uint32_t v[16];
        ld16(trow + c0, v);
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float2 u = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), l2e, nm2);
            float2 e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(u.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(u.y));
            vs2 = __fadd2_rn(vs2, e);
            v[j] = __float_as_uint(e.x);
            v[j + 1] = __float_as_uint(e.y);
        }
        st16_nowait(trow + c0, v);
    }
    float vsum = __fadd_rn(vs2.x, vs2.y);
    xchg[2 * AT + hs * AT + row] = vsum;
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // both threads of a row add the two partial sums in the same order (low half + high half)
    vsum = xchg[2 * AT + row] + xchg[3 * AT + row];
    const float rsum = __frcp_rn(vsum);
    const bool ieee = (__float_as_uint(vsum) & 0x7fffffu) == 0x7fffffu;
    const float2 r2 = make_float2(rsum, rsum), nd2 = make_float2(-vsum, -vsum);
    // pass 3: probs -> QDQ -> centred integers into this thread's swizzle span of the K-major A tile
    uint4* prow = reinterpret_cast<uint4*>(pP + hs * 16384 + row * 128);
#pragma unroll 1
    for (int c0 = k0; c0 < k0 + 64; c0 += 16) {
        uint32_t v[16];
        ld16(trow + c0, v);
        float c[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float2 e2 = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            float2 pr;                                                              // softmax, :197
            if (ieee) {
                pr = make_float2(__fdiv_rn(e2.x, vsum), __fdiv_rn(e2.y, vsum));
            } else {                                                                // exact quotient, see div_by
                const float2 q0 = __fmul2_rn(e2, r2);
                const float2 q1 = __ffma2_rn(__ffma2_rn(q0, nd2, e2), r2, q0);
                pr = __ffma2_rn(__ffma2_rn(q1, nd2, e2), r2, q1);
            }
            float2 ci;                                                              // :198
            if (FAST) ci = quant_ctr2_finite(pr, qp2);
            else ci = make_float2(__fsub_rn(quant_int_t<false>(pr.x, qp), qp.zp), __fsub_rn(quant_int_t<false>(pr.y, qp), qp.zp));
            c[j] = ci.x;
            c[j + 1] = ci.y;
        }
        const int ch0 = (c0 & 63) >> 3;                  // first 16-byte chunk of this row piece
        uint4 w0, w1;
        w0.x = pack2(c[0], c[1]); w0.y = pack2(c[2], c[3]); w0.z = pack2(c[4], c[5]); w0.w = pack2(c[6], c[7]);
        w1.x = pack2(c[8], c[9]); w1.y = pack2(c[10], c[11]); w1.z = pack2(c[12], c[13]); w1.w = pack2(c[14], c[15]);
        prow[(ch0) ^ (row & 7)] = w0;
        prow[(ch0 + 1) ^ (row & 7)] = w1;
    }
}

// spv = s_p * s_v; ooff: element offset of this thread's 32 context values ((b * AT + row) * dmodel + h * AD + hs * 32)
template <bool FAST>
__device__ __forceinline__ void context_rows(const RowArgs& a, const QP& qc, float spv, uint32_t trow, int hs, int64_t ooff, int32_t dmodel) {
    const QP2 qc2 = pair_of(qc);
    const float2 spv2 = make_float2(spv, spv);
    __nv_bfloat16* orow = a.c_ctr + ooff;
    const bool wide = ((((uintptr_t)a.c_ctr) & 31u) == 0) && ((dmodel & 15) == 0);
    uint32_t bytes[8];                                  // 32 context values as x_int bytes (8-bit operand mode)
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
        const int c0 = hs * 32 + i * 16;
        uint32_t v[16];
        ld16(trow + c0, v);
        uint32_t w[8];
        uint32_t xi[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {                                           // :201-213
            const float2 cv = __fmul2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), spv2);
            float2 ci;
            if (FAST) ci = quant_ctr2_finite(cv, qc2);
            else ci = make_float2(__fsub_rn(quant_int_t<false>(cv.x, qc), qc.zp), __fsub_rn(quant_int_t<false>(cv.y, qc), qc.zp));
            w[j >> 1] = pack2(ci.x, ci.y);
            // x_int bytes without F2I: v + 1.5 * 2^23 keeps the integer in the low mantissa bits
            xi[j] = __float_as_uint(__fadd_rn(ci.x, __fadd_rn(qc.zp, 12582912.0f)));
            xi[j + 1] = __float_as_uint(__fadd_rn(ci.y, __fadd_rn(qc.zp, 12582912.0f)));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t packed = __byte_perm(__byte_perm(xi[4 * q], xi[4 * q + 1], 0x0040),
                                                __byte_perm(xi[4 * q + 2], xi[4 * q + 3], 0x0040), 0x5410);
            if (i == 0) bytes[q] = packed; else bytes[4 + q] = packed;
        }
        if (a.c_ctr == nullptr) continue;
        if (wide) {
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + i * 16), "r"(w[0]),
                         "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                         : "memory");
        } else {
            *reinterpret_cast<uint4*>(orow + i * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(orow + i * 16 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
    if (a.c_u8 != nullptr) {
        unsigned char* o8 = a.c_u8 + ooff;               // 32 bytes: one sector
        if ((((uintptr_t)a.c_u8) & 31u) == 0 && (dmodel & 31) == 0) {
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o8), "r"(bytes[0]),
                         "r"(bytes[1]), "r"(bytes[2]), "r"(bytes[3]), "r"(bytes[4]), "r"(bytes[5]), "r"(bytes[6]),
                         "r"(bytes[7])
                         : "memory");
        } else {
            *reinterpret_cast<uint4*>(o8) = make_uint4(bytes[0], bytes[1], bytes[2], bytes[3]);
            *reinterpret_cast<uint4*>(o8 + 16) = make_uint4(bytes[4], bytes[5], bytes[6], bytes[7]);
        }
    }
}

}  // namespace attn
}  // namespace tq
