// Shared device helpers for libtq_b200 (sm_100a).  Parity-critical arithmetic uses explicit
// round-to-nearest intrinsics (__fdiv_rn, __fmul_rn, __fadd_rn, __fsub_rn) so that nvcc can never
// contract it into FMAs or replace the division by a reciprocal multiply: the reference computes
// round(x / scale) with IEEE division (quantizers.py:184) and the integer result must be bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/tq_b200.h"

#define TQ_SM_COUNT_FALLBACK 148

namespace tq {

// ---- resolved quantizer parameters ----------------------------------------------------------
struct QP {
    float scale, zp, lo, hi;
    float rcp;     // RN(1 / scale)
    int exact;     // 1: use the IEEE division instruction (see div_rn)
};

// integer grid of the quantizer (quantizers.py:131-140, 321-328)
__device__ __forceinline__ void grid_of(const tq_qspec& q, float& lo, float& hi) {
    if (q.zero_float != nullptr) {
        lo = 0.0f;
        hi = (float)(1u << q.n_bits) - 1.0f;          // exact: n_bits <= 16
    } else {
        const bool sg = q.is_signed != nullptr && (*q.is_signed != 0);
        lo = sg ? -(float)(1u << (q.n_bits - 1)) : 0.0f;
        hi = (float)(1u << (q.n_bits - (sg ? 1 : 0))) - 1.0f;
    }
}

__device__ __forceinline__ QP make_qp(float scale, float zp, float lo, float hi) {
    QP p;
    p.scale = scale;
    p.zp = zp;
    p.lo = lo;
    p.hi = hi;
    p.rcp = __frcp_rn(scale);
    // Markstein's correction (div_rn) is proven for a correctly rounded reciprocal of a NORMAL
    // divisor whose significand is not all ones; anything else takes the division instruction.
    const uint32_t b = __float_as_uint(scale);
    const uint32_t ex = (b >> 23) & 0xffu;
    p.exact = ((b & 0x7fffffu) == 0x7fffffu) || ex == 0u || ex >= 0xfeu || ex <= 2u ? 1 : 0;
    return p;
}

// scale / zero_point of parameter slot i (quantizers.py:142-153, 330-332)
__device__ __forceinline__ QP resolve(const tq_qspec& q, int64_t i, float lo, float hi) {
    const float d = __ldg(q.delta + i);
    const float scale = q.log_domain ? expf(d) : (d < q.eps ? q.eps : d);   // torch.clamp(min=eps) keeps NaN
    float zp = 0.0f;
    if (q.zero_float != nullptr) {
        float z = rintf(__ldg(q.zero_float + i));
        z = z < lo ? lo : z;
        z = z > hi ? hi : z;
        zp = z;
    }
    return make_qp(scale, zp, lo, hi);
}

// RN(x / s) without the division instruction.  The reference computes round(x / scale) with a true
// IEEE division (quantizers.py:184) and the integer must be bit-exact, so a reciprocal multiply is
// not enough: q0 = RN(x*r) can be 2 ulp off.  Two FMA residual corrections with r = RN(1/s)
// (Markstein 1990; Muller et al., Handbook of FP Arithmetic, thm. on division by FMA iterations):
//   q1 = RN(q0 + (x - q0*s)*r) is a faithful quotient, q2 = RN(q1 + (x - q1*s)*r) == RN(x/s).
// The residuals are exact (FMA).  Outside |q| < 2^22 (or NaN/inf) q0 is returned: those values are
// clamped to the grid edge (<= 2^16) no matter how they round, and inf - inf must not appear.
// Verified bit-for-bit against __fdiv_rn on the GPU by tq_selftest_div (tests/test_gpu_parity.py).
// This keeps the XU pipe (MUFU.RCP, 16 lanes/SM) out of the per-element path.
//
// FAST is a COMPILE-TIME switch: a per-element run-time test of p.exact puts a branch (BSSY/BSYNC)
// between the independent element chains and stops ptxas from interleaving them (measured: 4-5x
// slower epilogues).  Kernels evaluate `p.exact` once (it is uniform) and pick the instantiation.
template <bool FAST>
__device__ __forceinline__ float div_rn_t(float x, const QP& p) {
    if (!FAST) return __fdiv_rn(x, p.scale);
    const float q0 = __fmul_rn(x, p.rcp);
    const float e0 = __fmaf_rn(-q0, p.scale, x);
    const float q1 = __fmaf_rn(e0, p.rcp, q0);
    const float e1 = __fmaf_rn(-q1, p.scale, x);
    const float q2 = __fmaf_rn(e1, p.rcp, q1);
    return fabsf(q0) < 4194304.0f ? q2 : q0;
}

// round-half-to-even == torch.round for |v| < 2^22 via the 1.5 * 2^23 constant (two FADDs on the
// FMA pipe instead of FRND on the XU pipe); larger |v|, inf and NaN pass through to the clamp.
__device__ __forceinline__ float rint_even(float v) {
    return __fsub_rn(__fadd_rn(v, 12582912.0f), 12582912.0f);
}

// clamp(rint(x / scale) + zp, lo, hi)  -- quantizers.py:184-185.  NaN propagates (torch.clamp).
template <bool FAST>
__device__ __forceinline__ float quant_int_t(float x, const QP& p) {
    const float d = div_rn_t<FAST>(x, p);
    // |d| >= 2^22: rint_even would return a non-integer for 2^22 <= |d| < 2^23 -- irrelevant, the
    // clamp maps all of them to the grid edge
    float q = __fadd_rn(rint_even(d), p.zp);
    q = q < p.lo ? p.lo : q;
    q = q > p.hi ? p.hi : q;
    return q;
}
// Same integer for FINITE x inside div_rn's domain, 5 instructions shorter: no inf/NaN guard on the
// quotient (a finite x cannot produce inf - inf in the residuals) and min/max instead of the
// NaN-propagating compare/select clamp.  Used by the GEMM / attention epilogues, whose inputs are
// products of finite integer grids and scales; a NaN bias or scale would saturate instead of
// propagating (documented deviation from torch.clamp for non-finite layer outputs).
__device__ __forceinline__ float quant_int_finite(float x, const QP& p) {
    const float q0 = __fmul_rn(x, p.rcp);
    const float q1 = __fmaf_rn(__fmaf_rn(-q0, p.scale, x), p.rcp, q0);
    const float q2 = __fmaf_rn(__fmaf_rn(-q1, p.scale, x), p.rcp, q1);
    return fminf(fmaxf(__fadd_rn(rint_even(q2), p.zp), p.lo), p.hi);
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2: two IEEE round-to-nearest fp32
// operations per issue slot).  The quantizer chain is FMA-pipe bound in every kernel of this library
// (~12 FP32 instructions per element); processing elements in pairs halves the issue slots.  Each
// component is rounded exactly like the scalar instruction, so results are bit-identical.
struct QP2 {
    float2 scale, nscale, rcp, zp, nzp;   // per-component parameters (nscale = -scale, nzp = -zp)
    float lo, hi;
};
__device__ __forceinline__ QP2 pair_of(const QP& a, const QP& b) {
    QP2 p;
    p.scale = make_float2(a.scale, b.scale);
    p.nscale = make_float2(-a.scale, -b.scale);
    p.rcp = make_float2(a.rcp, b.rcp);
    p.zp = make_float2(a.zp, b.zp);
    p.nzp = make_float2(-a.zp, -b.zp);
    p.lo = a.lo;
    p.hi = a.hi;
    return p;
}
__device__ __forceinline__ QP2 pair_of(const QP& a) { return pair_of(a, a); }

__device__ __forceinline__ float clamp_nan(float q, float lo, float hi) {   // torch.clamp: NaN propagates
    q = q < lo ? lo : q;
    return q > hi ? hi : q;
}

// RN(x / scale) for both components, division-free (see div_rn_t)
__device__ __forceinline__ float2 quot2(float2 x, const QP2& p) {
    const float2 q0 = __fmul2_rn(x, p.rcp);
    const float2 q1 = __ffma2_rn(__ffma2_rn(q0, p.nscale, x), p.rcp, q0);
    return __ffma2_rn(__ffma2_rn(q1, p.nscale, x), p.rcp, q1);
}
__device__ __forceinline__ float2 rint2_plus(float2 q, float2 zp) {          // rint_even(q) + zp
    const float2 M = make_float2(12582912.0f, 12582912.0f), nM = make_float2(-12582912.0f, -12582912.0f);
    return __fadd2_rn(__fadd2_rn(__fadd2_rn(q, M), nM), zp);
}

// clamp(rint(x / scale) + zp, lo, hi) for a pair; FAST as in quant_int_t
template <bool FAST>
__device__ __forceinline__ float2 quant_int2_t(float2 x, const QP2& p) {
    float2 q;
    if (FAST) {
        const float2 q0 = __fmul2_rn(x, p.rcp);
        float2 q2 = quot2(x, p);
        q2.x = fabsf(q0.x) < 4194304.0f ? q2.x : q0.x;       // inf / NaN / huge: see div_rn_t
        q2.y = fabsf(q0.y) < 4194304.0f ? q2.y : q0.y;
        q = rint2_plus(q2, p.zp);
    } else {
        q = rint2_plus(make_float2(__fdiv_rn(x.x, p.scale.x), __fdiv_rn(x.y, p.scale.y)), p.zp);
    }
    q.x = clamp_nan(q.x, p.lo, p.hi);
    q.y = clamp_nan(q.y, p.lo, p.hi);
    return q;
}
// finite inputs only (GEMM / attention epilogues), see quant_int_finite
__device__ __forceinline__ float2 quant_int2_finite(float2 x, const QP2& p) {
    float2 q = rint2_plus(quot2(x, p), p.zp);
    q.x = fminf(fmaxf(q.x, p.lo), p.hi);
    q.y = fminf(fmaxf(q.y, p.lo), p.hi);
    return q;
}
__device__ __forceinline__ float2 centre2(float2 xi, const QP2& p) { return __fadd2_rn(xi, p.nzp); }       // x_int - zp
// x_int - zp directly: clamp(rint(q) + zp, lo, hi) - zp == clamp(rint(q), lo - zp, hi - zp) (all
// operands are integers below 2^24, so both forms are exact) -- two FP32 ops fewer per element
__device__ __forceinline__ float2 quant_ctr2_finite(float2 x, const QP2& p) {
    const float2 M = make_float2(12582912.0f, 12582912.0f), nM = make_float2(-12582912.0f, -12582912.0f);
    float2 q = __fadd2_rn(__fadd2_rn(quot2(x, p), M), nM);
    q.x = fminf(fmaxf(q.x, p.lo + p.nzp.x), p.hi + p.nzp.x);
    q.y = fminf(fmaxf(q.y, p.lo + p.nzp.y), p.hi + p.nzp.y);
    return q;
}
__device__ __forceinline__ float2 dequant2(float2 xi, const QP2& p) { return __fmul2_rn(p.scale, centre2(xi, p)); }
template <bool FAST>
__device__ __forceinline__ float2 qdq2_t(float2 x, const QP2& p) { return dequant2(quant_int2_t<FAST>(x, p), p); }

// scale * (x_int - zp)  -- quantizers.py:209
__device__ __forceinline__ float dequant(float xi, const QP& p) {
    return __fmul_rn(p.scale, __fsub_rn(xi, p.zp));
}
template <bool FAST>
__device__ __forceinline__ float qdq_t(float x, const QP& p) { return dequant(quant_int_t<FAST>(x, p), p); }

// run-time dispatch (cold paths only: scalar tails, generic fallbacks)
__device__ __forceinline__ float div_rn(float x, const QP& p) { return p.exact ? div_rn_t<false>(x, p) : div_rn_t<true>(x, p); }
__device__ __forceinline__ float quant_int(float x, const QP& p) {
    return p.exact ? quant_int_t<false>(x, p) : quant_int_t<true>(x, p);
}
__device__ __forceinline__ float qdq(float x, const QP& p) { return dequant(quant_int(x, p), p); }

// ---- 128-bit global access --------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// ---- order-preserving float <-> uint (for atomicMin/Max on floats) ---------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}
#define TQ_ORD_MIN_IDENTITY 0xffffffffu   // > every key
#define TQ_ORD_MAX_IDENTITY 0x00000000u   // < every key

// ---- warp / block reductions -------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LayerNorm statistics of a row of a fake-quantized tensor x_i = s * k_i from the EXACT integer sums S1 = sum k_i,
// S2 = sum k_i^2 (k = centred integers, |k| <= 2^8): mean = s * S1 / n, var = s^2 * (S2 - S1^2 / n) / n evaluated
// in fp64 and rounded once.  Independent of the summation order and of how a row is tiled over threads / CTAs,
// so the fused GEMM epilogues (any tile width, bf16 or int8 operands) and the stand-alone LayerNorm kernel give
// bit-identical results; closer to the true statistics than any fp32 summation.
__device__ __forceinline__ void ln_stats_from_sums(long long S1, long long S2, int64_t n, float s, float eps, float& mean,
                                                   float& rstd) {
    const double m = (double)S1 / (double)n;
    double var = ((double)S2 - (double)S1 * m) / (double)n;
    var = var < 0.0 ? 0.0 : var;
    mean = (float)((double)s * m);
    const float varf = (float)((double)s * (double)s * var);
    rstd = __fdiv_rn(1.0f, sqrtf(__fadd_rn(varf, eps)));
}

inline int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = TQ_SM_COUNT_FALLBACK;
    }
    return cached;
}

inline bool pdl_enabled() {          // programmatic dependent launch for the tcgen05 / LN kernels (TQ_PDL=0: off).
    static int v = -1;               // Measured on the BERT-base chain: ~1 us less per launch, +2.5 % tokens/s.
    if (v < 0) {
        const char* e = getenv("TQ_PDL");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int check_qspec(const tq_qspec& q) {
    if (q.delta == nullptr) return TQ_EINVAL;
    if (q.n_bits < 1 || q.n_bits > 16) return TQ_EINVAL;
    if (q.zero_float == nullptr && q.is_signed == nullptr) return TQ_EINVAL;
    return TQ_OK;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// The fused engine is a chain of ~90 short kernels per forward (7 per encoder layer).  Each kernel
// releases its dependents immediately (pdl_trigger) and blocks before its first dependent global
// access (pdl_wait) until the previous kernel has completed and flushed: the next kernel's launch
// latency, barrier init, TMEM allocation and descriptor prefetch overlap the current kernel's tail.
// Both instructions are no-ops when the kernel was launched without the PDL attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                      Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x > 1) {                       // thread-block cluster (CTA pair of a 2-CTA tcgen05 kernel)
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    return e == cudaSuccess ? TQ_OK : (int)e;
}

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TQ_OK : (int)e;
}

}  // namespace tq
