// Shared device helpers for libtq_b200 (sm_100a).  Parity-critical arithmetic uses explicit
// round-to-nearest intrinsics (__fdiv_rn, __fmul_rn, __fadd_rn, __fsub_rn) so that nvcc can never
// contract it into FMAs or replace the division by a reciprocal multiply: the reference computes
// round(x / scale) with IEEE division (quantizers.py:184) and the integer result must be bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tq_b200.h"

#define TQ_SM_COUNT_FALLBACK 148

namespace tq {

// ---- resolved quantizer parameters ----------------------------------------------------------
struct QP {
    float scale, zp, lo, hi;
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

// scale / zero_point of parameter slot i (quantizers.py:142-153, 330-332)
__device__ __forceinline__ QP resolve(const tq_qspec& q, int64_t i, float lo, float hi) {
    QP p;
    const float d = __ldg(q.delta + i);
    p.scale = q.log_domain ? expf(d) : (d < q.eps ? q.eps : d);   // torch.clamp(min=eps) keeps NaN
    if (q.zero_float != nullptr) {
        float z = rintf(__ldg(q.zero_float + i));
        z = z < lo ? lo : z;
        z = z > hi ? hi : z;
        p.zp = z;
    } else {
        p.zp = 0.0f;
    }
    p.lo = lo;
    p.hi = hi;
    return p;
}

// clamp(rint(x / scale) + zp, lo, hi)  -- quantizers.py:184-185.  NaN propagates (torch.clamp).
__device__ __forceinline__ float quant_int(float x, const QP& p) {
    float q = __fadd_rn(rintf(__fdiv_rn(x, p.scale)), p.zp);
    q = q < p.lo ? p.lo : q;
    q = q > p.hi ? p.hi : q;
    return q;
}
// scale * (x_int - zp)  -- quantizers.py:209
__device__ __forceinline__ float dequant(float xi, const QP& p) {
    return __fmul_rn(p.scale, __fsub_rn(xi, p.zp));
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

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int check_qspec(const tq_qspec& q) {
    if (q.delta == nullptr) return TQ_EINVAL;
    if (q.n_bits < 1 || q.n_bits > 16) return TQ_EINVAL;
    if (q.zero_float == nullptr && q.is_signed == nullptr) return TQ_EINVAL;
    return TQ_OK;
}

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TQ_OK : (int)e;
}

}  // namespace tq
