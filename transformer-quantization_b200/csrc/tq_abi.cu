// Library-info entry points of the C ABI (include/tq_b200.h).
#include "tq_common.cuh"

namespace tq {

__device__ __forceinline__ uint64_t splitmix(uint64_t& st) {
    uint64_t z = (st += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// reference formulation with the division instruction and rintf (what tq::quant_int must equal)
__device__ __forceinline__ float quant_int_ref(float x, float s, float zp, float lo, float hi) {
    float q = __fadd_rn(rintf(__fdiv_rn(x, s)), zp);
    q = q < lo ? lo : q;
    q = q > hi ? hi : q;
    return q;
}

// Brute-force check of the division-free quotient (tq::div_rn) and of tq::quant_int against the
// IEEE division instruction on adversarial inputs: x near (k + 1/2) * s ties, random x, random s.
__global__ void selftest_div_kernel(uint64_t seed, int iters, unsigned long long* out) {
    uint64_t st = seed ^ ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0xD1B54A32D192ED03ULL);
    unsigned long long bad_div = 0, bad_q = 0, bad_tiny = 0, bad_div2 = 0, bad_q2 = 0;
    for (int it = 0; it < iters; ++it) {
        const uint64_t r0 = splitmix(st), r1 = splitmix(st);
        // scale: random significand, exponent so that s in ~[1e-8, 1e4]
        const uint32_t sexp = 100u + (uint32_t)(r0 % 41u);
        const float s = __uint_as_float((sexp << 23) | (uint32_t)((r0 >> 16) & 0x7fffffu));
        const int n_bits = 2 + (int)((r0 >> 40) % 15u);
        const float hi = (float)(1u << n_bits) - 1.0f;
        const float zp = (float)((r0 >> 48) % (uint32_t)(hi + 1.0f));
        const QP p = make_qp(s, zp, 0.0f, hi);
        float x;
        const uint32_t mode = (uint32_t)(r1 & 3u);
        if (mode == 0) {                          // fully random finite float
            uint32_t b = (uint32_t)(r1 >> 32);
            if (((b >> 23) & 0xffu) == 0xffu) b &= ~(1u << 30);
            x = __uint_as_float(b);
        } else {                                  // near a rounding tie of the integer grid
            const int k = (int)((r1 >> 8) % 140001u) - 70000;
            const float frac = mode == 1 ? 0.5f : (mode == 2 ? 0.0f : 0.25f + 0.5f * (float)((r1 >> 40) & 1u));
            x = __fmul_rn((float)k + frac, s);
            const int ulps = (int)((r1 >> 44) % 9u) - 4;
            x = __uint_as_float(__float_as_uint(x) + (uint32_t)ulps);
        }
        const float a = div_rn(x, p), b = __fdiv_rn(x, s);
        if (fabsf(b) < 4194304.0f && __float_as_uint(a) != __float_as_uint(b) && !(a == 0.0f && b == 0.0f)) {
            // below 2^-60 the FMA residuals can underflow; such quotients round to integer 0 anyway
            if (fabsf(b) >= 8.67361737988e-19f) ++bad_div;
            else ++bad_tiny;
        }
        const float qa = quant_int(x, p), qb = quant_int_ref(x, s, zp, 0.0f, hi);
        if (!(qa == qb) && !(qa != qa && qb != qb)) ++bad_q;
        // the PACKED path the fused epilogues use (quot2 / quant_int2_finite / quant_ctr2_finite on FFMA2): both components,
        // the second one a neighbour of x, finite inputs with |x / s| < 2^22 (what the epilogues guarantee)
        const float x2 = __uint_as_float(__float_as_uint(x) ^ (uint32_t)((r1 >> 20) & 7u));
        const float b2 = __fdiv_rn(x2, s);
        if (fabsf(b) < 4194304.0f && fabsf(b2) < 4194304.0f && fabsf(b) >= 8.67361737988e-19f && fabsf(b2) >= 8.67361737988e-19f) {
            const QP2 p2 = pair_of(p);
            const float2 xx = make_float2(x, x2);
            const float2 qq = quot2(xx, p2);
            if (__float_as_uint(qq.x) != __float_as_uint(b) || __float_as_uint(qq.y) != __float_as_uint(b2)) ++bad_div2;
            const float2 ki = quant_int2_finite(xx, p2), kc = quant_ctr2_finite(xx, p2);
            const float rb2 = quant_int_ref(x2, s, zp, 0.0f, hi);
            if (ki.x != qb || ki.y != rb2 || kc.x != qb - zp || kc.y != rb2 - zp) ++bad_q2;
        }
    }
    if (bad_div) atomicAdd(out, bad_div);
    if (bad_q) atomicAdd(out + 1, bad_q);
    if (bad_tiny) atomicAdd(out + 2, bad_tiny);
    if (bad_div2) atomicAdd(out + 3, bad_div2);
    if (bad_q2) atomicAdd(out + 4, bad_q2);
}

// ---- copy-bandwidth probe with this library's access patterns ---------------------------------------------
// What an 8 B / element streaming kernel of this library can reach at most on the device at hand: the same
// 128-bit grid-stride loop as the quant-dequant kernels without the arithmetic.  flags select the cache policy
// of the loads / stores and persistent (grid-stride, <= 4 CTAs per SM) vs one 16 KB chunk per CTA.
template <bool LD_NA, bool ST_NA>
__global__ void __launch_bounds__(256, 4)
probe_copy_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t nvec) {
    const int64_t stride = (int64_t)gridDim.x * 256 * 4;
    for (int64_t base = (int64_t)blockIdx.x * 256 * 4 + threadIdx.x; base < nvec; base += stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t idx = base + (int64_t)u * 256;
            if (idx < nvec) v[u] = LD_NA ? ld_stream(x + idx) : x[idx];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t idx = base + (int64_t)u * 256;
            if (idx < nvec) {
                if (ST_NA) st_stream(y + idx, v[u]);
                else y[idx] = v[u];
            }
        }
    }
}

}  // namespace tq

extern "C" {

int tq_probe_copy_f32(const float* x, float* y, int64_t n, int32_t flags, void* stream) {
    if (x == nullptr || y == nullptr || n < 0 || (n & 3) != 0 || !tq::aligned16(x) || !tq::aligned16(y)) return TQ_EINVAL;
    const int64_t nvec = n >> 2;
    if (nvec == 0) return TQ_OK;
    const int64_t chunks = (nvec + 1023) / 1024;
    const int64_t cap = (int64_t)tq::sm_count() * 4;
    const int grid = (int)((flags & 4) && chunks > cap ? cap : chunks);
    const float4* xv = reinterpret_cast<const float4*>(x);
    float4* yv = reinterpret_cast<float4*>(y);
    cudaStream_t st = (cudaStream_t)stream;
    switch (flags & 3) {
        case 0: tq::probe_copy_kernel<false, false><<<grid, 256, 0, st>>>(xv, yv, nvec); break;
        case 1: tq::probe_copy_kernel<true, false><<<grid, 256, 0, st>>>(xv, yv, nvec); break;
        case 2: tq::probe_copy_kernel<false, true><<<grid, 256, 0, st>>>(xv, yv, nvec); break;
        default: tq::probe_copy_kernel<true, true><<<grid, 256, 0, st>>>(xv, yv, nvec); break;
    }
    return tq::launch_status();
}

int tq_selftest_div(uint64_t seed, int32_t blocks, int32_t iters, uint64_t* mismatches, void* stream) {
    if (mismatches == nullptr || blocks < 1 || iters < 1) return TQ_EINVAL;
    tq::selftest_div_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(seed, iters,
                                                                     (unsigned long long*)mismatches);
    return tq::launch_status();
}

int tq_version(void) { return 4; }

const char* tq_error_string(int code) {
    switch (code) {
        case TQ_OK: return "ok";
        case TQ_EINVAL: return "tq: invalid argument";
        case TQ_EALIGN: return "tq: pointer alignment requirement not met";
        case TQ_EWORKSPACE: return "tq: workspace too small";
        case TQ_EUNSUPPORTED: return "tq: unsupported configuration";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "tq: unknown error";
}

int tq_device_sm_count(void) { return tq::sm_count(); }

}  // extern "C"
