// Hijacked nn.Linear as ONE Blackwell kernel:
//     y = act_quant( act_fn( x @ Wq.T + bias ) )       (reference hijacker.py:66-116,
//                                                        autoquant_utils.py:16-21)
// The reference runs an fp32 cuBLAS SGEMM on dequantized tensors plus separate bias / activation /
// six QDQ kernels.  Here the GEMM consumes the INTEGER grids of the fake-quantized operands carried
// in bf16 (exact for |v| <= 256) on the 5th-gen tensor cores:
//
//   warp 0      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) -> 4-6 stage smem ring
//   warp 1      MMA issuer     tcgen05.mma.cta_group::1.kind::f16, M=128 x N=BN x K=16, fp32
//                              accumulators in TMEM, double buffered (2 x BN columns)
//   warps 2-9   epilogue       tcgen05.ld 32x32b.x32 -> acc * (s_a * s_w[n]) + bias[n] -> act_fn
//                              -> per-tensor or per-column (PEG / fused-QKV) QDQ -> fp32 and/or
//                              bf16 centred-integer output (operand format of the next GEMM)
//
// Persistent: grid = min(#tiles, #SMs); tile order keeps the A row-panel hot in L2.  Three
// pipelines (smem full/empty, TMEM full/empty, tile loop) synchronised with mbarriers only.
#include "tq_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>

namespace tq {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 B: one SWIZZLE_128B span
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 12;    // three per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31)
constexpr int kThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA, 12 x epilogue
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kTmemCols = 512;

template <int BN>
struct Cfg {
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BN > 192) ? 3 : (BN > 128 ? 4 : (BN > 96 ? 5 : 6));
    static constexpr int kParamBytes = 12 * BN * 4;   // colscale | bias | {scale,-scale,1/scale,zp,-zp} x {out_q, out2_q}
    static constexpr int kStoreBytes = kEpiWarps * 2048;            // per-epilogue-warp 32x16 fp32 transpose tile
    static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + kParamBytes + kBarBytes + 1024;
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
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
        if (clock64() - t0 > 4000000000LL) {   // ~2 s: a pipeline bug must not hang the GPU
            printf("tq_linear: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
                   threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1,
                                            uint32_t bar) {
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
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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

// UMMA shared-memory descriptor: K-major operand, SWIZZLE_128B, 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address       bits [0,14)
    d |= (uint64_t)1 << 16;                               // leading byte offset (unused for SW128 K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset  bits [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}
// instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, dense
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case 1: return (v * 0.5f) * (1.0f + erff(v * 0.70710678118654752440f));   // nn.GELU (erf form)
        case 2: return v > 0.0f ? v : (v != v ? v : 0.0f);                         // nn.ReLU
        case 3: return tanhf(v);                                                   // nn.Tanh
        default: return v;
    }
}

// erf with one code path (no range split -> no divergence between the elements of a warp):
// Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 plus fp32 evaluation noise (~3e-7 total, the same
// order as the spread between libm / Sleef / CUDA erff).  GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) then
// carries an absolute error <= 1e-6 for |x| <= 6, far below the 8-bit output step that follows.
__device__ __forceinline__ float erf_as(float x) {
    const float ax = fabsf(x);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float y = 1.0f - p * t * __expf(-ax * ax);
    return copysignf(y, x);
}

template <int ACT>
__device__ __forceinline__ float act_fn(float v) {
    if (ACT == 1) return (v * 0.5f) * (1.0f + erf_as(v * 0.70710678118654752440f));   // nn.GELU (erf form)
    if (ACT == 2) return v > 0.0f ? v : (v != v ? v : 0.0f);                         // nn.ReLU
    if (ACT == 3) return tanhf(v);                                                   // nn.Tanh
    return v;
}

// per-column epilogue parameters in shared memory (one float per column each)
struct ColParams {
    const float *cs, *cb;                         // s_a * s_w[n], bias[n]
    const float *qs, *qns, *qr, *qz, *qnz;        // output quantizer: scale, -scale, 1/scale, zp, -zp
};
__device__ __forceinline__ QP2 qp2_at(const float* s, const float* ns, const float* r, const float* z, const float* nz,
                                      int j, float lo, float hi) {
    QP2 p;
    p.scale = *reinterpret_cast<const float2*>(s + j);
    p.nscale = *reinterpret_cast<const float2*>(ns + j);
    p.rcp = *reinterpret_cast<const float2*>(r + j);
    p.zp = *reinterpret_cast<const float2*>(z + j);
    p.nzp = *reinterpret_cast<const float2*>(nz + j);
    p.lo = lo;
    p.hi = hi;
    return p;
}

// 16 accumulators of one row -> scale, bias, activation, optional output quantizer, two columns per
// instruction (FMUL2 / FADD2 / FFMA2).  One compact, branch-free body per (activation, quantized?)
// pair so the executed path is contiguous in the instruction cache.  On return o[] holds the fp32
// outputs, v[] the centred integers (if HASQ).
template <int ACT, bool HASQ>
__device__ __forceinline__ void epi_math16(uint32_t (&v)[16], float (&o)[16], const ColParams& c, float qlo, float qhi) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        // acc * (s_a * s_w) + bias as ONE fused multiply-add per column (single rounding; ptxas
        // contracts packed mul + add into FFMA2 anyway, so the fusion is made explicit and is part of
        // the kernel's contract -- tests/test_gpu_linear.py checks against the fused formula)
        float2 f = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                              *reinterpret_cast<const float2*>(c.cs + j), *reinterpret_cast<const float2*>(c.cb + j));
        f.x = act_fn<ACT>(f.x);
        f.y = act_fn<ACT>(f.y);
        if (HASQ) {
            const QP2 p = qp2_at(c.qs, c.qns, c.qr, c.qz, c.qnz, j, qlo, qhi);
            const float2 ctr = quant_ctr2_finite(f, p);                      // centred integers x_int - zp
            v[j] = __float_as_uint(ctr.x);
            v[j + 1] = __float_as_uint(ctr.y);
            f = __fmul2_rn(p.scale, ctr);                                    // scale * (x_int - zp)
        }
        o[j] = f.x;
        o[j + 1] = f.y;
    }
}

struct EpiArgs {
    const float* bias;      // [N] or null
    float* y;               // [M, N] fp32 or null
    __nv_bfloat16* y_ctr;   // [M, N] bf16 centred integer grid or null
    tq_qspec a_q;           // input activation quantizer (delta == null: scale 1)
    tq_qspec w_q;           // weight quantizer
    int64_t w_q_params;     // 1 or N
    tq_qspec out_q;         // output quantizer (delta == null: no output quantization)
    int64_t out_q_params;   // 1 or N
    int act_fn;
    float* tile_minmax;     // optional calibration side reduction (ordered-int encoded, 2 words)
    long long* trace;       // optional: clock64 timeline of CTA 0 (16 slots), for tools/trace_linear.py
    // residual branch (attention-output / FFN-output blocks of the encoder, reference
    // models/quantized_bert.py:238-245, 264-277):  y = Q2( dequant(Q1(linear)) + res_scale * res_ctr )
    const __nv_bfloat16* res_ctr;   // [M, N] centred integer grid of the residual input, or null
    tq_qspec res_q;                 // its (per-tensor) quantizer
    tq_qspec out2_q;                // quantizer of the residual sum
    int64_t out2_params;            // 1 or N
};

#define TQ_TRACE(slot) do { if (ep.trace != nullptr && blockIdx.x == 0) ep.trace[slot] = clock64(); } while (0)

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
linear_qdq_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  int64_t M, int64_t N, int64_t K, int k_split, EpiArgs ep) {
    using C = Cfg<BN>;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;       // SWIZZLE_128B: 1024 B alignment
    unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    unsigned char* store_stage = base_ptr + C::kStages * C::kStageBytes;
    float* params = reinterpret_cast<float*>(store_stage + C::kStoreBytes);
    const uint32_t bar0 = base + C::kStages * C::kStageBytes + C::kStoreBytes + C::kParamBytes;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::kStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + 2 + s); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * C::kStages + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
        base_ptr + C::kStages * C::kStageBytes + C::kStoreBytes + C::kParamBytes + 8 * (2 * C::kStages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m_tiles = (M + BM - 1) / BM, n_tiles = (N + BN - 1) / BN;
    const int64_t tiles = m_tiles * n_tiles;
    const int kb_per_pass = (int)(K / BK);
    const int num_kb = kb_per_pass * k_split;

    if (threadIdx.x == 0) TQ_TRACE(0);
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < C::kStages; ++s) {
                mbar_init(full_bar(s), 1);
                mbar_init(empty_bar(s), 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(tfull_bar(s), 1);
                mbar_init(tempty_bar(s), kEpiWarps);      // one arrival per epilogue warp
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_trigger();
    pdl_wait();                       // A / residual tiles are produced by the previous kernel
    if (threadIdx.x == 0) TQ_TRACE(1);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int32_t m0 = (int32_t)((t / n_tiles) * BM), n0 = (int32_t)((t % n_tiles) * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (kb == 0) TQ_TRACE(2);
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_expect_tx(full_bar(stage), C::kStageBytes);
                    const uint32_t sa = base + stage * C::kStageBytes;
                    tma_load_2d(sa, &map_a, kb * BK, m0, full_bar(stage));
                    tma_load_2d(sa + C::kABytes, &map_w, (kb % kb_per_pass) * BK, n0, full_bar(stage));
                    if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                }
                TQ_TRACE(3);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    if (kb == 0) TQ_TRACE(4);
                    if (kb == 1) TQ_TRACE(5);
                    tc_fence_after();
                    const uint32_t sa = base + stage * C::kStageBytes;
                    const uint64_t adesc = make_desc_sw128(sa), bdesc = make_desc_sw128(sa + C::kABytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 B (16 bf16) inside the 128 B swizzle span: +2 in 16-byte units
                        tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                    (uint32_t)((kb | k) != 0));
                    }
                    tc_commit(empty_bar(stage));               // smem slot free once these MMAs retire
                    if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
                }
                tc_commit(tfull_bar(acc));                     // accumulator complete -> epilogue
                TQ_TRACE(6);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int et = threadIdx.x - 64;                       // 0..383
        const int quarter = warp & 3;                          // TMEM lane quarter this warp may access
        const int third = (warp - 2) >> 2;                     // 0..2: which of the quarter's three warps
        float* colscale = params;
        float* cbias = params + BN;
        float* qscale = params + 2 * BN;
        float* qzp = params + 3 * BN;
        float* qrcp = params + 4 * BN;
        float* qnscale = params + 5 * BN;
        float* qnzp = params + 6 * BN;
        float* q2scale = params + 7 * BN;
        float* q2zp = params + 8 * BN;
        float* q2rcp = params + 9 * BN;
        float* q2nscale = params + 10 * BN;
        float* q2nzp = params + 11 * BN;
        const bool has_q = ep.out_q.delta != nullptr;
        const bool has_res = ep.res_ctr != nullptr;
        float q2lo = 0.0f, q2hi = 0.0f, res_scale = 1.0f;
        if (has_res) {
            grid_of(ep.out2_q, q2lo, q2hi);
            float lo, hi;
            grid_of(ep.res_q, lo, hi);
            res_scale = resolve(ep.res_q, 0, lo, hi).scale;
        }
        float qlo = 0.0f, qhi = 0.0f;
        if (has_q) grid_of(ep.out_q, qlo, qhi);
        float a_scale = 1.0f;
        if (ep.a_q.delta != nullptr) {
            float lo, hi;
            grid_of(ep.a_q, lo, hi);
            a_scale = resolve(ep.a_q, 0, lo, hi).scale;
        }
        float wlo = 0.0f, whi = 0.0f;
        if (ep.w_q.delta != nullptr) grid_of(ep.w_q, wlo, whi);
        int acc = 0;
        uint32_t acc_phase = 0;
        float run_min = __int_as_float(0x7f800000), run_max = __int_as_float(0xff800000);
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int64_t m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
            asm volatile("bar.sync 1, 384;" ::: "memory");      // previous tile's parameter reads done
            int need_exact = 0;
            for (int j = et; j < BN; j += kEpiThreads) {
                const int64_t n = n0 + j;
                float cs = 0.0f, b = 0.0f, qs = 1.0f, qz = 0.0f, qr = 1.0f;
                if (n < N) {
                    const float ws = ep.w_q.delta != nullptr
                                         ? resolve(ep.w_q, ep.w_q_params > 1 ? n : 0, wlo, whi).scale
                                         : 1.0f;
                    cs = a_scale * ws;
                    b = ep.bias != nullptr ? ep.bias[n] : 0.0f;
                    if (has_q) {
                        const QP p = resolve(ep.out_q, ep.out_q_params > 1 ? n : 0, qlo, qhi);
                        qs = p.scale;
                        qz = p.zp;
                        qr = p.rcp;
                        need_exact |= p.exact;
                    }
                }
                colscale[j] = cs;
                cbias[j] = b;
                qscale[j] = qs;
                qzp[j] = qz;
                qrcp[j] = qr;
                qnscale[j] = -qs;
                qnzp[j] = -qz;
                if (has_res) {
                    float s2 = 1.0f, z2 = 0.0f, r2 = 1.0f;
                    if (n < N) {
                        const QP p2 = resolve(ep.out2_q, ep.out2_params > 1 ? n : 0, q2lo, q2hi);
                        s2 = p2.scale;
                        z2 = p2.zp;
                        r2 = p2.rcp;
                        need_exact |= p2.exact;
                    }
                    q2scale[j] = s2;
                    q2zp[j] = z2;
                    q2rcp[j] = r2;
                    q2nscale[j] = -s2;
                    q2nzp[j] = -z2;
                }
            }
            // barrier + OR-reduction over the 384 epilogue threads (named barrier 1)
            int exact;
            asm volatile(
                "{\n\t.reg .pred p, q;\n\t"
                "setp.ne.b32 p, %1, 0;\n\t"
                "bar.red.or.pred q, 1, 384, p;\n\t"
                "selp.u32 %0, 1, 0, q;\n\t}"
                : "=r"(exact)
                : "r"(need_exact)
                : "memory");
            if (et == 0) TQ_TRACE(7);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            if (et == 0) TQ_TRACE(8);
            const int64_t row = m0 + quarter * 32 + lane;
            const bool row_ok = row < M;
            const int64_t grow0 = m0 + quarter * 32;
            float4* stg = reinterpret_cast<float4*>(store_stage + (warp - 2) * 2048);
            // The three warps of a lane quarter take the tile's 16-column slices round robin: the loop
            // body (16 elements) stays small enough for the instruction cache -- a fully unrolled
            // 32-wide body (~60 KB of SASS) made instruction fetch the top stall.
#pragma unroll 1
            for (int c0 = third * 16; c0 < BN; c0 += 48) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + c0), v);
                float o[16];
                if (has_q && exact) {
                    // rare (a column scale outside div_rn's proven domain): IEEE divide, one element
                    // at a time through a compact loop
#pragma unroll 1
                    for (int j = 0; j < 16; ++j) {
                        float f = 0.0f;
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            if (k == j) f = __uint_as_float(v[k]);
                        f = apply_act(__fmaf_rn(f, colscale[c0 + j], cbias[c0 + j]), ep.act_fn);
                        const QP p{qscale[c0 + j], qzp[c0 + j], qlo, qhi, qrcp[c0 + j], 1};
                        const float ctr = __fsub_rn(quant_int_t<false>(f, p), p.zp);
                        f = __fmul_rn(p.scale, ctr);
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            if (k == j) {
                                v[k] = __float_as_uint(ctr);
                                o[k] = f;
                            }
                    }
                } else {
                    const ColParams cp{colscale + c0, cbias + c0, qscale + c0, qnscale + c0, qrcp + c0, qzp + c0,
                                       qnzp + c0};
                    switch (ep.act_fn * 2 + (has_q ? 1 : 0)) {      // warp-uniform
                        case 0: epi_math16<0, false>(v, o, cp, qlo, qhi); break;
                        case 1: epi_math16<0, true>(v, o, cp, qlo, qhi); break;
                        case 2: epi_math16<1, false>(v, o, cp, qlo, qhi); break;
                        case 3: epi_math16<1, true>(v, o, cp, qlo, qhi); break;
                        case 4: epi_math16<2, false>(v, o, cp, qlo, qhi); break;
                        case 5: epi_math16<2, true>(v, o, cp, qlo, qhi); break;
                        case 6: epi_math16<3, false>(v, o, cp, qlo, qhi); break;
                        default: epi_math16<3, true>(v, o, cp, qlo, qhi); break;
                    }
                }
                if (ep.tile_minmax != nullptr && row_ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (n0 + c0 + j < N) {
                            run_min = fminf(run_min, o[j]);
                            run_max = fmaxf(run_max, o[j]);
                        }
                    }
                }
                if (has_res) {
                    // residual rows: coalesced 32 B pieces -> swizzled smem -> one row per lane
                    uint4* s2 = reinterpret_cast<uint4*>(stg);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int r = i * 16 + (lane >> 1), ch = lane & 1;
                        const int64_t grow = grow0 + r, gcol = n0 + c0 + ch * 8;
                        uint4 val = make_uint4(0u, 0u, 0u, 0u);
                        if (grow < M && gcol < N) val = *reinterpret_cast<const uint4*>(ep.res_ctr + grow * N + gcol);
                        s2[r * 2 + (ch ^ ((r >> 2) & 1))] = val;
                    }
                    __syncwarp();
                    uint32_t rw[8];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint4 t4 = s2[lane * 2 + (c ^ ((lane >> 2) & 1))];
                        rw[4 * c] = t4.x; rw[4 * c + 1] = t4.y; rw[4 * c + 2] = t4.z; rw[4 * c + 3] = t4.w;
                    }
                    __syncwarp();
                    if (exact) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const uint32_t pair = rw[j >> 1];
                            const float rc = __uint_as_float((j & 1) ? (pair & 0xffff0000u) : (pair << 16));
                            const float sum = __fadd_rn(o[j], __fmul_rn(res_scale, rc));
                            const QP p2{q2scale[c0 + j], q2zp[c0 + j], q2lo, q2hi, q2rcp[c0 + j], 1};
                            const float ctr = __fsub_rn(quant_int_t<false>(sum, p2), p2.zp);
                            v[j] = __float_as_uint(ctr);
                            o[j] = __fmul_rn(p2.scale, ctr);
                        }
                    } else {
                        const float2 rs2 = make_float2(res_scale, res_scale);
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const uint32_t pair = rw[j >> 1];
                            const float2 rc = make_float2(__uint_as_float(pair << 16), __uint_as_float(pair & 0xffff0000u));
                            // residual value = fl(scale * ctr) as the reference materialises it: scalar
                            // multiplies (a packed mul feeding a packed add would be contracted to FFMA2)
                            const float2 sum = __fadd2_rn(make_float2(o[j], o[j + 1]),
                                                          make_float2(__fmul_rn(rs2.x, rc.x), __fmul_rn(rs2.y, rc.y)));
                            const QP2 p2 = qp2_at(q2scale, q2nscale, q2rcp, q2zp, q2nzp, c0 + j, q2lo, q2hi);
                            const float2 ctr = quant_ctr2_finite(sum, p2);
                            v[j] = __float_as_uint(ctr.x);
                            v[j + 1] = __float_as_uint(ctr.y);
                            const float2 dq = __fmul2_rn(p2.scale, ctr);
                            o[j] = dq.x;
                            o[j + 1] = dq.y;
                        }
                    }
                }
                // ---- coalesced stores: 32 x 16 transpose through this warp's private smem tile ----
                // registers hold one ROW per lane; a direct store would touch 32 different lines per
                // instruction.  Swizzled 16-byte chunks keep both smem phases bank-conflict free.
                const int64_t gcol0 = n0 + c0;
                if (ep.y != nullptr) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        stg[lane * 4 + (c ^ ((lane >> 1) & 3))] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                    __syncwarp();
                    float4 vals[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                 // all shared-memory reads first ...
                        const int r = i * 8 + (lane >> 2), ch = lane & 3;
                        vals[i] = stg[r * 4 + (ch ^ ((r >> 1) & 3))];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                 // ... then the (predicated) global stores
                        const int r = i * 8 + (lane >> 2), ch = lane & 3;
                        const int64_t grow = grow0 + r, gcol = gcol0 + ch * 4;
                        if (grow < M && gcol < N) *reinterpret_cast<float4*>(ep.y + grow * N + gcol) = vals[i];
                    }
                    __syncwarp();
                }
                if (ep.y_ctr != nullptr) {
                    uint4* s2 = reinterpret_cast<uint4*>(stg);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint4 w;
                        __nv_bfloat162 h;
                        h = __floats2bfloat162_rn(__uint_as_float(v[8 * c]), __uint_as_float(v[8 * c + 1]));
                        w.x = *reinterpret_cast<uint32_t*>(&h);
                        h = __floats2bfloat162_rn(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3]));
                        w.y = *reinterpret_cast<uint32_t*>(&h);
                        h = __floats2bfloat162_rn(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5]));
                        w.z = *reinterpret_cast<uint32_t*>(&h);
                        h = __floats2bfloat162_rn(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7]));
                        w.w = *reinterpret_cast<uint32_t*>(&h);
                        s2[lane * 2 + (c ^ ((lane >> 2) & 1))] = w;
                    }
                    __syncwarp();
                    uint4 cv[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int r = i * 16 + (lane >> 1), ch = lane & 1;
                        cv[i] = s2[r * 2 + (ch ^ ((r >> 2) & 1))];
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int r = i * 16 + (lane >> 1), ch = lane & 1;
                        const int64_t grow = grow0 + r, gcol = gcol0 + ch * 8;
                        if (grow < M && gcol < N) *reinterpret_cast<uint4*>(ep.y_ctr + grow * N + gcol) = cv[i];
                    }
                    __syncwarp();
                }
            }
            if (et == 0) TQ_TRACE(9);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (ep.tile_minmax != nullptr) {
            run_min = warp_min(run_min);
            run_max = warp_max(run_max);
            if (lane == 0) {
                uint32_t* w = reinterpret_cast<uint32_t*>(ep.tile_minmax);
                atomicMax(w, ~f2ord(run_min));
                atomicMax(w + 1, f2ord(run_max));
            }
        }
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TQ_TRACE(10);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)kTmemCols)
                     : "memory");
    }
}

// hi | mid | lo bf16 split: x = hi + mid + lo up to 2^-24 relative (three 8-bit mantissa pieces)
__global__ void __launch_bounds__(256, 4)
split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t M, int64_t K) {
    const int64_t n = M * K;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = i / K, c = i - r * K;
        const float v = x[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const float r1 = __fsub_rn(v, __bfloat162float(h));
        const __nv_bfloat16 m = __float2bfloat16_rn(r1);
        const float r2 = __fsub_rn(r1, __bfloat162float(m));
        const __nv_bfloat16 l = __float2bfloat16_rn(r2);
        __nv_bfloat16* o = out + r * 3 * K;
        o[c] = h;
        o[K + c] = m;
        o[2 * K + c] = l;
    }
}

// ---- host side ------------------------------------------------------------------------------------
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

// row-major [rows, cols] bf16 -> 2-D tensor map with a [box_rows, 64] box, 128 B swizzle
static int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return TQ_EUNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? TQ_OK : TQ_EINVAL;
}

template <int BN>
static int launch(const void* a, const void* w, int64_t M, int64_t N, int64_t K, int k_split, const EpiArgs& ep,
                  cudaStream_t st) {
    using C = Cfg<BN>;
    static_assert(C::kSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert(2 * BN <= kTmemCols, "TMEM budget");
    CUtensorMap map_a, map_w;
    if (int e = make_map(&map_a, a, M, K * k_split, BM)) return e;
    if (int e = make_map(&map_w, w, N, K, BN)) return e;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linear_qdq_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    return launch_pdl(linear_qdq_kernel<BN>, dim3(grid), dim3(kThreads), C::kSmemBytes, st, map_a, map_w, M, N, K,
                      k_split, ep);
}

// Tile width.  Cycle model from the clock64 timeline of the kernel (tools/trace_linear.py): the main
// loop advances one k-block per ~600 cycles whatever the tile width, the epilogue is bound by FP32 issue (~75 cycles per
// 16-column slice per warp: ~2.1 k cycles, ~3.4 k with GELU/tanh -- FMA-pipe and store bound); with double-buffered TMEM a CTA that owns t
// tiles takes  setup + main + (t-1) * max(main, epi) + epi.
static int pick_bn(int64_t M, int64_t N, int64_t K, int k_split, int act_fn) {
    const int cands[5] = {256, 192, 128, 96, 64};
    const int64_t m_tiles = (M + BM - 1) / BM;
    const int sms = sm_count();
    int best = 64;
    double best_cost = 1e30;
    for (int i = 0; i < 5; ++i) {
        const int bn = cands[i];
        if (bn > 64 && N < bn) continue;                 // TMA box must fit inside the weight matrix
        const int64_t tiles = m_tiles * ((N + bn - 1) / bn);
        const double per_cta = (double)((tiles + sms - 1) / sms);
        // per 64-wide k-block: ~600 cycles of TMA->MMA hand-off latency (measured, independent of the
        // tile width up to 192) or the MMA time itself (128 x bn x 64 MACs at 4096 MAC/clk)
        const double mma_c = bn * 2.05;
        const double main_c = (double)(K / BK) * k_split * (mma_c > 600.0 ? mma_c : 600.0);
        const double epi_c = (bn / 16.0 / 3.0) * (act_fn == 1 || act_fn == 3 ? 3400.0 : 2100.0);
        const double cost = 800.0 + main_c + (per_cta - 1.0) * (main_c > epi_c ? main_c : epi_c) + epi_c;
        if (cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

}  // namespace gemm
}  // namespace tq

extern "C" {

size_t tq_linear_workspace_bytes(int64_t, int64_t, int64_t) { return 256; }

static int linear_impl(const void* a_ctr_bf16, const void* w_ctr_bf16, const float* bias, float* y,
                       void* y_ctr_bf16, int64_t M, int64_t N, int64_t K, int32_t k_split, tq_qspec a_q,
                       tq_qspec w_q, int64_t w_q_params, int32_t act_fn, tq_qspec out_q, int64_t out_q_params,
                       const void* res_ctr_bf16, tq_qspec res_q, tq_qspec out2_q, int64_t out2_q_params,
                       float* tile_minmax, void* ws, size_t ws_bytes, void* stream) {
    using namespace tq::gemm;
    if (res_ctr_bf16 != nullptr) {
        if (out_q.delta == nullptr || res_q.delta == nullptr) return TQ_EINVAL;
        if (int e = tq::check_qspec(out2_q)) return e;
        if (int e = tq::check_qspec(res_q)) return e;
        if (out2_q_params != 1 && out2_q_params != N) return TQ_EINVAL;
        if (!tq::aligned16(res_ctr_bf16)) return TQ_EALIGN;
    }
    if (a_ctr_bf16 == nullptr || w_ctr_bf16 == nullptr || (y == nullptr && y_ctr_bf16 == nullptr)) return TQ_EINVAL;
    if (M < 1 || N < 1 || K < 1 || (k_split != 1 && k_split != 3)) return TQ_EINVAL;
    if (K % BK != 0 || N % 8 != 0) return TQ_EUNSUPPORTED;
    if (!tq::aligned16(a_ctr_bf16) || !tq::aligned16(w_ctr_bf16)) return TQ_EALIGN;
    if ((y != nullptr && !tq::aligned16(y)) || (y_ctr_bf16 != nullptr && !tq::aligned16(y_ctr_bf16))) return TQ_EALIGN;
    if (act_fn < 0 || act_fn > 3) return TQ_EINVAL;
    if (out_q.delta != nullptr) {
        if (int e = tq::check_qspec(out_q)) return e;
        if (out_q_params != 1 && out_q_params != N) return TQ_EINVAL;
    }
    if (w_q.delta != nullptr && w_q_params != 1 && w_q_params != N) return TQ_EINVAL;
    EpiArgs ep;
    ep.bias = bias;
    ep.y = y;
    ep.y_ctr = reinterpret_cast<__nv_bfloat16*>(y_ctr_bf16);
    ep.a_q = a_q;
    ep.w_q = w_q;
    ep.w_q_params = w_q_params;
    ep.out_q = out_q;
    ep.out_q_params = out_q_params;
    ep.act_fn = act_fn;
    ep.tile_minmax = tile_minmax;
    ep.trace = (ws != nullptr && ws_bytes >= 16 * sizeof(long long)) ? reinterpret_cast<long long*>(ws) : nullptr;
    ep.res_ctr = reinterpret_cast<const __nv_bfloat16*>(res_ctr_bf16);
    ep.res_q = res_q;
    ep.out2_q = out2_q;
    ep.out2_params = out2_q_params;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pick_bn(M, N, K, k_split, act_fn)) {
        case 256: return launch<256>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
        case 192: return launch<192>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
        case 128: return launch<128>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
        case 96: return launch<96>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
        default: return launch<64>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
    }
}

int tq_linear_qdq_bf16(const void* a_ctr_bf16, const void* w_ctr_bf16, const float* bias, float* y,
                       void* y_ctr_bf16, int64_t M, int64_t N, int64_t K, int32_t k_split, tq_qspec a_q,
                       tq_qspec w_q, int64_t w_q_params, int32_t act_fn, tq_qspec out_q, int64_t out_q_params,
                       float* tile_minmax, void* ws, size_t ws_bytes, void* stream) {
    tq_qspec none;
    none.delta = nullptr; none.zero_float = nullptr; none.is_signed = nullptr;
    none.n_bits = 8; none.log_domain = 0; none.eps = 1e-8f;
    return linear_impl(a_ctr_bf16, w_ctr_bf16, bias, y, y_ctr_bf16, M, N, K, k_split, a_q, w_q, w_q_params, act_fn,
                       out_q, out_q_params, nullptr, none, none, 1, tile_minmax, ws, ws_bytes, stream);
}

int tq_linear_res_qdq_bf16(const void* a_ctr_bf16, const void* w_ctr_bf16, const float* bias, float* y,
                           void* y_ctr_bf16, int64_t M, int64_t N, int64_t K, tq_qspec a_q, tq_qspec w_q,
                           int64_t w_q_params, tq_qspec out_q, int64_t out_q_params, const void* res_ctr_bf16,
                           tq_qspec res_q, tq_qspec out2_q, int64_t out2_q_params, void* stream) {
    if (res_ctr_bf16 == nullptr) return TQ_EINVAL;
    return linear_impl(a_ctr_bf16, w_ctr_bf16, bias, y, y_ctr_bf16, M, N, K, 1, a_q, w_q, w_q_params, 0, out_q,
                       out_q_params, res_ctr_bf16, res_q, out2_q, out2_q_params, nullptr, nullptr, 0, stream);
}

int tq_split3_bf16(const float* x, void* out_bf16, int64_t M, int64_t K, void* stream) {
    if (x == nullptr || out_bf16 == nullptr || M < 1 || K < 1) return TQ_EINVAL;
    int64_t blocks = (M * K + 255) / 256;
    const int64_t cap = (int64_t)tq::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    tq::gemm::split3_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, M, K);
    return tq::launch_status();
}

}  // extern "C"
