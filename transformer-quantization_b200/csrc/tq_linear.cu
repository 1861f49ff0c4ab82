// Hijacked nn.Linear as ONE Blackwell kernel:
//     y = act_quant( act_fn( x @ Wq.T + bias ) )       (reference hijacker.py:66-116,
//                                                        autoquant_utils.py:16-21)
// The reference runs an fp32 cuBLAS SGEMM on dequantized tensors plus separate bias / activation /
// six QDQ kernels.  Here the GEMM consumes the INTEGER grids of the fake-quantized operands on the
// 5th-gen tensor cores -- centred grids in bf16 (exact for |v| <= 256, kind::f16) or the raw 8-bit
// grids x_int (kind::i8, int32 accumulators, zero point removed with the weight row sums):
//
//   warp 12     TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) -> 4-8 stage smem ring
//   warp 13     MMA issuer     tcgen05.mma.cta_group::1, M=128 x N=BN x K=16 (bf16) / K=32 (int8),
//                              fp32 / int32 accumulators in TMEM, double buffered (2 x BN columns)
//   warps 0-11  epilogue       tcgen05.ld 32x32b.x16 -> acc * (s_a * s_w[n]) + bias[n] -> act_fn
//                              -> per-tensor or per-column (PEG / fused-QKV) QDQ [-> + residual -> QDQ]
//                              -> fp32 and/or bf16 centred-integer output (operand format of the next
//                              GEMM), one 32-byte row piece per lane (STG.256), no smem staging
//
// Persistent: grid = min(#tiles, #SMs); tile order keeps the A row-panel hot in L2.  Three
// pipelines (smem full/empty, TMEM full/empty, tile loop) synchronised with mbarriers only.
#include "tq_common.cuh"
#include "tq_attn.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstring>
#include <vector>

namespace tq {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 B: one SWIZZLE_128B span
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 12;    // three per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31)
constexpr int kThreads = 64 + 32 * kEpiWarps;   // 12 x epilogue, then the TMA producer and the MMA issuer
// The two single-lane roles sit in the HIGHEST warp ids: the warp scheduler favours higher ids, and an
// issue-bound epilogue (GELU) sharing a scheduler with a low-id producer / MMA warp starved the main
// loop of the next tile (measured: time per tile = main + epilogue instead of max(main, epilogue)).
constexpr int kProdWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kTmemCols = 512;

// CTAS = 2: a CTA PAIR (2-CTA cluster on one TPC) computes a 256 x BN tile with tcgen05 cta_group::2 --
// each CTA stages its own 128 rows of A and HALF of the weight tile, so the L2 -> SM traffic per
// flop (what bounds the main loop: ~43 B/clk/SM when every SM streams) drops by a third at BN = 256.
template <int BN, int CTAS>
struct Cfg {
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBRows = BN / CTAS;          // weight rows staged by this CTA
    static constexpr int kBBytes = kBRows * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kXchgFloats = 6 * BM + 2 * 8 * BM;       // LayerNorm row statistics (fused LN epilogue)
    static constexpr int kParamBytes = 16 * BN * 4 + kXchgFloats * 4;   // eight float4 arrays per column pair + exchange
    // as many ring stages as fit beside the parameters (227 KB per CTA): the main loop is bound by
    // the TMA -> MMA hand-off latency, so depth is what buys throughput
    static constexpr int kStagesFit = (227 * 1024 - 1024 - kParamBytes - 256) / kStageBytes;
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
    static constexpr int kSmemBytes = kStages * kStageBytes + kParamBytes + kBarBytes + 1024;
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
template <int CTAS>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1,
                                            uint32_t bar) {
    if (CTAS == 1) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
            : "memory");
    } else {
        // data lands in THIS CTA's shared memory, the transaction bytes are counted on the LEADER's
        // barrier (bit 24 of a shared::cluster address = CTA rank inside the pair)
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar & 0xFEFFFFFFu)
            : "memory");
    }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// store one float at the same shared-memory offset of CTA `rank` of the cluster (distributed smem)
__device__ __forceinline__ void st_remote_f32(uint32_t local_addr, uint32_t rank, float v) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.f32 [ra], %2;\n\t}"
        ::"r"(local_addr), "r"(rank), "f"(v)
        : "memory");
}
// arrive on the same barrier of CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(rank)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int CTAS>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    if (CTAS == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    } else {                                          // the same barrier in BOTH CTAs of the pair
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(bar), "h"((uint16_t)3)
            : "memory");
    }
}
template <int CTAS>
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    if (CTAS == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
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

// 8-bit integer operands (unsigned or two's complement), int32 accumulators: K = 32 per instruction
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
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

// GELU (erf form) for two columns with ONE code path: erf by Abramowitz & Stegun 7.1.26
// (|error| <= 1.5e-7 plus fp32 evaluation noise, ~3e-7 total -- the same order as the spread between
// libm / Sleef / CUDA erff), so GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) carries an absolute error
// <= 1e-6 for |x| <= 6, far below the 8-bit output step that follows.  No range split -> no
// divergence; the polynomial runs on packed FFMA2, only 1/d and exp2 use the XU pipe.
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
__device__ __forceinline__ float rcp_approx(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float ex2_approx(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float2 gelu2(float2 x) {
    const float2 z = __fmul2_rn(x, splat(0.70710678118654752440f));
    const float2 az = make_float2(fabsf(z.x), fabsf(z.y));
    const float2 d = __ffma2_rn(splat(0.3275911f), az, splat(1.0f));
    const float2 t = make_float2(rcp_approx(d.x), rcp_approx(d.y));      // (a Newton reciprocal on the FMA pipe measured slower:
    // MUFU runs beside the FMA pipe, six more FFMA2 per pair do not -- tools/microbench/fp32_pipes.cu)
    float2 p = __ffma2_rn(splat(-1.061405429f), t, splat(1.453152027f));     // -(a5 t + a4) ...
    p = __ffma2_rn(p, t, splat(-1.421413741f));
    p = __ffma2_rn(p, t, splat(0.284496736f));
    p = __ffma2_rn(p, t, splat(-0.254829592f));                               // = -(polynomial)
    const float2 npt = __fmul2_rn(p, t);
    const float2 u = __fmul2_rn(__fmul2_rn(az, az), splat(-1.4426950408889634f));   // -z^2 log2 e
    const float2 e = make_float2(ex2_approx(u.x), ex2_approx(u.y));
    const float2 y = __ffma2_rn(npt, e, splat(1.0f));                         // erf(|z|)
    const float2 erf = make_float2(copysignf(y.x, z.x), copysignf(y.y, z.y));
    const float2 h = __fmul2_rn(x, splat(0.5f));
    return __ffma2_rn(h, erf, h);
}
template <int ACT>
__device__ __forceinline__ float2 act2(float2 f) {
    if (ACT == 1) return gelu2(f);
    if (ACT == 2) return make_float2(fmaxf(f.x, 0.0f), fmaxf(f.y, 0.0f));      // nn.ReLU (finite inputs)
    return f;
}

// ---- 256-bit global access (LDG / STG.E.ENL2.256 on sm_100): 32 bytes = one full sector per lane.
// The epilogue keeps one output ROW per lane (the TMEM layout); with 32-byte accesses a row-per-lane
// load or store moves whole sectors, so no shared-memory transpose is needed for coalescing.
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// 32 bytes of one row: `wide` (32-byte aligned row pieces) -> one 256-bit access, else two 128-bit
// ones; `full` = false: only the first 16 bytes exist (ragged last column block, N % 16 == 8)
__device__ __forceinline__ void ld_row32(const void* p, bool wide, bool full, uint32_t (&r)[8]) {
    if (wide && full) {
        ldg256(p, r);
    } else {
        const uint4 a = *reinterpret_cast<const uint4*>(p);
        uint4 b = make_uint4(0u, 0u, 0u, 0u);
        if (full) b = *(reinterpret_cast<const uint4*>(p) + 1);
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
        r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
    }
}
__device__ __forceinline__ void st_row32(void* p, bool wide, bool full, const uint32_t (&r)[8]) {
    if (wide && full) {
        stg256(p, r);
    } else {
        *reinterpret_cast<uint4*>(p) = make_uint4(r[0], r[1], r[2], r[3]);
        if (full) *(reinterpret_cast<uint4*>(p) + 1) = make_uint4(r[4], r[5], r[6], r[7]);
    }
}

// ---- 8-bit integer operand mode (I8) ------------------------------------------------------------------
// A and W are the quantizers' integer grids x_int / w_int in one byte per element (kind::i8 MMA,
// int32 accumulators: always exact, twice the K per shared-memory byte and per MMA cycle).  The
// zero point of A is taken out per output column with the exact integer identity
//     sum_k (a_int - zp) * w[n,k]  =  acc[n] - zp * sum_k w[n,k]        (w row sums precomputed once).
template <bool I8>
__device__ __forceinline__ float2 acc_pair(uint32_t a, uint32_t b, const float4& corr) {
    if (!I8) return make_float2(__uint_as_float(a), __uint_as_float(b));
    return make_float2(__int2float_rn((int)a - __float_as_int(corr.x)), __int2float_rn((int)b - __float_as_int(corr.y)));
}
// residual row piece of 16 columns: bf16 centred grid (32 bytes) or 8-bit x_int (16 bytes)
template <bool I8>
__device__ __forceinline__ void ld_res16(const void* base, int64_t elem, bool wide, bool full, uint32_t (&r)[8]) {
    if (!I8) {
        ld_row32(reinterpret_cast<const __nv_bfloat16*>(base) + elem, wide, full, r);
    } else {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(base) + elem;
        if (full) {
            const uint4 a = *reinterpret_cast<const uint4*>(p);
            r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
        } else {
            const uint2 a = *reinterpret_cast<const uint2*>(p);
            r[0] = a.x; r[1] = a.y; r[2] = 0u; r[3] = 0u;
        }
    }
}
// columns 2 jp, 2 jp + 1 of that piece as centred values (x_int - zp)
template <bool I8>
__device__ __forceinline__ float2 res_pair(const uint32_t (&rw)[8], int jp, float res_zp) {
    if (!I8) return make_float2(__uint_as_float(rw[jp] << 16), __uint_as_float(rw[jp] & 0xffff0000u));
    // byte -> float without the conversion unit: 0x4B000000 | b is the float 2^23 + b; (2^23 + b) - (2^23 + zp)
    // is exact.  One PRMT + one FADD per element (I2F / F2I run on the quarter-rate XU pipe).
    const uint32_t w = rw[jp >> 1];
    const float off = __fadd_rn(8388608.0f, res_zp);
    const uint32_t lo = __byte_perm(w, 0x4B000000u, (jp & 1) ? 0x7652 : 0x7650);
    const uint32_t hi = __byte_perm(w, 0x4B000000u, (jp & 1) ? 0x7653 : 0x7651);
    return make_float2(__fsub_rn(__uint_as_float(lo), off), __fsub_rn(__uint_as_float(hi), off));
}
// Slice order of an epilogue warp.  bf16 outputs: 16-column slices third*16 + 48*it (a lane stores 32 bytes
// per slice).  8-bit outputs: a 16-column slice is only 16 bytes, half a sector, so a warp takes PAIRS of
// adjacent slices (third*32 + 96*(it/2) + 16*(it%2)), keeps the four packed words of the even one and
// stores the 32 bytes of both with the odd one.
template <bool I8>
__device__ __forceinline__ int slice_col(int third, int it) {
    return I8 ? third * 32 + 96 * (it >> 1) + 16 * (it & 1) : third * 16 + 48 * it;
}
// 16 centred integers -> x_int = k - (lo - zp) + lo, one byte each (two's complement low byte): four words
__device__ __forceinline__ void pack_u8x16(uint32_t (&w)[4], const float2 (&k)[8], float2 clo0, const float4* Pclo,
                                           bool percol, float lo) {
    uint32_t b[16];
    const float2 off0 = make_float2(__fsub_rn(__fadd_rn(lo, 12582912.0f), clo0.x), __fsub_rn(__fadd_rn(lo, 12582912.0f), clo0.y));
#pragma unroll
    for (int jp = 0; jp < 8; ++jp) {
        float2 off = off0;
        if (percol) off = make_float2(__fsub_rn(__fadd_rn(lo, 12582912.0f), Pclo[jp].z), __fsub_rn(__fadd_rn(lo, 12582912.0f), Pclo[jp].w));
        const float2 t = __fadd2_rn(k[jp], off);       // v + 1.5 * 2^23 keeps v (two's complement) in the low mantissa bits
        b[2 * jp] = __float_as_uint(t.x);
        b[2 * jp + 1] = __float_as_uint(t.y);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        w[i] = __byte_perm(__byte_perm(b[4 * i], b[4 * i + 1], 0x0040), __byte_perm(b[4 * i + 2], b[4 * i + 3], 0x0040), 0x5410);
}
// store the slice pair: `it` even -> hold (store alone when there is no odd partner), odd -> 32 bytes
__device__ __forceinline__ void st_u8_pair(unsigned char* row_ptr, int64_t gcol, int64_t N, int it, uint32_t (&held)[4],
                                           const uint32_t (&w)[4], bool wide32) {
    if ((it & 1) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) held[i] = w[i];
        if (gcol + 16 >= N) *reinterpret_cast<uint4*>(row_ptr + gcol) = make_uint4(w[0], w[1], w[2], w[3]);
    } else if (wide32) {
        const uint32_t r[8] = {held[0], held[1], held[2], held[3], w[0], w[1], w[2], w[3]};
        stg256(row_ptr + gcol - 16, r);
    } else {
        *reinterpret_cast<uint4*>(row_ptr + gcol - 16) = make_uint4(held[0], held[1], held[2], held[3]);
        *reinterpret_cast<uint4*>(row_ptr + gcol) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
__device__ __forceinline__ void st_u8x16(void* p, const float2 (&k)[8], float2 clo0, const float4* Pclo, bool percol,
                                         float lo, bool full) {
    // float integer -> byte without F2I: v + 1.5 * 2^23 keeps v (two's complement) in the low mantissa bits
    uint32_t b[16];
    const float2 off0 = make_float2(__fsub_rn(__fadd_rn(lo, 12582912.0f), clo0.x), __fsub_rn(__fadd_rn(lo, 12582912.0f), clo0.y));
#pragma unroll
    for (int jp = 0; jp < 8; ++jp) {
        float2 off = off0;
        if (percol) off = make_float2(__fsub_rn(__fadd_rn(lo, 12582912.0f), Pclo[jp].z), __fsub_rn(__fadd_rn(lo, 12582912.0f), Pclo[jp].w));
        const float2 t = __fadd2_rn(k[jp], off);
        b[2 * jp] = __float_as_uint(t.x);
        b[2 * jp + 1] = __float_as_uint(t.y);
    }
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        w[i] = __byte_perm(__byte_perm(b[4 * i], b[4 * i + 1], 0x0040), __byte_perm(b[4 * i + 2], b[4 * i + 3], 0x0040), 0x5410);
    if (full) *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    else *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
}

// ---- epilogue parameters -------------------------------------------------------------------------
// Shared memory: six float4 arrays indexed by COLUMN PAIR jp (= column / 2), two columns per entry:
//   P[0][jp] = {cs.x, cs.y, cb.x, cb.y}      acc scale s_a * s_w[n], bias[n]
//   P[1][jp] = {s.x,  s.y,  -s.x, -s.y}      output quantizer scale
//   P[2][jp] = {r.x,  r.y,  clo.x, clo.y}    RN(1/s), lower clamp bound lo - zp (centred domain)
//   P[3][jp] = {chi.x, chi.y, chi2.x, chi2.y}  upper clamp bounds hi - zp of both quantizers
//   P[4][jp] = {s2, -s2}, P[5][jp] = {r2, clo2}   quantizer of the residual sum
//   P[6][jp] = {gamma.x, gamma.y, beta.x, beta.y}  LayerNorm weight (fake-quantized) and bias (fused LN)
//   P[7][jp] = {corr.x, corr.y, -, -}  int32 bits: zero point of A times the weight row sums (I8 mode)
// A per-tensor quantizer (the common case) is read once per tile into registers (QReg).
__device__ __forceinline__ int pidx(int bn, int arr, int slot, int j) {       // float index
    return ((arr * (bn >> 1) + (j >> 1)) << 2) + (slot << 1) + (j & 1);
}
struct QReg {
    float2 s, ns, r, clo, chi;
};

// centred integers x_int - zp = clamp(rint(RN(x / s)), lo - zp, hi - zp) for two columns (finite x:
// see quant_int_finite; every operand of the clamp is an integer below 2^24, so shifting the clamp
// by zp is exact)
__device__ __forceinline__ float2 ctr2(float2 x, const QReg& q) {
    const float2 q0 = __fmul2_rn(x, q.r);
    const float2 q1 = __ffma2_rn(__ffma2_rn(q0, q.ns, x), q.r, q0);
    const float2 q2 = __ffma2_rn(__ffma2_rn(q1, q.ns, x), q.r, q1);
    float2 k = __fadd2_rn(__fadd2_rn(q2, splat(12582912.0f)), splat(-12582912.0f));
    k.x = fminf(fmaxf(k.x, q.clo.x), q.chi.x);
    k.y = fminf(fmaxf(k.y, q.clo.y), q.chi.y);
    return k;
}
__device__ __forceinline__ uint32_t pack_bf16(float2 v) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    return *reinterpret_cast<const uint32_t*>(&h);
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
    long long* trace_all;   // optional: {globaltimer start, end, smid, clock64 span} per CTA (tools/trace_ctas.py)
    long long* trace_tiles; // optional: per-tile clock64 stamps of CTA 0, [8 tiles][8 slots] (tools/trace_tiles.py)
    // residual branch (attention-output / FFN-output blocks of the encoder, reference
    // models/quantized_bert.py:238-245, 264-277):  y = Q2( dequant(Q1(linear)) + res_scale * res_ctr )
    const __nv_bfloat16* res_ctr;   // [M, N] centred integer grid of the residual input, or null
    tq_qspec res_q;                 // its (per-tensor) quantizer
    tq_qspec out2_q;                // quantizer of the residual sum
    int64_t out2_params;            // 1 or N
    // fused LayerNorm of the residual sum (reference models/quantized_bert.py:245, 277 -> QuantLayerNorm,
    // autoquant_utils.py:55-66):  z = ln_q( LayerNorm(y; gamma_q, beta, eps) );  ln_gamma != null selects it
    // 8-bit operand mode
    void* y_u8;                     // [M, N] x_int of the output (unsigned or two's complement byte) or null
    const int32_t* w_rowsum;        // [N] sum_k w_int[n, k]
    const float* ln_gamma;          // [N] fake-quantized LayerNorm weight
    const float* ln_beta;           // [N]
    float ln_eps;
    tq_qspec ln_q;                  // per-tensor output quantizer
};

#define TQ_TRACE(slot) do { if (ep.trace != nullptr && blockIdx.x == 0) ep.trace[slot] = clock64(); } while (0)
// per-tile stamps of CTA 0: slot 0 params start, 1 params done, 2 accumulator ready, 3 epilogue done,
// 4 MMA got a free accumulator, 5 first operands landed, 6 MMAs issued, 7 producer issued its last load
#define TQ_TTRACE(tile_no, slot) do { if (ep.trace_tiles != nullptr && blockIdx.x == 0 && (tile_no) < 8) \
        ep.trace_tiles[(tile_no) * 8 + (slot)] = clock64(); } while (0)


// ---- epilogue of one accumulator tile ---------------------------------------------------------------
// Fast path: activation ACT in {none, GELU}, an output quantizer whose scales are inside div_rn's
// domain, optionally the residual add + second quantizer.  The three warps of a TMEM lane quarter take
// the tile's 16-column slices round robin (c0 = third * 16, + 48, ...); a lane owns one output row:
//   tcgen05.ld 16 columns -> FFMA2 (scale, bias) -> act -> centred integers (ctr2) [-> + residual ->
//   ctr2] -> bf16 pack -> ONE 32-byte store per lane (and/or two for the fp32 output).
// No shared-memory staging and no cross-lane traffic; the residual row piece of the NEXT slice is
// requested before the math of the current one.
template <int BN, int ACT, bool PERCOL, bool RES, bool I8>
__device__ __forceinline__ void epi_tile_fast(const EpiArgs& ep, const float* params, uint32_t tmem_tile, int third,
                                              int64_t row, bool row_ok, int64_t n0, int64_t N, float res_scale,
                                              float res_zp, float qlo, uint32_t tfull, uint32_t tphase, int tno) {
    constexpr int HP = BN / 2;
    const float4* P = reinterpret_cast<const float4*>(params);
    QReg q1, q2;
    if (!PERCOL) {
        const float4 a = P[HP], b = P[2 * HP], c = P[3 * HP];
        q1.s = make_float2(a.x, a.y); q1.ns = make_float2(a.z, a.w);
        q1.r = make_float2(b.x, b.y); q1.clo = make_float2(b.z, b.w);
        q1.chi = make_float2(c.x, c.y);
        if (RES) {
            const float4 d = P[4 * HP], e = P[5 * HP];
            q2.s = make_float2(d.x, d.y); q2.ns = make_float2(d.z, d.w);
            q2.r = make_float2(e.x, e.y); q2.clo = make_float2(e.z, e.w);
            q2.chi = make_float2(c.z, c.w);
        }
    }
    const bool wide_c = ((((uintptr_t)ep.y_ctr) | ((uintptr_t)ep.res_ctr)) & 31u) == 0 && (N & 15) == 0;
    const bool wide_y = (((uintptr_t)ep.y) & 31u) == 0;
    const int64_t res_base = row * N + n0;             // element offset of this lane's residual row piece
    uint32_t rnext[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rnext[i] = 0u;
    int c0 = slice_col<I8>(third, 0);
    if (RES && row_ok && n0 + c0 < N) ld_res16<I8>(ep.res_ctr, res_base + c0, wide_c, n0 + c0 + 16 <= N, rnext);
    mbar_wait(tfull, tphase);
    tc_fence_after();
    if (threadIdx.x == 0) TQ_TRACE(8);
    if (threadIdx.x == 0) TQ_TTRACE(tno, 2);
    uint32_t held[4] = {0u, 0u, 0u, 0u};
    const bool wide8 = I8 && ((((uintptr_t)ep.y_u8) & 31u) == 0) && (N & 31) == 0;
#pragma unroll 1
    for (int it = 0; (c0 = slice_col<I8>(third, it)) < BN; ++it) {
        const int64_t gcol = n0 + c0;
        if (gcol >= N) break;                                    // warp-uniform
        const bool full = gcol + 16 <= N;
        uint32_t v[16];
        tmem_ld16(tmem_tile + (uint32_t)c0, v);
        uint32_t rw[8];
        if (RES) {
#pragma unroll
            for (int i = 0; i < 8; ++i) rw[i] = rnext[i];
            const int cn = slice_col<I8>(third, it + 1);
            if (row_ok && cn < BN && n0 + cn < N) ld_res16<I8>(ep.res_ctr, res_base + cn, wide_c, n0 + cn + 16 <= N, rnext);
        }
        float2 k[8];
#pragma unroll
        for (int jp = 0; jp < 8; ++jp) {
            const int gp = (c0 >> 1) + jp;
            const float4 p0 = P[gp];
            float2 f = __ffma2_rn(acc_pair<I8>(v[2 * jp], v[2 * jp + 1], P[7 * HP + (I8 ? gp : 0)]),
                                  make_float2(p0.x, p0.y), make_float2(p0.z, p0.w));
            f = act2<ACT>(f);
            float4 pc;
            if (PERCOL) {
                const float4 a = P[HP + gp], b = P[2 * HP + gp];
                pc = P[3 * HP + gp];
                q1.s = make_float2(a.x, a.y); q1.ns = make_float2(a.z, a.w);
                q1.r = make_float2(b.x, b.y); q1.clo = make_float2(b.z, b.w);
                q1.chi = make_float2(pc.x, pc.y);
            }
            float2 c = ctr2(f, q1);
            if (RES) {
                if (PERCOL) {
                    const float4 d = P[4 * HP + gp], e = P[5 * HP + gp];
                    q2.s = make_float2(d.x, d.y); q2.ns = make_float2(d.z, d.w);
                    q2.r = make_float2(e.x, e.y); q2.clo = make_float2(e.z, e.w);
                    q2.chi = make_float2(pc.z, pc.w);
                }
                // dequantized linear output and residual value as the reference materialises them
                // (fl(scale * ctr) each, then one add): SCALAR multiplies -- a packed multiply feeding
                // a packed add would be contracted into FFMA2 by ptxas
                const float2 rc = res_pair<I8>(rw, jp, res_zp);
                const float2 sum = __fadd2_rn(
                    make_float2(__fmul_rn(q1.s.x, c.x), __fmul_rn(q1.s.y, c.y)),
                    make_float2(__fmul_rn(res_scale, rc.x), __fmul_rn(res_scale, rc.y)));
                c = ctr2(sum, q2);
            }
            k[jp] = c;
        }
        if (row_ok) {
            if (ep.y_u8 != nullptr) {
                uint32_t w8[4];
                pack_u8x16(w8, k, RES ? q2.clo : q1.clo, P + (RES ? 5 : 2) * HP + (c0 >> 1), PERCOL, qlo);
                unsigned char* o8 = reinterpret_cast<unsigned char*>(ep.y_u8) + row * N;
                if (I8) st_u8_pair(o8, gcol, N, it, held, w8, wide8);            // paired slices: 32-byte stores
                else if (full) *reinterpret_cast<uint4*>(o8 + gcol) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                else *reinterpret_cast<uint2*>(o8 + gcol) = make_uint2(w8[0], w8[1]);
            }
            if (ep.y_ctr != nullptr) {
                uint32_t w[8];
#pragma unroll
                for (int jp = 0; jp < 8; ++jp) w[jp] = pack_bf16(k[jp]);
                st_row32(ep.y_ctr + row * N + gcol, wide_c, full, w);
            }
            if (ep.y != nullptr) {
                uint32_t w[16];
#pragma unroll
                for (int jp = 0; jp < 8; ++jp) {
                    float2 s = RES ? q2.s : q1.s;
                    if (PERCOL) {
                        const float4 a = P[(RES ? 4 : 1) * HP + (c0 >> 1) + jp];
                        s = make_float2(a.x, a.y);
                    }
                    const float2 o = __fmul2_rn(s, k[jp]);                  // scale * (x_int - zp)
                    w[2 * jp] = __float_as_uint(o.x);
                    w[2 * jp + 1] = __float_as_uint(o.y);
                }
                float* yr = ep.y + row * N + gcol;
                st_row32(yr, wide_y, true, *reinterpret_cast<uint32_t(*)[8]>(&w[0]));
                if (full) st_row32(yr + 8, wide_y, true, *reinterpret_cast<uint32_t(*)[8]>(&w[8]));
            }
        }
    }
}

// centred integers for a pair, FAST (division-free, see ctr2) or with the IEEE division instruction
template <bool FAST>
__device__ __forceinline__ float2 ctr2_t(float2 x, const QReg& q) {
    if (FAST) return ctr2(x, q);
    float2 k;
    k.x = fminf(fmaxf(rint_even(__fdiv_rn(x.x, q.s.x)), q.clo.x), q.chi.x);
    k.y = fminf(fmaxf(rint_even(__fdiv_rn(x.y, q.s.y)), q.clo.y), q.chi.y);
    return k;
}

// 8-column TMEM accesses (32 lanes x 8 words): the fused LayerNorm epilogue parks its packed
// intermediate in the accumulator columns it has already consumed
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8_nowait(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Residual block with the LayerNorm fused in (one tile per CTA, the CTAs of a cluster cover the N
// columns of their 128 rows).  Three compact loops over the warp's 16-column slices (a fully unrolled
// body that kept the intermediate in registers was instruction-fetch bound: every instruction ran
// once):
//   loop 1   y = Q2( dequant(Q1(acc * cs + bias)) + res )  -> packed bf16 centred integers, parked in
//            the first 8 of the 16 accumulator columns just read (TMEM); lane sum of scale * y
//   loop 2   squared deviations from the LANE's own mean
//   exchange lane partials -> smem (3 warps per row) -> every CTA of the cluster (distributed shared
//            memory) -> one cluster barrier -> mean, variance by pairwise combination (Chan et al.):
//            as accurate as a global two-pass variance
//   loop 3   z = Q3( (v - mean) * rstd * gamma + beta ) -> bf16 grid / fp32
// All three quantizers are per-tensor here (the host routes anything else to the unfused kernels).
// (Both operand modes walk the slices in the paired order of slice_col<true>, so the fp32 row statistics
// are summed in the same order and the two modes give bit-identical results.)
template <int BN, bool FAST, bool I8>
__device__ __forceinline__ void epi_tile_res_ln(const EpiArgs& ep, float* params, uint32_t tmem_tile, int third, int quarter,
                                                int lane, int64_t row, bool row_ok, int64_t n0, int64_t N,
                                                float res_scale, float res_zp, uint32_t tfull, uint32_t tphase, int tno) {
    constexpr int HP = BN / 2;
    const float4* P = reinterpret_cast<const float4*>(params);
    float* part = params + 16 * BN;                    // [2][3][BM] lane partials (sum | squared deviations)
    float* xs = part + 6 * BM;                         // [8][BM][2] (sum, M2) of every CTA of the cluster
    QReg q1, q2, q3;
    float ln_lo = 0.0f;
    {
        const float4 a = P[HP], b = P[2 * HP], c = P[3 * HP], d = P[4 * HP], e = P[5 * HP];
        q1.s = make_float2(a.x, a.y); q1.ns = make_float2(a.z, a.w);
        q1.r = make_float2(b.x, b.y); q1.clo = make_float2(b.z, b.w);
        q1.chi = make_float2(c.x, c.y);
        q2.s = make_float2(d.x, d.y); q2.ns = make_float2(d.z, d.w);
        q2.r = make_float2(e.x, e.y); q2.clo = make_float2(e.z, e.w);
        q2.chi = make_float2(c.z, c.w);
        float lo, hi;
        grid_of(ep.ln_q, lo, hi);
        const QP p3 = resolve(ep.ln_q, 0, lo, hi);
        q3.s = splat(p3.scale); q3.ns = splat(-p3.scale); q3.r = splat(p3.rcp);
        q3.clo = splat(lo - p3.zp); q3.chi = splat(hi - p3.zp);
        ln_lo = lo;
    }
    const uint32_t cn = cluster_nctarank(), my = cluster_ctarank();
    const int rl = quarter * 32 + lane;                // row inside the tile
    const bool wide_c = ((((uintptr_t)ep.y_ctr) | ((uintptr_t)ep.res_ctr)) & 31u) == 0 && (N & 15) == 0;
    const bool wide_y = (((uintptr_t)ep.y) & 31u) == 0;
    const int64_t res_base = row * N + n0;
    uint32_t rnext[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rnext[i] = 0u;
    if (row_ok) ld_res16<I8>(ep.res_ctr, res_base + slice_col<true>(third, 0), wide_c, true, rnext);
    mbar_wait(tfull, tphase);
    tc_fence_after();
    if (threadIdx.x == 0) TQ_TRACE(8);
    if (threadIdx.x == 0) TQ_TTRACE(tno, 2);

    // ---- loop 1 ----
    float S1 = 0.0f, S2 = 0.0f;           // sum k, sum k^2 of this lane's centred integers: exact in fp32 (< 2^24)
#pragma unroll 1
    for (int it = 0, c0; (c0 = slice_col<true>(third, it)) < BN; ++it) {
        uint32_t v[16];
        tmem_ld16(tmem_tile + (uint32_t)c0, v);
        uint32_t rw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) rw[j] = rnext[j];
        const int cn = slice_col<true>(third, it + 1);
        if (row_ok && cn < BN) ld_res16<I8>(ep.res_ctr, res_base + cn, wide_c, true, rnext);
        uint32_t kp[8];
#pragma unroll
        for (int jp = 0; jp < 8; ++jp) {
            const float4 p0 = P[(c0 >> 1) + jp];
            const float2 f = __ffma2_rn(acc_pair<I8>(v[2 * jp], v[2 * jp + 1], P[7 * HP + (I8 ? (c0 >> 1) + jp : 0)]),
                                        make_float2(p0.x, p0.y), make_float2(p0.z, p0.w));
            const float2 c = ctr2_t<FAST>(f, q1);
            const float2 rc = res_pair<I8>(rw, jp, res_zp);
            const float2 sum = __fadd2_rn(
                make_float2(__fmul_rn(q1.s.x, c.x), __fmul_rn(q1.s.y, c.y)),
                make_float2(__fmul_rn(res_scale, rc.x), __fmul_rn(res_scale, rc.y)));
            const float2 k = ctr2_t<FAST>(sum, q2);
            kp[jp] = pack_bf16(k);
            S1 = __fadd_rn(S1, __fadd_rn(k.x, k.y));
            S2 = __fmaf_rn(k.x, k.x, S2);
            S2 = __fmaf_rn(k.y, k.y, S2);
        }
        tmem_st8_nowait(tmem_tile + (uint32_t)c0, kp);
    }
    tmem_st_wait();
    if (threadIdx.x == 0) TQ_TRACE(11);
    // ---- exchange: exact integer sums (see ln_stats_from_sums) ----
    int* parti = reinterpret_cast<int*>(part);
    int* xsi = reinterpret_cast<int*>(xs);
    parti[third * BM + rl] = __float2int_rn(S1);
    parti[(3 + third) * BM + rl] = __float2int_rn(S2);
    asm volatile("bar.sync 1, 384;" ::: "memory");
    if (third == 0) {
        int t1 = 0, t2 = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            t1 += parti[t * BM + rl];
            t2 += parti[(3 + t) * BM + rl];
        }
        const uint32_t dst = smem_u32(xsi + (my * BM + rl) * 2);
        for (uint32_t r = 0; r < cn; ++r) {
            st_remote_f32(dst, r, __int_as_float(t1));
            st_remote_f32(dst + 4, r, __int_as_float(t2));
        }
    }
    cluster_sync_all();
    if (threadIdx.x == 0) TQ_TRACE(12);
    long long T1 = 0, T2 = 0;
    for (uint32_t r = 0; r < cn; ++r) {
        T1 += xsi[(r * BM + rl) * 2];
        T2 += xsi[(r * BM + rl) * 2 + 1];
    }
    float mean, rstd;
    ln_stats_from_sums(T1, T2, N, q2.s.x, ep.ln_eps, mean, rstd);
    // ---- loop 3: normalise, affine, output quantizer ----
    const float2 nmean = splat(-mean), rstd2 = splat(rstd);
    uint32_t held[4] = {0u, 0u, 0u, 0u};
    const bool wide8 = I8 && ((((uintptr_t)ep.y_u8) & 31u) == 0) && (N & 31) == 0;
#pragma unroll 1
    for (int it = 0, c0; (c0 = slice_col<true>(third, it)) < BN; ++it) {
        uint32_t kp[8];
        tmem_ld8(tmem_tile + (uint32_t)c0, kp);
        float2 k3[8];
#pragma unroll
        for (int jp = 0; jp < 8; ++jp) {
            const float2 v = make_float2(__fmul_rn(q2.s.x, __uint_as_float(kp[jp] << 16)),
                                         __fmul_rn(q2.s.y, __uint_as_float(kp[jp] & 0xffff0000u)));
            const float4 gb = P[6 * HP + (c0 >> 1) + jp];
            float2 y = __fmul2_rn(__fadd2_rn(v, nmean), rstd2);
            y = __ffma2_rn(y, make_float2(gb.x, gb.y), make_float2(gb.z, gb.w));
            k3[jp] = ctr2_t<FAST>(y, q3);
        }
        if (row_ok) {
            const int64_t gcol = n0 + c0;
            if (I8 && ep.y_u8 != nullptr) {
                uint32_t w8[4];
                pack_u8x16(w8, k3, q3.clo, P, false, ln_lo);
                st_u8_pair(reinterpret_cast<unsigned char*>(ep.y_u8) + row * N, gcol, N, it, held, w8, wide8);
            }
            if (ep.y_ctr != nullptr) {
                uint32_t w[8];
#pragma unroll
                for (int jp = 0; jp < 8; ++jp) w[jp] = pack_bf16(k3[jp]);
                st_row32(ep.y_ctr + row * N + gcol, wide_c, true, w);
            }
            if (ep.y != nullptr) {
                uint32_t w[16];
#pragma unroll
                for (int jp = 0; jp < 8; ++jp) {
                    const float2 o = __fmul2_rn(q3.s, k3[jp]);
                    w[2 * jp] = __float_as_uint(o.x);
                    w[2 * jp + 1] = __float_as_uint(o.y);
                }
                float* yr = ep.y + row * N + gcol;
                st_row32(yr, wide_y, true, *reinterpret_cast<uint32_t(*)[8]>(&w[0]));
                st_row32(yr + 8, wide_y, true, *reinterpret_cast<uint32_t(*)[8]>(&w[8]));
            }
        }
    }
}

// Generic path (cold): any activation, no output quantizer, scales that need the IEEE division
// instruction, calibration min/max side reduction.  One element at a time through a compact loop;
// NaN propagates through the clamp like torch.clamp.
template <int BN, bool I8>
__device__ __forceinline__ void epi_tile_generic(const EpiArgs& ep, const float* params, uint32_t tmem_tile, int third,
                                                 int64_t row, bool row_ok, int64_t n0, int64_t N, float res_scale,
                                                 float res_zp, float qlo, bool has_q, bool has_res, uint32_t tfull,
                                                 uint32_t tphase, float& run_min, float& run_max) {
    const bool wide_c = ((((uintptr_t)ep.y_ctr) | ((uintptr_t)ep.res_ctr)) & 31u) == 0 && (N & 15) == 0;
    const bool wide_y = (((uintptr_t)ep.y) & 31u) == 0;
    mbar_wait(tfull, tphase);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = third * 16; c0 < BN; c0 += 48) {
        const int64_t gcol = n0 + c0;
        if (gcol >= N) break;
        const bool full = gcol + 16 <= N;
        uint32_t v[16];
        tmem_ld16(tmem_tile + (uint32_t)c0, v);
        uint32_t rw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rw[i] = 0u;
        if (has_res && row_ok) ld_res16<I8>(ep.res_ctr, row * N + gcol, wide_c, full, rw);
        uint32_t o[16];
#pragma unroll 1
        for (int j = 0; j < 16; ++j) {
            float acc = 0.0f;
            float rc = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i == j) {
                    const float2 a2 = acc_pair<I8>(v[i], v[i], reinterpret_cast<const float4*>(params)[7 * (BN / 2) + (I8 ? (c0 + i) >> 1 : 0)]);
                    acc = (i & 1) ? a2.y : a2.x;
                    const float2 r2 = res_pair<I8>(rw, i >> 1, res_zp);
                    rc = (i & 1) ? r2.y : r2.x;
                }
            const int col = c0 + j;
            float f = apply_act(__fmaf_rn(acc, params[pidx(BN, 0, 0, col)], params[pidx(BN, 0, 1, col)]), ep.act_fn);
            float c = f;
            if (has_q) {
                const float s = params[pidx(BN, 1, 0, col)];
                c = clamp_nan(rint_even(__fdiv_rn(f, s)), params[pidx(BN, 2, 1, col)], params[pidx(BN, 3, 0, col)]);
                f = __fmul_rn(s, c);
            }
            if (has_res) {
                const float sum = __fadd_rn(f, __fmul_rn(res_scale, rc));
                const float s2 = params[pidx(BN, 4, 0, col)];
                c = clamp_nan(rint_even(__fdiv_rn(sum, s2)), params[pidx(BN, 5, 1, col)], params[pidx(BN, 3, 1, col)]);
                f = __fmul_rn(s2, c);
            }
            if (ep.tile_minmax != nullptr && row_ok && gcol + j < N) {
                run_min = fminf(run_min, f);
                run_max = fmaxf(run_max, f);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i == j) {
                    v[i] = __float_as_uint(c);
                    o[i] = __float_as_uint(f);
                }
        }
        if (row_ok) {
            if (I8 && ep.y_u8 != nullptr && has_q) {
                float2 k[8];
#pragma unroll
                for (int jp = 0; jp < 8; ++jp) k[jp] = make_float2(__uint_as_float(v[2 * jp]), __uint_as_float(v[2 * jp + 1]));
                st_u8x16(reinterpret_cast<unsigned char*>(ep.y_u8) + row * N + gcol, k, make_float2(0.0f, 0.0f),
                         reinterpret_cast<const float4*>(params) + (has_res ? 5 : 2) * (BN / 2) + (c0 >> 1), true, qlo, full);
            }
            if (ep.y_ctr != nullptr) {
                uint32_t w[8];
#pragma unroll
                for (int jp = 0; jp < 8; ++jp)
                    w[jp] = pack_bf16(make_float2(__uint_as_float(v[2 * jp]), __uint_as_float(v[2 * jp + 1])));
                st_row32(ep.y_ctr + row * N + gcol, wide_c, full, w);
            }
            if (ep.y != nullptr) {
                float* yr = ep.y + row * N + gcol;
                st_row32(yr, wide_y, true, *reinterpret_cast<uint32_t(*)[8]>(&o[0]));
                if (full) st_row32(yr + 8, wide_y, true, *reinterpret_cast<uint32_t(*)[8]>(&o[8]));
            }
        }
    }
}

// LNF: residual block with the LayerNorm fused into the epilogue -- one tile per CTA, launched as
// clusters of N / BN CTAs that exchange the row statistics through distributed shared memory.
template <int BN, int CTAS, bool LNF, bool I8>
__global__ void __launch_bounds__(kThreads, 1)
linear_qdq_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  int64_t M, int64_t N, int64_t K, int k_split, int ring, EpiArgs ep) {
    using C = Cfg<BN, CTAS>;                          // ring: stages of the smem ring in use (<= C::kStages)
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;       // SWIZZLE_128B: 1024 B alignment
    unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    float* params = reinterpret_cast<float*>(base_ptr + C::kStages * C::kStageBytes);
    const uint32_t bar0 = base + C::kStages * C::kStageBytes + C::kParamBytes;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::kStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + 2 + s); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * C::kStages + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
        base_ptr + C::kStages * C::kStageBytes + C::kParamBytes + 8 * (2 * C::kStages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a work item is a (BM * CTAS) x BN tile; CTA `cta_rank` of the pair owns rows cta_rank*BM.. of it
    const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
    const int64_t m_tiles = (M + BM * CTAS - 1) / (BM * CTAS), n_tiles = (N + BN - 1) / BN;
    const int64_t tiles = m_tiles * n_tiles;
    const int64_t tile0 = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;
    constexpr int BKe = I8 ? 2 * BK : BK;             // elements per 128-byte k-block row
    const int kb_per_pass = (int)(K / BKe);
    const int num_kb = kb_per_pass * k_split;

    if (threadIdx.x == 0) TQ_TRACE(0);
    long long t_start_ns = 0, t_start_clk = 0;
    if (ep.trace_all != nullptr && threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start_ns));
        t_start_clk = clock64();
    }
    // producer state; a single CTA starts its first loads BEFORE the TMEM allocation / block barrier
    // below (the first TMA round trip, ~1.5 k cycles, then overlaps the rest of the prologue)
    int p_stage = 0, p_pre = 0;
    uint32_t p_phase = 0;
    if (warp == kProdWarp && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (CTAS == 1 && tile0 < tiles) {             // (a pair must wait for the peer's barriers)
            pdl_wait();                               // A is produced by the previous kernel
            const int32_t m0 = (int32_t)((tile0 / n_tiles) * BM), n0 = (int32_t)((tile0 % n_tiles) * BN);
            p_pre = num_kb < ring ? num_kb : ring;
            for (int kb = 0; kb < p_pre; ++kb) {      // ring slots are free: no empty-barrier wait
                mbar_expect_tx(full_bar(p_stage), C::kStageBytes);
                const uint32_t sa = base + p_stage * C::kStageBytes;
                tma_load_2d<1>(sa, &map_a, kb * BKe, m0, full_bar(p_stage));
                tma_load_2d<1>(sa + C::kABytes, &map_w, (kb % kb_per_pass) * BKe, n0, full_bar(p_stage));
                if (++p_stage == ring) { p_stage = 0; p_phase ^= 1u; }
            }
        }
    }
    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(tfull_bar(s), 1);
                mbar_init(tempty_bar(s), kEpiWarps * CTAS);   // one arrival per epilogue warp (of both CTAs)
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (CTAS == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"((uint32_t)kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {                                      // the same warp of both CTAs, same slot address
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"((uint32_t)kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();   // barriers of BOTH CTAs initialised
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_trigger();
    pdl_wait();                       // A / residual tiles are produced by the previous kernel
    if (threadIdx.x == 0) TQ_TRACE(1);

    if (warp == kProdWarp) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = p_stage;
            uint32_t phase = p_phase;
            int tno = 0;
            for (int64_t t = tile0; t < tiles; t += tile_step, ++tno) {
                const int32_t m0 = (int32_t)((t / n_tiles) * (BM * CTAS) + cta_rank * BM);
                const int32_t n0 = (int32_t)((t % n_tiles) * BN + cta_rank * C::kBRows);
                for (int kb = (t == tile0 ? p_pre : 0); kb < num_kb; ++kb) {
                    if (kb == 0) TQ_TRACE(2);
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    // the leader's barrier counts the bytes of both CTAs (the MMA issuer waits on it)
                    if (cta_rank == 0) mbar_expect_tx(full_bar(stage), C::kStageBytes * CTAS);
                    const uint32_t sa = base + stage * C::kStageBytes;
                    tma_load_2d<CTAS>(sa, &map_a, kb * BKe, m0, full_bar(stage));
                    tma_load_2d<CTAS>(sa + C::kABytes, &map_w, (kb % kb_per_pass) * BKe, n0, full_bar(stage));
                    if (++stage == ring) { stage = 0; phase ^= 1u; }
                }
                TQ_TRACE(3);
                TQ_TTRACE(tno, 7);
            }
            if (CTAS == 2) {
                // tail: every slot handed out has been consumed -- no multicast arrival may target
                // this CTA's barriers after it has left
                for (int s = 0; s < ring; ++s) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    if (++stage == ring) { stage = 0; phase ^= 1u; }
                }
            }
        }
        if (LNF) {                                    // the epilogue's cluster barrier counts every thread
            __syncwarp();
            cluster_sync_all();
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        if (lane == 0 && cta_rank == 0) {             // the pair's leader issues for both CTAs
            uint32_t idesc = make_idesc(BM * CTAS, BN);
            if (I8) {
                // kind::i8: int32 accumulators; operand signedness from the quantizers (asymmetric grids
                // are unsigned, symmetric ones carry a device-side `signed` flag)
                const uint32_t a_s8 = (ep.a_q.zero_float == nullptr && ep.a_q.is_signed != nullptr && *ep.a_q.is_signed) ? 1u : 0u;
                const uint32_t w_s8 = (ep.w_q.zero_float == nullptr && ep.w_q.is_signed != nullptr && *ep.w_q.is_signed) ? 1u : 0u;
                idesc = (2u << 4) | (a_s8 << 7) | (w_s8 << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CTAS) >> 4) << 24);
            }
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int tno = 0;
            for (int64_t t = tile0; t < tiles; t += tile_step, ++tno) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                TQ_TTRACE(tno, 4);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    if (kb == 0) TQ_TRACE(4);
                    if (kb == 0) TQ_TTRACE(tno, 5);
                    if (kb == 1) TQ_TRACE(5);
                    tc_fence_after();
                    const uint32_t sa = base + stage * C::kStageBytes;
                    const uint64_t adesc = make_desc_sw128(sa), bdesc = make_desc_sw128(sa + C::kABytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 B (16 bf16) inside the 128 B swizzle span: +2 in 16-byte units
                        if (I8) tc_mma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                          (uint32_t)((kb | k) != 0));
                        else tc_mma_bf16<CTAS>(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                               (uint32_t)((kb | k) != 0));
                    }
                    tc_commit<CTAS>(empty_bar(stage));         // smem slot free once these MMAs retire
                    if (++stage == ring) { stage = 0; phase ^= 1u; }
                }
                tc_commit<CTAS>(tfull_bar(acc));               // accumulator complete -> epilogue
                TQ_TRACE(6);
                TQ_TTRACE(tno, 6);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
        if (LNF) {
            __syncwarp();
            cluster_sync_all();
        }
    } else {
        // ===================== epilogue (warps 0..11) =====================
        const int et = threadIdx.x;                            // 0..383
        const int quarter = warp & 3;                          // TMEM lane quarter this warp may access
        const int third = warp >> 2;                           // 0..2: which of the quarter's three warps
        const bool has_q = ep.out_q.delta != nullptr;
        const bool has_res = ep.res_ctr != nullptr;
        const bool percol = (has_q && ep.out_q_params > 1) || (has_res && ep.out2_params > 1);
        float q2lo = 0.0f, q2hi = 0.0f, res_scale = 1.0f, res_zp = 0.0f;
        if (has_res) {
            grid_of(ep.out2_q, q2lo, q2hi);
            float lo, hi;
            grid_of(ep.res_q, lo, hi);
            const QP rq = resolve(ep.res_q, 0, lo, hi);
            res_scale = rq.scale;
            res_zp = rq.zp;
        }
        float qlo = 0.0f, qhi = 0.0f;
        if (has_q) grid_of(ep.out_q, qlo, qhi);
        float a_scale = 1.0f;
        int a_zp = 0;
        if (ep.a_q.delta != nullptr) {
            float lo, hi;
            grid_of(ep.a_q, lo, hi);
            const QP aq = resolve(ep.a_q, 0, lo, hi);
            a_scale = aq.scale;
            a_zp = (int)aq.zp;
        }
        float wlo = 0.0f, whi = 0.0f;
        if (ep.w_q.delta != nullptr) grid_of(ep.w_q, wlo, whi);
        int acc = 0;
        uint32_t acc_phase = 0;
        float run_min = __int_as_float(0x7f800000), run_max = __int_as_float(0xff800000);
        int tno = 0;
        for (int64_t t = tile0; t < tiles; t += tile_step, ++tno) {
            const int64_t m0 = (t / n_tiles) * (BM * CTAS) + cta_rank * BM, n0 = (t % n_tiles) * BN;
            asm volatile("bar.sync 1, 384;" ::: "memory");      // previous tile's parameter reads done
            if (et == 0) TQ_TTRACE(tno, 0);
            int need_exact = 0;
            if (LNF) {
                float lo, hi;
                grid_of(ep.ln_q, lo, hi);
                need_exact |= resolve(ep.ln_q, 0, lo, hi).exact;
            }
            for (int j = et; j < BN; j += kEpiThreads) {
                const int64_t n = n0 + j;
                float cs = 0.0f, b = 0.0f, qs = 1.0f, qr = 1.0f, clo = 0.0f, chi = 0.0f;
                if (n < N) {
                    const float ws = ep.w_q.delta != nullptr
                                         ? resolve(ep.w_q, ep.w_q_params > 1 ? n : 0, wlo, whi).scale
                                         : 1.0f;
                    cs = a_scale * ws;
                    b = ep.bias != nullptr ? ep.bias[n] : 0.0f;
                    if (has_q) {
                        const QP p = resolve(ep.out_q, ep.out_q_params > 1 ? n : 0, qlo, qhi);
                        qs = p.scale;
                        qr = p.rcp;
                        clo = qlo - p.zp;                       // integers: exact
                        chi = qhi - p.zp;
                        need_exact |= p.exact;
                    }
                }
                params[pidx(BN, 0, 0, j)] = cs;
                params[pidx(BN, 0, 1, j)] = b;
                params[pidx(BN, 1, 0, j)] = qs;
                params[pidx(BN, 1, 1, j)] = -qs;
                params[pidx(BN, 2, 0, j)] = qr;
                params[pidx(BN, 2, 1, j)] = clo;
                params[pidx(BN, 3, 0, j)] = chi;
                float s2 = 1.0f, r2 = 1.0f, clo2 = 0.0f, chi2 = 0.0f;
                if (has_res && n < N) {
                    const QP p2 = resolve(ep.out2_q, ep.out2_params > 1 ? n : 0, q2lo, q2hi);
                    s2 = p2.scale;
                    r2 = p2.rcp;
                    clo2 = q2lo - p2.zp;
                    chi2 = q2hi - p2.zp;
                    need_exact |= p2.exact;
                }
                params[pidx(BN, 3, 1, j)] = chi2;
                params[pidx(BN, 4, 0, j)] = s2;
                params[pidx(BN, 4, 1, j)] = -s2;
                params[pidx(BN, 5, 0, j)] = r2;
                params[pidx(BN, 5, 1, j)] = clo2;
                if (LNF) {
                    params[pidx(BN, 6, 0, j)] = n < N ? ep.ln_gamma[n] : 0.0f;
                    params[pidx(BN, 6, 1, j)] = n < N ? ep.ln_beta[n] : 0.0f;
                }
                if (I8) params[pidx(BN, 7, 0, j)] = __int_as_float(n < N ? a_zp * ep.w_rowsum[n] : 0);
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
            if (et == 0) TQ_TTRACE(tno, 1);
            const int64_t row = m0 + quarter * 32 + lane;
            const bool row_ok = row < M;
            const uint32_t tmem_tile = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
            const uint32_t tf = tfull_bar(acc);
            const bool fast = has_q && !exact && ep.tile_minmax == nullptr && ep.act_fn <= 1;
            const int mode = !fast ? -1 : (has_res ? 4 : ep.act_fn * 2) + (percol ? 1 : 0);
            const float out_lo = has_res ? q2lo : qlo;           // lower edge of the integer grid that is stored
            if (LNF) {
                if (exact) epi_tile_res_ln<BN, false, I8>(ep, params, tmem_tile, third, quarter, lane, row, row_ok, n0, N, res_scale, res_zp, tf, acc_phase, tno);
                else epi_tile_res_ln<BN, true, I8>(ep, params, tmem_tile, third, quarter, lane, row, row_ok, n0, N, res_scale, res_zp, tf, acc_phase, tno);
            } else
            switch (mode) {                                      // warp-uniform
                case 0: epi_tile_fast<BN, 0, false, false, I8>(ep, params, tmem_tile, third, row, row_ok, n0, N, res_scale, res_zp, out_lo, tf, acc_phase, tno); break;
                case 1: epi_tile_fast<BN, 0, true, false, I8>(ep, params, tmem_tile, third, row, row_ok, n0, N, res_scale, res_zp, out_lo, tf, acc_phase, tno); break;
                case 2: epi_tile_fast<BN, 1, false, false, I8>(ep, params, tmem_tile, third, row, row_ok, n0, N, res_scale, res_zp, out_lo, tf, acc_phase, tno); break;
                case 3: epi_tile_fast<BN, 1, true, false, I8>(ep, params, tmem_tile, third, row, row_ok, n0, N, res_scale, res_zp, out_lo, tf, acc_phase, tno); break;
                case 4: epi_tile_fast<BN, 0, false, true, I8>(ep, params, tmem_tile, third, row, row_ok, n0, N, res_scale, res_zp, out_lo, tf, acc_phase, tno); break;
                case 5: epi_tile_fast<BN, 0, true, true, I8>(ep, params, tmem_tile, third, row, row_ok, n0, N, res_scale, res_zp, out_lo, tf, acc_phase, tno); break;
                default:
                    epi_tile_generic<BN, I8>(ep, params, tmem_tile, third, row, row_ok, n0, N, res_scale, res_zp, out_lo, has_q,
                                             has_res, tf, acc_phase, run_min, run_max);
                    break;
            }
            if (et == 0) TQ_TRACE(9);
            if (et == 0) TQ_TTRACE(tno, 3);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTAS == 2) mbar_arrive_remote(tempty_bar(acc), 0u);      // the leader's barrier
                else mbar_arrive(tempty_bar(acc));
            }
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
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x == 0) TQ_TRACE(10);
    if (ep.trace_all != nullptr && threadIdx.x == 0) {
        long long t_end_ns;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end_ns));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        long long* o = ep.trace_all + 4 * (long long)blockIdx.x;
        o[0] = t_start_ns;
        o[1] = t_end_ns;
        o[2] = (long long)smid;
        o[3] = clock64() - t_start_clk;
    }
    if (warp == kMmaWarp) {
        if (CTAS == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"((uint32_t)kTmemCols)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"((uint32_t)kTmemCols)
                         : "memory");
    }
}

// =====================================================================================================
// LEAN int8 kernels (the fused engine's hot path).  Same TMA -> tcgen05 kind::i8 -> TMEM pipeline as
// linear_qdq_kernel; what is different is everything around the accumulators.  The per-tile timelines
// and the ncu source view of the general kernel (profiles/r2_trace_tiles.txt, r2_ncu_engine_layer0.json)
// showed the int8 GEMMs to be bound by their EPILOGUE, not by the tensor pipe: ~2.5 k cycles of
// per-tile parameter set-up (dependent global loads behind a block barrier), ~9 k cycles per 128 x 256
// tile in an epilogue loop that reloaded spilled address registers before every TMEM load and kept
// run-time output-format branches in the loop, against a 3-4.6 k cycle main loop.  Here:
//   * quantizer parameters are per SEGMENT of output columns (1 segment, or 3 for the fused Q|K|V GEMM):
//     resolved once per CTA by a dedicated parameter warp, which also streams the per-column {bias,
//     zero-point correction} pairs of the NEXT tile into a double-buffered shared-memory table while the
//     epilogue warps work on the current one (mbarrier hand-off, no block barrier per tile);
//   * 8 epilogue warps (two per TMEM lane quarter) with up to 168 registers each take 32-column slices:
//     one tcgen05.ld.x32 per slice, the next slice's load in flight during the arithmetic of the current
//     one, output format and activation fixed at compile time, one 32-byte (u8) or two 32-byte (bf16)
//     row pieces per store;
//   * the fused LayerNorm takes its row statistics from EXACT integer sums (sum k, sum k^2 of the centred
//     integers k, |k| <= 255), so they do not depend on the summation order or the tiling: two passes
//     over the tile instead of three, and bit-identical to the general kernels.
// Arithmetic per element is the same operation chain as the general kernel (bit-identical outputs).
// =====================================================================================================
namespace lean {

constexpr int kEpiWarps = 8;
constexpr int kProdWarp = 8, kMmaWarp = 9, kParWarp = 10;
constexpr int kThreads = 32 * 11;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kMaxSeg = 4;
constexpr int kSegFloats = 24;

// MODE: 0 plain, 1 residual + LayerNorm (cluster), 2 NoNorm (elementwise affine), 3 residual + NoNorm
template <int BN, int MODE>
struct Cfg {
    static constexpr bool LNF = MODE == 1;
    static constexpr bool AFF = MODE >= 2;
    static constexpr int kABytes = BM * 128;
    static constexpr int kBBytes = BN * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kColBytes = 2 * (BN / 2) * 16;                 // 2 x float4 {bias.x, bias.y, corr.x, corr.y} per column pair
    static constexpr int kLnBytes = LNF ? (BN / 2) * 16 + 2 * BM * 8 + 8 * BM * 8           // gamma|beta pairs, half partials, cluster partials
                                        : (AFF ? 2 * (BN / 2) * 16 : 0);                    // NoNorm: gamma|beta pairs per tile (double-buffered)
    static constexpr int kSegBytes = kMaxSeg * kSegFloats * 4;
    static constexpr int kParamBytes = kColBytes + kLnBytes + kSegBytes;
    static constexpr int kStagesFit = (227 * 1024 - 1024 - kParamBytes - 256) / kStageBytes;
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr int kBarBytes = (2 * kStages + 8) * 8 + 16;
    static constexpr int kSmemBytes = kStages * kStageBytes + kParamBytes + kBarBytes + 1024;
};

struct Args {
    const float* bias;              // [N] or null
    const int32_t* w_rowsum;        // [N]
    tq_qspec a_q;                   // per-tensor input quantizer
    tq_qspec w_q, out_q;            // nseg parameter slots each
    int64_t seg_width;              // columns per segment (N / nseg)
    int32_t nseg;
    void* y_u8;                     // [M, N] x_int bytes          (exactly one of the two outputs ...
    __nv_bfloat16* y_ctr;           // [M, N] centred bf16 grid     ... unless LNF, where y_ctr is optional)
    int64_t ldc;                    // output row stride in elements (>= N: lets two GEMMs fill one [M, ldc] buffer)
    const unsigned char* res_u8;    // LNF: residual x_int bytes [M, N]
    tq_qspec res_q, out2_q, ln_q;   // LNF: per-tensor
    const float* ln_gamma;
    const float* ln_beta;
    float ln_eps;
    long long* trace;
    long long* trace_tiles;
};

#define TQL_TRACE(slot) do { if (ep.trace != nullptr && blockIdx.x == 0) ep.trace[slot] = clock64(); } while (0)
#define TQL_TTRACE(tile_no, slot) do { if (ep.trace_tiles != nullptr && blockIdx.x == 0 && (tile_no) < 8) \
        ep.trace_tiles[(tile_no) * 8 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
// tcgen05.wait::ld, then "touch" the destination registers: the compiler may not use them before the wait
template <int NV>
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&v)[NV]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < NV; ++i) asm volatile("" : "+r"(v[i]));
}

// resolved per-tensor quantizer -> QReg (clamp bounds in the centred domain)
__device__ __forceinline__ QReg qreg_of(float s, float r, float clo, float chi) {
    QReg q;
    q.s = splat(s); q.ns = splat(-s); q.r = splat(r); q.clo = splat(clo); q.chi = splat(chi);
    return q;
}
// 16 float bit patterns (x_int + 1.5 * 2^23: the integer sits in the low mantissa byte) -> 4 packed words
__device__ __forceinline__ void pack_bytes16(uint32_t (&w)[4], const uint32_t (&b)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        w[i] = __byte_perm(__byte_perm(b[4 * i], b[4 * i + 1], 0x0040), __byte_perm(b[4 * i + 2], b[4 * i + 3], 0x0040), 0x5410);
}

// Loop shape of both epilogues.  The ncu source view of the first version (32-column bodies, 1100 instructions in
// the LayerNorm pass, each executed three times per warp) showed `no_instruction` -- instruction fetch -- as the top
// stall of the hot loop: straight-line code that runs only a few times per CTA has to fit the instruction caches.
// So: a 32-column slice is LOADED at once (tcgen05.ld.x32, the next slice in flight behind it), but PROCESSED as two
// 16-column halves by one compact loop body (8 column pairs, ~200-400 instructions) that always works on registers
// v[0..15]; the second half is rotated down by register moves.

// ---- plain epilogue: y = Q(act(acc * cs + bias)) for one 128 x BN accumulator tile ---------------------
// segment parameters sg[]: 0 cs, 1 s, 2 r, 3 clo, 4 chi, 5 zp + 1.5 * 2^23, 6 exact flag
// Loop shape: a ROLLED loop over the 32-column slices of this warp's half of the tile, one 16-pair body, the next slice's
// tcgen05.ld in flight behind it.  History (all measured): fully unrolled over the tile (2400 instructions = 39 KB with
// GELU) 40-60 % of the samples stalled on `no_instruction` (profiles/r2_ncu_chain_layer0.json, first capture); an 8-pair
// body lost 20 % on the GELU stage (too little independent work for the 12-deep dependent FFMA2 chain of a quantizer);
// this form keeps the ILP of the unrolled one with a quarter of the code.
template <int BN, int ACT, bool FAST, bool OUT8>
__device__ __forceinline__ void epi_plain(const Args& ep, const float4* __restrict__ Pcol, const float* __restrict__ sg,
                                                  uint32_t tmem_tile, int half, int64_t row, bool row_ok, int64_t n0, int64_t N) {
    constexpr int NIT = BN / 64;
    const QReg q = qreg_of(sg[1], sg[2], sg[3], sg[4]);
    const float2 cs2 = splat(sg[0]), off2 = splat(sg[5]);
    unsigned char* o8 = OUT8 ? reinterpret_cast<unsigned char*>(ep.y_u8) + row * ep.ldc + n0 : nullptr;
    __nv_bfloat16* oc = OUT8 ? nullptr : ep.y_ctr + row * ep.ldc + n0;
    uint32_t v[32], vn[32];
    tmem_ld32_nowait(tmem_tile + (uint32_t)(half * 32), vn);
#pragma unroll 1
    for (int it = 0; it < NIT; ++it) {
        tmem_ld_fence(vn);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = vn[i];
        const int c0 = half * 32 + it * 64;
        if (it + 1 < NIT) tmem_ld32_nowait(tmem_tile + (uint32_t)(c0 + 64), vn);
        const float4* P = Pcol + (c0 >> 1);
        uint32_t w[16], b[16];
#pragma unroll
        for (int jp = 0; jp < 16; ++jp) {
            const float4 pp = P[jp];
            const float2 a = make_float2(__int2float_rn((int)v[2 * jp] - __float_as_int(pp.z)),
                                         __int2float_rn((int)v[2 * jp + 1] - __float_as_int(pp.w)));
            float2 f = __ffma2_rn(a, cs2, make_float2(pp.x, pp.y));
            f = act2<ACT>(f);
            const float2 k = ctr2_t<FAST>(f, q);
            if (OUT8) {
                const float2 t = __fadd2_rn(k, off2);
                b[2 * (jp & 7)] = __float_as_uint(t.x);
                b[2 * (jp & 7) + 1] = __float_as_uint(t.y);
                if ((jp & 7) == 7) {
                    uint32_t w4[4];
                    pack_bytes16(w4, b);
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[(jp >> 3) * 4 + i] = w4[i];
                }
            } else {
                w[jp] = pack_bf16(k);
            }
        }
        if (row_ok) {
            if (OUT8) {
                stg256(o8 + c0, *reinterpret_cast<uint32_t(*)[8]>(&w[0]));
            } else {
                stg256(oc + c0, *reinterpret_cast<uint32_t(*)[8]>(&w[0]));
                stg256(oc + c0 + 16, *reinterpret_cast<uint32_t(*)[8]>(&w[8]));
            }
        }
    }
}

// ---- residual + LayerNorm epilogue (one tile per CTA, cluster over the N tiles of a row panel) ----------
//   pass 1   k = Q2( s1 * Q1(acc * cs + bias) + s_r * (r_int - zp_r) )  centred integers, parked as bf16 pairs in
//            the accumulator columns just read; S1 += k, S2 += k * k  (exact: |k| <= 255, <= 128 columns per thread)
//   exchange the two halves of a row -> every CTA of the cluster (DSMEM, int32) -> one cluster barrier
//   mean = s2 * S1 / N, var = s2^2 * (S2 - S1^2 / N) / N  in fp64 from the exact totals, rounded once to fp32
//   pass 2   z = Q3( (s2 * k - mean) * rstd * gamma + beta )
// segment parameters sg[]: 0 cs, 1-4 q1 {s, r, clo, chi}, 6 exact, 7-10 q2 {s, r, clo, chi}, 11 res scale,
// 12 2^23 + res zp, 13-16 q3 {s, r, clo, chi}, 17 zp3 + 1.5 * 2^23
// r0 / r1: the residual row pieces of the first two slices, requested before the accumulator was waited for
template <int BN, bool FAST>
__device__ __forceinline__ void epi_res_ln(const Args& ep, const float4* __restrict__ Pcol, const float4* __restrict__ Pgb,
                                           const float* __restrict__ sg, int2* part, int2* xs, uint32_t tmem_tile, int half,
                                           int quarter, int lane, int64_t row, bool row_ok, int64_t n0, int64_t N,
                                           uint32_t (&rn)[8], uint32_t (&rn2)[8]) {
    constexpr int NIT = BN / 64;
    const QReg q1 = qreg_of(sg[1], sg[2], sg[3], sg[4]);
    const QReg q2 = qreg_of(sg[7], sg[8], sg[9], sg[10]);
    const float2 cs2 = splat(sg[0]);
    const float s1 = sg[1], rs = sg[11], roff = sg[12], s2 = sg[7];
    const int rl = quarter * 32 + lane;
    const unsigned char* rrow = ep.res_u8 + row * N + n0;
    uint32_t v[32], vn[32];
    uint32_t rw[8];
    tmem_ld32_nowait(tmem_tile + (uint32_t)(half * 32), vn);
    float2 S1 = make_float2(0.0f, 0.0f), S2 = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int it = 0; it < NIT; ++it) {
        tmem_ld_fence(vn);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = vn[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) { rw[i] = rn[i]; rn[i] = rn2[i]; }
        const int c0 = half * 32 + it * 64;
        if (it + 1 < NIT) tmem_ld32_nowait(tmem_tile + (uint32_t)(c0 + 64), vn);
        if (it + 2 < NIT && row_ok) ldg256(rrow + c0 + 128, rn2);
        {
            const float4* P = Pcol + (c0 >> 1);
            uint32_t kp[16];
#pragma unroll
            for (int jp = 0; jp < 16; ++jp) {
                const float4 pp = P[jp];
                const float2 a = make_float2(__int2float_rn((int)v[2 * jp] - __float_as_int(pp.z)),
                                             __int2float_rn((int)v[2 * jp + 1] - __float_as_int(pp.w)));
                const float2 f = __ffma2_rn(a, cs2, make_float2(pp.x, pp.y));
                const float2 c = ctr2_t<FAST>(f, q1);
                // residual bytes -> floats without the conversion unit: 0x4B000000 | b = 2^23 + b, minus (2^23 + zp): exact
                const uint32_t wd = rw[jp >> 1];
                const uint32_t lo = __byte_perm(wd, 0x4B000000u, (jp & 1) ? 0x7652 : 0x7650);
                const uint32_t hi = __byte_perm(wd, 0x4B000000u, (jp & 1) ? 0x7653 : 0x7651);
                const float2 rc = make_float2(__fsub_rn(__uint_as_float(lo), roff), __fsub_rn(__uint_as_float(hi), roff));
                // fl(s1 * c) + fl(s_r * r) as the reference materialises them: scalar multiplies (a packed multiply
                // feeding a packed add would be contracted into FFMA2)
                const float2 sum = __fadd2_rn(make_float2(__fmul_rn(s1, c.x), __fmul_rn(s1, c.y)),
                                              make_float2(__fmul_rn(rs, rc.x), __fmul_rn(rs, rc.y)));
                const float2 k = ctr2_t<FAST>(sum, q2);
                S1 = __fadd2_rn(S1, k);                      // integers below 2^24: exact
                S2 = __ffma2_rn(k, k, S2);
                kp[jp] = pack_bf16(k);
            }
            tmem_st16_nowait(tmem_tile + (uint32_t)c0, kp);
        }
    }
    tmem_st_wait();
    if (threadIdx.x == 0) TQL_TRACE(11);
    // ---- exchange ----
    part[half * BM + rl] = make_int2(__float2int_rn(S1.x) + __float2int_rn(S1.y), __float2int_rn(S2.x) + __float2int_rn(S2.y));
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const uint32_t cn = cluster_nctarank(), my = cluster_ctarank();
    if (half == 0) {
        const int2 pa = part[rl], pb = part[BM + rl];
        const int t1 = pa.x + pb.x, t2 = pa.y + pb.y;
        const uint32_t dst = smem_u32(xs + (my * BM + rl));
        for (uint32_t r = 0; r < cn; ++r) {
            st_remote_f32(dst, r, __int_as_float(t1));
            st_remote_f32(dst + 4, r, __int_as_float(t2));
        }
    }
    cluster_sync_all();
    if (threadIdx.x == 0) TQL_TRACE(12);
    long long T1 = 0, T2 = 0;
    for (uint32_t r = 0; r < cn; ++r) {
        const int2 p = xs[r * BM + rl];
        T1 += p.x;
        T2 += p.y;
    }
    float mean, rstd;
    ln_stats_from_sums(T1, T2, N, s2, ep.ln_eps, mean, rstd);
    // ---- pass 2: normalise, affine, output quantizer ----
    const QReg q3 = qreg_of(sg[13], sg[14], sg[15], sg[16]);
    const float2 nmean = splat(-mean), rstd2 = splat(rstd), off3 = splat(sg[17]);
    unsigned char* o8 = reinterpret_cast<unsigned char*>(ep.y_u8) + row * N + n0;
    uint32_t kq[16], kn[16];
    tmem_ld16_nowait(tmem_tile + (uint32_t)(half * 32), kn);
#pragma unroll 1
    for (int it = 0; it < NIT; ++it) {
        tmem_ld_fence(kn);
#pragma unroll
        for (int i = 0; i < 16; ++i) kq[i] = kn[i];
        const int c0 = half * 32 + it * 64;
        if (it + 1 < NIT) tmem_ld16_nowait(tmem_tile + (uint32_t)(c0 + 64), kn);
        {
            const float4* G = Pgb + (c0 >> 1);
            uint32_t wc[16], b[32];
#pragma unroll
            for (int jp = 0; jp < 16; ++jp) {
                const float2 x = make_float2(__fmul_rn(s2, __uint_as_float(kq[jp] << 16)),
                                             __fmul_rn(s2, __uint_as_float(kq[jp] & 0xffff0000u)));
                const float4 gb = G[jp];
                float2 y = __fmul2_rn(__fadd2_rn(x, nmean), rstd2);
                y = __ffma2_rn(y, make_float2(gb.x, gb.y), make_float2(gb.z, gb.w));
                const float2 k3 = ctr2_t<FAST>(y, q3);
                wc[jp] = pack_bf16(k3);
                const float2 t = __fadd2_rn(k3, off3);
                b[2 * jp] = __float_as_uint(t.x);
                b[2 * jp + 1] = __float_as_uint(t.y);
            }
            if (row_ok) {
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    o[i] = __byte_perm(__byte_perm(b[4 * i], b[4 * i + 1], 0x0040), __byte_perm(b[4 * i + 2], b[4 * i + 3], 0x0040), 0x5410);
                stg256(o8 + c0, o);
                if (ep.y_ctr != nullptr) {
                    stg256(ep.y_ctr + row * N + n0 + c0, *reinterpret_cast<uint32_t(*)[8]>(&wc[0]));
                    stg256(ep.y_ctr + row * N + n0 + c0 + 16, *reinterpret_cast<uint32_t(*)[8]>(&wc[8]));
                }
            }
        }
    }
}

// ---- NoNorm epilogue (MobileBERT, reference models/quantized_mobilebert.py:58-72): QuantNoNorm is the elementwise
// affine y * weight + bias with fake-quantized parameters and a quantized output, always applied to a quantized dense
// output [+ quantized residual sum]:   z = Q3( s * k (*) gamma + beta ),  k = Q1(acc * cs + bias)  or, with the residual,
// k = Q2( s1 * Q1(.) + s_r * (r_int - zp_r) ).  One pass, no statistics, no cluster.  Multiply and add are separate
// roundings like the reference's two tensor ops.  Segment parameters as in epi_res_ln.
template <int BN, bool FAST, bool RES>
__device__ __forceinline__ void epi_affine(const Args& ep, const float4* __restrict__ Pcol, const float4* __restrict__ Pgb,
                                           const float* __restrict__ sg, uint32_t tmem_tile, int half, int64_t row, bool row_ok,
                                           int64_t n0, int64_t N) {
    constexpr int NIT = BN / 64;
    const QReg q1 = qreg_of(sg[1], sg[2], sg[3], sg[4]);
    const QReg q2 = qreg_of(sg[7], sg[8], sg[9], sg[10]);
    const QReg q3 = qreg_of(sg[13], sg[14], sg[15], sg[16]);
    const float2 cs2 = splat(sg[0]), off3 = splat(sg[17]);
    const float s1 = sg[1], rs = sg[11], roff = sg[12];
    const float2 sk = splat(RES ? sg[7] : sg[1]);
    const unsigned char* rrow = RES ? ep.res_u8 + row * N + n0 : nullptr;
    unsigned char* o8 = reinterpret_cast<unsigned char*>(ep.y_u8) + row * ep.ldc + n0;
    uint32_t va[32], vb[32], rw[8], rn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rn[i] = 0u;
    tmem_ld32_nowait(tmem_tile + (uint32_t)(half * 32), va);
    if (RES && row_ok) ldg256(rrow + half * 32, rn);
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        uint32_t (&v)[32] = (it & 1) ? vb : va;
        uint32_t (&vn)[32] = (it & 1) ? va : vb;
        tmem_ld_fence(v);
#pragma unroll
        for (int i = 0; i < 8; ++i) rw[i] = rn[i];
        const int c0 = half * 32 + it * 64;
        if (it + 1 < NIT) {
            tmem_ld32_nowait(tmem_tile + (uint32_t)(c0 + 64), vn);
            if (RES && row_ok) ldg256(rrow + c0 + 64, rn);
        }
        const float4* P = Pcol + (c0 >> 1);
        const float4* G = Pgb + (c0 >> 1);
        uint32_t w[8], b[16];
#pragma unroll
        for (int jp = 0; jp < 16; ++jp) {
            const float4 pp = P[jp];
            const float2 a = make_float2(__int2float_rn((int)v[2 * jp] - __float_as_int(pp.z)),
                                         __int2float_rn((int)v[2 * jp + 1] - __float_as_int(pp.w)));
            const float2 f = __ffma2_rn(a, cs2, make_float2(pp.x, pp.y));
            float2 k = ctr2_t<FAST>(f, q1);
            if (RES) {
                const uint32_t wd = rw[jp >> 1];
                const uint32_t lo = __byte_perm(wd, 0x4B000000u, (jp & 1) ? 0x7652 : 0x7650);
                const uint32_t hi = __byte_perm(wd, 0x4B000000u, (jp & 1) ? 0x7653 : 0x7651);
                const float2 rc = make_float2(__fsub_rn(__uint_as_float(lo), roff), __fsub_rn(__uint_as_float(hi), roff));
                const float2 sum = __fadd2_rn(make_float2(__fmul_rn(s1, k.x), __fmul_rn(s1, k.y)),
                                              make_float2(__fmul_rn(rs, rc.x), __fmul_rn(rs, rc.y)));
                k = ctr2_t<FAST>(sum, q2);
            }
            const float4 gb = G[jp];
            const float2 x = __fmul2_rn(sk, k);                                        // dequantized NoNorm input
            // x * weight, then + bias: scalar multiplies (a packed multiply feeding a packed add would be contracted)
            const float2 y = __fadd2_rn(make_float2(__fmul_rn(x.x, gb.x), __fmul_rn(x.y, gb.y)), make_float2(gb.z, gb.w));
            const float2 k3 = ctr2_t<FAST>(y, q3);
            const float2 t = __fadd2_rn(k3, off3);
            b[2 * (jp & 7)] = __float_as_uint(t.x);
            b[2 * (jp & 7) + 1] = __float_as_uint(t.y);
            if ((jp & 7) == 7) {
                uint32_t w4[4];
                pack_bytes16(w4, b);
#pragma unroll
                for (int i = 0; i < 4; ++i) w[(jp >> 3) * 4 + i] = w4[i];
            }
        }
        if (row_ok) stg256(o8 + c0, w);
    }
}

template <int BN, int ACT, int MODE, bool OUT8>
__global__ void __launch_bounds__(kThreads, 1)
linear_lean_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                   int64_t M, int64_t N, int64_t K, int ring, Args ep) {
    using C = Cfg<BN, MODE>;
    constexpr bool LNF = MODE == 1, AFF = MODE >= 2;
    constexpr bool QX = LNF || AFF;               // three quantizers (+ residual) instead of one
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    unsigned char* par_ptr = base_ptr + C::kStages * C::kStageBytes;
    float4* Pcol = reinterpret_cast<float4*>(par_ptr);                       // [2][BN / 2]
    float4* Pgb = reinterpret_cast<float4*>(par_ptr + C::kColBytes);         // LNF: [BN / 2] {gamma.x, gamma.y, beta.x, beta.y}; AFF: [2][BN / 2]
    int2* part = reinterpret_cast<int2*>(par_ptr + C::kColBytes + (LNF ? (BN / 2) * 16 : 0));     // [2][BM]
    int2* xs = part + 2 * BM;                                                 // [8][BM]
    float* segp = reinterpret_cast<float*>(par_ptr + C::kColBytes + C::kLnBytes);   // [kMaxSeg][kSegFloats]
    const uint32_t bar0 = base + C::kStages * C::kStageBytes + C::kParamBytes;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::kStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + 2 + s); };
    auto pfull_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + 4 + s); };
    auto pempty_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + 6 + s); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * C::kStages + 8);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
        base_ptr + C::kStages * C::kStageBytes + C::kParamBytes + 8 * (2 * C::kStages + 8));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m_tiles = (M + BM - 1) / BM, n_tiles = N / BN;
    const int64_t tiles = m_tiles * n_tiles;
    const int64_t tile0 = blockIdx.x, tile_step = gridDim.x;
    const int num_kb = (int)(K / 128);

    if (threadIdx.x == 0) TQL_TRACE(0);
    int p_stage = 0, p_pre = 0;
    uint32_t p_phase = 0;
    if (warp == kProdWarp && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tile0 < tiles) {
            pdl_wait();                               // A is produced by the previous kernel
            const int32_t m0 = (int32_t)((tile0 / n_tiles) * BM), n0 = (int32_t)((tile0 % n_tiles) * BN);
            p_pre = num_kb < ring ? num_kb : ring;
            for (int kb = 0; kb < p_pre; ++kb) {      // ring slots are free: no empty-barrier wait
                mbar_expect_tx(full_bar(p_stage), C::kStageBytes);
                const uint32_t sa = base + p_stage * C::kStageBytes;
                tma_load_2d<1>(sa, &map_a, kb * 128, m0, full_bar(p_stage));
                tma_load_2d<1>(sa + C::kABytes, &map_w, kb * 128, n0, full_bar(p_stage));
                if (++p_stage == ring) { p_stage = 0; p_phase ^= 1u; }
            }
        }
    }
    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(tfull_bar(s), 1);
                mbar_init(tempty_bar(s), kEpiWarps);
                mbar_init(pfull_bar(s), 1);
                mbar_init(pempty_bar(s), kEpiWarps);
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
    if (threadIdx.x == 0) TQL_TRACE(1);

    if (warp == kProdWarp) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = p_stage;
            uint32_t phase = p_phase;
            int tno = 0;
            for (int64_t t = tile0; t < tiles; t += tile_step, ++tno) {
                const int32_t m0 = (int32_t)((t / n_tiles) * BM), n0 = (int32_t)((t % n_tiles) * BN);
                for (int kb = (t == tile0 ? p_pre : 0); kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_expect_tx(full_bar(stage), C::kStageBytes);
                    const uint32_t sa = base + stage * C::kStageBytes;
                    tma_load_2d<1>(sa, &map_a, kb * 128, m0, full_bar(stage));
                    tma_load_2d<1>(sa + C::kABytes, &map_w, kb * 128, n0, full_bar(stage));
                    if (++stage == ring) { stage = 0; phase ^= 1u; }
                }
                TQL_TTRACE(tno, 7);
            }
        }
        if (LNF) {                                    // the epilogue's cluster barrier counts every thread
            __syncwarp();
            cluster_sync_all();
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t a_s8 = (ep.a_q.zero_float == nullptr && ep.a_q.is_signed != nullptr && *ep.a_q.is_signed) ? 1u : 0u;
            const uint32_t w_s8 = (ep.w_q.zero_float == nullptr && ep.w_q.is_signed != nullptr && *ep.w_q.is_signed) ? 1u : 0u;
            const uint32_t idesc = (2u << 4) | (a_s8 << 7) | (w_s8 << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int tno = 0;
            for (int64_t t = tile0; t < tiles; t += tile_step, ++tno) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                TQL_TTRACE(tno, 4);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    if (kb == 0) TQL_TTRACE(tno, 5);
                    tc_fence_after();
                    const uint32_t sa = base + stage * C::kStageBytes;
                    const uint64_t adesc = make_desc_sw128(sa), bdesc = make_desc_sw128(sa + C::kABytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                    tc_commit<1>(empty_bar(stage));
                    if (++stage == ring) { stage = 0; phase ^= 1u; }
                }
                tc_commit<1>(tfull_bar(acc));
                TQL_TTRACE(tno, 6);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
        if (LNF) {
            __syncwarp();
            cluster_sync_all();
        }
    } else if (warp == kParWarp) {
        // ===================== parameter warp =====================
        // One lane per quantizer: lane 0 a_q, 1 res_q, 2 out2_q, 3 ln_q, 4.. w_q[j], 8.. out_q[j] -- their (independent)
        // global loads are in flight together; the first tile's column loads are issued before any of them is used.
        QP mine = make_qp(1.0f, 0.0f, 0.0f, 0.0f);
        {
            float lo = 0.0f, hi = 0.0f;
            const tq_qspec* qs = nullptr;
            int slot = 0;
            if (lane == 0) qs = &ep.a_q;
            else if ((LNF || MODE == 3) && lane == 1) qs = &ep.res_q;
            else if ((LNF || MODE == 3) && lane == 2) qs = &ep.out2_q;
            else if (QX && lane == 3) qs = &ep.ln_q;
            else if (lane >= 4 && lane < 4 + ep.nseg) { qs = &ep.w_q; slot = lane - 4; }
            else if (lane >= 8 && lane < 8 + ep.nseg) { qs = &ep.out_q; slot = lane - 8; }
            if (qs != nullptr) {
                grid_of(*qs, lo, hi);
                mine = resolve(*qs, slot, lo, hi);
            }
        }
        float4 first[(BN / 2 + 31) / 32];
        float4 gb[(BN / 2 + 31) / 32];
        if (tile0 < tiles) {
            const int64_t n0 = (tile0 % n_tiles) * BN;
#pragma unroll
            for (int i = 0; i < (BN / 2 + 31) / 32; ++i) {
                const int jp = lane + 32 * i;
                first[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                gb[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (jp < BN / 2) {
                    const int64_t n = n0 + 2 * jp;
                    first[i] = make_float4(ep.bias != nullptr ? ep.bias[n] : 0.0f, ep.bias != nullptr ? ep.bias[n + 1] : 0.0f,
                                           __int_as_float(ep.w_rowsum[n]), __int_as_float(ep.w_rowsum[n + 1]));
                    if (LNF) gb[i] = make_float4(ep.ln_gamma[n], ep.ln_gamma[n + 1], ep.ln_beta[n], ep.ln_beta[n + 1]);
                }
            }
        }
        // quantizers -> every lane (shuffles), segment parameters by lanes < nseg
        const float a_scale = __shfl_sync(0xffffffffu, mine.scale, 0);
        const int a_zp = (int)__shfl_sync(0xffffffffu, mine.zp, 0);
        {
            const int j = lane < ep.nseg ? lane : 0;
            const float w_scale = __shfl_sync(0xffffffffu, mine.scale, 4 + j);
            const float o_scale = __shfl_sync(0xffffffffu, mine.scale, 8 + j), o_rcp = __shfl_sync(0xffffffffu, mine.rcp, 8 + j);
            const float o_zp = __shfl_sync(0xffffffffu, mine.zp, 8 + j), o_lo = __shfl_sync(0xffffffffu, mine.lo, 8 + j);
            const float o_hi = __shfl_sync(0xffffffffu, mine.hi, 8 + j);
            int exact = __shfl_sync(0xffffffffu, mine.exact, 8 + j);
            const float r_scale = __shfl_sync(0xffffffffu, mine.scale, 1), r_zp = __shfl_sync(0xffffffffu, mine.zp, 1);
            const float s2 = __shfl_sync(0xffffffffu, mine.scale, 2), r2 = __shfl_sync(0xffffffffu, mine.rcp, 2);
            const float z2 = __shfl_sync(0xffffffffu, mine.zp, 2), l2 = __shfl_sync(0xffffffffu, mine.lo, 2);
            const float h2 = __shfl_sync(0xffffffffu, mine.hi, 2);
            const int e2 = __shfl_sync(0xffffffffu, mine.exact, 2);
            const float s3 = __shfl_sync(0xffffffffu, mine.scale, 3), r3 = __shfl_sync(0xffffffffu, mine.rcp, 3);
            const float z3 = __shfl_sync(0xffffffffu, mine.zp, 3), l3 = __shfl_sync(0xffffffffu, mine.lo, 3);
            const float h3 = __shfl_sync(0xffffffffu, mine.hi, 3);
            const int e3 = __shfl_sync(0xffffffffu, mine.exact, 3);
            if (lane < ep.nseg && lane < kMaxSeg) {
                float* sg = segp + lane * kSegFloats;
                sg[0] = __fmul_rn(a_scale, w_scale);
                sg[1] = o_scale; sg[2] = o_rcp; sg[3] = o_lo - o_zp; sg[4] = o_hi - o_zp;
                sg[5] = __fadd_rn(o_zp, 12582912.0f);               // x_int = k + zp, + 1.5 * 2^23
                if (QX) {
                    exact |= (MODE == 2 ? 0 : e2) | e3;
                    sg[7] = s2; sg[8] = r2; sg[9] = l2 - z2; sg[10] = h2 - z2;
                    sg[11] = r_scale; sg[12] = __fadd_rn(8388608.0f, r_zp);
                    sg[13] = s3; sg[14] = r3; sg[15] = l3 - z3; sg[16] = h3 - z3;
                    sg[17] = __fadd_rn(z3, 12582912.0f);
                }
                sg[6] = __int_as_float(exact);
            }
        }
        if (LNF && tile0 < tiles) {
#pragma unroll
            for (int i = 0; i < (BN / 2 + 31) / 32; ++i)
                if (lane + 32 * i < BN / 2) Pgb[lane + 32 * i] = gb[i];
        }
        int pb = 0;
        uint32_t pphase = 0;
        for (int64_t t = tile0; t < tiles; t += tile_step) {
            const int64_t n0 = (t % n_tiles) * BN;
            mbar_wait(pempty_bar(pb), pphase ^ 1u);
            float4* P = Pcol + pb * (BN / 2);
            if (t == tile0) {
#pragma unroll
                for (int i = 0; i < (BN / 2 + 31) / 32; ++i)
                    if (lane + 32 * i < BN / 2)
                        P[lane + 32 * i] = make_float4(first[i].x, first[i].y, __int_as_float(a_zp * __float_as_int(first[i].z)),
                                                       __int_as_float(a_zp * __float_as_int(first[i].w)));
            } else {
                for (int jp = lane; jp < BN / 2; jp += 32) {
                    const int64_t n = n0 + 2 * jp;
                    const float b0 = ep.bias != nullptr ? ep.bias[n] : 0.0f, b1 = ep.bias != nullptr ? ep.bias[n + 1] : 0.0f;
                    P[jp] = make_float4(b0, b1, __int_as_float(a_zp * ep.w_rowsum[n]), __int_as_float(a_zp * ep.w_rowsum[n + 1]));
                }
            }
            if (AFF) {
                float4* Gt = Pgb + pb * (BN / 2);
                for (int jp = lane; jp < BN / 2; jp += 32) {
                    const int64_t n = n0 + 2 * jp;
                    Gt[jp] = make_float4(ep.ln_gamma[n], ep.ln_gamma[n + 1], ep.ln_beta[n], ep.ln_beta[n + 1]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(pfull_bar(pb));
            if (++pb == 2) { pb = 0; pphase ^= 1u; }
        }
        if (LNF) {
            __syncwarp();
            cluster_sync_all();
        }
    } else {
        // ===================== epilogue (warps 0..7) =====================
        const int quarter = warp & 3, half = warp >> 2;
        int acc = 0, pb = 0;
        uint32_t acc_phase = 0, pphase = 0;
        int tno = 0;
        for (int64_t t = tile0; t < tiles; t += tile_step, ++tno) {
            const int64_t m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
            if (threadIdx.x == 0) TQL_TTRACE(tno, 0);
            mbar_wait(pfull_bar(pb), pphase);
            if (threadIdx.x == 0) TQL_TTRACE(tno, 1);
            const float* sg = segp + (int)(n0 / ep.seg_width) * kSegFloats;
            const int exact = __float_as_int(sg[6]);
            const int64_t row = m0 + quarter * 32 + lane;
            const bool row_ok = row < M;
            const uint32_t tmem_tile = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
            const float4* P = Pcol + pb * (BN / 2);
            uint32_t r0[8], r1[8];
            if (LNF) {                                 // the residual does not depend on the MMAs: request it now
#pragma unroll
                for (int i = 0; i < 8; ++i) r0[i] = r1[i] = 0u;
                if (row_ok) {
                    const unsigned char* rrow = ep.res_u8 + row * N + n0 + half * 32;
                    ldg256(rrow, r0);
                    if (BN > 64) ldg256(rrow + 64, r1);
                }
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            if (threadIdx.x == 0) { TQL_TRACE(8); TQL_TTRACE(tno, 2); }
            if (LNF) {
                if (exact) epi_res_ln<BN, false>(ep, P, Pgb, sg, part, xs, tmem_tile, half, quarter, lane, row, row_ok, n0, N, r0, r1);
                else epi_res_ln<BN, true>(ep, P, Pgb, sg, part, xs, tmem_tile, half, quarter, lane, row, row_ok, n0, N, r0, r1);
            } else if (AFF) {
                const float4* Gt = Pgb + pb * (BN / 2);
                if (exact) epi_affine<BN, false, MODE == 3>(ep, P, Gt, sg, tmem_tile, half, row, row_ok, n0, N);
                else epi_affine<BN, true, MODE == 3>(ep, P, Gt, sg, tmem_tile, half, row, row_ok, n0, N);
            } else {
                if (exact) epi_plain<BN, ACT, false, OUT8>(ep, P, sg, tmem_tile, half, row, row_ok, n0, N);
                else epi_plain<BN, ACT, true, OUT8>(ep, P, sg, tmem_tile, half, row, row_ok, n0, N);
            }
            if (threadIdx.x == 0) { TQL_TRACE(9); TQL_TTRACE(tno, 3); }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(tempty_bar(acc));
                mbar_arrive(pempty_bar(pb));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            if (++pb == 2) { pb = 0; pphase ^= 1u; }
        }
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TQL_TRACE(10);
    if (warp == kMmaWarp)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                     : "memory");
}

// =====================================================================================================
// PEG kernels: per-embedding-group activations through the lean int8 pipeline (BASELINE config 3,
// reference utils/per_embd_quant_utils.py:54-68, quantization/range_estimators.py:82-112, quantizers.py:213-217).
// With K groups along the hidden dimension the A operand's scale and zero point change every d / K columns of the
// CONTRACTION, so one integer accumulator per output element is not enough.  The k-loop runs GROUP BY GROUP into two
// ping-pong TMEM accumulators (128 columns each); the epilogue warps drain group g while the tensor core works on
// group g + 1 and keep the running fp32 sum  sum_g (s_a[g] s_w) * (acc_g[n] - zp_g * rowsum_g[n])  in registers
// (every partial product sum is an exact integer; one fp32 FMA per group and element).  The OUTPUT side of PEG
// (per-group quantizers of q / k / v / g / u / x / h / y / z, per-group residual) costs nothing extra: tiles are 128
// columns wide = aligned to the groups, so every quantizer is a per-TILE constant ("segment").
// =====================================================================================================
namespace peg {


constexpr int BN = 128;
constexpr int kMaxGroups = 8;
constexpr int kSegFloats = 24;
constexpr int kTmemColsPeg = 256;

template <bool LNF>
struct Cfg {
    static constexpr int kABytes = BM * 128;
    static constexpr int kBBytes = BN * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    // per-tile parameter buffer: bias pairs [BN / 2] float2 | corr [G][BN] int32 | segment parameters | cs per group
    static constexpr int kTileParBytes = (BN / 2) * 8 + kMaxGroups * BN * 4 + kSegFloats * 4 + kMaxGroups * 4 + 32;
    static constexpr int kLnBytes = LNF ? (BN / 2) * 16 + 2 * BM * 8 + 8 * BM * 8 + 64 : 0;   // gamma|beta, half partials, cluster partials, cluster scales
    static constexpr int kParamBytes = 2 * kTileParBytes + kLnBytes;
    static constexpr int kStagesFit = (227 * 1024 - 1024 - kParamBytes - 256) / kStageBytes;
    static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
    static constexpr int kSmemBytes = kStages * kStageBytes + kParamBytes + (2 * kStages + 8) * 8 + 16 + 1024;
};

struct Args {
    const float* bias;              // [N] or null
    const int32_t* w_grp_rowsum;    // [G][N] sum of w_int over the K columns of group g
    tq_qspec a_q;                   // G parameter slots (delta[G], zero_float[G]); asymmetric / unsigned grid
    int32_t a_groups;               // G >= 1, K / G a multiple of 128
    tq_qspec w_q, out_q;            // w_params / out_params slots (1 or nseg)
    int32_t w_params, out_params;
    int64_t seg_width;              // columns per segment (a multiple of 128)
    void* y_u8;
    __nv_bfloat16* y_ctr;
    const unsigned char* res_u8;    // LNF: residual x_int bytes [M, N]
    tq_qspec res_q, out2_q, ln_q;   // LNF: res_params / out2_params / ln_params slots (1 or nseg)
    int32_t res_params, out2_params, ln_params;
    const float* ln_gamma;
    const float* ln_beta;
    float ln_eps;
    long long* trace;
    long long* trace_tiles;
};

// segment parameters sg[] as in the lean kernels: 0 (unused), 1-4 q1 {s, r, clo, chi}, 5 zp1 + 1.5 * 2^23, 6 exact,
// 7-10 q2, 11 res scale, 12 2^23 + res zp, 13-16 q3, 17 zp3 + 1.5 * 2^23

// drain one group's accumulator (this warp's 64 columns) into the running sums
__device__ __forceinline__ void drain_group(float (&run)[64], uint32_t tmem_buf, int half, const int32_t* __restrict__ corr,
                                            float csg) {
    // four 16-column chunks, the next one in flight during the arithmetic of the current one (two 16-register sets:
    // with the 64 running sums a 32-column double buffer spilled)
    uint32_t va[16], vb[16];
    tmem_ld16_nowait(tmem_buf + (uint32_t)(half * 32), va);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t (&v)[16] = (i & 1) ? vb : va;
        uint32_t (&vn)[16] = (i & 1) ? va : vb;
        tmem_ld_fence(v);
        const int c0 = half * 32 + (i >> 1) * 64 + (i & 1) * 16;
        if (i + 1 < 4) tmem_ld16_nowait(tmem_buf + (uint32_t)(half * 32 + ((i + 1) >> 1) * 64 + ((i + 1) & 1) * 16), vn);
        const int4* cr = reinterpret_cast<const int4*>(corr + c0);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const int4 c = cr[j4];
            const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = __int2float_rn((int)v[4 * j4 + j] - cc[j]);
                const int r = (i >> 1) * 32 + (i & 1) * 16 + 4 * j4 + j;
                run[r] = __fmaf_rn(csg, a, run[r]);
            }
        }
    }
}

template <int ACT, bool FAST, bool OUT8>
__device__ __forceinline__ void finish_plain(const Args& ep, const float (&run)[64], const float2* __restrict__ Pb,
                                             const float* __restrict__ sg, int half, int64_t row, bool row_ok, int64_t n0,
                                             int64_t N) {
    const QReg q = qreg_of(sg[1], sg[2], sg[3], sg[4]);
    const float2 off2 = splat(sg[5]);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int c0 = half * 32 + s * 64;
        uint32_t w[16], b[16];
#pragma unroll
        for (int jp = 0; jp < 16; ++jp) {
            float2 f = __fadd2_rn(make_float2(run[s * 32 + 2 * jp], run[s * 32 + 2 * jp + 1]), Pb[(c0 >> 1) + jp]);
            f = act2<ACT>(f);
            const float2 k = ctr2_t<FAST>(f, q);
            if (OUT8) {
                const float2 t = __fadd2_rn(k, off2);
                b[2 * (jp & 7)] = __float_as_uint(t.x);
                b[2 * (jp & 7) + 1] = __float_as_uint(t.y);
                if ((jp & 7) == 7) {
                    uint32_t w4[4];
                    pack_bytes16(w4, b);
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[(jp >> 3) * 4 + i] = w4[i];
                }
            } else {
                w[jp] = pack_bf16(k);
            }
        }
        if (row_ok) {
            if (OUT8) {
                stg256(reinterpret_cast<unsigned char*>(ep.y_u8) + row * N + n0 + c0, *reinterpret_cast<uint32_t(*)[8]>(&w[0]));
            } else {
                __nv_bfloat16* oc = ep.y_ctr + row * N + n0 + c0;
                stg256(oc, *reinterpret_cast<uint32_t(*)[8]>(&w[0]));
                stg256(oc + 16, *reinterpret_cast<uint32_t(*)[8]>(&w[8]));
            }
        }
    }
}

template <bool FAST>
__device__ __forceinline__ void finish_res_ln(const Args& ep, const float (&run)[64], const float2* __restrict__ Pb,
                                              const float4* __restrict__ Pgb, const float* __restrict__ sg, int2* part, int2* xs,
                                              const float* xscale, uint32_t tmem_park, int half, int quarter, int lane, int64_t row, bool row_ok,
                                              int64_t n0, int64_t N, const uint32_t (&r0)[8], const uint32_t (&r1)[8]) {
    const QReg q1 = qreg_of(sg[1], sg[2], sg[3], sg[4]);
    const QReg q2 = qreg_of(sg[7], sg[8], sg[9], sg[10]);
    const float s1 = sg[1], rs = sg[11], roff = sg[12], s2 = sg[7];
    const int rl = quarter * 32 + lane;
    float2 S1 = make_float2(0.0f, 0.0f), S2 = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int c0 = half * 32 + s * 64;
        uint32_t kp[16];
#pragma unroll
        for (int jp = 0; jp < 16; ++jp) {
            const float2 f = __fadd2_rn(make_float2(run[s * 32 + 2 * jp], run[s * 32 + 2 * jp + 1]), Pb[(c0 >> 1) + jp]);
            const float2 c = ctr2_t<FAST>(f, q1);
            const uint32_t wd = s == 0 ? r0[jp >> 1] : r1[jp >> 1];
            const uint32_t lo = __byte_perm(wd, 0x4B000000u, (jp & 1) ? 0x7652 : 0x7650);
            const uint32_t hi = __byte_perm(wd, 0x4B000000u, (jp & 1) ? 0x7653 : 0x7651);
            const float2 rc = make_float2(__fsub_rn(__uint_as_float(lo), roff), __fsub_rn(__uint_as_float(hi), roff));
            const float2 sum = __fadd2_rn(make_float2(__fmul_rn(s1, c.x), __fmul_rn(s1, c.y)),
                                          make_float2(__fmul_rn(rs, rc.x), __fmul_rn(rs, rc.y)));
            const float2 k = ctr2_t<FAST>(sum, q2);
            S1 = __fadd2_rn(S1, k);
            S2 = __ffma2_rn(k, k, S2);
            kp[jp] = pack_bf16(k);
        }
        tmem_st16_nowait(tmem_park + (uint32_t)c0, kp);
    }
    tmem_st_wait();
    part[half * BM + rl] = make_int2(__float2int_rn(S1.x) + __float2int_rn(S1.y), __float2int_rn(S2.x) + __float2int_rn(S2.y));
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const uint32_t cn = cluster_nctarank(), my = cluster_ctarank();
    if (half == 0) {
        const int2 pa = part[rl], pb = part[BM + rl];
        const uint32_t dst = smem_u32(xs + (my * BM + rl));
        for (uint32_t r = 0; r < cn; ++r) {
            st_remote_f32(dst, r, __int_as_float(pa.x + pb.x));
            st_remote_f32(dst + 4, r, __int_as_float(pa.y + pb.y));
        }
    }
    cluster_sync_all();
    // per-GROUP output scale s2: the LayerNorm input x = s2[g(n)] * k differs per cluster member, so the row statistics
    // are combined from (S1, S2) of every member scaled by ITS s2 (exchanged with the sums)
    double m = 0.0, e2 = 0.0;
    for (uint32_t r = 0; r < cn; ++r) {
        const int2 p = xs[r * BM + rl];
        const double sr = (double)xscale[r];
        m += sr * (double)p.x;
        e2 += sr * sr * (double)p.y;
    }
    m /= (double)N;
    double var = e2 / (double)N - m * m;
    var = var < 0.0 ? 0.0 : var;
    const float mean = (float)m;
    const float rstd = __fdiv_rn(1.0f, sqrtf(__fadd_rn((float)var, ep.ln_eps)));
    const QReg q3 = qreg_of(sg[13], sg[14], sg[15], sg[16]);
    const float2 nmean = splat(-mean), rstd2 = splat(rstd), off3 = splat(sg[17]);
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const int c0 = half * 32 + s * 64;
        uint32_t kq[16];
        tmem_ld16_nowait(tmem_park + (uint32_t)c0, kq);
        tmem_ld_fence(kq);
        const float4* G = Pgb + (c0 >> 1);
        uint32_t w8[8], wc[16], b[16];
#pragma unroll
        for (int jp = 0; jp < 16; ++jp) {
            const float2 x = make_float2(__fmul_rn(s2, __uint_as_float(kq[jp] << 16)),
                                         __fmul_rn(s2, __uint_as_float(kq[jp] & 0xffff0000u)));
            const float4 gb = G[jp];
            float2 y = __fmul2_rn(__fadd2_rn(x, nmean), rstd2);
            y = __ffma2_rn(y, make_float2(gb.x, gb.y), make_float2(gb.z, gb.w));
            const float2 k3 = ctr2_t<FAST>(y, q3);
            wc[jp] = pack_bf16(k3);
            const float2 t = __fadd2_rn(k3, off3);
            b[2 * (jp & 7)] = __float_as_uint(t.x);
            b[2 * (jp & 7) + 1] = __float_as_uint(t.y);
            if ((jp & 7) == 7) {
                uint32_t w4[4];
                pack_bytes16(w4, b);
#pragma unroll
                for (int i = 0; i < 4; ++i) w8[(jp >> 3) * 4 + i] = w4[i];
            }
        }
        if (row_ok) {
            stg256(reinterpret_cast<unsigned char*>(ep.y_u8) + row * N + n0 + c0, w8);
            if (ep.y_ctr != nullptr) {
                __nv_bfloat16* oc = ep.y_ctr + row * N + n0 + c0;
                stg256(oc, *reinterpret_cast<uint32_t(*)[8]>(&wc[0]));
                stg256(oc + 16, *reinterpret_cast<uint32_t(*)[8]>(&wc[8]));
            }
        }
    }
}

template <int ACT, bool LNF, bool OUT8>
__global__ void __launch_bounds__(kThreads, 1)
linear_peg_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  int64_t M, int64_t N, int64_t K, int ring, Args ep) {
    using C = Cfg<LNF>;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    unsigned char* par_ptr = base_ptr + C::kStages * C::kStageBytes;
    auto tile_par = [&](int b) { return par_ptr + b * C::kTileParBytes; };
    float4* Pgb = reinterpret_cast<float4*>(par_ptr + 2 * C::kTileParBytes);
    int2* part = reinterpret_cast<int2*>(par_ptr + 2 * C::kTileParBytes + (BN / 2) * 16);
    int2* xs = part + 2 * BM;                                                 // [8][BM] (S1, S2) of every cluster member
    float* xscale = reinterpret_cast<float*>(xs + 8 * BM);                    // [8] output scale s2 of every cluster member
    const uint32_t bar0 = base + C::kStages * C::kStageBytes + C::kParamBytes;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::kStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + 2 + s); };
    auto pfull_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + 4 + s); };
    auto pempty_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + 6 + s); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * C::kStages + 8);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
        base_ptr + C::kStages * C::kStageBytes + C::kParamBytes + 8 * (2 * C::kStages + 8));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m_tiles = (M + BM - 1) / BM, n_tiles = N / BN;
    const int64_t tiles = m_tiles * n_tiles;
    const int64_t tile0 = blockIdx.x, tile_step = gridDim.x;
    const int G = ep.a_groups;
    const int kb_per_group = (int)(K / 128) / G;
    const int num_kb = kb_per_group * G;

    if (threadIdx.x == 0) TQL_TRACE(0);
    int p_stage = 0, p_pre = 0;
    uint32_t p_phase = 0;
    if (warp == kProdWarp && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tile0 < tiles) {
            pdl_wait();
            const int32_t m0 = (int32_t)((tile0 / n_tiles) * BM), n0 = (int32_t)((tile0 % n_tiles) * BN);
            p_pre = num_kb < ring ? num_kb : ring;
            for (int kb = 0; kb < p_pre; ++kb) {
                mbar_expect_tx(full_bar(p_stage), C::kStageBytes);
                const uint32_t sa = base + p_stage * C::kStageBytes;
                tma_load_2d<1>(sa, &map_a, kb * 128, m0, full_bar(p_stage));
                tma_load_2d<1>(sa + C::kABytes, &map_w, kb * 128, n0, full_bar(p_stage));
                if (++p_stage == ring) { p_stage = 0; p_phase ^= 1u; }
            }
        }
    }
    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(tfull_bar(s), 1);
                mbar_init(tempty_bar(s), kEpiWarps);
                mbar_init(pfull_bar(s), 1);
                mbar_init(pempty_bar(s), kEpiWarps);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)kTmemColsPeg)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) TQL_TRACE(1);

    if (warp == kProdWarp) {
        if (lane == 0) {
            int stage = p_stage;
            uint32_t phase = p_phase;
            for (int64_t t = tile0; t < tiles; t += tile_step) {
                const int32_t m0 = (int32_t)((t / n_tiles) * BM), n0 = (int32_t)((t % n_tiles) * BN);
                for (int kb = (t == tile0 ? p_pre : 0); kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_expect_tx(full_bar(stage), C::kStageBytes);
                    const uint32_t sa = base + stage * C::kStageBytes;
                    tma_load_2d<1>(sa, &map_a, kb * 128, m0, full_bar(stage));
                    tma_load_2d<1>(sa + C::kABytes, &map_w, kb * 128, n0, full_bar(stage));
                    if (++stage == ring) { stage = 0; phase ^= 1u; }
                }
            }
        }
        if (LNF) {
            __syncwarp();
            cluster_sync_all();
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0) {
            const uint32_t w_s8 = (ep.w_q.zero_float == nullptr && ep.w_q.is_signed != nullptr && *ep.w_q.is_signed) ? 1u : 0u;
            const uint32_t idesc = (2u << 4) | (w_s8 << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int stage = 0, buf = 0;
            uint32_t phase = 0, bphase = 0;
            int tno = 0;
            for (int64_t t = tile0; t < tiles; t += tile_step, ++tno) {
                for (int g = 0; g < G; ++g) {
                    mbar_wait(tempty_bar(buf), bphase ^ 1u);
                    tc_fence_after();
                    if (g == 0) TQL_TTRACE(tno, 4);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                    for (int kb = 0; kb < kb_per_group; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = base + stage * C::kStageBytes;
                        const uint64_t adesc = make_desc_sw128(sa), bdesc = make_desc_sw128(sa + C::kABytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                        tc_commit<1>(empty_bar(stage));
                        if (++stage == ring) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit<1>(tfull_bar(buf));
                    if (g == G - 1) TQL_TTRACE(tno, 6);
                    if (++buf == 2) { buf = 0; bphase ^= 1u; }
                }
            }
        }
        if (LNF) {
            __syncwarp();
            cluster_sync_all();
        }
    } else if (warp == kParWarp) {
        // lanes 0..4: w_q, out_q, out2_q, ln_q, res_q of the tile's segment; lanes 8..8+G-1: a_q of group g
        int pb = 0;
        uint32_t pphase = 0;
        bool first = true;
        for (int64_t t = tile0; t < tiles; t += tile_step) {
            const int64_t n0 = (t % n_tiles) * BN;
            const int seg = (int)(n0 / ep.seg_width);
            QP mine = make_qp(1.0f, 0.0f, 0.0f, 0.0f);
            {
                const tq_qspec* qs = nullptr;
                int slot = 0;
                if (lane == 0) { qs = &ep.w_q; slot = ep.w_params > 1 ? seg : 0; }
                else if (lane == 1) { qs = &ep.out_q; slot = ep.out_params > 1 ? seg : 0; }
                else if (LNF && lane == 2) { qs = &ep.out2_q; slot = ep.out2_params > 1 ? seg : 0; }
                else if (LNF && lane == 3) { qs = &ep.ln_q; slot = ep.ln_params > 1 ? seg : 0; }
                else if (LNF && lane == 4) { qs = &ep.res_q; slot = ep.res_params > 1 ? seg : 0; }
                else if (lane >= 8 && lane < 8 + G) { qs = &ep.a_q; slot = lane - 8; }
                if (qs != nullptr) {
                    float lo, hi;
                    grid_of(*qs, lo, hi);
                    mine = resolve(*qs, slot, lo, hi);
                }
            }
            mbar_wait(pempty_bar(pb), pphase ^ 1u);
            unsigned char* tp = tile_par(pb);
            float2* Pb = reinterpret_cast<float2*>(tp);
            int32_t* corr = reinterpret_cast<int32_t*>(tp + (BN / 2) * 8);
            float* sg = reinterpret_cast<float*>(tp + (BN / 2) * 8 + kMaxGroups * BN * 4);
            float* csg = sg + kSegFloats;
            for (int jp = lane; jp < BN / 2; jp += 32) {
                const int64_t n = n0 + 2 * jp;
                Pb[jp] = make_float2(ep.bias != nullptr ? ep.bias[n] : 0.0f, ep.bias != nullptr ? ep.bias[n + 1] : 0.0f);
                if (LNF && first) Pgb[jp] = make_float4(ep.ln_gamma[n], ep.ln_gamma[n + 1], ep.ln_beta[n], ep.ln_beta[n + 1]);
            }
            const float w_scale = __shfl_sync(0xffffffffu, mine.scale, 0);
            // all row sums of the tile first (independent loads in flight together), then the stores: a load -> store
            // per element made every one of the G * 4 loads of a lane wait for the previous one (12 k cycles per tile)
            int32_t rsv[kMaxGroups][BN / 32];
#pragma unroll
            for (int g = 0; g < kMaxGroups; ++g)
#pragma unroll
                for (int i = 0; i < BN / 32; ++i)
                    rsv[g][i] = g < G ? __ldg(ep.w_grp_rowsum + (int64_t)g * N + n0 + lane + 32 * i) : 0;
#pragma unroll
            for (int g = 0; g < kMaxGroups; ++g) {
                if (g < G) {
                    const int zp_g = (int)__shfl_sync(0xffffffffu, mine.zp, 8 + g);
                    const float s_g = __shfl_sync(0xffffffffu, mine.scale, 8 + g);
#pragma unroll
                    for (int i = 0; i < BN / 32; ++i) corr[g * BN + lane + 32 * i] = zp_g * rsv[g][i];
                    if (lane == 0) csg[g] = __fmul_rn(s_g, w_scale);
                }
            }
            {
                const float o_s = __shfl_sync(0xffffffffu, mine.scale, 1), o_r = __shfl_sync(0xffffffffu, mine.rcp, 1);
                const float o_z = __shfl_sync(0xffffffffu, mine.zp, 1), o_lo = __shfl_sync(0xffffffffu, mine.lo, 1);
                const float o_hi = __shfl_sync(0xffffffffu, mine.hi, 1);
                int exact = __shfl_sync(0xffffffffu, mine.exact, 1);
                const float s2 = __shfl_sync(0xffffffffu, mine.scale, 2), r2 = __shfl_sync(0xffffffffu, mine.rcp, 2);
                const float z2 = __shfl_sync(0xffffffffu, mine.zp, 2), l2 = __shfl_sync(0xffffffffu, mine.lo, 2);
                const float h2 = __shfl_sync(0xffffffffu, mine.hi, 2);
                const int e2 = __shfl_sync(0xffffffffu, mine.exact, 2);
                const float s3 = __shfl_sync(0xffffffffu, mine.scale, 3), r3 = __shfl_sync(0xffffffffu, mine.rcp, 3);
                const float z3 = __shfl_sync(0xffffffffu, mine.zp, 3), l3 = __shfl_sync(0xffffffffu, mine.lo, 3);
                const float h3 = __shfl_sync(0xffffffffu, mine.hi, 3);
                const int e3 = __shfl_sync(0xffffffffu, mine.exact, 3);
                const float r_s = __shfl_sync(0xffffffffu, mine.scale, 4), r_z = __shfl_sync(0xffffffffu, mine.zp, 4);
                if (lane == 0) {
                    sg[1] = o_s; sg[2] = o_r; sg[3] = o_lo - o_z; sg[4] = o_hi - o_z;
                    sg[5] = __fadd_rn(o_z, 12582912.0f);
                    if (LNF) {
                        exact |= e2 | e3;
                        sg[7] = s2; sg[8] = r2; sg[9] = l2 - z2; sg[10] = h2 - z2;
                        sg[11] = r_s; sg[12] = __fadd_rn(8388608.0f, r_z);
                        sg[13] = s3; sg[14] = r3; sg[15] = l3 - z3; sg[16] = h3 - z3;
                        sg[17] = __fadd_rn(z3, 12582912.0f);
                    }
                    sg[6] = __int_as_float(exact);
                }
            }
            first = false;
            __syncwarp();
            if (lane == 0) mbar_arrive(pfull_bar(pb));
            if (++pb == 2) { pb = 0; pphase ^= 1u; }
        }
        if (LNF) {
            __syncwarp();
            cluster_sync_all();
        }
    } else {
        // ===================== epilogue (warps 0..7) =====================
        const int quarter = warp & 3, half = warp >> 2;
        int buf = 0, pb = 0;
        uint32_t bphase = 0, pphase = 0;
        int tno = 0;
        for (int64_t t = tile0; t < tiles; t += tile_step, ++tno) {
            const int64_t m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
            if (threadIdx.x == 0) TQL_TTRACE(tno, 0);
            const int64_t row = m0 + quarter * 32 + lane;
            const bool row_ok = row < M;
            uint32_t r0[8], r1[8];
            if (LNF) {
#pragma unroll
                for (int i = 0; i < 8; ++i) r0[i] = r1[i] = 0u;
                if (row_ok) {
                    const unsigned char* rrow = ep.res_u8 + row * N + n0 + half * 32;
                    ldg256(rrow, r0);
                    ldg256(rrow + 64, r1);
                }
            }
            mbar_wait(pfull_bar(pb), pphase);
            if (threadIdx.x == 0) TQL_TTRACE(tno, 1);
            const unsigned char* tp = tile_par(pb);
            const float2* Pb = reinterpret_cast<const float2*>(tp);
            const int32_t* corr = reinterpret_cast<const int32_t*>(tp + (BN / 2) * 8);
            const float* sg = reinterpret_cast<const float*>(tp + (BN / 2) * 8 + kMaxGroups * BN * 4);
            const float* csg = sg + kSegFloats;
            const int exact = __float_as_int(sg[6]);
            float run[64];                    // (one drain body for every group: two inlined copies cost ~100 register moves per group)
#pragma unroll
            for (int i = 0; i < 64; ++i) run[i] = 0.0f;
            uint32_t last_buf = 0;
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
                mbar_wait(tfull_bar(buf), bphase);
                tc_fence_after();
                if (threadIdx.x == 0 && g == 0) TQL_TTRACE(tno, 2);
                const uint32_t tb = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN);
                drain_group(run, tb, half, corr + g * BN, csg[g]);
                last_buf = tb;
                tc_fence_before();
                __syncwarp();
                // (LNF parks its intermediate in the LAST group's buffer: that one is released after the epilogue)
                if (lane == 0 && !(LNF && g == G - 1)) mbar_arrive(tempty_bar(buf));
                if (++buf == 2) { buf = 0; bphase ^= 1u; }
            }
            if (threadIdx.x == 0) TQL_TTRACE(tno, 5);          // all groups drained
            if (LNF) {
                if (warp == 0 && lane == 0) {       // this CTA's output scale s2 travels with its sums (see finish_res_ln)
                    const uint32_t my = cluster_ctarank(), cn = cluster_nctarank();
                    const uint32_t dst = smem_u32(xscale + my);
                    for (uint32_t r = 0; r < cn; ++r) st_remote_f32(dst, r, sg[7]);
                }
                if (exact) finish_res_ln<false>(ep, run, Pb, Pgb, sg, part, xs, xscale, last_buf, half, quarter, lane, row, row_ok, n0, N, r0, r1);
                else finish_res_ln<true>(ep, run, Pb, Pgb, sg, part, xs, xscale, last_buf, half, quarter, lane, row, row_ok, n0, N, r0, r1);
            } else {
                if (exact) finish_plain<ACT, false, OUT8>(ep, run, Pb, sg, half, row, row_ok, n0, N);
                else finish_plain<ACT, true, OUT8>(ep, run, Pb, sg, half, row, row_ok, n0, N);
            }
            if (threadIdx.x == 0) TQL_TTRACE(tno, 3);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pempty_bar(pb));
            if (++pb == 2) { pb = 0; pphase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TQL_TRACE(10);
    if (warp == kMmaWarp)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemColsPeg)
                     : "memory");
}

}  // namespace peg

// =====================================================================================================
// CHAIN kernel: a sequence of int8 GEMM stages of one encoder layer in ONE launch.
// A sequence of 128 tokens never interacts with another sequence outside attention, so a 128-row panel can be
// carried through  attention-out + LN -> FFN-in + GELU -> FFN-out + LN -> next layer's Q | K | V  by one cluster of
// N_hidden / 192 CTAs with cluster-level synchronisation only: every stage's A operand is the panel its own cluster
// wrote in the stage before (global memory, L2 resident).  What this removes per stage is the dependent-launch cost
// of the one-kernel-per-GEMM form (pipeline fill after a grid-wide dependency, tail of the slowest CTA, TMEM
// allocation, barrier set-up: ~4.5 us per launch, ~20 % of the BERT-base step).
// Roles, pipelines and epilogues are those of the lean kernels (same functions -> bit-identical results).
//   stage kinds   0  plain, per-segment quantizers, bf16 centred output, tiles of 192      (Q | K | V)
//                 1  plain + GELU, byte output, tiles of 256                               (FFN-in)
//                 2  residual + LayerNorm, byte output, one 192-column tile per CTA        (attention-out, FFN-out)
// Between stages: every epilogue thread fences its global stores (gpu scope + async proxy), then one cluster
// barrier; the producer issues the next stage's TMA loads after it.
// =====================================================================================================
namespace chain {

// One launch carries every 128-row panel (= one sequence of 128 tokens) through a LIST of stages; a cluster of
// hidden / 192 CTAs owns the panel, so stage i + 1 reads what the same cluster wrote in stage i and only cluster barriers
// separate the stages.  Stage kinds:
//   0  int8 GEMM, per-segment quantizers, bf16 centred-grid output (Q | K | V)           tile 128 x 192
//   1  int8 GEMM + GELU, x_int byte output (FFN-in)                                       tile 128 x 256
//   2  int8 GEMM + residual + LayerNorm over the cluster, byte output                     tile 128 x 192, one per CTA
//   3  attention of heads / cluster heads per CTA (bf16 tensor-core products, tq_attn.cuh), byte output
// The stage descriptors live in global memory (a plan built once per engine): the whole encoder -- Q|K|V of layer 0, then
// per layer attention, attention-output + LN, FFN-in, FFN-out + LN, next Q|K|V -- is ONE launch.
struct alignas(16) StageTail {           // everything but the tensor maps: copied to shared memory one stage ahead
    Args ep;                             // kind 3: a_q / w_q / res_q = Q / K / V, out2_q = scores, ln_q = probs, out_q = context, bias = mask
    int64_t N, K;                        // kind 3: N = hidden size, K = heads
    int32_t kind, tiles;                 // tiles (kind 3: heads) of this stage per CTA
    uint32_t idesc;                      // kinds 0-2: the kind::i8 instruction descriptor (operand signedness resolved by the host)
    int32_t ovl;                         // 1: GELU stage whose successor (ovl == 2, a LayerNorm stage with K = this N) starts its
                                         // main loop on the columns of this stage's first tiles while the last tile is in its epilogue
};
static_assert(sizeof(StageTail) % 16 == 0, "stage tail is copied in 16-byte pieces");
struct alignas(64) StageDesc {
    CUtensorMap map_a, map_w;            // kind 3: map_a = the Q | K | V buffer [M, 3 hidden] bf16, box 64 x 128
    StageTail t;
};
constexpr int BNL = 192;
constexpr int BNF = 256;
constexpr int kStageBytes = BM * 128 + BNF * 128;        // 48 KB: a GEMM k-block (A 16 KB | W <= 32 KB) or one head's Q | K | V
constexpr int kRing = 4;
constexpr int kColBytes = 2 * (BNF / 2) * 16;
constexpr int kGbBytes = (BNL / 2) * 16;                 // LayerNorm gamma | beta pairs; attention: the sequence's additive mask
constexpr int kXchgBytes = 2 * BM * 8 + 8 * BM * 8;      // LayerNorm partial sums; attention: row max / sum exchange
constexpr int kSegBytes = 2 * kMaxSeg * kSegFloats * 4;  // two stages in flight; attention: [6][8] resolved quantizers
constexpr int kTailBytes = 2 * ((int)sizeof(StageTail) + 15) / 16 * 16;
constexpr int kParamBytes = kColBytes + kGbBytes + kXchgBytes + kSegBytes + kTailBytes;
constexpr int kNumBars = 2 * kRing + 8 + 4;
constexpr int kSmemBytes = kRing * kStageBytes + kParamBytes + kNumBars * 8 + 16 + 1024;
static_assert(kSmemBytes <= 227 * 1024, "chain kernel shared memory budget");
static_assert(6 * 8 <= kMaxSeg * kSegFloats, "attention quantizer table fits a segment-parameter slot");

struct Params {
    const StageDesc* st;                 // [n] in global memory
    int64_t M;
    int32_t n;
    long long* trace;                    // tools: [CTA][stage][4] clock64 stamps (stage top, first accumulator, epilogue end, past barrier)
    int32_t gpu_fence;                   // 1: __threadfence() at every stage end as well
};

__device__ __forceinline__ int bn_of(int kind) { return kind == 1 ? BNF : BNL; }
// MN-major SWIZZLE_128B operand (V of the PV product): 64 MN elements (128 B) contiguous per K row, 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ QP load_qp(const float* o) {
    QP p;
    p.scale = o[0]; p.zp = o[1]; p.lo = o[2]; p.hi = o[3]; p.rcp = o[4]; p.exact = __float_as_int(o[5]);
    return p;
}

// Overlapped pair (FFN-in + GELU -> FFN-out + LayerNorm): CTA c, tile j of the GELU stage produces the k-blocks
// (c T + j) * 2 + {0, 1} of the next stage's A operand.  Consumption order of the next stage: first the k-blocks of every
// member's tiles 0 .. T - 2 (ready at the MID barrier, after the epilogue of tile T - 2), then those of the last tiles.
__device__ __forceinline__ int ovl_kb(int i, int T, int n_early) {
    int cc, j, h;
    if (i < n_early) {
        const int per = (T - 1) * 2;
        cc = i / per;
        const int r = i - cc * per;
        j = r >> 1;
        h = r & 1;
    } else {
        const int r = i - n_early;
        cc = r >> 1;
        j = T - 1;
        h = r & 1;
    }
    return (cc * T + j) * 2 + h;
}

// Measured and dropped (profiles/r2_trace_chain_mcast.txt, r2_trace_chain_12warps.txt): (a) TMA-multicasting the A
// k-blocks across the cluster (each member loads 128 / cluster rows) leaves every main loop unchanged -- the bound is the
// bytes LANDING in an SM's shared memory per clock (~60 B/clk: 16 KB + 24 KB per k-block in ~670 cycles), which multicast
// does not reduce; (b) twelve epilogue warps (three per scheduler, setmaxnreg 152 / 48) leave every epilogue unchanged --
// the epilogues are bound by the FMA + ALU pipe time of their instruction mix (packed FP32 2.1-2.3 cycles, FMNMX / PRMT
// 2 cycles, MUFU 8 cycles per warp instruction), not by latency.
// ATT: the plan has attention stages (kind 3); plans of GEMM stages only run the instantiation without that code
template <bool ATT>
__global__ void __launch_bounds__(kThreads, 1) linear_chain_kernel(const Params P) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    unsigned char* par_ptr = base_ptr + kRing * kStageBytes;
    float4* Pcol = reinterpret_cast<float4*>(par_ptr);                                  // [2][BNF / 2]
    float4* Pgb = reinterpret_cast<float4*>(par_ptr + kColBytes);                       // [BNL / 2]
    int2* part = reinterpret_cast<int2*>(par_ptr + kColBytes + kGbBytes);               // [2][BM]
    int2* xs = part + 2 * BM;                                                            // [8][BM]
    float* segp = reinterpret_cast<float*>(par_ptr + kColBytes + kGbBytes + kXchgBytes); // [2][kMaxSeg][kSegFloats]
    const StageTail* tails = reinterpret_cast<const StageTail*>(par_ptr + kColBytes + kGbBytes + kXchgBytes + kSegBytes);   // [2]
    float* smask = reinterpret_cast<float*>(Pgb);                                        // attention stages
    float* xchg = reinterpret_cast<float*>(xs);
    const uint32_t bar0 = base + kRing * kStageBytes + kParamBytes;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kRing + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kRing + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kRing + 2 + s); };
    auto pfull_bar = [&](int s) { return bar0 + 8u * (2 * kRing + 4 + s); };
    auto pempty_bar = [&](int s) { return bar0 + 8u * (2 * kRing + 6 + s); };
    auto pready_bar = [&](int s) { return bar0 + 8u * (2 * kRing + 8 + s); };           // attention: P tile written
    auto oready_bar = [&](int s) { return bar0 + 8u * (2 * kRing + 10 + s); };          // attention: O accumulated
    const uint32_t tmem_slot = bar0 + 8u * kNumBars;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
        base_ptr + kRing * kStageBytes + kParamBytes + 8 * kNumBars);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
    const int64_t m0 = (int64_t)(blockIdx.x / csize) * BM;        // this cluster's row panel
    const int nst = P.n;
    const StageDesc* __restrict__ ST = P.st;

    if (warp == kProdWarp && lane == 0) {
        for (int s = 0; s < kRing; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(tfull_bar(s), 1);
                mbar_init(tempty_bar(s), kEpiWarps);
                mbar_init(pfull_bar(s), 1);
                mbar_init(pempty_bar(s), kEpiWarps);
                mbar_init(pready_bar(s), kEpiThreads);
                mbar_init(oready_bar(s), 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Stage descriptors: the tensor maps are used from global memory, everything else is copied into one of two shared
    // slots by the parameter warp ONE STAGE AHEAD (before it arrives at the stage-end barrier) -- reading them from
    // global memory at every use costs an L2 round trip per 128-byte line per warp per stage (measured: +3-4 k cycles
    // on every stage).
    auto copy_tail = [&](int s) {                     // parameter warp
        const uint4* src = reinterpret_cast<const uint4*>(&ST[s].t);
        uint4* dst = reinterpret_cast<uint4*>(const_cast<StageTail*>(tails + (s & 1)));
        for (int i = lane; i < (int)(sizeof(StageTail) / 16); i += 32) dst[i] = __ldg(src + i);
    };
    if (warp == kParWarp) copy_tail(0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_trigger();
    // The WEIGHT halves of a stage's first ring slots do not depend on the stage before: they are requested while the
    // previous stage (or the previous kernel) is still in its epilogue; only the A halves wait for the boundary.
    int p_stage = 0, p_pre = 0;
    uint32_t p_phase = 0;
    auto preissue_w = [&](int s) {                    // producer lane 0: W of the first k-blocks of stage s
        const StageDesc& S = ST[s];                   // (ahead of the shared copy: global memory, off the critical path)
        p_pre = 0;
        const int kind = S.t.kind;
        if (ATT && kind == 3) return;
        if (S.t.ovl == 2) return;                     // its first k-blocks were requested whole inside the stage before
        const int bn = bn_of(kind);
        const int num_kb = (int)(S.t.K / 128);
        const uint32_t bytes = (uint32_t)(BM * 128 + bn * 128);
        const int32_t n0 = (int32_t)((crank * S.t.tiles) * bn);
        int st = p_stage;
        uint32_t ph = p_phase;
        p_pre = num_kb < kRing ? num_kb : kRing;
        for (int kb = 0; kb < p_pre; ++kb) {
            mbar_wait(empty_bar(st), ph ^ 1u);
            mbar_expect_tx(full_bar(st), bytes);
            tma_load_2d<1>(base + st * kStageBytes + BM * 128, &S.map_w, kb * 128, n0, full_bar(st));
            if (++st == kRing) { st = 0; ph ^= 1u; }
        }
    };
    if (warp == kProdWarp && lane == 0) preissue_w(0);
    pdl_wait();                                       // the first stage's A / residual come from the previous kernel

    if (warp == kProdWarp) {
        // ===================== TMA producer =====================
        for (int s = 0; s < nst; ++s) {
            const StageDesc& G = ST[s];
            const StageTail& S = tails[s & 1];
            const int kind = S.kind;
            if (lane == 0) {
                if (ATT && kind == 3) {
                    const int32_t D = (int32_t)S.N;
                    for (int j = 0; j < S.tiles; ++j) {                 // one ring slot per head: Q | K | V, 16 KB each
                        const int32_t h = (int32_t)(crank * S.tiles + j);
                        mbar_wait(empty_bar(p_stage), p_phase ^ 1u);
                        mbar_expect_tx(full_bar(p_stage), 3u * 16384u);
                        const uint32_t sa = base + p_stage * kStageBytes;
                        tma_load_2d<1>(sa, &G.map_a, h * 64, (int32_t)m0, full_bar(p_stage));
                        tma_load_2d<1>(sa + 16384, &G.map_a, D + h * 64, (int32_t)m0, full_bar(p_stage));
                        tma_load_2d<1>(sa + 32768, &G.map_a, 2 * D + h * 64, (int32_t)m0, full_bar(p_stage));
                        if (++p_stage == kRing) { p_stage = 0; p_phase ^= 1u; }
                    }
                } else if (S.ovl == 2) {
                    // second part of an overlapped pair: the k-blocks of the predecessor's LAST tiles (the early ones went
                    // out before the stage boundary, see below)
                    const int Tp = tails[(s - 1) & 1].tiles, total = (int)(S.K / 128), n_early = (int)csize * (Tp - 1) * 2;
                    const uint32_t bytes = (uint32_t)(BM * 128 + BNL * 128);
                    const int32_t n0 = (int32_t)(crank * BNL);
                    for (int i = n_early; i < total; ++i) {
                        const int kb = ovl_kb(i, Tp, n_early);
                        const uint32_t sa = base + p_stage * kStageBytes;
                        mbar_wait(empty_bar(p_stage), p_phase ^ 1u);
                        mbar_expect_tx(full_bar(p_stage), bytes);
                        tma_load_2d<1>(sa + BM * 128, &G.map_w, kb * 128, n0, full_bar(p_stage));
                        tma_load_2d<1>(sa, &G.map_a, kb * 128, (int32_t)m0, full_bar(p_stage));
                        if (++p_stage == kRing) { p_stage = 0; p_phase ^= 1u; }
                    }
                } else {
                    const int bn = bn_of(kind);
                    const int num_kb = (int)(S.K / 128);
                    const uint32_t bytes = (uint32_t)(BM * 128 + bn * 128);
                    for (int j = 0; j < S.tiles; ++j) {
                        const int32_t n0 = (int32_t)((crank * S.tiles + j) * bn);
                        for (int kb = 0; kb < num_kb; ++kb) {
                            const uint32_t sa = base + p_stage * kStageBytes;
                            if (j > 0 || kb >= p_pre) {
                                mbar_wait(empty_bar(p_stage), p_phase ^ 1u);
                                mbar_expect_tx(full_bar(p_stage), bytes);
                                tma_load_2d<1>(sa + BM * 128, &G.map_w, kb * 128, n0, full_bar(p_stage));
                            }
                            tma_load_2d<1>(sa, &G.map_a, kb * 128, (int32_t)m0, full_bar(p_stage));
                            if (++p_stage == kRing) { p_stage = 0; p_phase ^= 1u; }
                        }
                    }
                }
                if (s + 1 < nst && S.ovl != 1) preissue_w(s + 1);
            }
            __syncwarp();
            if (S.ovl == 1) {
                // overlapped pair, first part: after the MID barrier (every member has written its tiles 0 .. T - 2) the
                // successor's main loop starts on those columns -- under the epilogue of this stage's last tile
                cluster_sync_all();
                if (lane == 0) {
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    const StageDesc& GN = ST[s + 1];
                    const int n_early = (int)csize * (S.tiles - 1) * 2;
                    const uint32_t bytes = (uint32_t)(BM * 128 + BNL * 128);
                    const int32_t n0 = (int32_t)(crank * BNL);
                    for (int i = 0; i < n_early; ++i) {
                        const int kb = ovl_kb(i, S.tiles, n_early);
                        const uint32_t sa = base + p_stage * kStageBytes;
                        mbar_wait(empty_bar(p_stage), p_phase ^ 1u);
                        mbar_expect_tx(full_bar(p_stage), bytes);
                        tma_load_2d<1>(sa + BM * 128, &GN.map_w, kb * 128, n0, full_bar(p_stage));
                        tma_load_2d<1>(sa, &GN.map_a, kb * 128, (int32_t)m0, full_bar(p_stage));
                        if (++p_stage == kRing) { p_stage = 0; p_phase ^= 1u; }
                    }
                    p_pre = 0;
                }
                __syncwarp();
            }
            if (kind == 2) cluster_sync_all();                // the stage's LayerNorm statistics exchange
            if (s + 1 < nst) {
                cluster_sync_all();                           // stage boundary: the panel of stage s is complete
                asm volatile("fence.proxy.async.global;" ::: "memory");
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        uint32_t att_ph0 = 0, att_ph1 = 0;                    // phases of pready[0 / 1]
        for (int s = 0; s < nst; ++s) {
            const StageTail& S = tails[s & 1];
            const int kind = S.kind;
            if (ATT && lane == 0 && kind == 3) {
                // per head: S = Q K^T (K-major operands, 4 x k16), then -- once the softmax warps have written P over
                // Q | K -- O = P V (V MN-major, 8 x k16) into the first 64 score columns.  The scores of head j + 1 are
                // issued BEFORE waiting for P of head j: its softmax overlaps the next head's loads and products.
                constexpr uint32_t idS = make_idesc(128, 128), idO = make_idesc(128, 64) | (1u << 16);
                auto issue_scores = [&]() {
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * kStageBytes;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BNF);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_bf16<1>(d_tmem, make_desc_sw128(sa) + (uint64_t)(2 * k), make_desc_sw128(sa + 16384) + (uint64_t)(2 * k), idS, (uint32_t)(k != 0));
                    tc_commit<1>(tfull_bar(acc));
                    if (++stage == kRing) { stage = 0; phase ^= 1u; }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                };
                int sj = stage, aj = acc;
                issue_scores();
                for (int j = 0; j < S.tiles; ++j) {
                    const int sn = stage, an = acc;
                    if (j + 1 < S.tiles) issue_scores();
                    mbar_wait(pready_bar(aj), aj ? att_ph1 : att_ph0);
                    tc_fence_after();
                    const uint32_t sa = base + sj * kStageBytes;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(aj * BNF);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        tc_mma_bf16<1>(d_tmem, make_desc_sw128(sa + (k >> 2) * 16384) + (uint64_t)(2 * (k & 3)), desc_mn_sw128(sa + 32768 + k * 2048), idO,
                                       (uint32_t)(k != 0));
                    tc_commit<1>(oready_bar(aj));
                    tc_commit<1>(empty_bar(sj));
                    if (aj) att_ph1 ^= 1u; else att_ph0 ^= 1u;
                    sj = sn;
                    aj = an;
                }
            } else if (lane == 0 && S.ovl == 2) {
                // second part of an overlapped pair: the accumulator already holds the early k-blocks
                const int Tp = tails[(s - 1) & 1].tiles, total = (int)(S.K / 128), n_early = (int)csize * (Tp - 1) * 2;
                const uint32_t idesc = S.idesc;
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BNF);
                for (int i = n_early; i < total; ++i) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * kStageBytes;
                    const uint64_t adesc = make_desc_sw128(sa), bdesc = make_desc_sw128(sa + BM * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                    tc_commit<1>(empty_bar(stage));
                    if (++stage == kRing) { stage = 0; phase ^= 1u; }
                }
                tc_commit<1>(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            } else if (lane == 0) {
                const int bn = bn_of(kind);
                const int num_kb = (int)(S.K / 128);
                const uint32_t idesc = S.idesc;
                for (int j = 0; j < S.tiles; ++j) {
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BNF);
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = base + stage * kStageBytes;
                        const uint64_t adesc = make_desc_sw128(sa), bdesc = make_desc_sw128(sa + BM * 128);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                        tc_commit<1>(empty_bar(stage));
                        if (++stage == kRing) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit<1>(tfull_bar(acc));
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
            }
            // (the other lanes follow the ring / accumulator counters: they do not use them)
            __syncwarp();
            if (S.ovl == 1) {
                cluster_sync_all();                           // MID barrier of an overlapped pair
                if (lane == 0) {                              // the successor's early k-blocks into the free accumulator
                    const int n_early = (int)csize * (S.tiles - 1) * 2;
                    const uint32_t idesc = ST[s + 1].t.idesc;
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BNF);
                    for (int i = 0; i < n_early; ++i) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = base + stage * kStageBytes;
                        const uint64_t adesc = make_desc_sw128(sa), bdesc = make_desc_sw128(sa + BM * 128);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((i | k) != 0));
                        tc_commit<1>(empty_bar(stage));
                        if (++stage == kRing) { stage = 0; phase ^= 1u; }
                    }
                }
                __syncwarp();
            }
            if (kind == 2) cluster_sync_all();
            if (s + 1 < nst) cluster_sync_all();
        }
    } else if (warp == kParWarp) {
        // ===================== parameter warp =====================
        int pb = 0;
        uint32_t pphase = 0;
        bool pending = false;                         // arrived at a stage-end cluster barrier, not yet waited
        for (int s = 0; s < nst; ++s) {
            const StageTail& S = tails[s & 1];
            const Args& ep = S.ep;
            const int kind = S.kind;
            const bool lnf = kind == 2;
            const int bn = bn_of(kind);
            float* sgs = segp + (s & 1) * kMaxSeg * kSegFloats;
            QP mine = make_qp(1.0f, 0.0f, 0.0f, 0.0f);
            int a_zp = 0;
            if (ATT && kind == 3) {
                if (lane < 6) {
                    const tq_qspec& q = lane == 0 ? ep.out2_q : lane == 1 ? ep.ln_q : lane == 2 ? ep.out_q : lane == 3 ? ep.a_q : lane == 4 ? ep.w_q : ep.res_q;
                    float lo, hi;
                    grid_of(q, lo, hi);
                    const QP p = resolve(q, 0, lo, hi);
                    float* o = sgs + lane * 8;
                    o[0] = p.scale; o[1] = p.zp; o[2] = p.lo; o[3] = p.hi; o[4] = p.rcp; o[5] = __int_as_float(p.exact);
                }
            } else {
                float lo = 0.0f, hi = 0.0f;
                const tq_qspec* qs = nullptr;
                int slot = 0;
                if (lane == 0) qs = &ep.a_q;
                else if (lnf && lane == 1) qs = &ep.res_q;
                else if (lnf && lane == 2) qs = &ep.out2_q;
                else if (lnf && lane == 3) qs = &ep.ln_q;
                else if (lane >= 4 && lane < 4 + ep.nseg) { qs = &ep.w_q; slot = lane - 4; }
                else if (lane >= 8 && lane < 8 + ep.nseg) { qs = &ep.out_q; slot = lane - 8; }
                if (qs != nullptr) {
                    grid_of(*qs, lo, hi);
                    mine = resolve(*qs, slot, lo, hi);
                }
                const float a_scale = __shfl_sync(0xffffffffu, mine.scale, 0);
                a_zp = (int)__shfl_sync(0xffffffffu, mine.zp, 0);
                const int j = lane < ep.nseg ? lane : 0;
                const float w_scale = __shfl_sync(0xffffffffu, mine.scale, 4 + j);
                const float o_scale = __shfl_sync(0xffffffffu, mine.scale, 8 + j), o_rcp = __shfl_sync(0xffffffffu, mine.rcp, 8 + j);
                const float o_zp = __shfl_sync(0xffffffffu, mine.zp, 8 + j), o_lo = __shfl_sync(0xffffffffu, mine.lo, 8 + j);
                const float o_hi = __shfl_sync(0xffffffffu, mine.hi, 8 + j);
                int exact = __shfl_sync(0xffffffffu, mine.exact, 8 + j);
                const float r_scale = __shfl_sync(0xffffffffu, mine.scale, 1), r_zp = __shfl_sync(0xffffffffu, mine.zp, 1);
                const float s2 = __shfl_sync(0xffffffffu, mine.scale, 2), r2 = __shfl_sync(0xffffffffu, mine.rcp, 2);
                const float z2 = __shfl_sync(0xffffffffu, mine.zp, 2), l2 = __shfl_sync(0xffffffffu, mine.lo, 2);
                const float h2 = __shfl_sync(0xffffffffu, mine.hi, 2);
                const int e2 = __shfl_sync(0xffffffffu, mine.exact, 2);
                const float s3 = __shfl_sync(0xffffffffu, mine.scale, 3), r3 = __shfl_sync(0xffffffffu, mine.rcp, 3);
                const float z3 = __shfl_sync(0xffffffffu, mine.zp, 3), l3 = __shfl_sync(0xffffffffu, mine.lo, 3);
                const float h3 = __shfl_sync(0xffffffffu, mine.hi, 3);
                const int e3 = __shfl_sync(0xffffffffu, mine.exact, 3);
                if (lane < ep.nseg && lane < kMaxSeg) {
                    float* sg = sgs + lane * kSegFloats;
                    sg[0] = __fmul_rn(a_scale, w_scale);
                    sg[1] = o_scale; sg[2] = o_rcp; sg[3] = o_lo - o_zp; sg[4] = o_hi - o_zp;
                    sg[5] = __fadd_rn(o_zp, 12582912.0f);
                    if (lnf) {
                        exact |= e2 | e3;
                        sg[7] = s2; sg[8] = r2; sg[9] = l2 - z2; sg[10] = h2 - z2;
                        sg[11] = r_scale; sg[12] = __fadd_rn(8388608.0f, r_zp);
                        sg[13] = s3; sg[14] = r3; sg[15] = l3 - z3; sg[16] = h3 - z3;
                        sg[17] = __fadd_rn(z3, 12582912.0f);
                    }
                    sg[6] = __int_as_float(exact);
                }
            }
            if (pending) {
                asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
                pending = false;
            }
            if (ATT && kind == 3) {
                const float* mask = ep.bias;                             // [sequences][128] additive mask or null
                for (int i = lane; i < BM; i += 32) smask[i] = mask != nullptr ? __ldg(mask + m0 + i) : 0.0f;
                for (int j = 0; j < S.tiles; ++j) {
                    mbar_wait(pempty_bar(pb), pphase ^ 1u);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pfull_bar(pb));
                    if (++pb == 2) { pb = 0; pphase ^= 1u; }
                }
            } else {
                for (int j = 0; j < S.tiles; ++j) {
                    const int64_t n0 = (int64_t)(crank * S.tiles + j) * bn;
                    mbar_wait(pempty_bar(pb), pphase ^ 1u);
                    float4* Pt = Pcol + pb * (BNF / 2);
                    for (int jp = lane; jp < bn / 2; jp += 32) {
                        const int64_t n = n0 + 2 * jp;
                        const float b0 = ep.bias != nullptr ? __ldg(ep.bias + n) : 0.0f, b1 = ep.bias != nullptr ? __ldg(ep.bias + n + 1) : 0.0f;
                        const int r0 = __ldg(ep.w_rowsum + n), r1 = __ldg(ep.w_rowsum + n + 1);
                        float4 gb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        if (lnf) gb = make_float4(__ldg(ep.ln_gamma + n), __ldg(ep.ln_gamma + n + 1), __ldg(ep.ln_beta + n), __ldg(ep.ln_beta + n + 1));
                        Pt[jp] = make_float4(b0, b1, __int_as_float(a_zp * r0), __int_as_float(a_zp * r1));
                        if (lnf) Pgb[jp] = gb;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pfull_bar(pb));
                    if (++pb == 2) { pb = 0; pphase ^= 1u; }
                }
            }
            __syncwarp();
            if (S.ovl == 1) cluster_sync_all();               // MID barrier of an overlapped pair
            if (lnf) cluster_sync_all();
            if (s + 1 < nst) {
                // the next stage's descriptor goes to the other shared slot BEFORE this warp arrives at the stage-end barrier
                // (every warp reads it after the barrier); arrive now, wait at the top of the next stage's tile loop: the next
                // stage's quantizers are resolved while the epilogue warps finish this stage
                copy_tail(s + 1);
                __syncwarp();
                asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
                pending = true;
            }
        }
    } else {
        // ===================== epilogue / softmax (warps 0..7) =====================
        const int quarter = warp & 3, half = warp >> 2;
        int acc = 0, pb = 0, e_ring = 0;
        uint32_t acc_phase = 0, pphase = 0, att_ph0 = 0, att_ph1 = 0;
        const int rl = quarter * 32 + lane;
        const int64_t row = m0 + rl;
        const bool row_ok = row < P.M;
        long long* tr = (P.trace != nullptr && threadIdx.x == 0) ? P.trace + (int64_t)blockIdx.x * nst * 4 : nullptr;
        for (int s = 0; s < nst; ++s) {
            const StageTail& S = tails[s & 1];
            const Args& ep = S.ep;
            const int kind = S.kind;
            const int bn = bn_of(kind);
            const int64_t N = S.N;
            const float* sgs = segp + (s & 1) * kMaxSeg * kSegFloats;
            if (tr != nullptr) tr[s * 4 + 0] = clock64();
            if (ATT && kind == 3) {
                attn::RowArgs ra;
                ra.inv_sqrt_d = 0.125f; ra.sqrt_d = 0.0f; ra.c_ctr = nullptr; ra.c_u8 = reinterpret_cast<unsigned char*>(ep.y_u8);
                for (int j = 0; j < S.tiles; ++j) {
                    const int h = (int)(crank * S.tiles) + j;
                    int slot = e_ring + j;
                    slot -= slot >= kRing ? kRing : 0;
                    mbar_wait(pfull_bar(pb), pphase);
                    const QP qs = load_qp(sgs), qp = load_qp(sgs + 8), qc = load_qp(sgs + 16);
                    const float sqk = sgs[24] * sgs[32], spv = qp.scale * sgs[40];
                    const bool exact = (qs.exact | qp.exact | qc.exact) != 0;
                    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BNF);
                    unsigned char* pP = base_ptr + slot * kStageBytes;
                    mbar_wait(tfull_bar(acc), acc_phase);
                    tc_fence_after();
                    if (tr != nullptr && j == 0) tr[s * 4 + 1] = clock64();
                    if (exact) attn::softmax_rows<false, false>(ra, qs, qp, sqk, trow, rl, half, smask, pP, xchg);
                    else attn::softmax_rows<true, false>(ra, qs, qp, sqk, trow, rl, half, smask, pP, xchg);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tensor core reads
                    tc_fence_before();
                    mbar_arrive(pready_bar(acc));
                    mbar_wait(oready_bar(acc), acc ? att_ph1 : att_ph0);
                    tc_fence_after();
                    const int64_t ooff = row * N + h * 64 + half * 32;
                    if (exact) attn::context_rows<false>(ra, qc, spv, trow, half, ooff, (int32_t)N);
                    else attn::context_rows<true>(ra, qc, spv, trow, half, ooff, (int32_t)N);
                    if (acc) att_ph1 ^= 1u; else att_ph0 ^= 1u;
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(tempty_bar(acc));
                        mbar_arrive(pempty_bar(pb));
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                    if (++pb == 2) { pb = 0; pphase ^= 1u; }
                }
                e_ring = (e_ring + S.tiles) % kRing;
            } else {
                for (int j = 0; j < S.tiles; ++j) {
                    const int64_t n0 = (int64_t)(crank * S.tiles + j) * bn;
                    uint32_t r0[8], r1[8];
                    if (kind == 2) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) r0[i] = r1[i] = 0u;
                        if (row_ok) {
                            const unsigned char* rrow = ep.res_u8 + row * N + n0 + half * 32;
                            ldg256(rrow, r0);
                            ldg256(rrow + 64, r1);
                        }
                    }
                    mbar_wait(pfull_bar(pb), pphase);
                    const float* sg = sgs + (int)(n0 / ep.seg_width) * kSegFloats;
                    const int exact = __float_as_int(sg[6]);
                    const uint32_t tmem_tile = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BNF);
                    const float4* Pt = Pcol + pb * (BNF / 2);
                    mbar_wait(tfull_bar(acc), acc_phase);
                    tc_fence_after();
                    if (tr != nullptr && j == 0) tr[s * 4 + 1] = clock64();
                    if (kind == 2) {
                        if (exact) epi_res_ln<BNL, false>(ep, Pt, Pgb, sg, part, xs, tmem_tile, half, quarter, lane, row, row_ok, n0, N, r0, r1);
                        else epi_res_ln<BNL, true>(ep, Pt, Pgb, sg, part, xs, tmem_tile, half, quarter, lane, row, row_ok, n0, N, r0, r1);
                    } else if (kind == 1) {
                        if (exact) epi_plain<BNF, 1, false, true>(ep, Pt, sg, tmem_tile, half, row, row_ok, n0, N);
                        else epi_plain<BNF, 1, true, true>(ep, Pt, sg, tmem_tile, half, row, row_ok, n0, N);
                    } else {
                        if (exact) epi_plain<BNL, 0, false, false>(ep, Pt, sg, tmem_tile, half, row, row_ok, n0, N);
                        else epi_plain<BNL, 0, true, false>(ep, Pt, sg, tmem_tile, half, row, row_ok, n0, N);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(tempty_bar(acc));
                        mbar_arrive(pempty_bar(pb));
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                    if (++pb == 2) { pb = 0; pphase ^= 1u; }
                    if (S.ovl == 1 && j == S.tiles - 2) {
                        // MID barrier of an overlapped pair: tiles 0 .. T - 2 of every member are in memory, the successor's
                        // main loop may read them (same fence protocol as a stage end)
                        if (P.gpu_fence) __threadfence();
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                        cluster_sync_all();
                    }
                }
                e_ring = (e_ring + S.tiles * (int)(S.K / 128)) % kRing;
            }
            if (tr != nullptr) tr[s * 4 + 2] = clock64();
            if (s + 1 < nst) {
                // the next stage reads this stage's panel through TMA (async proxy) issued by a thread of this cluster:
                // generic -> async proxy fence by every writer, then the cluster barrier's release / acquire pair.  The
                // gpu-scope fence in front of it is belt and braces: TQ_CHAIN_GPU_FENCE=0 drops it (300 forwards of
                // tools/stress_chain.py bit-identical without it, step -0.7 %) -- kept by default
                if (P.gpu_fence) __threadfence();
                asm volatile("fence.proxy.async.global;" ::: "memory");
                cluster_sync_all();
            }
            if (tr != nullptr) tr[s * 4 + 3] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                     : "memory");
}

}  // namespace chain

}  // namespace lean

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

// row-major [rows, cols] bf16 (or 8-bit) -> 2-D tensor map with a [box_rows, 128 bytes] box, 128 B swizzle
static int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int box_rows, bool i8 = false) {
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return TQ_EUNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * (i8 ? 1 : 2)};
    const cuuint32_t box[2] = {(cuuint32_t)(i8 ? 2 * BK : BK), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, i8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? TQ_OK : TQ_EINVAL;
}

template <int BN, int CTAS, bool LNF = false, bool I8 = false>
static int launch(const void* a, const void* w, int64_t M, int64_t N, int64_t K, int k_split, const EpiArgs& ep,
                  cudaStream_t st) {
    using C = Cfg<BN, CTAS>;
    static_assert(C::kSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert(2 * BN <= kTmemCols, "TMEM budget");
    CUtensorMap map_a, map_w;
    if (int e = make_map(&map_a, a, M, K * k_split, BM, I8)) return e;
    if (int e = make_map(&map_w, w, N, K, C::kBRows, I8)) return e;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linear_qdq_kernel<BN, CTAS, LNF, I8>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int64_t tiles = ((M + BM * CTAS - 1) / (BM * CTAS)) * ((N + BN - 1) / BN);
    const int64_t slots = sm_count() / CTAS;                       // CTAs (pairs) resident at once
    int grid = (int)(tiles < slots ? tiles : slots) * CTAS;
    int cluster = CTAS;
    if (LNF) {                                                     // one tile per CTA, a cluster per 128-row panel
        cluster = (int)(N / BN);
        grid = (int)tiles;
    }
    int ring = C::kStages;
    if (const char* e = getenv("TQ_LINEAR_STAGES")) {              // tuning aid (tools/mainloop_probe.py)
        const int f = atoi(e);
        if (f >= 1 && f < ring) ring = f;
    }
    return launch_pdl(linear_qdq_kernel<BN, CTAS, LNF, I8>, dim3(grid), dim3(kThreads), C::kSmemBytes, st, cluster, map_a,
                      map_w, M, N, K, k_split, ring, ep);
}

// Tile shape.  Cycle model from the clock64 timelines of the kernel (tools/mainloop_probe.py,
// tools/trace_linear.py, tools/sweep_linear.py; every SM streaming operands):
//   * one 64-wide k-block takes ~650 (bn = 64) ... ~820 (bn = 256) cycles for a single CTA and ~715
//     for a CTA pair whatever bn -- a lone CTA needs ~530, the MMA itself 2 * bn;
//   * first operands land ~3.5 k cycles after launch (pair: ~6 k, cluster barrier + peer hand-shake);
//   * the epilogue costs ~0.9 k cycles per 16-column slice per warp (1.05 k with the residual branch,
//     1.75 k with GELU).
// With double-buffered TMEM a CTA that owns t tiles takes  setup + main + (t-1) * max(main, epi) + epi.
struct TileShape {
    int bn, ctas;
};
static TileShape pick_tile(int64_t M, int64_t N, int64_t K, int k_split, int act_fn, bool has_res, bool i8 = false) {
    const int cands[5] = {256, 192, 128, 96, 64};
    int force_bn = 0, force_ctas = 0;
    if (const char* e = getenv("TQ_LINEAR_BN")) force_bn = atoi(e);          // tuning aids
    if (const char* e = getenv("TQ_LINEAR_CTAS")) force_ctas = atoi(e);
    if (force_ctas != 2) force_ctas = 1;      // CTA pairs measured no faster on any BERT GEMM: opt-in only
    const int sms = sm_count();
    TileShape best = {64, 1};
    double best_cost = 1e30;
    for (int ctas = 1; ctas <= 2; ++ctas) {
        if (force_ctas != 0 && ctas != force_ctas) continue;
        for (int i = 0; i < 5; ++i) {
            const int bn = cands[i];
            if (force_bn != 0 && bn != force_bn && !(force_bn > N && bn == 64)) continue;
            if (ctas == 2 && (bn < 128 || M < 2 * BM)) continue;
            if (bn > 64 && N < bn) continue;             // TMA box must fit inside the weight matrix
            const int64_t tiles = ((M + BM * ctas - 1) / (BM * ctas)) * ((N + bn - 1) / bn);
            const int64_t slots = sms / ctas;
            const double per_cta = (double)((tiles + slots - 1) / slots);
            double kb_c = ctas == 2 ? 715.0 : 595.0 + 0.9 * bn;
            if (kb_c < bn * 2.05) kb_c = bn * 2.05;
            const double main_c = (double)(K / (i8 ? 2 * BK : BK)) * k_split * kb_c;
            const double epi_c = (bn / 16.0 / 3.0) * (has_res ? 1050.0 : (act_fn == 1 || act_fn == 3 ? 1750.0 : 900.0));
            const double cost = (ctas == 2 ? 6000.0 : 3500.0) + main_c + (per_cta - 1.0) * (main_c > epi_c ? main_c : epi_c) + epi_c;
            if (cost < best_cost) {
                best_cost = cost;
                best.bn = bn;
                best.ctas = ctas;
            }
        }
    }
    return best;
}


template <int BN, int ACT, int MODE, bool OUT8>
static int launch_lean(const void* a, const void* w, int64_t M, int64_t N, int64_t K, const lean::Args& ep, cudaStream_t st) {
    using C = lean::Cfg<BN, MODE>;
    constexpr bool LNF = MODE == 1;
    static_assert(C::kSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert(2 * BN <= kTmemCols && BN % 64 == 0, "TMEM budget / slice width");
    CUtensorMap map_a, map_w;
    if (int e = make_map(&map_a, a, M, K, BM, true)) return e;
    if (int e = make_map(&map_w, w, N, K, BN, true)) return e;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(lean::linear_lean_kernel<BN, ACT, MODE, OUT8>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int64_t tiles = ((M + BM - 1) / BM) * (N / BN);
    int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    int cluster = 1;
    if (LNF) {
        cluster = (int)(N / BN);
        grid = (int)tiles;
    }
    int ring = C::kStages;
    if (const char* e = getenv("TQ_LINEAR_STAGES")) {
        const int f = atoi(e);
        if (f >= 1 && f < ring) ring = f;
    }
    return launch_pdl(lean::linear_lean_kernel<BN, ACT, MODE, OUT8>, dim3(grid), dim3(lean::kThreads), C::kSmemBytes, st, cluster,
                      map_a, map_w, M, N, K, ring, ep);
}


template <int ACT, bool LNF, bool OUT8>
static int launch_peg(const void* a, const void* w, int64_t M, int64_t N, int64_t K, const lean::peg::Args& ep, cudaStream_t st) {
    using C = lean::peg::Cfg<LNF>;
    static_assert(C::kSmemBytes <= 227 * 1024, "shared memory budget");
    CUtensorMap map_a, map_w;
    if (int e = make_map(&map_a, a, M, K, BM, true)) return e;
    if (int e = make_map(&map_w, w, N, K, lean::peg::BN, true)) return e;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(lean::peg::linear_peg_kernel<ACT, LNF, OUT8>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int64_t tiles = ((M + BM - 1) / BM) * (N / lean::peg::BN);
    int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    int cluster = 1;
    if (LNF) {
        cluster = (int)(N / lean::peg::BN);
        grid = (int)tiles;
    }
    return launch_pdl(lean::peg::linear_peg_kernel<ACT, LNF, OUT8>, dim3(grid), dim3(lean::kThreads), C::kSmemBytes, st, cluster,
                      map_a, map_w, M, N, K, (int)C::kStages, ep);
}


static inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

struct ChainPlan {
    lean::chain::StageDesc* d_st;
    int32_t n, csize;
    int64_t M;
    bool attention;                       // any kind-3 stage
};
constexpr int kMaxPlanStages = 256;

static int chain_plan_create(const tq_chain_stage* stages, int32_t n, int64_t M, void** out) {
    using namespace lean::chain;
    if (stages == nullptr || out == nullptr || n < 1 || n > kMaxPlanStages || M < 1) return TQ_EINVAL;
    *out = nullptr;
    int64_t csize = 0;
    for (int i = 0; i < n; ++i)
        if (stages[i].kind == 2) {
            const int64_t c = stages[i].N / BNL;
            if (stages[i].N % BNL != 0 || c < 1 || c > 8 || (csize != 0 && c != csize)) return TQ_EUNSUPPORTED;
            csize = c;
        }
    if (csize == 0) return TQ_EUNSUPPORTED;            // (a chain without a LayerNorm stage: use the single kernels)
    std::vector<StageDesc> host((size_t)n);
    memset(host.data(), 0, sizeof(StageDesc) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        const tq_chain_stage& g = stages[i];
        StageDesc& S = host[(size_t)i];
        if (g.kind < 0 || g.kind > 3 || g.a_i8 == nullptr || g.out == nullptr) return TQ_EINVAL;
        lean::Args& a = S.t.ep;
        a.a_q = g.a_q; a.w_q = g.w_q; a.out_q = g.out_q; a.res_q = g.res_q; a.out2_q = g.out2_q; a.ln_q = g.ln_q;
        a.bias = g.bias;
        a.trace = nullptr; a.trace_tiles = nullptr;
        S.t.N = g.N;
        S.t.K = g.K;
        S.t.kind = g.kind;
        if (g.kind == 3) {
            // attention of K heads of 64 dimensions over [M, 3 N] bf16 centred grids, one sequence = one 128-row panel
            const tq_qspec* need[6] = {&g.a_q, &g.w_q, &g.res_q, &g.out2_q, &g.ln_q, &g.out_q};
            for (int k = 0; k < 6; ++k) {
                if (int e = tq::check_qspec(*need[k])) return e;
                if (need[k]->n_bits > 8) return TQ_EUNSUPPORTED;
            }
            if (g.K < 1 || g.N != g.K * 64 || g.K % csize != 0 || M % BM != 0 || g.N != csize * BNL) return TQ_EUNSUPPORTED;
            if (!tq::aligned16(g.a_i8) || !aligned32(g.out) || (g.N & 31) != 0) return TQ_EALIGN;
            if (int e = make_map(&S.map_a, g.a_i8, M, 3 * g.N, BM, false)) return e;
            S.t.tiles = (int32_t)(g.K / csize);
            a.y_u8 = g.out;
            a.ldc = g.N;
            continue;
        }
        if (g.w_i8 == nullptr || g.w_rowsum == nullptr) return TQ_EINVAL;
        if (g.N < 1 || g.K < 1 || g.K % 128 != 0 || g.nseg < 1 || g.nseg > lean::kMaxSeg || g.N % g.nseg != 0) return TQ_EUNSUPPORTED;
        const int bn = g.kind == 1 ? BNF : BNL;
        if (g.N % bn != 0 || (g.N / bn) % csize != 0 || (g.N / g.nseg) % bn != 0) return TQ_EUNSUPPORTED;
        if (g.kind == 2 && (g.N / bn != csize || g.res_i8 == nullptr || g.ln_gamma_q == nullptr || g.ln_beta == nullptr || g.nseg != 1)) return TQ_EINVAL;
        const tq_qspec* need[3] = {&g.a_q, &g.w_q, &g.out_q};
        for (int k = 0; k < 3; ++k) {
            if (int e = tq::check_qspec(*need[k])) return e;
            if (need[k]->n_bits > 8) return TQ_EUNSUPPORTED;
        }
        if (g.kind == 2) {
            const tq_qspec* more[3] = {&g.res_q, &g.out2_q, &g.ln_q};
            for (int k = 0; k < 3; ++k) {
                if (int e = tq::check_qspec(*more[k])) return e;
                if (more[k]->n_bits > 8) return TQ_EUNSUPPORTED;
            }
        }
        if (!tq::aligned16(g.a_i8) || !tq::aligned16(g.w_i8) || !aligned32(g.out) || (g.res_i8 != nullptr && !aligned32(g.res_i8)) || (g.N & 31) != 0) return TQ_EALIGN;
        if (int e = make_map(&S.map_a, g.a_i8, M, g.K, BM, true)) return e;
        if (int e = make_map(&S.map_w, g.w_i8, g.N, g.K, bn, true)) return e;
        S.t.tiles = (int32_t)((g.N / bn) / csize);
        {   // operand signedness of the kind::i8 products (symmetric quantizers carry a device-side flag): resolved once, here
            auto is_s8 = [](const tq_qspec& q, uint32_t& out) -> int {
                out = 0u;
                if (q.zero_float != nullptr || q.is_signed == nullptr) return TQ_OK;
                unsigned char flag = 0;
                if (cudaMemcpy(&flag, q.is_signed, 1, cudaMemcpyDeviceToHost) != cudaSuccess) return TQ_EINVAL;
                out = flag ? 1u : 0u;
                return TQ_OK;
            };
            uint32_t a_s8 = 0u, w_s8 = 0u;
            if (int e = is_s8(g.a_q, a_s8)) return e;
            if (int e = is_s8(g.w_q, w_s8)) return e;
            S.t.idesc = (2u << 4) | (a_s8 << 7) | (w_s8 << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        }
        a.w_rowsum = g.w_rowsum;
        a.seg_width = g.N / g.nseg; a.nseg = g.nseg; a.ldc = g.N;
        a.y_u8 = g.kind == 0 ? nullptr : g.out;
        a.y_ctr = g.kind == 0 ? reinterpret_cast<__nv_bfloat16*>(g.out) : nullptr;
        a.res_u8 = reinterpret_cast<const unsigned char*>(g.res_i8);
        a.ln_gamma = g.ln_gamma_q; a.ln_beta = g.ln_beta; a.ln_eps = g.ln_eps;
    }
    // overlapped pairs: a GELU stage with >= 2 tiles per CTA followed by the LayerNorm stage that consumes its output
    static const bool ovl_on = [] { const char* e = getenv("TQ_CHAIN_OVERLAP"); return e == nullptr || e[0] != '0'; }();
    for (int i = 0; ovl_on && i + 1 < n; ++i) {
        StageTail &b = host[(size_t)i].t, &c = host[(size_t)i + 1].t;
        if (b.kind == 1 && c.kind == 2 && c.K == b.N && b.tiles >= 2 && stages[i + 1].a_i8 == stages[i].out &&
            (int64_t)csize * b.tiles * 2 == c.K / 128) {
            b.ovl = 1;
            c.ovl = 2;
        }
    }
    StageDesc* dev = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&dev), sizeof(StageDesc) * (size_t)n);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(dev, host.data(), sizeof(StageDesc) * (size_t)n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dev); return (int)e; }
    bool att = false;
    for (int i = 0; i < n; ++i) att = att || stages[i].kind == 3;
    ChainPlan* plan = new ChainPlan{dev, n, (int32_t)csize, M, att};
    *out = plan;
    return TQ_OK;
}

static int chain_plan_run(const ChainPlan* plan, cudaStream_t st) {
    using namespace lean::chain;
    if (plan == nullptr || plan->d_st == nullptr) return TQ_EINVAL;
    Params P;
    P.st = plan->d_st;
    P.M = plan->M;
    P.n = plan->n;
    P.trace = nullptr;
    static const int gpu_fence = [] { const char* e = getenv("TQ_CHAIN_GPU_FENCE"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
    P.gpu_fence = gpu_fence;
    if (const char* env = getenv("TQ_LINEAR_TRACE_CHAIN"))         // tools/trace_chain.py: device pointer of a zeroed int64 buffer
        P.trace = reinterpret_cast<long long*>(strtoull(env, nullptr, 0));
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linear_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(linear_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int64_t panels = (plan->M + BM - 1) / BM;
    const dim3 grid((unsigned)(panels * plan->csize)), block(lean::kThreads);
    if (plan->attention) return launch_pdl(linear_chain_kernel<true>, grid, block, kSmemBytes, st, plan->csize, P);
    return launch_pdl(linear_chain_kernel<false>, grid, block, kSmemBytes, st, plan->csize, P);
}

static void lean_trace(lean::Args& ep) {
    ep.trace = nullptr;
    ep.trace_tiles = nullptr;
    if (const char* e = getenv("TQ_LINEAR_TRACE_PTR")) ep.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
    if (const char* e = getenv("TQ_LINEAR_TRACE_TILES")) ep.trace_tiles = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
}
static bool lean_enabled() {          // TQ_LINEAR_LEAN=0: keep the general kernels (tools / A-B timing)
    const char* e = getenv("TQ_LINEAR_LEAN");
    return !(e != nullptr && e[0] == '0');
}

}  // namespace gemm
}  // namespace tq

extern "C" {

size_t tq_linear_workspace_bytes(int64_t, int64_t, int64_t) { return 256; }

static int linear_impl(const void* a_ctr_bf16, const void* w_ctr_bf16, const float* bias, float* y,
                       void* y_ctr_bf16, int64_t M, int64_t N, int64_t K, int32_t k_split, tq_qspec a_q,
                       tq_qspec w_q, int64_t w_q_params, int32_t act_fn, tq_qspec out_q, int64_t out_q_params,
                       const void* res_ctr_bf16, tq_qspec res_q, tq_qspec out2_q, int64_t out2_q_params,
                       float* tile_minmax, void* ws, size_t ws_bytes, void* stream, const float* ln_gamma = nullptr,
                       const float* ln_beta = nullptr, float ln_eps = 0.0f, const tq_qspec* ln_q = nullptr,
                       bool i8 = false, const int32_t* w_rowsum = nullptr, void* y_u8 = nullptr) {
    using namespace tq::gemm;
    if (i8) {
        // 8-bit operand mode: A / W / residual are x_int bytes; K in 128-element blocks; rows 16-byte aligned
        if (w_rowsum == nullptr || a_q.delta == nullptr || w_q.delta == nullptr) return TQ_EINVAL;
        if (a_q.n_bits > 8 || w_q.n_bits > 8 || (res_ctr_bf16 != nullptr && res_q.n_bits > 8)) return TQ_EUNSUPPORTED;
        if (k_split != 1 || K % (2 * BK) != 0 || N % 16 != 0) return TQ_EUNSUPPORTED;
        if (y_u8 != nullptr && (out_q.delta == nullptr || !tq::aligned16(y_u8))) return TQ_EINVAL;
    }
    if (res_ctr_bf16 != nullptr) {
        if (out_q.delta == nullptr || res_q.delta == nullptr) return TQ_EINVAL;
        if (int e = tq::check_qspec(out2_q)) return e;
        if (int e = tq::check_qspec(res_q)) return e;
        if (out2_q_params != 1 && out2_q_params != N) return TQ_EINVAL;
        if (!tq::aligned16(res_ctr_bf16)) return TQ_EALIGN;
    }
    if (a_ctr_bf16 == nullptr || w_ctr_bf16 == nullptr || (y == nullptr && y_ctr_bf16 == nullptr && y_u8 == nullptr)) return TQ_EINVAL;
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
    if (ws == nullptr) {
        if (const char* e = getenv("TQ_LINEAR_TRACE_PTR")) {     // tools: device buffer of >= 16 int64 for the timeline
            ws = reinterpret_cast<void*>(strtoull(e, nullptr, 0));
            ws_bytes = ws != nullptr ? 16 * sizeof(long long) : 0;
        }
    }
    ep.trace = (ws != nullptr && ws_bytes >= 16 * sizeof(long long)) ? reinterpret_cast<long long*>(ws) : nullptr;
    ep.trace_all = (ws != nullptr && ws_bytes >= (16 + 4 * 160) * sizeof(long long))
                       ? reinterpret_cast<long long*>(ws) + 16
                       : nullptr;
    ep.trace_tiles = nullptr;
    if (const char* e = getenv("TQ_LINEAR_TRACE_TILES"))          // tools/trace_tiles.py: device buffer of 64 int64
        ep.trace_tiles = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
    ep.res_ctr = reinterpret_cast<const __nv_bfloat16*>(res_ctr_bf16);
    ep.res_q = res_q;
    ep.out2_q = out2_q;
    ep.out2_params = out2_q_params;
    cudaStream_t st = (cudaStream_t)stream;
    ep.y_u8 = y_u8;
    ep.w_rowsum = w_rowsum;
    ep.ln_gamma = ln_gamma;
    ep.ln_beta = ln_beta;
    ep.ln_eps = ln_eps;
    ep.ln_q = ln_q != nullptr ? *ln_q : out_q;
    if (ln_gamma != nullptr) {
        // fused LayerNorm: clusters of N / bn CTAs, one tile each (see epi_tile_res_ln)
        if (res_ctr_bf16 == nullptr || ln_beta == nullptr || ln_q == nullptr) return TQ_EINVAL;
        if (int e = tq::check_qspec(*ln_q)) return e;
        if (out_q_params != 1 || out2_q_params != 1 || k_split != 1 || (N & 15) != 0) return TQ_EUNSUPPORTED;
        const int cands[3] = {256, 192, 128};
        int best = 0;
        double best_cost = 1e30;
        int force_bn = 0;
        if (const char* e = getenv("TQ_LINEAR_BN")) force_bn = atoi(e);
        for (int i = 0; i < 3; ++i) {
            const int bn = cands[i];
            if (N % bn != 0 || N / bn > 8) continue;
            if (force_bn != 0 && force_bn != bn && N % force_bn == 0 && N / force_bn <= 8 && force_bn >= 128) continue;
            const int64_t tiles = ((M + BM - 1) / BM) * (N / bn);
            const double waves = (double)((tiles + tq::sm_count() - 1) / tq::sm_count());
            const double cost = waves * ((double)(K / (i8 ? 2 * BK : BK)) * (595.0 + 0.9 * bn) + (bn / 16.0 / 3.0) * 1600.0 + 4500.0);
            if (cost < best_cost) {
                best_cost = cost;
                best = bn;
            }
        }
        if (i8 && lean_enabled() && y == nullptr && y_u8 != nullptr && w_q_params == 1 && out2_q.n_bits <= 8 &&
            out_q.n_bits <= 8 && ln_q->n_bits <= 8 && (N & 31) == 0 && aligned32(y_u8) && aligned32(res_ctr_bf16) &&
            (y_ctr_bf16 == nullptr || aligned32(y_ctr_bf16)) && (best == 256 || best == 192 || best == 128)) {
            // lean int8 kernel: same arithmetic, parameter warp + 32-column epilogue slices (see namespace lean)
            lean::Args la;
            la.bias = bias; la.w_rowsum = w_rowsum; la.a_q = a_q; la.w_q = w_q; la.out_q = out_q;
            la.seg_width = N; la.nseg = 1;
            la.y_u8 = y_u8; la.y_ctr = reinterpret_cast<__nv_bfloat16*>(y_ctr_bf16);
            la.res_u8 = reinterpret_cast<const unsigned char*>(res_ctr_bf16);
            la.res_q = res_q; la.out2_q = out2_q; la.ln_q = *ln_q;
            la.ln_gamma = ln_gamma; la.ln_beta = ln_beta; la.ln_eps = ln_eps; la.ldc = N;
            lean_trace(la);
            switch (best) {
                case 256: return launch_lean<256, 0, 1, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, la, st);
                case 192: return launch_lean<192, 0, 1, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, la, st);
                default: return launch_lean<128, 0, 1, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, la, st);
            }
        }
        if (i8) {
            switch (best) {
                case 256: return launch<256, 1, true, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
                case 192: return launch<192, 1, true, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
                case 128: return launch<128, 1, true, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
                default: return TQ_EUNSUPPORTED;
            }
        }
        switch (best) {
            case 256: return launch<256, 1, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
            case 192: return launch<192, 1, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
            case 128: return launch<128, 1, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
            default: return TQ_EUNSUPPORTED;
        }
    }
    const TileShape ts = pick_tile(M, N, K, k_split, act_fn, res_ctr_bf16 != nullptr, i8);
    if (i8) {
        switch (ts.bn) {
            case 256: return launch<256, 1, false, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
            case 192: return launch<192, 1, false, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
            case 128: return launch<128, 1, false, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
            case 96: return launch<96, 1, false, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
            default: return launch<64, 1, false, true>(a_ctr_bf16, w_ctr_bf16, M, N, K, 1, ep, st);
        }
    }
    if (ts.ctas == 2) {
        switch (ts.bn) {
            case 256: return launch<256, 2>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
            case 192: return launch<192, 2>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
            default: return launch<128, 2>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
        }
    }
    switch (ts.bn) {
        case 256: return launch<256, 1>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
        case 192: return launch<192, 1>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
        case 128: return launch<128, 1>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
        case 96: return launch<96, 1>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
        default: return launch<64, 1>(a_ctr_bf16, w_ctr_bf16, M, N, K, k_split, ep, st);
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

int tq_linear_res_ln_qdq_bf16(const void* a_ctr_bf16, const void* w_ctr_bf16, const float* bias, float* z,
                              void* z_ctr_bf16, int64_t M, int64_t N, int64_t K, tq_qspec a_q, tq_qspec w_q,
                              int64_t w_q_params, tq_qspec out_q, const void* res_ctr_bf16, tq_qspec res_q,
                              tq_qspec out2_q, const float* ln_gamma_q, const float* ln_beta, float ln_eps,
                              tq_qspec ln_q, void* stream) {
    if (res_ctr_bf16 == nullptr || ln_gamma_q == nullptr || ln_beta == nullptr) return TQ_EINVAL;
    return linear_impl(a_ctr_bf16, w_ctr_bf16, bias, z, z_ctr_bf16, M, N, K, 1, a_q, w_q, w_q_params, 0, out_q, 1,
                       res_ctr_bf16, res_q, out2_q, 1, nullptr, nullptr, 0, stream, ln_gamma_q, ln_beta, ln_eps, &ln_q);
}

int tq_linear_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_rowsum, const float* bias, float* y,
                     void* y_ctr_bf16, void* y_i8, int64_t M, int64_t N, int64_t K, tq_qspec a_q, tq_qspec w_q,
                     int64_t w_q_params, int32_t act_fn, tq_qspec out_q, int64_t out_q_params, void* stream) {
    tq_qspec none;
    none.delta = nullptr; none.zero_float = nullptr; none.is_signed = nullptr;
    none.n_bits = 8; none.log_domain = 0; none.eps = 1e-8f;
    return linear_impl(a_i8, w_i8, bias, y, y_ctr_bf16, M, N, K, 1, a_q, w_q, w_q_params, act_fn, out_q, out_q_params,
                       nullptr, none, none, 1, nullptr, nullptr, 0, stream, nullptr, nullptr, 0.0f, nullptr, true, w_rowsum,
                       y_i8);
}

int tq_linear_qdq_bf16_o8(const void* a_ctr_bf16, const void* w_ctr_bf16, const float* bias, void* y_i8, int64_t M,
                          int64_t N, int64_t K, tq_qspec a_q, tq_qspec w_q, int64_t w_q_params, int32_t act_fn,
                          tq_qspec out_q, int64_t out_q_params, void* stream) {
    tq_qspec none;
    none.delta = nullptr; none.zero_float = nullptr; none.is_signed = nullptr;
    none.n_bits = 8; none.log_domain = 0; none.eps = 1e-8f;
    if (y_i8 == nullptr || out_q.delta == nullptr || out_q.n_bits > 8 || N % 16 != 0 || act_fn > 1) return TQ_EUNSUPPORTED;
    return linear_impl(a_ctr_bf16, w_ctr_bf16, bias, nullptr, nullptr, M, N, K, 1, a_q, w_q, w_q_params, act_fn, out_q,
                       out_q_params, nullptr, none, none, 1, nullptr, nullptr, 0, stream, nullptr, nullptr, 0.0f, nullptr,
                       false, nullptr, y_i8);
}

int tq_linear_res_ln_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_rowsum, const float* bias, float* z,
                            void* z_ctr_bf16, void* z_i8, int64_t M, int64_t N, int64_t K, tq_qspec a_q, tq_qspec w_q,
                            int64_t w_q_params, tq_qspec out_q, const void* res_i8, tq_qspec res_q, tq_qspec out2_q,
                            const float* ln_gamma_q, const float* ln_beta, float ln_eps, tq_qspec ln_q, void* stream) {
    if (res_i8 == nullptr || ln_gamma_q == nullptr || ln_beta == nullptr) return TQ_EINVAL;
    return linear_impl(a_i8, w_i8, bias, z, z_ctr_bf16, M, N, K, 1, a_q, w_q, w_q_params, 0, out_q, 1, res_i8, res_q,
                       out2_q, 1, nullptr, nullptr, 0, stream, ln_gamma_q, ln_beta, ln_eps, &ln_q, true, w_rowsum, z_i8);
}

int tq_linear_seg_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_rowsum, const float* bias, void* y_ctr_bf16,
                         void* y_i8, int64_t M, int64_t N, int64_t K, tq_qspec a_q, tq_qspec w_q, tq_qspec out_q,
                         int32_t nseg, int32_t act_fn, int64_t ldc, void* stream) {
    using namespace tq::gemm;
    if (a_i8 == nullptr || w_i8 == nullptr || w_rowsum == nullptr || ((y_ctr_bf16 == nullptr) == (y_i8 == nullptr))) return TQ_EINVAL;
    if (M < 1 || N < 1 || K < 1 || nseg < 1 || nseg > lean::kMaxSeg || N % nseg != 0) return TQ_EINVAL;
    if (act_fn < 0 || act_fn > 2) return TQ_EUNSUPPORTED;
    if (int e = tq::check_qspec(a_q)) return e;
    if (int e = tq::check_qspec(w_q)) return e;
    if (int e = tq::check_qspec(out_q)) return e;
    if (a_q.n_bits > 8 || w_q.n_bits > 8 || out_q.n_bits > 8 || K % 128 != 0) return TQ_EUNSUPPORTED;
    if (!tq::aligned16(a_i8) || !tq::aligned16(w_i8)) return TQ_EALIGN;
    if (ldc == 0) ldc = N;
    void* out = y_i8 != nullptr ? y_i8 : y_ctr_bf16;
    if (ldc < N || !aligned32(out) || (N & 31) != 0 || (ldc & 31) != 0) return TQ_EALIGN;
    const int64_t seg = N / nseg;
    const int bn = seg % 256 == 0 ? 256 : (seg % 192 == 0 ? 192 : (seg % 128 == 0 ? 128 : 0));
    if (bn == 0) return TQ_EUNSUPPORTED;
    lean::Args la = {};
    la.bias = bias; la.w_rowsum = w_rowsum; la.a_q = a_q; la.w_q = w_q; la.out_q = out_q;
    la.seg_width = seg; la.nseg = nseg; la.ldc = ldc;
    la.y_u8 = y_i8; la.y_ctr = reinterpret_cast<__nv_bfloat16*>(y_ctr_bf16);
    lean_trace(la);
    cudaStream_t st = (cudaStream_t)stream;
    const bool o8 = y_i8 != nullptr;
#define TQ_LEAN_ACT(BN_, ACT_)                                                                                     \
    return o8 ? launch_lean<BN_, ACT_, 0, true>(a_i8, w_i8, M, N, K, la, st)                                        \
              : launch_lean<BN_, ACT_, 0, false>(a_i8, w_i8, M, N, K, la, st);
#define TQ_LEAN_CASE(BN_)                                                                                          \
    if (bn == BN_) {                                                                                               \
        if (act_fn == 1) { TQ_LEAN_ACT(BN_, 1) }                                                                   \
        if (act_fn == 2) { TQ_LEAN_ACT(BN_, 2) }                                                                   \
        TQ_LEAN_ACT(BN_, 0)                                                                                        \
    }
    TQ_LEAN_CASE(256)
    TQ_LEAN_CASE(192)
    TQ_LEAN_CASE(128)
#undef TQ_LEAN_CASE
#undef TQ_LEAN_ACT
    return TQ_EUNSUPPORTED;
}

int tq_linear_nonorm_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_rowsum, const float* bias, void* z_i8,
                            int64_t M, int64_t N, int64_t K, tq_qspec a_q, tq_qspec w_q, tq_qspec out_q, const void* res_i8,
                            tq_qspec res_q, tq_qspec out2_q, const float* nn_weight_q, const float* nn_bias_q, tq_qspec nn_q,
                            int64_t ldc, void* stream) {
    using namespace tq::gemm;
    if (a_i8 == nullptr || w_i8 == nullptr || w_rowsum == nullptr || z_i8 == nullptr || nn_weight_q == nullptr || nn_bias_q == nullptr) return TQ_EINVAL;
    if (M < 1 || N < 1 || K < 1) return TQ_EINVAL;
    const tq_qspec* need[4] = {&a_q, &w_q, &out_q, &nn_q};
    for (int i = 0; i < 4; ++i) {
        if (int e = tq::check_qspec(*need[i])) return e;
        if (need[i]->n_bits > 8) return TQ_EUNSUPPORTED;
    }
    if (res_i8 != nullptr) {
        if (int e = tq::check_qspec(res_q)) return e;
        if (int e = tq::check_qspec(out2_q)) return e;
        if (res_q.n_bits > 8 || out2_q.n_bits > 8) return TQ_EUNSUPPORTED;
        if (!aligned32(res_i8)) return TQ_EALIGN;
    }
    if (K % 128 != 0) return TQ_EUNSUPPORTED;
    if (ldc == 0) ldc = N;
    if (ldc < N || !aligned32(z_i8) || (N & 31) != 0 || (ldc & 31) != 0 || !tq::aligned16(a_i8) || !tq::aligned16(w_i8)) return TQ_EALIGN;
    const int bn = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 0);
    if (bn == 0) return TQ_EUNSUPPORTED;
    lean::Args la = {};
    la.bias = bias; la.w_rowsum = w_rowsum; la.a_q = a_q; la.w_q = w_q; la.out_q = out_q;
    la.seg_width = N; la.nseg = 1; la.ldc = ldc;
    la.y_u8 = z_i8; la.y_ctr = nullptr;
    la.res_u8 = reinterpret_cast<const unsigned char*>(res_i8);
    la.res_q = res_q; la.out2_q = out2_q; la.ln_q = nn_q;
    la.ln_gamma = nn_weight_q; la.ln_beta = nn_bias_q; la.ln_eps = 0.0f;
    lean_trace(la);
    cudaStream_t st = (cudaStream_t)stream;
    if (res_i8 != nullptr)
        return bn == 256 ? launch_lean<256, 0, 3, true>(a_i8, w_i8, M, N, K, la, st) : launch_lean<128, 0, 3, true>(a_i8, w_i8, M, N, K, la, st);
    return bn == 256 ? launch_lean<256, 0, 2, true>(a_i8, w_i8, M, N, K, la, st) : launch_lean<128, 0, 2, true>(a_i8, w_i8, M, N, K, la, st);
}

static int peg_common_checks(const void* a_i8, const void* w_i8, const int32_t* w_grp_rowsum, int64_t M, int64_t N, int64_t K,
                             const tq_qspec& a_q, int32_t a_groups, const tq_qspec& w_q, int32_t w_params, const tq_qspec& out_q,
                             int32_t out_params, int64_t seg_width) {
    using namespace tq::gemm;
    if (a_i8 == nullptr || w_i8 == nullptr || w_grp_rowsum == nullptr || M < 1 || N < 1 || K < 1) return TQ_EINVAL;
    if (a_groups < 1 || a_groups > lean::peg::kMaxGroups || seg_width < 1) return TQ_EINVAL;
    if (int e = tq::check_qspec(a_q)) return e;
    if (int e = tq::check_qspec(w_q)) return e;
    if (int e = tq::check_qspec(out_q)) return e;
    if (a_q.zero_float == nullptr && a_q.is_signed == nullptr) return TQ_EINVAL;
    if (a_q.n_bits > 8 || w_q.n_bits > 8 || out_q.n_bits > 8) return TQ_EUNSUPPORTED;
    if (N % lean::peg::BN != 0 || seg_width % lean::peg::BN != 0 || N % seg_width != 0 || K % (128 * (int64_t)a_groups) != 0) return TQ_EUNSUPPORTED;
    const int64_t nseg = N / seg_width;
    if ((w_params != 1 && w_params != nseg) || (out_params != 1 && out_params != nseg)) return TQ_EINVAL;
    if (!tq::aligned16(a_i8) || !tq::aligned16(w_i8)) return TQ_EALIGN;
    return TQ_OK;
}

int tq_linear_peg_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_grp_rowsum, const float* bias, void* y_ctr_bf16,
                         void* y_i8, int64_t M, int64_t N, int64_t K, tq_qspec a_q, int32_t a_groups, tq_qspec w_q,
                         int32_t w_params, tq_qspec out_q, int32_t out_params, int64_t seg_width, int32_t act_fn, void* stream) {
    using namespace tq::gemm;
    if ((y_ctr_bf16 == nullptr) == (y_i8 == nullptr)) return TQ_EINVAL;
    if (act_fn < 0 || act_fn > 1) return TQ_EUNSUPPORTED;
    if (int e = peg_common_checks(a_i8, w_i8, w_grp_rowsum, M, N, K, a_q, a_groups, w_q, w_params, out_q, out_params, seg_width)) return e;
    if (!aligned32(y_i8 != nullptr ? y_i8 : y_ctr_bf16)) return TQ_EALIGN;
    lean::peg::Args pa = {};
    if (const char* e = getenv("TQ_LINEAR_TRACE_PTR")) pa.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
    if (const char* e = getenv("TQ_LINEAR_TRACE_TILES")) pa.trace_tiles = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
    pa.bias = bias; pa.w_grp_rowsum = w_grp_rowsum; pa.a_q = a_q; pa.a_groups = a_groups;
    pa.w_q = w_q; pa.out_q = out_q; pa.w_params = w_params; pa.out_params = out_params; pa.seg_width = seg_width;
    pa.y_u8 = y_i8; pa.y_ctr = reinterpret_cast<__nv_bfloat16*>(y_ctr_bf16);
    cudaStream_t st = (cudaStream_t)stream;
    if (y_i8 != nullptr) return act_fn == 1 ? launch_peg<1, false, true>(a_i8, w_i8, M, N, K, pa, st) : launch_peg<0, false, true>(a_i8, w_i8, M, N, K, pa, st);
    return act_fn == 1 ? launch_peg<1, false, false>(a_i8, w_i8, M, N, K, pa, st) : launch_peg<0, false, false>(a_i8, w_i8, M, N, K, pa, st);
}

int tq_linear_peg_res_ln_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_grp_rowsum, const float* bias,
                                void* z_ctr_bf16, void* z_i8, int64_t M, int64_t N, int64_t K, tq_qspec a_q, int32_t a_groups,
                                tq_qspec w_q, int32_t w_params, tq_qspec out_q, int32_t out_params, const void* res_i8,
                                tq_qspec res_q, int32_t res_params, tq_qspec out2_q, int32_t out2_params,
                                const float* ln_gamma_q, const float* ln_beta, float ln_eps, tq_qspec ln_q, int32_t ln_params,
                                int64_t seg_width, void* stream) {
    using namespace tq::gemm;
    if (z_i8 == nullptr || res_i8 == nullptr || ln_gamma_q == nullptr || ln_beta == nullptr) return TQ_EINVAL;
    if (int e = peg_common_checks(a_i8, w_i8, w_grp_rowsum, M, N, K, a_q, a_groups, w_q, w_params, out_q, out_params, seg_width)) return e;
    if (int e = tq::check_qspec(res_q)) return e;
    if (int e = tq::check_qspec(out2_q)) return e;
    if (int e = tq::check_qspec(ln_q)) return e;
    if (res_q.n_bits > 8 || out2_q.n_bits > 8 || ln_q.n_bits > 8 || N / lean::peg::BN > 8) return TQ_EUNSUPPORTED;
    const int64_t nseg = N / seg_width;
    if ((res_params != 1 && res_params != nseg) || (out2_params != 1 && out2_params != nseg) || (ln_params != 1 && ln_params != nseg)) return TQ_EINVAL;
    if (!aligned32(z_i8) || !aligned32(res_i8) || (z_ctr_bf16 != nullptr && !aligned32(z_ctr_bf16))) return TQ_EALIGN;
    lean::peg::Args pa = {};
    if (const char* e = getenv("TQ_LINEAR_TRACE_PTR")) pa.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
    if (const char* e = getenv("TQ_LINEAR_TRACE_TILES")) pa.trace_tiles = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
    pa.bias = bias; pa.w_grp_rowsum = w_grp_rowsum; pa.a_q = a_q; pa.a_groups = a_groups;
    pa.w_q = w_q; pa.out_q = out_q; pa.w_params = w_params; pa.out_params = out_params; pa.seg_width = seg_width;
    pa.y_u8 = z_i8; pa.y_ctr = reinterpret_cast<__nv_bfloat16*>(z_ctr_bf16);
    pa.res_u8 = reinterpret_cast<const unsigned char*>(res_i8);
    pa.res_q = res_q; pa.out2_q = out2_q; pa.ln_q = ln_q;
    pa.res_params = res_params; pa.out2_params = out2_params; pa.ln_params = ln_params;
    pa.ln_gamma = ln_gamma_q; pa.ln_beta = ln_beta; pa.ln_eps = ln_eps;
    return launch_peg<0, true, true>(a_i8, w_i8, M, N, K, pa, (cudaStream_t)stream);
}

int tq_chain_plan_create(const tq_chain_stage* stages, int32_t n_stages, int64_t M, void** plan) {
    return tq::gemm::chain_plan_create(stages, n_stages, M, plan);
}
int tq_chain_plan_run(const void* plan, void* stream) {
    return tq::gemm::chain_plan_run(static_cast<const tq::gemm::ChainPlan*>(plan), (cudaStream_t)stream);
}
int tq_chain_plan_destroy(void* plan) {
    if (plan == nullptr) return TQ_OK;
    tq::gemm::ChainPlan* p = static_cast<tq::gemm::ChainPlan*>(plan);
    cudaFree(p->d_st);
    delete p;
    return TQ_OK;
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
