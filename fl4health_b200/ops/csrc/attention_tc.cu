// Self-attention for short sequences (T <= 128, head dimension 64, bf16) on the 5th-generation tensor cores, sm_100a.
//
// The reference's transformer workload (examples/bert_finetuning_example: bert-base, sequence 128, 12 heads x 64) has a
// whole (batch, head) attention problem that fits ONE CTA: S = Q K^T is a single 128 x 128 x 64 tile.  So there is no
// online-softmax loop and no running rescale: one CTA per (batch, head) does
//
//   forward :  S = Q K^T (tcgen05.mma, fp32 in TMEM) -> row softmax in registers (one thread per query row = one TMEM
//              lane) -> P (bf16) into shared memory in the canonical K-major SWIZZLE_128B layout -> O = P V (second MMA;
//              V is consumed MN-major exactly as TMA delivered it) -> O / rowsum, bf16, straight into the [B, T, H]
//              activation layout the output projection reads; log-sum-exp per row kept for the backward.
//   backward:  recompute S and P; dP = dO V^T; dS = P o (dP - rowsum(dO o O)) * scale; dV = P^T dO; dK = dS^T Q;
//              dQ = dS K -- five MMAs whose operands are the SAME five shared-memory tiles read K-major or MN-major as
//              each contraction needs (no transposes anywhere), accumulators side by side in TMEM (320 of 512 columns).
//
// Q, K and V are read in place from the fused projection output [B, T, 3, heads, 64] through one 3-D TMA tensor map
// (no split / permute copies); sequence padding to 128 rows is the TMA unit's out-of-bounds zero fill, key padding is a
// -inf bias from the [B, T] attention mask.  Gradients are written in the same packed layout.

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace {

constexpr int kT = 128;          // rows per tile (queries / keys), TMEM lanes
constexpr int kD = 64;           // head dimension: one 128-byte swizzle row of bf16
constexpr int kTile = kT * kD * 2;        // 16 KB: a [128, 64] bf16 operand tile
constexpr int kProb = kT * kT * 2;        // 32 KB: a [128, 128] bf16 probability tile (two 64-column halves)
constexpr int kThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "ATT_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra ATT_WAIT_DONE;\n\t"
        "bra ATT_WAIT_LOOP;\n\t"
        "ATT_WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// SWIZZLE_128B shared-memory matrix descriptor (see conv_tc.cu for the field layout).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    desc |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    desc |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    desc |= static_cast<uint64_t>(1) << 46;
    desc |= static_cast<uint64_t>(2) << 61;
    return desc;
}
// operand whose contraction index is the contiguous one: k-th 16-element slice of a [rows, 64] tile
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int k16) { return make_desc(tile + k16 * 32, 16, 1024); }
// operand whose M/N index is the contiguous one: contraction rows [16 k16, 16 k16 + 16) of a [rows, 64 * chunks] tile
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int k16, uint32_t chunk_stride) {
    return make_desc(tile + k16 * 16 * 128, chunk_stride, 1024);
}
// bf16 x bf16 -> fp32; bits 15 / 16: A / B is MN-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// byte address of the 16-byte chunk holding columns [8 c8, 8 c8 + 8) of `row` in a [128, 128] bf16 probability tile
__device__ __forceinline__ uint32_t prob_chunk(uint32_t tile, int row, int c8) {
    return tile + (c8 >> 3) * kTile + row * 128 + (((c8 & 7) ^ (row & 7)) << 4);
}
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const float (&v)[32], int offset) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        uint4 w = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                             pack_bf16(v[j + 6], v[j + 7]));
        *reinterpret_cast<uint4*>(dst + offset + j) = w;
    }
}

struct AttnMaps {
    CUtensorMap qkv;   // dims (3 * H, T, B), box (64, 128, 1)
    CUtensorMap out;   // dims (H, T, B): O (backward) -- unused in the forward
    CUtensorMap dout;  // dims (H, T, B): dO
};

struct AttnGeom {
    int batch, seq, heads;   // H = heads * 64
    float scale;             // 1 / sqrt(64) unless the caller says otherwise
};

// key bias: 0 for a key that takes part, -inf for sequence padding (k >= seq) or masked-out keys
__device__ __forceinline__ void load_key_bias(float* bias, const uint8_t* mask, int b, int seq) {
    const int k = threadIdx.x;
    const bool on = k < seq && (mask == nullptr || mask[(int64_t)b * seq + k] != 0);
    bias[k] = on ? 0.f : -CUDART_INF_F;
}

__global__ void __launch_bounds__(kThreads)
attention_fwd_kernel(const __grid_constant__ AttnMaps maps, const AttnGeom g, const uint8_t* __restrict__ mask,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ lse) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *q_s = smem, *k_s = smem + kTile, *v_s = smem + 2 * kTile, *p_s = smem + 3 * kTile;
    float* bias = reinterpret_cast<float*>(p_s + kProb);
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias + kT);        // [0] Q,K landed  [1] V landed  [2] S ready  [3] O ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.x / g.heads, h = blockIdx.x % g.heads;
    const int hidden = g.heads * kD;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    load_key_bias(bias, mask, b, g.seq);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);

    if (tid == 0) {
        mbar_expect_tx(bars + 0, 2 * kTile);
        tma_load_3d(q_s, &maps.qkv, h * kD, 0, b, bars + 0);
        tma_load_3d(k_s, &maps.qkv, hidden + h * kD, 0, b, bars + 0);
        mbar_expect_tx(bars + 1, kTile);
        tma_load_3d(v_s, &maps.qkv, 2 * hidden + h * kD, 0, b, bars + 1);
        mbar_wait(bars + 0, 0);
        tc_fence_after();
        constexpr uint32_t idesc_s = make_idesc(kT, kT, false, false);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
            umma(tmem, desc_kmajor(smem_u32(q_s), k), desc_kmajor(smem_u32(k_s), k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bars + 2);
    }
    mbar_wait(bars + 2, 0);
    tc_fence_after();

    // ---- softmax over this thread's row: pass 1 the maximum, pass 2 the (unnormalised) probabilities ----
    const float c = g.scale * 1.4426950408889634f;                  // scores are used as s * scale, in base 2
    float row_max = -CUDART_INF_F;
#pragma unroll 1
    for (int c0 = 0; c0 < kT; c0 += 32) {
        float s[32];
        tmem_ld32(lane_base + c0, s);
#pragma unroll
        for (int j = 0; j < 32; ++j) row_max = fmaxf(row_max, s[j] * c + bias[c0 + j]);
    }
    const float m = row_max == -CUDART_INF_F ? 0.f : row_max;      // a fully masked row produces zeros, not NaNs
    float row_sum = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < kT; c0 += 32) {
        float s[32];
        tmem_ld32(lane_base + c0, s);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            s[j] = exp2f(s[j] * c + bias[c0 + j] - m);
            row_sum += s[j];
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8)
            sts128(prob_chunk(smem_u32(p_s), tid, (c0 + j) >> 3), pack_bf16(s[j], s[j + 1]), pack_bf16(s[j + 2], s[j + 3]),
                   pack_bf16(s[j + 4], s[j + 5]), pack_bf16(s[j + 6], s[j + 7]));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of P -> visible to the MMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        mbar_wait(bars + 1, 0);
        tc_fence_after();
        constexpr uint32_t idesc_o = make_idesc(kT, kD, false, true);
#pragma unroll
        for (int k = 0; k < kT / 16; ++k)                           // contraction over the 128 keys
            umma(tmem + kT, desc_kmajor(smem_u32(p_s) + (k >> 2) * kTile, k & 3), desc_mnmajor(smem_u32(v_s), k, kTile), idesc_o,
                 k > 0 ? 1u : 0u);
        umma_commit(bars + 3);
    }
    mbar_wait(bars + 3, 0);
    tc_fence_after();
    const float inv = row_sum > 0.f ? 1.f / row_sum : 0.f;
    __nv_bfloat16* dst = out + ((int64_t)b * g.seq + tid) * hidden + h * kD;
#pragma unroll 1
    for (int c0 = 0; c0 < kD; c0 += 32) {                           // (warp-collective TMEM loads: rows past the sequence take part)
        float o[32];
        tmem_ld32(lane_base + kT + c0, o);
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] *= inv;
        if (tid < g.seq) store_row_bf16(dst, o, c0);
    }
    // natural-log sum of exp of the SCALED scores: log(sum_j exp(scale * s_j)) = (m + log2(row_sum)) * ln 2
    if (tid < g.seq) lse[((int64_t)b * g.heads + h) * g.seq + tid] = (m + log2f(row_sum)) * 0.6931471805599453f;
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

// TMEM columns of the backward accumulators
constexpr int kColS = 0, kColDV = 128, kColDK = 192, kColDQ = 256;

__global__ void __launch_bounds__(kThreads)
attention_bwd_kernel(const __grid_constant__ AttnMaps maps, const AttnGeom g, const uint8_t* __restrict__ mask,
                     const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                     const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *q_s = smem, *k_s = smem + kTile, *v_s = smem + 2 * kTile, *do_s = smem + 3 * kTile;
    uint8_t *p_s = smem + 4 * kTile, *ds_s = p_s + kProb;
    float* bias = reinterpret_cast<float*>(ds_s + kProb);
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias + kT);        // [0] Q,K  [1] V,dO  [2] S  [3] dP,dV  [4] dK,dQ
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.x / g.heads, h = blockIdx.x % g.heads;
    const int hidden = g.heads * kD;
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    load_key_bias(bias, mask, b, g.seq);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t q_a = smem_u32(q_s), k_a = smem_u32(k_s), v_a = smem_u32(v_s), do_a = smem_u32(do_s), p_a = smem_u32(p_s),
                   ds_a = smem_u32(ds_s);

    if (tid == 0) {
        mbar_expect_tx(bars + 0, 2 * kTile);
        tma_load_3d(q_s, &maps.qkv, h * kD, 0, b, bars + 0);
        tma_load_3d(k_s, &maps.qkv, hidden + h * kD, 0, b, bars + 0);
        mbar_expect_tx(bars + 1, 2 * kTile);
        tma_load_3d(v_s, &maps.qkv, 2 * hidden + h * kD, 0, b, bars + 1);
        tma_load_3d(do_s, &maps.dout, h * kD, 0, b, bars + 1);
        mbar_wait(bars + 0, 0);
        tc_fence_after();
        constexpr uint32_t idesc_s = make_idesc(kT, kT, false, false);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma(tmem + kColS, desc_kmajor(q_a, k), desc_kmajor(k_a, k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bars + 2);
    }

    // rowsum(dO o O) and the row's log-sum-exp while the first MMA runs (rows past the sequence: zeros)
    float delta = 0.f, row_lse = 0.f;
    if (tid < g.seq) {
        const int64_t off = ((int64_t)b * g.seq + tid) * hidden + h * kD;
#pragma unroll
        for (int j = 0; j < kD; j += 8) {
            const uint4 a = *reinterpret_cast<const uint4*>(out + off + j), d = *reinterpret_cast<const uint4*>(dout + off + j);
            const __nv_bfloat162 *ah = reinterpret_cast<const __nv_bfloat162*>(&a), *dh = reinterpret_cast<const __nv_bfloat162*>(&d);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 x = __bfloat1622float2(ah[e]), y = __bfloat1622float2(dh[e]);
                delta += x.x * y.x + x.y * y.y;
            }
        }
        row_lse = lse[((int64_t)b * g.heads + h) * g.seq + tid];
    }
    const float c = g.scale * 1.4426950408889634f, lse2 = row_lse * 1.4426950408889634f;
    const bool live = tid < g.seq && row_lse > -CUDART_INF_F;       // padding rows and fully masked rows carry no probability

    mbar_wait(bars + 2, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < kT; c0 += 32) {                           // P = exp(scale * S - lse), bf16, into shared memory
        float s[32];
        tmem_ld32(lane_base + kColS + c0, s);
#pragma unroll
        for (int j = 0; j < 32; ++j) s[j] = live ? exp2f(s[j] * c + bias[c0 + j] - lse2) : 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 8)
            sts128(prob_chunk(p_a, tid, (c0 + j) >> 3), pack_bf16(s[j], s[j + 1]), pack_bf16(s[j + 2], s[j + 3]),
                   pack_bf16(s[j + 4], s[j + 5]), pack_bf16(s[j + 6], s[j + 7]));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();                                                // every row of S has been read: its columns are free
    tc_fence_after();
    if (tid == 0) {
        mbar_wait(bars + 1, 0);
        tc_fence_after();
        constexpr uint32_t idesc_dp = make_idesc(kT, kT, false, false);   // dP[q, k] = sum_d dO[q, d] V[k, d]
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma(tmem + kColS, desc_kmajor(do_a, k), desc_kmajor(v_a, k), idesc_dp, k > 0 ? 1u : 0u);
        constexpr uint32_t idesc_t = make_idesc(kT, kD, true, true);      // dV[k, d] = sum_q P[q, k] dO[q, d]
#pragma unroll
        for (int k = 0; k < kT / 16; ++k)
            umma(tmem + kColDV, desc_mnmajor(p_a, k, kTile), desc_mnmajor(do_a, k, kTile), idesc_t, k > 0 ? 1u : 0u);
        umma_commit(bars + 3);
    }
    mbar_wait(bars + 3, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < kT; c0 += 32) {                           // dS = P o (dP - delta) * scale
        float dp[32];
        tmem_ld32(lane_base + kColS + c0, dp);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const uint4 pw = lds128(prob_chunk(p_a, tid, (c0 + j) >> 3));
            const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&pw);
            float ds[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 p = __bfloat1622float2(ph[e]);
                ds[2 * e] = p.x * (dp[j + 2 * e] - delta) * g.scale;
                ds[2 * e + 1] = p.y * (dp[j + 2 * e + 1] - delta) * g.scale;
            }
            sts128(prob_chunk(ds_a, tid, (c0 + j) >> 3), pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]),
                   pack_bf16(ds[6], ds[7]));
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        constexpr uint32_t idesc_t = make_idesc(kT, kD, true, true);      // dK[k, d] = sum_q dS[q, k] Q[q, d]
#pragma unroll
        for (int k = 0; k < kT / 16; ++k)
            umma(tmem + kColDK, desc_mnmajor(ds_a, k, kTile), desc_mnmajor(q_a, k, kTile), idesc_t, k > 0 ? 1u : 0u);
        constexpr uint32_t idesc_q = make_idesc(kT, kD, false, true);     // dQ[q, d] = sum_k dS[q, k] K[k, d]
#pragma unroll
        for (int k = 0; k < kT / 16; ++k)
            umma(tmem + kColDQ, desc_kmajor(ds_a + (k >> 2) * kTile, k & 3), desc_mnmajor(k_a, k, kTile), idesc_q, k > 0 ? 1u : 0u);
        umma_commit(bars + 4);
    }
    mbar_wait(bars + 4, 0);
    tc_fence_after();
    // rows of dV / dK are keys, rows of dQ are queries; all three go to the packed [B, T, 3, heads, 64] gradient
    __nv_bfloat16* row_out = dqkv + ((int64_t)b * g.seq + tid) * 3 * hidden + h * kD;
    const int cols[3] = {kColDQ, kColDK, kColDV};
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
#pragma unroll 1
        for (int c0 = 0; c0 < kD; c0 += 32) {
            float v[32];
            tmem_ld32(lane_base + cols[which] + c0, v);
            if (tid < g.seq) store_row_bf16(row_out + which * hidden, v, c0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &status) == cudaSuccess &&
            status == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// [B, T, cols] bf16 tensor seen as (cols, T, B); tiles of 64 columns x 128 rows of one batch element; rows past T read 0.
int token_map(CUtensorMap* map, const void* base, int cols, int seq, int batch) {
    EncodeTiledFn encode = encode_fn();
    if (encode == nullptr) return (int)cudaErrorNotSupported;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)seq, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * (cuuint64_t)seq};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

constexpr int kFwdSmem = 3 * kTile + kProb + kT * 4 + 4 * 8 + 16 + 1024;
constexpr int kBwdSmem = 4 * kTile + 2 * kProb + kT * 4 + 5 * 8 + 16 + 1024;

}  // namespace

extern "C" {

// qkv: [B, T, 3, heads, 64] bf16; mask: [B, T] bytes (non-zero = attend) or null; out: [B, T, heads * 64] bf16;
// lse: [B, heads, T] fp32.
int fl4h_attention_fwd(const void* qkv, const uint8_t* mask, void* out, float* lse, int batch, int seq, int heads, float scale,
                       cudaStream_t stream) {
    if (seq < 1 || seq > kT || heads < 1) return (int)cudaErrorInvalidValue;
    AttnMaps maps;
    const int hidden = heads * kD;
    int err = token_map(&maps.qkv, qkv, 3 * hidden, seq, batch);
    if (err) return err;
    maps.out = maps.qkv;
    maps.dout = maps.qkv;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
        configured = true;
    }
    const AttnGeom g{batch, seq, heads, scale};
    attention_fwd_kernel<<<batch * heads, kThreads, kFwdSmem, stream>>>(maps, g, mask, reinterpret_cast<__nv_bfloat16*>(out), lse);
    return (int)cudaGetLastError();
}

// out / dout: [B, T, heads * 64] bf16 (contiguous); dqkv: [B, T, 3, heads, 64] bf16 (every element is written).
int fl4h_attention_bwd(const void* qkv, const uint8_t* mask, const void* out, const void* dout, const float* lse, void* dqkv,
                       int batch, int seq, int heads, float scale, cudaStream_t stream) {
    if (seq < 1 || seq > kT || heads < 1) return (int)cudaErrorInvalidValue;
    AttnMaps maps;
    const int hidden = heads * kD;
    int err = token_map(&maps.qkv, qkv, 3 * hidden, seq, batch);
    if (err) return err;
    err = token_map(&maps.dout, dout, hidden, seq, batch);
    if (err) return err;
    maps.out = maps.dout;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
        configured = true;
    }
    const AttnGeom g{batch, seq, heads, scale};
    attention_bwd_kernel<<<batch * heads, kThreads, kBwdSmem, stream>>>(
        maps, g, mask, reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), lse,
        reinterpret_cast<__nv_bfloat16*>(dqkv));
    return (int)cudaGetLastError();
}

}  // extern "C"
