// Fused Linear forward on the 5th-generation tensor cores:  C[M,N] = act(A[M,K] . W[N,K]^T + bias)   (bf16 in, fp32
// accumulate, bf16 out) -- hand-written tcgen05 / TMEM / TMA kernel for sm_100a.
//
//   * operands: A (activations) and W (nn.Linear weight) are both K-major, so one TMA tensor map each with a
//     [128 rows x 64 bf16] box and the 128-byte swizzle lands tiles in shared memory in exactly the canonical
//     K-major SWIZZLE_128B layout tcgen05.mma consumes (no smem re-layout, no register staging);
//   * pipeline: kStages-deep ring of (A, W) tiles guarded by full/empty mbarriers; warp 0 (one lane) is the TMA
//     producer, warp 1 (one lane) issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128x128x16, 4 per 64-wide K block)
//     and releases ring slots with tcgen05.commit; the fp32 accumulator tile (128 lanes x 128 columns) lives in
//     TMEM (allocated by warp 2);
//   * epilogue: warps 4-7 wait on the accumulator-full mbarrier, pull their 32-lane quarter out of TMEM with
//     tcgen05.ld.32x32b.x32 (one output row per thread, 32 columns per instruction), add the bias, apply the optional
//     ReLU, convert to bf16 and store 64-byte row segments.
//
// One CTA per 128x128 output tile; ragged M / N / K edges are handled by TMA zero fill on loads and guards on stores.

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kThreads = 256;
constexpr int kTileBytes = BM * BK * 2;                 // 16 KiB per operand tile
constexpr int kTmemCols = BN;                           // fp32 accumulator: one column per output column
constexpr int kSmemBytes = 2 * kStages * kTileBytes + 256 + 1024;  // tiles + barriers + 1024 B alignment slack

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
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c_inner, int c_outer, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (=1, unused for swizzled K-major) |
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B between 8-row groups) | [46,48) version = 1 (sm_100) |
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    desc |= static_cast<uint64_t>(1) << 16;
    desc |= static_cast<uint64_t>(1024 >> 4) << 32;
    desc |= static_cast<uint64_t>(1) << 46;
    desc |= static_cast<uint64_t>(2) << 61;
    return desc;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, both operands K-major.
__host__ __device__ constexpr uint32_t make_instr_desc(int n = BN) {
    return (1u << 4)                    // c_format  = F32
           | (1u << 7)                  // a_format  = BF16
           | (1u << 10)                 // b_format  = BF16
           | (0u << 15) | (0u << 16)    // a_major = b_major = K
           | (uint32_t(n >> 3) << 17)   // n_dim
           | (uint32_t(BM >> 4) << 24); // m_dim
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
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
}

// Epilogue of one 32-column chunk of an accumulator row: (+bias) -> optional store of the pre-activation (what a GELU
// backward needs) -> activation -> bf16, 16-byte stores.
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_BIAS_BF16 = 0x100 };  // flag: persistent variants only

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16 resolution): one reciprocal, one exp2, a degree-5
// Horner chain -- about a third of libdevice erff's instructions.  With erff the GELU epilogue of a 128x256 tile was
// instruction-bound at ~5 us against a 3.2 us mainloop at K = 768 (ncu: 13 M instructions for 4096x3072x768).
__device__ __forceinline__ float erf_as(float x) {
    const float ax = fabsf(x);
    const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = exp2f(-1.4426950408889634f * ax * ax);
    return copysignf(1.f - p * t * e, x);
}

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == ACT_RELU) return x > 0.f ? x : 0.f;
    if (act == ACT_GELU) return 0.5f * x * (1.f + erf_as(x * 0.70710678118654752f));  // erf GELU, as nn.GELU()
    return x;
}

__device__ __forceinline__ void store8_bf16(__nv_bfloat16* dst, const float (&f)[8], int valid) {
    if (valid >= 8) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)                        // unrolled + predicated: keeps f[] in registers
            if (j < valid) dst[j] = __float2bfloat16(f[j]);
    }
}

__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], __nv_bfloat16* __restrict__ C,
                                               __nv_bfloat16* __restrict__ pre, const float* __restrict__ bias, int row,
                                               int col0, int N, int act) {
    const int64_t base = static_cast<int64_t>(row) * N + col0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {                          // 4 x 8 columns -> 16-byte stores
        const int col = col0 + q * 8;
        if (col >= N) break;
        float z[8], f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float x = __uint_as_float(v[q * 8 + j]);
            if (bias != nullptr && col + j < N) x += bias[col + j];
            z[j] = x;
            f[j] = apply_act(x, act);
        }
        store8_bf16(C + base + q * 8, f, N - col);
        if (pre != nullptr) store8_bf16(pre + base + q * 8, z, N - col);
    }
}

// Same epilogue with the 32x32 block transposed through a warp-private shared-memory buffer, so that every global
// store instruction writes 8 rows x 64 contiguous bytes (full 32-byte sectors) instead of 32 rows x 16 bytes.  With one
// thread per accumulator row, the direct version issues 16-byte partial-sector writes only: measured on B200 that
// costs ~10-15 us per 128x256 tile, i.e. the epilogue -- not the tensor pipe -- bounds every GEMM with K <= 1024.
// Phase timestamps of CTA 0 (SM clock), read back with fl4h_tc_debug_read: a handful of global stores per launch.
//   [0] kernel entry  [1] setup done (barriers, TMEM)  [2] first operands landed  [3 + t] MMA issue of tile t done (t < 8)
//   [16 + 2t] accumulator of tile t visible to the epilogue  [17 + 2t] epilogue of tile t done  [40] last TMA issued  [41] exit
//   [42..46] first chunk of tile 0, epilogue warp 0: TMEM load done / bias+activation done / staged / read back / stored
__device__ long long g_tc_phase_clock[48];
#define TC_STAMP(slot) do { if (blockIdx.x == 0) g_tc_phase_clock[(slot)] = clock64(); } while (0)

constexpr int kStageStride = 80;                            // 64 B of payload + 16 B pad: conflict-free 16-byte writes
constexpr int kStageBytesPerWarp = 32 * kStageStride;

// Explicit shared-space accesses.  The staging / bias buffers are carved out of the dynamically aligned `smem` pointer,
// for which the compiler had emitted GENERIC loads and stores (`LD.E.128 ... [R.64+0x30100]`, long-scoreboard class): one
// 32x32 epilogue chunk took ~2.0 us of pure latency (phase timestamps, 4096x2304x768).
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void stage_and_store(const float (&f)[32], uint32_t stage, __nv_bfloat16* __restrict__ dst, int row_base,
                                                int lane, int col0, int M, int N, bool stamp = false) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[q * 8 + 2 * j], f[q * 8 + 2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        sts128(stage + lane * kStageStride + q * 16, w[0], w[1], w[2], w[3]);
    }
    __syncwarp();
    if (stamp) TC_STAMP(45);
    const int seg = lane & 3, col = col0 + seg * 8;
    uint4 val[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) val[it] = lds128(stage + (it * 8 + (lane >> 2)) * kStageStride + seg * 16);
    if (stamp) TC_STAMP(46);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int grow = row_base + it * 8 + (lane >> 2);
        if (grow < M && col < N) *reinterpret_cast<uint4*>(dst + static_cast<int64_t>(grow) * N + col) = val[it];
    }
    __syncwarp();
    if (stamp) TC_STAMP(47);
}

// `bias_s`: this tile's bias slice staged in shared memory by the caller (nullptr = no bias), indexed from the chunk's
// first column.  Reading it from global here instead put ~40 % of the epilogue's stall samples on the first FADD after
// each LDG (ncu source view, 4096x2304x768).
__device__ __forceinline__ void epilogue_chunk_staged(const uint32_t (&v)[32], uint32_t stage, __nv_bfloat16* __restrict__ C,
                                                      __nv_bfloat16* __restrict__ pre, uint32_t bias_s /* shared address, 0 = none */,
                                                      int row_base, int lane, int col0, int M, int N, int act, bool stamp = false) {
    float z[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        float4 b4 = bias_s != 0 ? lds128f(bias_s + j * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        z[j] = __uint_as_float(v[j]) + b4.x;
        z[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
        z[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
        z[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
    }
    if (pre != nullptr) stage_and_store(z, stage, pre, row_base, lane, col0, M, N);
    // activation branch hoisted out of the element loop by hand: with the per-element `if (act == ...)` the compiler
    // kept a predicated GELU chain in the ReLU path (phase timestamps: 1.3 us of "bias + ReLU" per 32x32 chunk)
    if (act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = fmaxf(z[j], 0.f);
    } else if (act == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = 0.5f * z[j] * (1.f + erf_as(z[j] * 0.70710678118654752f));
    }
    if (stamp) TC_STAMP(44);
    stage_and_store(z, stage, C, row_base, lane, col0, M, N, stamp);
}

__global__ void __launch_bounds__(kThreads, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 __nv_bfloat16* __restrict__ C, __nv_bfloat16* __restrict__ pre, const float* __restrict__ bias, int M, int N,
                 int K, int act) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kTileBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * kStages * kTileBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.y, n_blk = blockIdx.x;
    const int num_k_blocks = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {  // whole warp: allocate the accumulator columns, publish the base address through smem
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int stage = kb % kStages;
                const uint32_t phase = (kb / kStages) & 1;
                mbar_wait(empty_bar + stage, phase ^ 1);
                mbar_expect_tx(full_bar + stage, 2 * kTileBytes);
                tma_load_2d(smem_a + stage * kTileBytes, &tma_a, kb * BK, m_blk * BM, full_bar + stage);
                tma_load_2d(smem_b + stage * kTileBytes, &tma_b, kb * BK, n_blk * BN, full_bar + stage);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer (single thread) =====
            constexpr uint32_t idesc = make_instr_desc();
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int stage = kb % kStages;
                const uint32_t phase = (kb / kStages) & 1;
                mbar_wait(full_bar + stage, phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + stage * kTileBytes);
                const uint32_t b_addr = smem_u32(smem_b + stage * kTileBytes);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // advancing K inside the 128-byte swizzle atom = advancing the start address by 32 bytes
                    const uint64_t desc_a = make_smem_desc(a_addr + k * UMMA_K * 2);
                    const uint64_t desc_b = make_smem_desc(b_addr + k * UMMA_K * 2);
                    umma_bf16(tmem_base, desc_a, desc_b, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar + stage);  // ring slot reusable once these MMAs have consumed it
            }
            umma_commit(tmem_full_bar);          // accumulator complete -> epilogue
        }
    } else if (warp >= 4) {  // ===== epilogue: TMEM -> registers -> (bias, ReLU, bf16) -> global =====
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int quarter = warp & 3;                      // warp w may only touch TMEM lanes [32 (w % 4), +32)
        const int row = m_blk * BM + quarter * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t acc[32];
            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0), acc);
            const int col0 = n_blk * BN + c0;
            if (row < M && col0 < N) epilogue_chunk(acc, C, pre, bias, row, col0, N, act);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

// cuTensorMapEncodeTiled is a driver-API entry point; resolve it through the runtime so the library does not link against
// libcuda (it must still dlopen on driver-less build machines).
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &status) == cudaSuccess &&
            status == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        }
    }
    return fn;
}

inline CUresult make_tensor_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols /*K, contiguous*/,
                                int box_rows = BM) {
    EncodeTiledFn encode = encode_tiled_fn();
    if (encode == nullptr) return CUDA_ERROR_NOT_SUPPORTED;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};                    // bytes between rows
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t elem_strides[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                                  elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}


// ---------------------------------------------------------------------------------------------------------------
// v2: persistent, 128 x kBN tiles (kBN = 256 doubles the math per shared-memory byte), double-buffered TMEM
// accumulators so the epilogue of tile i overlaps the main loop of tile i + 1.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int kBN, int kEpi = 4>
struct V2 {
    static constexpr int kStagesV2 = 4;
    static constexpr int kATile = BM * BK * 2;
    static constexpr int kBTile = kBN * BK * 2;
    static constexpr int kStageBytes = kATile + kBTile;
    static constexpr int kTmemColsV2 = 2 * kBN;           // two accumulator buffers
    static constexpr int kStagingOffset = kStagesV2 * kStageBytes + 256;       // after the barriers
    static constexpr int kEpiWarps = kEpi;                // 4 (one per TMEM lane quarter) or 8 (two per quarter, half the columns each)
    static constexpr int kThreadsV2 = 32 * (4 + kEpiWarps);
    static constexpr int kColsPerEpiWarp = kBN / (kEpiWarps / 4);
    static constexpr int kBiasOffset = kStagingOffset + kEpiWarps * kStageBytesPerWarp;  // per epilogue warp: its bias slice
    static constexpr int kSmemV2 = kBiasOffset + kEpiWarps * kColsPerEpiWarp * 4 + 1024;  // + staging + bias + alignment slack
};

template <int kBN, int kEpi>
__global__ void __launch_bounds__(V2<kBN, kEpi>::kThreadsV2, 1)
tc_linear_persistent_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                            __nv_bfloat16* __restrict__ C, __nv_bfloat16* __restrict__ pre, const float* __restrict__ bias, int M, int N,
                 int K, int act) {
    using Cfg = V2<kBN, kEpi>;
    const bool bias_bf16 = (act & ACT_BIAS_BF16) != 0;      // `bias` points at bf16 values (master-weight mode)
    act &= 0xff;
    if (threadIdx.x == 0) TC_STAMP(0);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStagesV2 * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + Cfg::kStagesV2;
    uint64_t* tmem_full_bar = empty_bar + Cfg::kStagesV2;     // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + kBN - 1) / kBN;
    const int num_tiles = tiles_m * tiles_n;
    const int num_k_blocks = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStagesV2; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar + a, 1);
            mbar_init(tmem_empty_bar + a, Cfg::kEpiWarps);    // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::kTmemColsV2));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    if (threadIdx.x == 0) TC_STAMP(1);

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer: one continuous ring across all of this CTA's tiles =====
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % tiles_m, n_blk = tile / tiles_m;   // concurrent CTAs share the W tile (L2 reuse)
                for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                    const int stage = it % Cfg::kStagesV2;
                    const uint32_t phase = (it / Cfg::kStagesV2) & 1;
                    uint8_t* a_dst = smem + stage * Cfg::kStageBytes;
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    mbar_expect_tx(full_bar + stage, Cfg::kStageBytes);
                    tma_load_2d(a_dst, &tma_a, kb * BK, m_blk * BM, full_bar + stage);
                    tma_load_2d(a_dst + Cfg::kATile, &tma_b, kb * BK, n_blk * kBN, full_bar + stage);
                }
            }
            TC_STAMP(40);
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            constexpr uint32_t idesc = make_instr_desc(kBN);
            uint32_t it = 0, local_tile = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
                const uint32_t acc = local_tile & 1, use = local_tile >> 1;
                mbar_wait(tmem_empty_bar + acc, (use & 1) ^ 1);  // epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kBN;
                for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                    const int stage = it % Cfg::kStagesV2;
                    const uint32_t phase = (it / Cfg::kStagesV2) & 1;
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after();
                    if (it == 0) TC_STAMP(2);
                    const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t b_addr = a_addr + Cfg::kATile;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        umma_bf16(tmem_d, make_smem_desc(a_addr + k * UMMA_K * 2), make_smem_desc(b_addr + k * UMMA_K * 2), idesc,
                                  (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar + stage);
                }
                umma_commit(tmem_full_bar + acc);
                if (local_tile < 8) TC_STAMP(3 + local_tile);
            }
        }
    } else if (warp >= 4) {  // ===== epilogue warps: warp e owns TMEM lanes [32 (e % 4), +32) x columns [half * kBN/2, +kBN/2) =====
        const int epi = warp - 4, quarter = epi & 3, col_begin = (epi >> 2) * Cfg::kColsPerEpiWarp;
        uint32_t local_tile = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
            const int m_blk = tile % tiles_m, n_blk = tile / tiles_m;
            const uint32_t acc = local_tile & 1, use = local_tile >> 1;
            // this warp's bias slice -> shared memory, issued BEFORE waiting for the accumulator so the global-load
            // latency hides behind the mainloop
            const uint32_t bias_s = smem_u32(smem + Cfg::kBiasOffset) + epi * Cfg::kColsPerEpiWarp * 4;
            if (bias != nullptr) {
                for (int c = lane; c < Cfg::kColsPerEpiWarp; c += 32) {
                    const int col = n_blk * kBN + col_begin + c;
                    float b = 0.f;
                    if (col < N) b = bias_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bias)[col]) : bias[col];
                    sts32f(bias_s + c * 4, b);
                }
                __syncwarp();
            }
            mbar_wait(tmem_full_bar + acc, use & 1);
            tc_fence_after();
            if (epi == 0 && lane == 0 && local_tile < 8) TC_STAMP(16 + 2 * local_tile);
            const int row_base = m_blk * BM + quarter * 32;
            const uint32_t stage = smem_u32(smem + Cfg::kStagingOffset) + epi * kStageBytesPerWarp;
#pragma unroll 1
            for (int c0 = col_begin; c0 < col_begin + Cfg::kColsPerEpiWarp; c0 += 32) {
                uint32_t v[32];
                const bool stamp = epi == 0 && lane == 0 && local_tile == 0 && c0 == col_begin;
                if (stamp) TC_STAMP(42);
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kBN + static_cast<uint32_t>(c0), v);
                if (stamp) TC_STAMP(43);
                const int col0 = n_blk * kBN + c0;
                if (row_base < M && col0 < N)
                    epilogue_chunk_staged(v, stage, C, pre, bias != nullptr ? bias_s + (c0 - col_begin) * 4 : 0u, row_base, lane, col0,
                                          M, N, act, stamp);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar + acc);     // this warp's quarter of the buffer is free again
            if (epi == 0 && lane == 0 && local_tile < 8) TC_STAMP(17 + 2 * local_tile);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemColsV2));
    }
    if (threadIdx.x == 0) TC_STAMP(41);
}

template <int kBN, int kEpi>
int launch_persistent(const void* a, const void* w, void* c, void* pre, const float* bias, int M, int N, int K, int act,
                      int sms, cudaStream_t stream) {
    using Cfg = V2<kBN, kEpi>;
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(tc_linear_persistent_kernel<kBN, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemV2);
        if (err != cudaSuccess) return static_cast<int>(err);
        configured = true;
    }
    CUtensorMap map_a, map_b;
    CUresult res = make_tensor_map(&map_a, a, static_cast<uint64_t>(M), static_cast<uint64_t>(K), BM);
    if (res != CUDA_SUCCESS) return -static_cast<int>(res);
    res = make_tensor_map(&map_b, w, static_cast<uint64_t>(N), static_cast<uint64_t>(K), kBN);
    if (res != CUDA_SUCCESS) return -static_cast<int>(res);
    const int tiles = ((M + BM - 1) / BM) * ((N + kBN - 1) / kBN);
    const int grid = tiles < sms ? tiles : sms;
    tc_linear_persistent_kernel<kBN, kEpi><<<grid, Cfg::kThreadsV2, Cfg::kSmemV2, stream>>>(
        map_a, map_b, static_cast<__nv_bfloat16*>(c), static_cast<__nv_bfloat16*>(pre), bias, M, N, K, act);
    return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// v4: CTA pair (thread-block cluster of 2, `tcgen05.mma.cta_group::2`).  One 256 x 256 output tile per pair: each CTA
// holds its own 128 rows of A and HALF of the W tile (128 of the 256 output columns) in shared memory; the leader's
// single MMA thread issues M = 256 instructions that run on both SMs' tensor cores, each SM reading the other's half
// of W over the pair's shared-memory path.  Per k-block a CTA now loads 32 KB instead of 48 KB (the L2 -> SM stream is
// what bounds the 1-CTA kernel at ~58 % of the tensor peak), and 5 stages fit instead of 4.
//
// Synchronisation (all mbarriers at identical offsets in both CTAs):
//   full[s]        lives in the LEADER: both CTAs' TMA loads complete_tx on it (peer bit of the address cleared);
//   empty[s]       in each CTA: `tcgen05.commit ... multicast::cluster` from the leader releases both producers;
//   tmem_full[a]   in each CTA: multicast commit, each CTA's epilogue warps wait locally;
//   tmem_empty[a]  in the LEADER: epilogue warps of both CTAs arrive (remote arrive from CTA 1).
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;              // shared::cluster address of the same offset in CTA rank 0

struct P2 {
    static constexpr int kBN = 256, kHalfN = 128;
    static constexpr int kStagesP = 5;
    static constexpr int kATile = BM * BK * 2;               // this CTA's 128 rows of A
    static constexpr int kBTile = kHalfN * BK * 2;           // this CTA's half of the W tile
    static constexpr int kStageBytes = kATile + kBTile;      // 32 KiB
    static constexpr int kEpiWarps = 8;
    static constexpr int kThreadsP = 32 * (4 + kEpiWarps);
    static constexpr int kColsPerEpiWarp = kBN / 2;
    static constexpr int kBarrierOffset = kStagesP * kStageBytes;
    static constexpr int kStagingOffset = kBarrierOffset + 256;
    static constexpr int kBiasOffset = kStagingOffset + kEpiWarps * kStageBytesPerWarp;
    static constexpr int kSmemP = kBiasOffset + kEpiWarps * kColsPerEpiWarp * 4 + 1024;
    static constexpr int kTmemColsP = 512;                   // two 256-column accumulator buffers = all of TMEM
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c_inner, int c_outer, uint64_t* leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c_inner), "r"(c_outer)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the leader CTA's copy of `bar`
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__host__ __device__ constexpr uint32_t make_instr_desc_pair() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(P2::kBN >> 3) << 17) | (uint32_t(256 >> 4) << 24);  // M = 256
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P2::kThreadsP, 1)
tc_linear_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                      __nv_bfloat16* __restrict__ C, __nv_bfloat16* __restrict__ pre, const float* __restrict__ bias, int M, int N,
                      int K, int act) {
    const bool bias_bf16 = (act & ACT_BIAS_BF16) != 0;
    act &= 0xff;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P2::kBarrierOffset);
    uint64_t* empty_bar = full_bar + P2::kStagesP;
    uint64_t* tmem_full_bar = empty_bar + P2::kStagesP;       // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;             // [2], used in the leader only
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + P2::kBN - 1) / P2::kBN;
    const int num_tiles = tiles_m * tiles_n;
    const int num_k_blocks = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < P2::kStagesP; ++s) {
            mbar_init(full_bar + s, 1);                      // leader: one expect_tx arrival covering both CTAs' bytes
            mbar_init(empty_bar + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar + a, 1);
            mbar_init(tmem_empty_bar + a, 2 * P2::kEpiWarps);  // every epilogue warp of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(P2::kTmemColsP));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                       // both CTAs' barriers are initialised, TMEM is allocated
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs): own A rows, own half of W =====
            uint32_t it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int m_blk = tile % tiles_m, n_blk = tile / tiles_m;
                const int row_a = m_blk * 2 * BM + static_cast<int>(rank) * BM;
                const int row_b = n_blk * P2::kBN + static_cast<int>(rank) * P2::kHalfN;
                for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                    const int stage = it % P2::kStagesP;
                    const uint32_t phase = (it / P2::kStagesP) & 1;
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    if (rank == 0) mbar_expect_tx(full_bar + stage, 2 * P2::kStageBytes);
                    uint8_t* a_dst = smem + stage * P2::kStageBytes;
                    tma_load_2d_pair(a_dst, &tma_a, kb * BK, row_a, full_bar + stage);
                    tma_load_2d_pair(a_dst + P2::kATile, &tma_b, kb * BK, row_b, full_bar + stage);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {  // ===== MMA issuer: leader CTA only =====
            constexpr uint32_t idesc = make_instr_desc_pair();
            uint32_t it = 0, local_tile = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local_tile) {
                const uint32_t acc = local_tile & 1, use = local_tile >> 1;
                mbar_wait(tmem_empty_bar + acc, (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * P2::kBN;
                for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                    const int stage = it % P2::kStagesP;
                    const uint32_t phase = (it / P2::kStagesP) & 1;
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * P2::kStageBytes);
                    const uint32_t b_addr = a_addr + P2::kATile;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        umma_bf16_pair(tmem_d, make_smem_desc(a_addr + k * UMMA_K * 2), make_smem_desc(b_addr + k * UMMA_K * 2), idesc,
                                       (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit_pair(empty_bar + stage);     // both CTAs' copies of this stage are free again
                }
                umma_commit_pair(tmem_full_bar + acc);       // both CTAs' epilogues may read their accumulator half
            }
        }
    } else if (warp >= 4) {  // ===== epilogue (both CTAs): rows [rank * 128, +128) of the pair's 256 x 256 tile =====
        const int epi = warp - 4, quarter = epi & 3, col_begin = (epi >> 2) * P2::kColsPerEpiWarp;
        uint32_t local_tile = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local_tile) {
            const int m_blk = tile % tiles_m, n_blk = tile / tiles_m;
            const uint32_t acc = local_tile & 1, use = local_tile >> 1;
            const uint32_t bias_s = smem_u32(smem + P2::kBiasOffset) + epi * P2::kColsPerEpiWarp * 4;
            if (bias != nullptr) {
                for (int c = lane; c < P2::kColsPerEpiWarp; c += 32) {
                    const int col = n_blk * P2::kBN + col_begin + c;
                    float b = 0.f;
                    if (col < N) b = bias_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bias)[col]) : bias[col];
                    sts32f(bias_s + c * 4, b);
                }
                __syncwarp();
            }
            mbar_wait(tmem_full_bar + acc, use & 1);
            tc_fence_after();
            const int row_base = m_blk * 2 * BM + static_cast<int>(rank) * BM + quarter * 32;
            const uint32_t stage = smem_u32(smem + P2::kStagingOffset) + epi * kStageBytesPerWarp;
#pragma unroll 1
            for (int c0 = col_begin; c0 < col_begin + P2::kColsPerEpiWarp; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * P2::kBN + static_cast<uint32_t>(c0), v);
                const int col0 = n_blk * P2::kBN + c0;
                if (row_base < M && col0 < N)
                    epilogue_chunk_staged(v, stage, C, pre, bias != nullptr ? bias_s + (c0 - col_begin) * 4 : 0u, row_base, lane, col0,
                                          M, N, act);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tmem_empty_bar + acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                       // neither CTA may free TMEM / exit while the peer still uses it
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(P2::kTmemColsP));
    }
}

int launch_pair(const void* a, const void* w, void* c, void* pre, const float* bias, int M, int N, int K, int act, int sms,
                cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(tc_linear_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2::kSmemP);
        if (err != cudaSuccess) return static_cast<int>(err);
        configured = true;
    }
    CUtensorMap map_a, map_b;
    CUresult res = make_tensor_map(&map_a, a, static_cast<uint64_t>(M), static_cast<uint64_t>(K), BM);
    if (res != CUDA_SUCCESS) return -static_cast<int>(res);
    res = make_tensor_map(&map_b, w, static_cast<uint64_t>(N), static_cast<uint64_t>(K), P2::kHalfN);
    if (res != CUDA_SUCCESS) return -static_cast<int>(res);
    const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + P2::kBN - 1) / P2::kBN);
    int pairs = sms / 2;
    if (tiles < pairs) pairs = tiles;
    tc_linear_pair_kernel<<<2 * pairs, P2::kThreadsP, P2::kSmemP, stream>>>(
        map_a, map_b, static_cast<__nv_bfloat16*>(c), static_cast<__nv_bfloat16*>(pre), bias, M, N, K, act);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace

extern "C" {

// act: 0 none, 1 ReLU, 2 GELU(erf); | 0x100 = `bias` holds bf16 instead of fp32 (variants 1 and 2 only).
// `pre` (optional, [M, N] bf16) receives x.W^T + b BEFORE the activation.
// variant: 0 = one 128x128 tile per CTA, 1 = persistent 128x128, 2 = persistent 128x256, 3 = 128x256 with 8 epilogue warps, 4 = CTA pair (cta_group::2), 256x256 per pair.
// Returns 0 on success; >0 cudaError; <0 = -CUresult of the tensor-map encoding.
int fl4h_tc_linear_ex(const void* a, const void* w, void* c, void* pre, const float* bias, int M, int N, int K, int act,
                      int variant, cudaStream_t stream) {
    static int sms = 0;
    if (sms == 0) {
        int device = 0;
        cudaGetDevice(&device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    }
    if (variant == 1) return launch_persistent<128, 4>(a, w, c, pre, bias, M, N, K, act, sms, stream);
    if (variant == 2) return launch_persistent<256, 4>(a, w, c, pre, bias, M, N, K, act, sms, stream);
    if (variant == 3) return launch_persistent<256, 8>(a, w, c, pre, bias, M, N, K, act, sms, stream);
    if (variant == 4) return launch_pair(a, w, c, pre, bias, M, N, K, act, sms, stream);
    if (act & ACT_BIAS_BF16) return static_cast<int>(cudaErrorInvalidValue);
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (err != cudaSuccess) return static_cast<int>(err);
        configured = true;
    }
    CUtensorMap map_a, map_b;
    CUresult res = make_tensor_map(&map_a, a, static_cast<uint64_t>(M), static_cast<uint64_t>(K));
    if (res != CUDA_SUCCESS) return -static_cast<int>(res);
    res = make_tensor_map(&map_b, w, static_cast<uint64_t>(N), static_cast<uint64_t>(K));
    if (res != CUDA_SUCCESS) return -static_cast<int>(res);
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    tc_linear_kernel<<<grid, kThreads, kSmemBytes, stream>>>(map_a, map_b, static_cast<__nv_bfloat16*>(c),
                                                             static_cast<__nv_bfloat16*>(pre), bias, M, N, K, act);
    return static_cast<int>(cudaGetLastError());
}

// Copies the 48 phase timestamps of the last persistent-variant launch (CTA 0, SM clock cycles) to `out`.
int fl4h_tc_debug_read(long long* out) {
    return static_cast<int>(cudaMemcpyFromSymbol(out, g_tc_phase_clock, sizeof(long long) * 48));
}

int fl4h_tc_linear_v(const void* a, const void* w, void* c, const float* bias, int M, int N, int K, int relu, int variant,
                     cudaStream_t stream) {
    return fl4h_tc_linear_ex(a, w, c, nullptr, bias, M, N, K, relu ? ACT_RELU : ACT_NONE, variant, stream);
}

int fl4h_tc_linear(const void* a, const void* w, void* c, const float* bias, int M, int N, int K, int relu, cudaStream_t stream) {
    return fl4h_tc_linear_ex(a, w, c, nullptr, bias, M, N, K, relu ? ACT_RELU : ACT_NONE, 0, stream);
}

}  // extern "C"
