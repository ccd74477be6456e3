// Implicit-GEMM 2-D convolution on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), NHWC, sm_100a.
//
// One "tap-GEMM" kernel covers forward and data-gradient convolutions:
//
//     out[n, h, w, :] = sum over taps t  in[map_t][n, h + dh_t, w + dw_t, :] . Wmat[:, wcol_t : wcol_t + Cin]^T
//
//   * A operand (activations): 4-D TMA tensor maps (C, W, H, N) over the NHWC tensor, box (128 B of channels, bw, bh,
//     bn) with bw*bh*bn = 128 output pixels.  A tap is just a coordinate shift; out-of-range coordinates are zero-filled
//     by the TMA unit, which IS the convolution padding.  Strided convolutions / their transposes use up to four maps
//     over parity sub-lattices of the tensor (base offset + doubled strides), so every tap stays a dense box load;
//   * B operand (filters): a K-major matrix [Cout_tile rows, taps*Cin columns] — PyTorch's channels_last weight layout
//     [Cout, R, S, Cin] as it sits in the parameter arena (forward), or its [Cin, R, S, Cout] permutation (dgrad);
//   * both land in shared memory in the canonical K-major SWIZZLE_128B layout and feed `tcgen05.mma` (kind::tf32 for
//     fp32 tensors = the reference's cuDNN-TF32 math, kind::f16 for bf16) from a 4-stage mbarrier ring; the fp32
//     accumulator tile (128 pixels x BN channels) lives in TMEM;
//   * split-K (deep layers: few pixels, K up to 4608): the K range is split over a thread-block CLUSTER along z; the
//     partial tiles meet in the leader CTA through distributed shared memory (`ld.shared::cluster`), so there is no
//     zero-init, no atomics on the output and a fixed summation order;
//   * epilogue: TMEM -> registers -> swizzled shared-memory tile -> ONE TMA store per 128-byte channel group
//     (`cp.async.bulk.tensor.4d.global.shared`, full-line writes, clipping at the batch edge), and — while the tile is in
//     shared memory — per-channel sum / sum-of-squares of the stored values, reduced over the tile and added to a
//     [2, Cout] statistics buffer: BatchNorm's first pass disappears (reference hot op:
//     examples/models/cnn_model.py:16-22, research/cifar10/model.py:38-47 -> torch.nn.Conv2d + BatchNorm2d).
//
// The weight gradient is a different contraction (pixels are the reduction dimension): see `wgrad_kernel` below,
// which reads BOTH operands MN-major straight from the NHWC tensors.

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int BM = 128;            // output pixels per tile (TMEM lanes)
constexpr int kRowBytes = 128;     // one K block = 128 bytes of channels of one tap (32 fp32 / 64 bf16)
constexpr int kThreads = 256;      // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-7: epilogue
constexpr int kMaxTaps = 9;
constexpr int kMaxSplit = 8;       // portable cluster size

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
        "CONV_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CONV_WAIT_DONE;\n\t"
        "bra CONV_WAIT_LOOP;\n\t"
        "CONV_WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 | [46,48) version = 1 |
//   [61,64) layout = 2 (SWIZZLE_128B).
// K-major : rows of 128 B, 8-row groups 1024 B apart (SBO = 1024; LBO unused).
// MN-major: each 128-B row runs along M/N for ONE k; 8 consecutive k rows form an atom (1024 B); the next 128-B chunk
//           along M/N sits LBO bytes away, the next group of 8 k rows SBO bytes away (canonical layout
//           ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)), cute/atom/mma_traits_sm100.hpp).
// 32-bit MN-major operands (TF32 weight gradient) need the 32-byte-atom flavour: layout = 1 (SWIZZLE_128B_BASE32B,
//           Swizzle<2,5,2>: 32-byte chunks XOR (row & 3)), atoms of 4 k rows (512 B), filled by TMA with
//           CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  (With the plain 128B layout the tf32 MN-major MMA returned zeros.)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    desc |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    desc |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    desc |= static_cast<uint64_t>(1) << 46;
    desc |= static_cast<uint64_t>(layout) << 61;
    return desc;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate; formats: kind::f16 {F16=0, BF16=1},
// kind::tf32 {TF32=2}; bit 15/16: A/B major (0 = K, 1 = MN); [17,23) N >> 3; [24,29) M >> 4.
__host__ __device__ constexpr uint32_t make_idesc(bool tf32, int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4) | ((tf32 ? 2u : 1u) << 7) | ((tf32 ? 2u : 1u) << 10) | ((a_mn ? 1u : 0u) << 15) |
           ((b_mn ? 1u : 0u) << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

template <bool kTf32>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (kTf32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    }
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

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta_rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(cta_rank));
    return remote;
}
__device__ __forceinline__ float4 ld_dsmem128f(uint32_t cluster_addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// Programmatic dependent launch: the kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization, so
// their prologue (barrier init, TMEM allocation, tensor-map prefetch) overlaps the tail of the previous kernel in the
// stream; `pdl_wait` blocks until that kernel's memory is visible, `pdl_trigger` lets the NEXT kernel start its own
// prologue early.  (In a CUDA graph these become programmatic dependency edges.)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // epilogue warps only

struct TapTable {
    int ntaps;
    int wcol[kMaxTaps];          // first column (elements) of the tap's block in the filter matrix
    signed char dh[kMaxTaps], dw[kMaxTaps], map[kMaxTaps];
};

// One launch serves up to four "classes" that share the input and the filter matrix but have their own tap list and
// output sub-lattice: the four output parities of a stride-2 data gradient (a forward convolution has one class).
struct ConvClasses {
    TapTable taps[4];
};

struct ConvMaps {
    CUtensorMap a[4];            // activation maps (parity sub-lattices for strided layers; a[0] otherwise)
    CUtensorMap b;               // filter matrix: K-major [rows, taps*Cin] (forward) or the forward matrix read MN-major (dgrad)
    CUtensorMap out[4];          // output sub-lattice per class: (C, W, H, N), box (128 B, bw, bh, bn), SWIZZLE_128B
};

struct ConvGeom {
    int tiles_w, tiles_h;        // tiles per image along w / h (tiles along n = gridDim.x / (tiles_w*tiles_h))
    int bw, bh, bn;              // pixel box of one tile
    int k_blocks;                // reduction channels * sizeof(T) / 128  (K blocks per tap)
    int cout;                    // output channels (statistics stride)
    int n_tiles;                 // output-channel tiles per class (blockIdx.y = class * n_tiles + n_blk)
};

// shared memory carve-up (after 1024-byte alignment), sized so that TWO CTAs are resident per SM (one CTA's epilogue
// overlaps the other's main loop, and 296 slots take the 256 tiles of a 32x32x32 layer in a single wave):
//   [0, kStages*16K)                 A ring           (reused as the split-K staging tile after the main loop)
//   [.., + kStages*BN*128)           B ring
//   [.., + BM*BN*4)                  output tile: (BN*sizeof(T)/128) groups of [128 px][128 B], swizzled
//   statistics scratch, barriers, TMEM slot
constexpr int kStages = 3;

template <int BN>
struct SmemPlan {
    static constexpr int a_bytes = kStages * BM * kRowBytes;
    static constexpr int b_bytes = kStages * BN * kRowBytes;
    static constexpr int out_bytes = BM * BN * 4;                   // sized for fp32 output
    static constexpr int stage_pitch = BN * 4 + 16;                 // split-K staging row pitch (bank spread)
    static constexpr int stats_bytes = 4 * BN * 2 * 4;
    static constexpr int total = a_bytes + b_bytes + out_bytes + stats_bytes + 256 + 1024;
    static_assert(BM * stage_pitch <= a_bytes + b_bytes, "split-K staging tile must fit in the operand rings");
    static_assert(2 * (total + 1024) <= 227 * 1024, "two CTAs must fit in one SM's shared memory");
};

// kBMn: the B operand is the FORWARD filter matrix [Cout, taps*Cin] consumed MN-major (data gradient: n = input
// channel is the contiguous dimension, k = output channel the row index) — no permuted copy of the filters exists.
template <typename T, int BN, bool kBMn>
__global__ void __launch_bounds__(kThreads, 2)
conv_tap_gemm_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvClasses classes, const ConvGeom g,
                     float* __restrict__ stats /* [2][cout] or null */) {
    constexpr bool kTf32 = sizeof(T) == 4;
    constexpr int kOutGroups = BN * sizeof(T) / kRowBytes;          // 128-byte channel groups of the output tile
    constexpr int kColsPerGroup = kRowBytes / sizeof(T);            // 32 (fp32) / 64 (bf16)
    constexpr int kBBoxes = kBMn ? kOutGroups : 1;                  // MN-major B: one box per 128 bytes of n
    constexpr int kBBoxBytes = BN * kRowBytes / kBBoxes;
    using Plan = SmemPlan<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Plan::a_bytes;
    uint8_t* smem_out = smem_b + Plan::b_bytes;
    float* smem_stats = reinterpret_cast<float*>(smem_out + Plan::out_bytes);       // [4 warps][2][BN]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_stats) + Plan::stats_bytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int splits = gridDim.z, split = blockIdx.z;              // cluster = (1, 1, splits)
    const int tile = blockIdx.x;
    const int tw = tile % g.tiles_w, th = (tile / g.tiles_w) % g.tiles_h, tn = tile / (g.tiles_w * g.tiles_h);
    const int w0 = tw * g.bw, h0 = th * g.bh, n0 = tn * g.bn;
    const int cls = blockIdx.y / g.n_tiles, n_blk = blockIdx.y - cls * g.n_tiles;
    const TapTable& taps = classes.taps[cls];

    const int total_kb = taps.ntaps * g.k_blocks;
    const int per = (total_kb + splits - 1) / splits;
    const int kb_begin = split * per;
    const int kb_end = min(total_kb, kb_begin + per);
    const int num_kb = max(0, kb_end - kb_begin);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a[0])) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.b)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.out[cls])) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_wait();                                                     // inputs written by the previous kernel are visible
    pdl_trigger();                                                  // the next kernel may begin its prologue

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            for (int i = 0; i < num_kb; ++i) {
                const int kb = kb_begin + i;
                const int t = kb / g.k_blocks, cb = kb - t * g.k_blocks;
                const int stage = i % kStages;
                mbar_wait(empty_bar + stage, ((i / kStages) & 1) ^ 1);
                mbar_expect_tx(full_bar + stage, (BM + BN) * kRowBytes);
                const int c0 = cb * kColsPerGroup;
                tma_load_4d(smem_a + stage * BM * kRowBytes, &maps.a[taps.map[t]], c0, w0 + taps.dw[t], h0 + taps.dh[t], n0,
                            full_bar + stage);
                uint8_t* b_dst = smem_b + stage * BN * kRowBytes;
                if constexpr (kBMn) {
#pragma unroll
                    for (int bx = 0; bx < kBBoxes; ++bx)            // (128 B of n) x (K-block rows of k)
                        tma_load_2d(b_dst + bx * kBBoxBytes, &maps.b, taps.wcol[t] + n_blk * BN + bx * kColsPerGroup, c0,
                                    full_bar + stage);
                } else {
                    tma_load_2d(b_dst, &maps.b, taps.wcol[t] + c0, n_blk * BN, full_bar + stage);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc(kTf32, BM, BN, false, kBMn);
            // MN-major B (see make_desc): k rows of 128 B; 32-bit types use the 32-byte-atom layout (4-row atoms)
            constexpr uint32_t b_layout = (kBMn && kTf32) ? 1u : 2u, b_sbo = (kBMn && kTf32) ? 512u : 1024u;
            constexpr uint32_t b_lbo = kBMn ? (uint32_t)kBBoxBytes : 16u;
            constexpr uint32_t b_step = kBMn ? (kTf32 ? 8u : 16u) * kRowBytes : 32u;  // bytes of B consumed per MMA
            for (int i = 0; i < num_kb; ++i) {
                const int stage = i % kStages;
                mbar_wait(full_bar + stage, (i / kStages) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + stage * BM * kRowBytes);
                const uint32_t b_addr = smem_u32(smem_b + stage * BN * kRowBytes);
#pragma unroll
                for (int k = 0; k < kRowBytes / 32; ++k) {          // one MMA consumes 32 bytes of K per A row
                    umma<kTf32>(tmem_base, make_desc(a_addr + k * 32, 16, 1024), make_desc(b_addr + k * b_step, b_lbo, b_sbo, b_layout),
                                idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar + stage);
            }
            umma_commit(tmem_full_bar);
        }
    }

    // ===== epilogue (warps 4-7) =====
    const bool epi = warp >= 4;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                            // pixel of this thread inside the tile
    const int et = threadIdx.x - 128;                               // 0..127 within the epilogue group
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t stage_base = smem_u32(smem);                     // operand rings are idle once tmem_full fired
    const uint32_t out_base = smem_u32(smem_out);
    if (epi && num_kb > 0) {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
    }
    int rows_here = BM;                                             // rows of the tile this CTA finalises and stores
    int slice_w = 0, slice_h = 0, slice_n = 0;                      // their offset inside the tile's pixel box
    if (splits > 1) {
        // split-K: reduce-scatter over distributed shared memory.  Every CTA parks its fp32 partial tile in its own
        // smem; after the cluster barrier CTA r sums rows [r*BM/S, (r+1)*BM/S) of all S tiles (its own locally, the others
        // through ld.shared::cluster, all loads of an element in flight together), then finishes that slice like a
        // whole tile: swizzled smem -> statistics -> TMA store of the sub-box.
        if (epi) {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                if (num_kb > 0) tmem_ld_32x32b_x32(tmem_row + c0, v);
                else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    sts128(stage_base + row * Plan::stage_pitch + (c0 + j) * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        __syncwarp();
        cluster_sync_all();                                          // partial tiles visible cluster-wide
        rows_here = BM / splits;
        const int m0 = split * rows_here;
        slice_n = m0 / (g.bh * g.bw);
        slice_h = (m0 / g.bw) % g.bh;
        slice_w = m0 % g.bw;
        if (epi) {
            const int items = rows_here * (BN / 4);                  // float4 elements of this CTA's slice
            for (int it = et; it < items; it += 128) {
                const int lrow = it / (BN / 4), c4 = it - lrow * (BN / 4);
                const uint32_t off = (m0 + lrow) * Plan::stage_pitch + c4 * 16;
                float4 part[kMaxSplit];
#pragma unroll
                for (int p = 0; p < kMaxSplit; ++p)
                    if (p < splits) part[p] = ld_dsmem128f(mapa(stage_base + off, p));
                float4 a = part[0];
#pragma unroll
                for (int p = 1; p < kMaxSplit; ++p)                 // fixed order: deterministic sum
                    if (p < splits) { a.x += part[p].x; a.y += part[p].y; a.z += part[p].z; a.w += part[p].w; }
                const int col = c4 * 4;
                if constexpr (kTf32) {
                    const uint32_t addr = out_base + (col / 32) * (rows_here * kRowBytes) + lrow * kRowBytes +
                                          ((((col % 32) / 4) ^ (lrow & 7)) << 4);
                    sts128(addr, __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w));
                } else {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
                    const uint32_t addr = out_base + (col / 64) * (rows_here * kRowBytes) + lrow * kRowBytes +
                                          ((((col % 64) / 8) ^ (lrow & 7)) << 4) + ((col % 8) / 4) * 8;
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t*>(&lo)),
                                 "r"(*reinterpret_cast<uint32_t*>(&hi)) : "memory");
                }
            }
        }
    } else if (epi) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            if (num_kb > 0) tmem_ld_32x32b_x32(tmem_row + c0, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            // registers -> swizzled output tile: a group holds kColsPerGroup channels of all 128 pixels as
            // [row][128 B] with the 16-byte chunk index XORed by (row & 7)  (= CU_TENSOR_MAP_SWIZZLE_128B)
            if constexpr (kTf32) {
                const uint32_t row_addr = out_base + (c0 / 32) * (BM * kRowBytes) + row * kRowBytes;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    sts128(row_addr + ((ch ^ (row & 7)) << 4), v[ch * 4], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
            } else {
                const uint32_t row_addr = out_base + (c0 / 64) * (BM * kRowBytes) + row * kRowBytes;
                const int ch0 = (c0 % 64) / 8;                       // 32 bf16 = four 16-byte chunks of the row
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[ch * 8 + 2 * e]), __uint_as_float(v[ch * 8 + 2 * e + 1]));
                        w[e] = *reinterpret_cast<uint32_t*>(&h);
                    }
                    sts128(row_addr + (((ch0 + ch) ^ (row & 7)) << 4), w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    if (epi) {
        epi_bar_sync();                                              // the (slice of the) tile is complete in smem_out
        if (stats != nullptr && et < BN) {
            // per-channel sum / sum of squares of the STORED values over this CTA's rows: thread c walks column c
            // (32-bit word index inside the row; consecutive threads hit consecutive banks)
            const int grp = kTf32 ? et / 32 : et / 64;
            const int word = kTf32 ? et % 32 : (et % 64) / 2;
            const uint32_t base = out_base + grp * (rows_here * kRowBytes);
            float s0 = 0.f, q0 = 0.f;
            for (int r = 0; r < rows_here; ++r) {
                const uint32_t wv = lds32(base + r * kRowBytes + ((((word >> 2) ^ (r & 7)) << 4) | ((word & 3) << 2)));
                const float x = kTf32 ? __uint_as_float(wv) : __uint_as_float((et & 1) ? (wv & 0xFFFF0000u) : (wv << 16));
                s0 += x;
                q0 = fmaf(x, x, q0);
            }
            atomicAdd(stats + n_blk * BN + et, s0);
            atomicAdd(stats + g.cout + n_blk * BN + et, q0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA unit
        epi_bar_sync();
        if (et == 0) {
#pragma unroll
            for (int gidx = 0; gidx < kOutGroups; ++gidx)
                tma_store_4d(&maps.out[cls], smem_out + gidx * (rows_here * kRowBytes), n_blk * BN + gidx * kColsPerGroup,
                             w0 + slice_w, h0 + slice_h, n0 + slice_n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the store's reads
        }
    }
    __syncwarp();
    if (splits > 1) cluster_sync_all();                              // peers' staging tiles stay mapped until consumed
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient:  dW[co, tap, ci] = sum over output pixels p  dy[p, co] * x[p + shift(tap), ci]
//
// The reduction runs over PIXELS, which is the slow dimension of both NHWC tensors — so both operands are consumed
// MN-major: a TMA box (128 B of channels, pw, ph, pn) lands as [64 pixels][128 B] rows; 8 consecutive pixel rows are one
// MN-major SWIZZLE_128B atom (8 k x 128 B along M/N), the next 128-byte channel chunk is the next box (LBO = 8 KB), the
// next 8 pixels follow 1 KB later.  No transposed copy of x or dy is ever materialised.
//   tile: 128 output channels (M) x 64 input channels (N) for ONE tap; K block = 64 pixels; grid.z = split of the pixel
//   range over a cluster, reduced in the leader through DSMEM like the forward kernel; TMA-store epilogue into the
//   [Cout, taps*Cin] gradient matrix (= channels_last [Cout, R, S, Cin]).
// ---------------------------------------------------------------------------------------------------------------
constexpr int WG_M = 128, WG_N = 64, WG_KP = 64, WG_STAGES = 3;

struct WgradMaps {
    CUtensorMap x[4];            // activation sub-lattices (as the forward A maps, 64-pixel box)
    CUtensorMap dy;              // output gradient (C, Wo, Ho, N), 64-pixel box
    CUtensorMap dw;              // gradient matrix [Cout, taps*Cin], box (128 B, 128 rows), SWIZZLE_128B
};

struct WgradGeom {
    int tiles_w, tiles_h, tiles_n;   // 64-pixel boxes per plane / images
    int pw, ph, pn;
    int ntaps;
};

template <typename T>
struct WgradPlan {
    static constexpr int a_box = WG_KP * kRowBytes;                               // 8 KB
    static constexpr int a_boxes = WG_M * sizeof(T) / kRowBytes;                  // 4 (fp32) / 2 (bf16)
    static constexpr int b_boxes = WG_N * sizeof(T) / kRowBytes;                  // 2 / 1
    static constexpr int stage_bytes = (a_boxes + b_boxes) * a_box;
    static constexpr int ring_bytes = WG_STAGES * stage_bytes;
    static constexpr int out_bytes = WG_M * WG_N * 4;
    static constexpr int stage_pitch = WG_N * 4 + 16;
    static constexpr int total = ring_bytes + out_bytes + 256 + 1024;
    static_assert(WG_M * stage_pitch <= ring_bytes, "split-K staging tile must fit in the operand ring");
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WgradMaps maps, const TapTable taps, const WgradGeom g) {
    constexpr bool kTf32 = sizeof(T) == 4;
    using Plan = WgradPlan<T>;
    constexpr int kColsPerGroup = kRowBytes / sizeof(T);
    constexpr int kOutGroups = WG_N * sizeof(T) / kRowBytes;
    constexpr int kRowsPerMma = kTf32 ? 8 : 16;                     // pixels (k) consumed by one MMA
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* smem_out = smem + Plan::ring_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_out + Plan::out_bytes);
    uint64_t* empty_bar = full_bar + WG_STAGES;
    uint64_t* tmem_full_bar = empty_bar + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int splits = gridDim.z, split = blockIdx.z;
    const int co_blk = blockIdx.x;
    const int tap = blockIdx.y % g.ntaps, ci_blk = blockIdx.y / g.ntaps;

    const int total_kb = g.tiles_w * g.tiles_h * g.tiles_n;
    const int per = (total_kb + splits - 1) / splits;
    const int kb_begin = split * per;
    const int num_kb = max(0, min(total_kb, kb_begin + per) - kb_begin);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.x[0])) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.dy)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.dw)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < WG_STAGES; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(WG_N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_wait();                                                     // inputs written by the previous kernel are visible
    pdl_trigger();                                                  // the next kernel may begin its prologue

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            const CUtensorMap* xmap = &maps.x[taps.map[tap]];
            for (int i = 0; i < num_kb; ++i) {
                const int kb = kb_begin + i;
                const int tw = kb % g.tiles_w, th = (kb / g.tiles_w) % g.tiles_h, tn = kb / (g.tiles_w * g.tiles_h);
                const int w0 = tw * g.pw, h0 = th * g.ph, n0 = tn * g.pn;
                const int stage = i % WG_STAGES;
                mbar_wait(empty_bar + stage, ((i / WG_STAGES) & 1) ^ 1);
                mbar_expect_tx(full_bar + stage, Plan::stage_bytes);
                uint8_t* a_dst = smem + stage * Plan::stage_bytes;
                uint8_t* b_dst = a_dst + Plan::a_boxes * Plan::a_box;
#pragma unroll
                for (int bx = 0; bx < Plan::a_boxes; ++bx)           // channels past Cout are zero-filled by the TMA unit
                    tma_load_4d(a_dst + bx * Plan::a_box, &maps.dy, co_blk * WG_M + bx * kColsPerGroup, w0, h0, n0, full_bar + stage);
#pragma unroll
                for (int bx = 0; bx < Plan::b_boxes; ++bx)
                    tma_load_4d(b_dst + bx * Plan::a_box, xmap, ci_blk * WG_N + bx * kColsPerGroup, w0 + taps.dw[tap], h0 + taps.dh[tap],
                                n0, full_bar + stage);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer: both operands MN-major =====
            constexpr uint32_t idesc = make_idesc(kTf32, WG_M, WG_N, true, true);
            for (int i = 0; i < num_kb; ++i) {
                const int stage = i % WG_STAGES;
                mbar_wait(full_bar + stage, (i / WG_STAGES) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * Plan::stage_bytes);
                const uint32_t b_addr = a_addr + Plan::a_boxes * Plan::a_box;
#pragma unroll
                for (int k = 0; k < WG_KP / kRowsPerMma; ++k) {
                    const uint32_t off = k * kRowsPerMma * kRowBytes;
                    constexpr uint32_t layout = kTf32 ? 1u : 2u, sbo = kTf32 ? 512u : 1024u;
                    umma<kTf32>(tmem_base, make_desc(a_addr + off, Plan::a_box, sbo, layout),
                                make_desc(b_addr + off, Plan::a_box, sbo, layout), idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar + stage);
            }
            umma_commit(tmem_full_bar);
        }
    }

    const bool epi = warp >= 4;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                            // output channel (row of dW) inside the tile
    const int et = threadIdx.x - 128;
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t stage_base = smem_u32(smem);
    const uint32_t out_base = smem_u32(smem_out);
    if (epi && num_kb > 0) {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
    }
    int rows_here = WG_M, row0 = 0;
    if (splits > 1) {                                               // reduce-scatter over DSMEM (see conv_tap_gemm_kernel)
        if (epi) {
#pragma unroll 1
            for (int c0 = 0; c0 < WG_N; c0 += 32) {
                uint32_t v[32];
                if (num_kb > 0) tmem_ld_32x32b_x32(tmem_row + c0, v);
                else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    sts128(stage_base + row * Plan::stage_pitch + (c0 + j) * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        __syncwarp();
        cluster_sync_all();
        rows_here = WG_M / splits;
        row0 = split * rows_here;
        if (epi) {
            const int items = rows_here * (WG_N / 4);
            for (int it = et; it < items; it += 128) {
                const int lrow = it / (WG_N / 4), c4 = it - lrow * (WG_N / 4);
                const uint32_t off = (row0 + lrow) * Plan::stage_pitch + c4 * 16;
                float4 part[kMaxSplit];
#pragma unroll
                for (int p = 0; p < kMaxSplit; ++p)
                    if (p < splits) part[p] = ld_dsmem128f(mapa(stage_base + off, p));
                float4 a = part[0];
#pragma unroll
                for (int p = 1; p < kMaxSplit; ++p)
                    if (p < splits) { a.x += part[p].x; a.y += part[p].y; a.z += part[p].z; a.w += part[p].w; }
                const int col = c4 * 4;
                if constexpr (kTf32) {
                    const uint32_t addr = out_base + (col / 32) * (rows_here * kRowBytes) + lrow * kRowBytes +
                                          ((((col % 32) / 4) ^ (lrow & 7)) << 4);
                    sts128(addr, __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w));
                } else {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
                    const uint32_t addr = out_base + (col / 64) * (rows_here * kRowBytes) + lrow * kRowBytes +
                                          ((((col % 64) / 8) ^ (lrow & 7)) << 4) + ((col % 8) / 4) * 8;
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t*>(&lo)),
                                 "r"(*reinterpret_cast<uint32_t*>(&hi)) : "memory");
                }
            }
        }
    } else if (epi) {
#pragma unroll 1
        for (int c0 = 0; c0 < WG_N; c0 += 32) {
            uint32_t v[32];
            if (num_kb > 0) tmem_ld_32x32b_x32(tmem_row + c0, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if constexpr (kTf32) {
                const uint32_t row_addr = out_base + (c0 / 32) * (WG_M * kRowBytes) + row * kRowBytes;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    sts128(row_addr + ((ch ^ (row & 7)) << 4), v[ch * 4], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
            } else {
                const uint32_t row_addr = out_base + (c0 / 64) * (WG_M * kRowBytes) + row * kRowBytes;
                const int ch0 = (c0 % 64) / 8;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[ch * 8 + 2 * e]), __uint_as_float(v[ch * 8 + 2 * e + 1]));
                        w[e] = *reinterpret_cast<uint32_t*>(&h);
                    }
                    sts128(row_addr + (((ch0 + ch) ^ (row & 7)) << 4), w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    if (epi) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        epi_bar_sync();
        if (et == 0) {
#pragma unroll
            for (int gidx = 0; gidx < kOutGroups; ++gidx)
                tma_store_2d(&maps.dw, smem_out + gidx * (rows_here * kRowBytes), taps.wcol[tap] + ci_blk * WG_N + gidx * kColsPerGroup,
                             co_blk * WG_M + row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    __syncwarp();
    if (splits > 1) cluster_sync_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(WG_N));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------------------------
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

// 4-D map over an NHWC tensor (or a strided sub-lattice of it): dims (C, W, H, N) in elements, strides in elements.
CUresult nhwc_map(CUtensorMap* map, const void* base, int esize, int C, int W, int H, int N, int64_t sw, int64_t sh, int64_t sn,
                  int box_c, int bw, int bh, int bn, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn encode = encode_fn();
    if (encode == nullptr) return CUDA_ERROR_NOT_SUPPORTED;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)(sw * esize), (cuuint64_t)(sh * esize), (cuuint64_t)(sn * esize)};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode(map, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

CUresult matrix_map(CUtensorMap* map, const void* base, int esize, int64_t rows, int64_t cols, int box_rows, int box_cols,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn encode = encode_fn();
    if (encode == nullptr) return CUDA_ERROR_NOT_SUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(cols * esize)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return encode(map, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// Pixel box of one 128-pixel tile for an (Ho, Wo) output plane: as wide as possible, then rows, then images.
bool tile_box(int Ho, int Wo, int* bw, int* bh, int* bn) {
    int w = Wo < BM ? Wo : BM;
    if (w <= 0 || (w & (w - 1)) != 0) return false;                 // power of two up to 128
    int h = BM / w;
    if (h > Ho) h = Ho;
    if ((h & (h - 1)) != 0 || Ho % h != 0 || Wo % w != 0) return false;
    int n = BM / (w * h);
    if (n < 1 || w * h * n != BM) return false;
    *bw = w; *bh = h; *bn = n;
    return true;
}

bool use_pdl() {
    static int on = -1;
    if (on < 0) {
        const char* v = getenv("FL4H_PDL");
        on = (v != nullptr && v[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

template <typename T, int BN, bool kBMn>
cudaError_t launch_tap_gemm(const ConvMaps& maps, const ConvClasses& classes, const ConvGeom& g, float* stats, int m_tiles,
                            int y_tiles, int splits, cudaStream_t stream) {
    auto kernel = conv_tap_gemm_kernel<T, BN, kBMn>;
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemPlan<BN>::total);
        if (err != cudaSuccess) return err;
        configured = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(m_tiles, y_tiles, splits);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = SmemPlan<BN>::total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = splits;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, maps, classes, g, stats);
}

template <typename T>
cudaError_t launch_wgrad(const WgradMaps& maps, const TapTable& taps, const WgradGeom& g, int co_tiles, int ci_tiles, int splits,
                         cudaStream_t stream) {
    auto kernel = conv_wgrad_kernel<T>;
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradPlan<T>::total);
        if (err != cudaSuccess) return err;
        configured = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(co_tiles, ci_tiles * g.ntaps, splits);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = WgradPlan<T>::total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = splits;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, maps, taps, g);
}

bool pixel_box(int Ho, int Wo, int total, int* bw, int* bh, int* bn) {
    int w = Wo < total ? Wo : total;
    if (w <= 0 || (w & (w - 1)) != 0) return false;
    int h = total / w;
    if (h > Ho) h = Ho;
    if ((h & (h - 1)) != 0 || Ho % h != 0 || Wo % w != 0) return false;
    int n = total / (w * h);
    if (n < 1 || w * h * n != total) return false;
    *bw = w; *bh = h; *bn = n;
    return true;
}

}  // namespace

extern "C" {

// Generic entry point.  dtype: 0 = fp32 tensors / TF32 math, 1 = bf16.
//   in        : NHWC activations; up to 4 sub-lattices (element offset, logical (hv, wv), strides (sw, sh), shared sn)
//   wmat      : filter matrix [wmat_rows, wmat_cols].  b_mn = 0: K-major, rows = output channels, tap t starts at column
//               wcol[t].  b_mn = 1 (data gradient): the forward matrix [Cout, taps*Cin] read MN-major: rows = reduction
//               channels, tap t's output-channel block starts at column wcol[t]
//   out       : NHWC output; class c writes the sub-lattice at element offset out_off[c] with strides (out_sw, out_sh,
//               out_sn) and logical plane (Ho, Wo); n_classes <= 4, class c has ntaps[c] taps at tap_*[c*9 ...]
//   stats     : optional [2, out_ch] fp32 accumulators (sum, sum of squares of the stored outputs), zero on entry
int fl4h_conv_tap_gemm(const void* in, const void* wmat, void* out, float* stats, int dtype, int N, int k_ch, int out_ch,
                       int wmat_rows, int wmat_cols, int b_mn, int n_maps, const long long* in_off, const int* in_hv,
                       const int* in_wv, const long long* in_sw, const long long* in_sh, long long in_sn, int Ho, int Wo,
                       int n_classes, const long long* out_off, long long out_sw, long long out_sh, long long out_sn,
                       const int* ntaps, const int* tap_dh, const int* tap_dw, const int* tap_map, const int* tap_wcol, int splits,
                       cudaStream_t stream) {
    const int esize = dtype == 0 ? 4 : 2;
    const int cpb = kRowBytes / esize;                              // channels per K block
    constexpr int BN = 64;
    if (k_ch % cpb != 0 || out_ch % BN != 0 || n_maps > 4 || n_maps < 1 || n_classes < 1 || n_classes > 4 || splits < 1 ||
        splits > kMaxSplit)
        return (int)cudaErrorInvalidValue;
    ConvGeom g;
    if (!tile_box(Ho, Wo, &g.bw, &g.bh, &g.bn)) return (int)cudaErrorInvalidValue;
    g.tiles_w = Wo / g.bw;
    g.tiles_h = Ho / g.bh;
    g.k_blocks = k_ch / cpb;
    g.cout = out_ch;
    g.n_tiles = out_ch / BN;
    const int tiles_n = (N + g.bn - 1) / g.bn;
    const int m_tiles = g.tiles_w * g.tiles_h * tiles_n;
    ConvMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int i = 0; i < n_maps; ++i) {
        const char* base = reinterpret_cast<const char*>(in) + in_off[i] * esize;
        if (nhwc_map(&maps.a[i], base, esize, k_ch, in_wv[i], in_hv[i], N, in_sw[i], in_sh[i], in_sn, cpb, g.bw, g.bh, g.bn) !=
            CUDA_SUCCESS)
            return (int)cudaErrorInvalidValue;
    }
    for (int i = n_maps; i < 4; ++i) maps.a[i] = maps.a[0];
    CUresult berr;
    if (b_mn)  // box = (128 B of output channels) x (one K block of reduction-channel rows)
        berr = matrix_map(&maps.b, wmat, esize, wmat_rows, wmat_cols, cpb, cpb,
                          dtype == 0 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
    else
        berr = matrix_map(&maps.b, wmat, esize, wmat_rows, wmat_cols, BN, cpb);
    if (berr != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
    // split-K clusters finish the tile as `splits` row slices (reduce-scatter): the store box is the slice's pixel box
    int min_kb = 1 << 30;
    for (int c = 0; c < n_classes; ++c) {
        const int kb = ntaps[c] * g.k_blocks;
        if (kb > 0 && kb < min_kb) min_kb = kb;
    }
    if (min_kb == (1 << 30)) min_kb = 1;
    while (splits > min_kb || (splits & (splits - 1)) != 0) --splits;   // power of two, at most the shortest K range
    int sub_w = g.bw, sub_h = g.bh, sub_n = g.bn;
    for (int rest = splits; rest > 1; rest >>= 1) {                     // halve the slowest dimension that still can be
        if (sub_n > 1) sub_n >>= 1;
        else if (sub_h > 1) sub_h >>= 1;
        else sub_w >>= 1;
    }
    if (sub_w < 1) return (int)cudaErrorInvalidValue;
    ConvClasses classes;
    memset(&classes, 0, sizeof(classes));
    for (int c = 0; c < n_classes; ++c) {
        char* obase = reinterpret_cast<char*>(out) + out_off[c] * esize;
        if (nhwc_map(&maps.out[c], obase, esize, out_ch, Wo, Ho, N, out_sw, out_sh, out_sn, cpb, sub_w, sub_h, sub_n) != CUDA_SUCCESS)
            return (int)cudaErrorInvalidValue;
        if (ntaps[c] < 0 || ntaps[c] > kMaxTaps) return (int)cudaErrorInvalidValue;
        TapTable& t = classes.taps[c];
        t.ntaps = ntaps[c];
        for (int i = 0; i < ntaps[c]; ++i) {
            t.dh[i] = (signed char)tap_dh[c * kMaxTaps + i];
            t.dw[i] = (signed char)tap_dw[c * kMaxTaps + i];
            t.map[i] = (signed char)tap_map[c * kMaxTaps + i];
            t.wcol[i] = tap_wcol[c * kMaxTaps + i];
        }
    }
    for (int c = n_classes; c < 4; ++c) maps.out[c] = maps.out[0];
    const int y_tiles = g.n_tiles * n_classes;
    cudaError_t err;
    if (dtype == 0)
        err = b_mn ? launch_tap_gemm<float, BN, true>(maps, classes, g, stats, m_tiles, y_tiles, splits, stream)
                   : launch_tap_gemm<float, BN, false>(maps, classes, g, stats, m_tiles, y_tiles, splits, stream);
    else
        err = b_mn ? launch_tap_gemm<__nv_bfloat16, BN, true>(maps, classes, g, stats, m_tiles, y_tiles, splits, stream)
                   : launch_tap_gemm<__nv_bfloat16, BN, false>(maps, classes, g, stats, m_tiles, y_tiles, splits, stream);
    return (int)err;
}

// Weight gradient.  x sub-lattices as in fl4h_conv_tap_gemm (their logical dims equal the output plane: the tap shift is
// applied in output-pixel units on the sub-lattice); dy is the dense NHWC output gradient [N, Ho, Wo, cout];
// dw is the [cout, dw_cols] gradient matrix, tap t is written at columns [wcol[t], wcol[t] + cin).
int fl4h_conv_wgrad(const void* x, const void* dy, void* dw, int dtype, int N, int cin, int cout, int dw_cols, int n_maps,
                    const long long* in_off, const int* in_hv, const int* in_wv, const long long* in_sw, const long long* in_sh,
                    long long in_sn, int Ho, int Wo, int ntaps, const int* tap_dh, const int* tap_dw, const int* tap_map,
                    const int* tap_wcol, int splits, cudaStream_t stream) {
    const int esize = dtype == 0 ? 4 : 2;
    const int cpb = kRowBytes / esize;
    if (cin % WG_N != 0 || cout % cpb != 0 || ntaps > kMaxTaps || ntaps < 1 || n_maps > 4 || splits < 1 || splits > kMaxSplit)
        return (int)cudaErrorInvalidValue;
    WgradGeom g;
    if (!pixel_box(Ho, Wo, WG_KP, &g.pw, &g.ph, &g.pn)) return (int)cudaErrorInvalidValue;
    g.tiles_w = Wo / g.pw;
    g.tiles_h = Ho / g.ph;
    g.tiles_n = (N + g.pn - 1) / g.pn;
    g.ntaps = ntaps;
    WgradMaps maps;
    memset(&maps, 0, sizeof(maps));
    // MN-major operands: 16-bit types use the plain 128B swizzle, 32-bit types the 32-byte-atom variant (see make_desc)
    const CUtensorMapSwizzle op_swizzle = dtype == 0 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    for (int i = 0; i < n_maps; ++i) {
        const char* base = reinterpret_cast<const char*>(x) + in_off[i] * esize;
        if (nhwc_map(&maps.x[i], base, esize, cin, in_wv[i], in_hv[i], N, in_sw[i], in_sh[i], in_sn, cpb, g.pw, g.ph, g.pn,
                     op_swizzle) != CUDA_SUCCESS)
            return (int)cudaErrorInvalidValue;
    }
    for (int i = n_maps; i < 4; ++i) maps.x[i] = maps.x[0];
    if (nhwc_map(&maps.dy, dy, esize, cout, Wo, Ho, N, cout, (int64_t)Wo * cout, (int64_t)Ho * Wo * cout, cpb, g.pw, g.ph, g.pn,
                 op_swizzle) != CUDA_SUCCESS)
        return (int)cudaErrorInvalidValue;
    const int total_kb = g.tiles_w * g.tiles_h * g.tiles_n;
    while (splits > total_kb || (splits & (splits - 1)) != 0) --splits;     // power of two: row slices of the tile
    if (matrix_map(&maps.dw, dw, esize, cout, dw_cols, WG_M / splits, cpb) != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
    TapTable taps;
    memset(&taps, 0, sizeof(taps));
    taps.ntaps = ntaps;
    for (int t = 0; t < ntaps; ++t) {
        taps.dh[t] = (signed char)tap_dh[t];
        taps.dw[t] = (signed char)tap_dw[t];
        taps.map[t] = (signed char)tap_map[t];
        taps.wcol[t] = tap_wcol[t];
    }
    const int co_tiles = (cout + WG_M - 1) / WG_M, ci_tiles = cin / WG_N;
    cudaError_t err = dtype == 0 ? launch_wgrad<float>(maps, taps, g, co_tiles, ci_tiles, splits, stream)
                                 : launch_wgrad<__nv_bfloat16>(maps, taps, g, co_tiles, ci_tiles, splits, stream);
    return (int)err;
}

}  // extern "C"
