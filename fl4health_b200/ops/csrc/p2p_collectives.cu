// Fused collective kernels over peer-mapped (NVLink 5 / NVSwitch) symmetric memory for sm_100a.
//
// The two communication hot paths of an FL round as ONE kernel each, issuing the peer loads/stores
// from inside the kernel and fusing the adjacent compute:
//
//   agg_fused   : reduce-scatter (weighted FedAvg sum, fixed rank order => bit-deterministic)
//                 -> strategy epilogue (FedAvg | FedAdam/FedAdagrad/FedYogi server step | SCAFFOLD server-lr | FedAvgM)
//                 -> all-gather (result slice stored into every rank's global buffer)
//   bcast_fused : scatter from the root + all-gather between peers (every NVLink carries 1/K of the payload),
//                 cross-GPU barrier, then the receiver-side unpack in the same kernel:
//                 w <- g, FedProx anchor w_t <- g, bf16 compute shadow <- g, SCAFFOLD d <- c - c_i.
//
// Two data paths, chosen per launch:
//   * NVLS (default when the symmetric segments are bound to an NVSwitch multicast object, see symm_mem.cu):
//       reduce-scatter = `multimem.ld_reduce.add.v4.f32` on the multicast address (the switch pulls the K copies and
//       returns their sum: K-1 fewer NVLink transfers into the GPU), all-gather / broadcast = `multimem.st.v4.f32`
//       (one store leaves the GPU, the switch replicates it into every rank's copy).  The switch can only ADD, so a
//       non-uniform weighted mean first scales this rank's contribution into a multicast-bound staging buffer
//       (phase A of the same kernel); uniform weights are applied after the reduction instead.
//   * P2P (deterministic mode, `FL4H_NVLS=0`, or no multicast support): fixed-rank-order `ld/st.volatile` over the
//       unicast peer mappings.
// The integer model buffers (BatchNorm `num_batches_tracked`) are reduced by the same launch (one warp, P2P loads).
//
// Synchronisation is done with flags in peer-mapped memory (release/acquire at system scope) + a grid barrier, so no
// host round trip and no NCCL call sits between the local training kernels and the aggregate.  Kernels are launched
// cooperatively (all CTAs co-resident: the grid barrier cannot deadlock).
//
// What the reference does for the same step: K gRPC messages of np.save bytes into a CPU server, reduce(np.add) per
// layer, K gRPC messages back (fl4health/strategies/aggregate_utils.py:23-32, parameter_exchange/full_exchanger.py).

#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace cg = cooperative_groups;

#define FL4H_MAX_RANKS 16

namespace {

constexpr int kThreads = 512;

struct PeerTable {
    float* contrib[FL4H_MAX_RANKS];     // each rank's contribution buffer (its client arena flat)
    float* result[FL4H_MAX_RANKS];      // each rank's global/result buffer
    uint32_t* flags[FL4H_MAX_RANKS];    // each rank's signal pad: [phase][FL4H_MAX_RANKS] epochs
    float coef[FL4H_MAX_RANKS];
    int rank;
    int world;
    // NVLS: multicast addresses of the contribution source and of the result buffer; `stage` is this rank's unicast
    // view of the multicast-bound staging buffer (null: contributions are reduced in place, weights are uniform)
    const float* mc_contrib;
    float* mc_result;
    float* stage;
    // integer buffers riding along (null when the model has none): ibuf[r] = rank r's int64 values (peer-mapped)
    const long long* ibuf[FL4H_MAX_RANKS];
    long long* ibuf_out;
    int n_int;
};

struct EpiArgs {
    float eta, beta1, beta2, tau, server_lr, momentum;
    int mode;  // same codes as flat_ops.cu
};
enum { EPI_NONE = 0, EPI_FEDADAM = 1, EPI_FEDADAGRAD = 2, EPI_FEDYOGI = 3, EPI_SERVER_LR = 4, EPI_MOMENTUM = 5 };

__device__ __forceinline__ float4 ld_peer4(const float* p) {
    // volatile => relaxed.sys semantics: never served from a stale L1 line, goes to the owner over NVLink
    float4 r;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_peer4(float* p, float4 v) {
    asm volatile("st.volatile.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Phase timeline of CTA 0 of the most recent collective kernel (global timer, ns): read back with
// fl4h_coll_debug_read.  [0] entry [1] staging pass done [2] peers arrived [3] grid released [4] main loop done
// [5] stores fenced + grid joined [6] peers finished [7] unpack done
__device__ unsigned long long g_coll_stamp[8];
__device__ __forceinline__ void coll_stamp(int slot) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_coll_stamp[slot] = t;
    }
}

// NVLS: in-switch reduction / replication on a multicast address (sm_90+; SASS: MULTIMEM / RED..MULTIMEM forms)
__device__ __forceinline__ float4 mm_ld_reduce4(const float* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(mc)
                 : "memory");
    return r;
}
__device__ __forceinline__ void mm_st4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// Cross-GPU barrier executed by CTA 0 (threads 0..world-1), bracketed by grid barriers by the caller.
// flags layout on every rank: flags[phase * FL4H_MAX_RANKS + src_rank] = epoch written by src_rank.
__device__ __forceinline__ void peer_barrier(const PeerTable& t, int phase, uint32_t epoch) {
    if (blockIdx.x == 0 && threadIdx.x < t.world) {
        const int peer = threadIdx.x;
        __threadfence_system();
        st_release_sys(t.flags[peer] + phase * FL4H_MAX_RANKS + t.rank, epoch);
        const uint32_t* mine = t.flags[t.rank] + phase * FL4H_MAX_RANKS + peer;
        while (ld_acquire_sys(mine) < epoch) { __nanosleep(64); }
    }
}

// Entry barrier without a grid-wide join: CTA 0 announces this rank (everything this rank contributes was written before
// the kernel started, or before the grid.sync preceding this call), EVERY CTA polls this rank's own flag row until all
// peers have announced.  Saves one grid.sync (~2-3 us of a ~25 us fixed cost) per collective.
__device__ __forceinline__ void peer_barrier_enter(const PeerTable& t, int phase, uint32_t epoch) {
    if (threadIdx.x < t.world) {
        const int peer = threadIdx.x;
        if (blockIdx.x == 0) {
            __threadfence_system();
            st_release_sys(t.flags[peer] + phase * FL4H_MAX_RANKS + t.rank, epoch);
        }
        const uint32_t* mine = t.flags[t.rank] + phase * FL4H_MAX_RANKS + peer;
        while (ld_acquire_sys(mine) < epoch) { __nanosleep(32); }
    }
    __syncthreads();
}

__device__ __forceinline__ float epi_apply(int mode, const EpiArgs& ea, float avg, float wcur, float& m, float& v) {
    switch (mode) {
        case EPI_FEDADAM: {
            float d = avg - wcur;
            m = ea.beta1 * m + (1.f - ea.beta1) * d;
            v = ea.beta2 * v + (1.f - ea.beta2) * d * d;
            return wcur + ea.eta * m / (sqrtf(v) + ea.tau);
        }
        case EPI_FEDADAGRAD: {
            float d = avg - wcur;
            m = ea.beta1 * m + (1.f - ea.beta1) * d;
            v = v + d * d;
            return wcur + ea.eta * m / (sqrtf(v) + ea.tau);
        }
        case EPI_FEDYOGI: {
            float d = avg - wcur;
            m = ea.beta1 * m + (1.f - ea.beta1) * d;
            float d2 = d * d;
            float sgn = (v - d2) > 0.f ? 1.f : ((v - d2) < 0.f ? -1.f : 0.f);
            v = v - (1.f - ea.beta2) * d2 * sgn;
            return wcur + ea.eta * m / (sqrtf(v) + ea.tau);
        }
        case EPI_SERVER_LR:
            return wcur + ea.server_lr * (avg - wcur);
        case EPI_MOMENTUM:
            m = ea.momentum * m + avg;
            return wcur + ea.server_lr * m;
        default:
            return avg;
    }
}

// Integer buffers (a handful of int64 counters): weighted mean over ranks, truncated like the reference's
// float-average -> int64 load (fl4health/parameter_exchange/full_exchanger.py:45-47).  One warp of CTA 0, P2P loads.
// Must run between the two peer barriers (peers' values final, nobody has overwritten them yet).
__device__ __forceinline__ void reduce_int_buffers(const PeerTable& t) {
    if (t.n_int == 0 || blockIdx.x != 0) return;
    for (int i = threadIdx.x; i < t.n_int; i += blockDim.x) {
        double acc = 0.0;
        for (int k = 0; k < t.world; ++k) {
            long long v;
            asm volatile("ld.volatile.global.s64 %0, [%1];" : "=l"(v) : "l"(t.ibuf[k] + i));
            acc += (double)t.coef[k] * (double)v;
        }
        t.ibuf_out[i] = (long long)acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// agg_fused: every rank owns slice [rank*slice, (rank+1)*slice) of the flat payload.
//   wcur/m/v are LOCAL full-length buffers (only the owned slice is touched): server optimizer state is sharded by
//   slice ownership, the updated weights are what gets all-gathered.
// ---------------------------------------------------------------------------------------------------------------
template <int kWorld>
__global__ void __launch_bounds__(kThreads, 1)
agg_fused_kernel(PeerTable t, const float* __restrict__ wcur, float* __restrict__ m, float* __restrict__ v,
                 EpiArgs ea, int64_t numel, int64_t slice, uint32_t epoch) {
    cg::grid_group grid = cg::this_grid();
    // (0) all ranks finished local training and are inside the kernel: contributions may be read.
    peer_barrier_enter(t, 0, epoch);
    reduce_int_buffers(t);

    const int64_t begin = (int64_t)t.rank * slice;
    int64_t end = begin + slice;
    if (end > numel) end = numel;
    const int64_t n4 = (end > begin) ? ((end - begin) >> 2) : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // kUnroll x kWorld 16-byte peer loads are issued before any is consumed: the NVLink round trip (~2 us) is
    // amortised over kUnroll*kWorld*16 B per thread (volatile accesses are never reordered by the compiler, so the
    // memory-level parallelism has to be explicit).
    constexpr int kUnroll = (kWorld <= 2) ? 8 : (kWorld <= 4 ? 4 : 2);
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += stride * kUnroll) {
        float4 vals[kUnroll][kWorld];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t i = base + (int64_t)u * stride;
            if (i < n4) {
                const int64_t e = begin + (i << 2);
#pragma unroll
                for (int k = 0; k < kWorld; ++k) vals[u][k] = ld_peer4(t.contrib[k] + e);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t i = base + (int64_t)u * stride;
            if (i >= n4) continue;
            const int64_t e = begin + (i << 2);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < kWorld; ++k) {                                    // fixed order => deterministic
                const float c = t.coef[k];
                acc.x = fmaf(c, vals[u][k].x, acc.x); acc.y = fmaf(c, vals[u][k].y, acc.y);
                acc.z = fmaf(c, vals[u][k].z, acc.z); acc.w = fmaf(c, vals[u][k].w, acc.w);
            }
            if (ea.mode != EPI_NONE) {
                float4 wv = *reinterpret_cast<const float4*>(wcur + e);
                float4 mv = m ? *reinterpret_cast<const float4*>(m + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 vv = v ? *reinterpret_cast<const float4*>(v + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                acc.x = epi_apply(ea.mode, ea, acc.x, wv.x, mv.x, vv.x);
                acc.y = epi_apply(ea.mode, ea, acc.y, wv.y, mv.y, vv.y);
                acc.z = epi_apply(ea.mode, ea, acc.z, wv.z, mv.z, vv.z);
                acc.w = epi_apply(ea.mode, ea, acc.w, wv.w, mv.w, vv.w);
                if (m) *reinterpret_cast<float4*>(m + e) = mv;
                if (v) *reinterpret_cast<float4*>(v + e) = vv;
            }
#pragma unroll
            for (int k = 0; k < kWorld; ++k) st_peer4(t.result[k] + e, acc);      // all-gather: push to every rank
        }
    }
    // (1) my slice is stored everywhere; wait until every peer's slice landed here before the kernel may complete.
    __threadfence_system();
    grid.sync();
    peer_barrier(t, 1, epoch);
}

// Receiver-side unpack shared by both broadcast kernels: one read of the landed global buffer, up to four writes.
__device__ __forceinline__ void unpack_landed(const float* g, float* __restrict__ w, float* __restrict__ anchor,
                                              __nv_bfloat16* __restrict__ shadow, const float* __restrict__ c_server,
                                              const float* __restrict__ c_local, float* __restrict__ cv_out,
                                              int64_t numel) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total4 = numel >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
        const int64_t e = i << 2;
        const float4 gv = ld_peer4(g + e);  // volatile: written by peers during this kernel, never in L1
        if (w) *reinterpret_cast<float4*>(w + e) = gv;
        if (anchor) *reinterpret_cast<float4*>(anchor + e) = gv;
        if (shadow) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(gv.x, gv.y), hi = __floats2bfloat162_rn(gv.z, gv.w);
            uint2 packed;
            packed.x = *reinterpret_cast<uint32_t*>(&lo);
            packed.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(shadow + e) = packed;
        }
        if (cv_out) {
            const float4 cs = *reinterpret_cast<const float4*>(c_server + e);
            const float4 cl = *reinterpret_cast<const float4*>(c_local + e);
            *reinterpret_cast<float4*>(cv_out + e) = make_float4(cs.x - cl.x, cs.y - cl.y, cs.z - cl.z, cs.w - cl.w);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// bcast_fused: root's `contrib[root]` buffer -> every rank's `result` buffer, then receiver-side unpack.
// ---------------------------------------------------------------------------------------------------------------
template <int kWorld>
__global__ void __launch_bounds__(kThreads, 1)
bcast_fused_kernel(PeerTable t, int root, float* __restrict__ w, float* __restrict__ anchor,
                   __nv_bfloat16* __restrict__ shadow, const float* __restrict__ c_server,
                   const float* __restrict__ c_local, float* __restrict__ cv_out, int64_t numel, int64_t slice,
                   uint32_t epoch) {
    cg::grid_group grid = cg::this_grid();
    peer_barrier_enter(t, 0, epoch);  // root's buffer is final; everyone's result buffer may be overwritten

    const int64_t begin = (int64_t)t.rank * slice;
    int64_t end = begin + slice;
    if (end > numel) end = numel;
    const int64_t n4 = (end > begin) ? ((end - begin) >> 2) : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float* src = t.contrib[root];
    constexpr int kUnroll = 8;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += stride * kUnroll) {
        float4 vals[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t i = base + (int64_t)u * stride;
            if (i < n4) vals[u] = ld_peer4(src + begin + (i << 2));               // scatter: my 1/K from the root
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t i = base + (int64_t)u * stride;
            if (i >= n4) continue;
            const int64_t e = begin + (i << 2);
#pragma unroll
            for (int k = 0; k < kWorld; ++k) st_peer4(t.result[k] + e, vals[u]);  // all-gather to every peer
        }
    }
    __threadfence_system();
    grid.sync();
    peer_barrier(t, 1, epoch);
    grid.sync();

    unpack_landed(t.result[t.rank], w, anchor, shadow, c_server, c_local, cv_out, numel);
}


// ---------------------------------------------------------------------------------------------------------------
// agg_nvls: the same contract as agg_fused, data path through the NVSwitch multicast object.
//   phase A (non-uniform weights only): stage <- coef[rank] * contrib            (local HBM, full length)
//   barrier 0
//   phase B: slice owner: acc = multimem.ld_reduce(mc_contrib + e)  [* coef if uniform] -> epilogue
//            -> multimem.st(mc_result + e)                                       (lands in every rank's result)
//   barrier 1
// ---------------------------------------------------------------------------------------------------------------
template <int kUnroll>
__global__ void __launch_bounds__(kThreads, 1)
agg_nvls_kernel(PeerTable t, const float* __restrict__ wcur, float* __restrict__ m, float* __restrict__ v,
                EpiArgs ea, int64_t numel, int64_t slice, uint32_t epoch, float uniform_coef) {
    cg::grid_group grid = cg::this_grid();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    coll_stamp(0);
    if (t.stage != nullptr) {
        const float c = t.coef[t.rank];
        const float4* src = reinterpret_cast<const float4*>(t.contrib[t.rank]);
        float4* dst = reinterpret_cast<float4*>(t.stage);
        const int64_t total4 = numel >> 2;
#pragma unroll 4
        for (int64_t i = tid; i < total4; i += stride) {
            float4 x = __ldcs(src + i);
            dst[i] = make_float4(c * x.x, c * x.y, c * x.z, c * x.w);
        }
        __threadfence_system();
        grid.sync();
    }
    coll_stamp(1);
    peer_barrier_enter(t, 0, epoch);
    coll_stamp(2);
    coll_stamp(3);
    reduce_int_buffers(t);

    const int64_t begin = (int64_t)t.rank * slice;
    int64_t end = begin + slice;
    if (end > numel) end = numel;
    const int64_t n4 = (end > begin) ? ((end - begin) >> 2) : 0;
    // kUnroll x 16 B per thread in flight: 148 CTAs x 512 thr x 4 x 16 B = 4.8 MB >= NVLink bandwidth-delay product
    for (int64_t base = tid; base < n4; base += stride * kUnroll) {
        float4 acc[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t i = base + (int64_t)u * stride;
            if (i < n4) acc[u] = mm_ld_reduce4(t.mc_contrib + begin + (i << 2));
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t i = base + (int64_t)u * stride;
            if (i >= n4) continue;
            const int64_t e = begin + (i << 2);
            float4 a = acc[u];
            a.x *= uniform_coef; a.y *= uniform_coef; a.z *= uniform_coef; a.w *= uniform_coef;
            if (ea.mode != EPI_NONE) {
                float4 wv = *reinterpret_cast<const float4*>(wcur + e);
                float4 mv = m ? *reinterpret_cast<const float4*>(m + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 vv = v ? *reinterpret_cast<const float4*>(v + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                a.x = epi_apply(ea.mode, ea, a.x, wv.x, mv.x, vv.x);
                a.y = epi_apply(ea.mode, ea, a.y, wv.y, mv.y, vv.y);
                a.z = epi_apply(ea.mode, ea, a.z, wv.z, mv.z, vv.z);
                a.w = epi_apply(ea.mode, ea, a.w, wv.w, mv.w, vv.w);
                if (m) *reinterpret_cast<float4*>(m + e) = mv;
                if (v) *reinterpret_cast<float4*>(v + e) = vv;
            }
            mm_st4(t.mc_result + e, a);
        }
    }
    coll_stamp(4);
    __threadfence_system();
    grid.sync();
    coll_stamp(5);
    peer_barrier(t, 1, epoch);
    coll_stamp(6);
}

// bcast_nvls: the root streams its buffer through ONE multimem.st per 16 bytes (the switch replicates it into every
// rank's result buffer: root egress = payload, no per-peer stores), barrier, then every rank unpacks locally.
template <int kUnroll>
__global__ void __launch_bounds__(kThreads, 1)
bcast_nvls_kernel(PeerTable t, int root, float* __restrict__ w, float* __restrict__ anchor,
                  __nv_bfloat16* __restrict__ shadow, const float* __restrict__ c_server,
                  const float* __restrict__ c_local, float* __restrict__ cv_out, int64_t numel, uint32_t epoch) {
    cg::grid_group grid = cg::this_grid();
    peer_barrier_enter(t, 0, epoch);  // everyone's result buffer may be overwritten
    if (t.rank == root) {
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        const int64_t total4 = numel >> 2;
        const float4* src = reinterpret_cast<const float4*>(t.contrib[root]);
        for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < total4; base += stride * kUnroll) {
            float4 vals[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t i = base + (int64_t)u * stride;
                if (i < total4) vals[u] = __ldcs(src + i);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t i = base + (int64_t)u * stride;
                if (i < total4) mm_st4(t.mc_result + (i << 2), vals[u]);
            }
        }
    }
    __threadfence_system();
    grid.sync();
    peer_barrier(t, 1, epoch);
    grid.sync();
    unpack_landed(t.result[t.rank], w, anchor, shadow, c_server, c_local, cv_out, numel);
}

// Tuning knobs (read once): FL4H_NVLS_UNROLL in {2,4,8} (default 4), FL4H_COLL_GRID = CTAs (default: one per SM).
int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return (v != nullptr && *v != 0) ? atoi(v) : fallback;
}

template <typename Kernel>
int coop_grid(Kernel kernel) {
    static int forced = env_int("FL4H_COLL_GRID", 0);
    if (forced > 0) return forced;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (per_sm < 1) return 0;  // cannot be co-resident: the caller's cooperative launch reports the error
    return sms;  // one CTA per SM (launch bounds guarantee >= 1 resident): 512 threads x unrolled 16-byte accesses
                 // already cover the NVLink bandwidth-delay product, and a single wave keeps the grid barrier cheap
}

}  // namespace

extern "C" {

int fl4h_coll_debug_read(unsigned long long* out8) {
    return (int)cudaMemcpyFromSymbol(out8, g_coll_stamp, sizeof(unsigned long long) * 8);
}

// ---- symmetric memory plumbing (CUDA IPC) -----------------------------------------------------------------------
int fl4h_ipc_alloc(size_t bytes, void** ptr) {
    cudaError_t err = cudaMalloc(ptr, bytes);
    if (err != cudaSuccess) return (int)err;
    return (int)cudaMemset(*ptr, 0, bytes);
}
int fl4h_ipc_free(void* ptr) { return (int)cudaFree(ptr); }
int fl4h_ipc_get_handle(void* ptr, void* handle_out /* 64 bytes */) {
    cudaIpcMemHandle_t handle;
    cudaError_t err = cudaIpcGetMemHandle(&handle, ptr);
    if (err != cudaSuccess) return (int)err;
    memcpy(handle_out, &handle, sizeof(handle));
    return 0;
}
int fl4h_ipc_open_handle(const void* handle_in, void** ptr) {
    cudaIpcMemHandle_t handle;
    memcpy(&handle, handle_in, sizeof(handle));
    return (int)cudaIpcOpenMemHandle(ptr, handle, cudaIpcMemLazyEnablePeerAccess);
}
int fl4h_ipc_close_handle(void* ptr) { return (int)cudaIpcCloseMemHandle(ptr); }
int fl4h_can_access_peer(int dev, int peer) {
    int ok = 0;
    cudaDeviceCanAccessPeer(&ok, dev, peer);
    return ok;
}

struct Fl4hPeerArgs {
    void* contrib[FL4H_MAX_RANKS];
    void* result[FL4H_MAX_RANKS];
    void* flags[FL4H_MAX_RANKS];
    float coef[FL4H_MAX_RANKS];
    int rank;
    int world;
    void* mc_contrib;   // multicast address the reduction reads (staging buffer, or the contribution itself)
    void* mc_result;    // multicast address of the result buffer
    void* stage;        // unicast address of this rank's staging buffer (null: reduce contributions in place)
    void* ibuf[FL4H_MAX_RANKS];
    void* ibuf_out;
    int n_int;
    int use_nvls;
    float uniform_coef; // applied after the in-switch sum (1.0 when the staging pass already applied the weights)
};

static PeerTable make_table(const Fl4hPeerArgs* a) {
    PeerTable t;
    for (int i = 0; i < FL4H_MAX_RANKS; ++i) {
        t.contrib[i] = reinterpret_cast<float*>(a->contrib[i]);
        t.result[i] = reinterpret_cast<float*>(a->result[i]);
        t.flags[i] = reinterpret_cast<uint32_t*>(a->flags[i]);
        t.coef[i] = a->coef[i];
    }
    t.rank = a->rank;
    t.world = a->world;
    t.mc_contrib = reinterpret_cast<const float*>(a->mc_contrib);
    t.mc_result = reinterpret_cast<float*>(a->mc_result);
    t.stage = reinterpret_cast<float*>(a->stage);
    for (int i = 0; i < FL4H_MAX_RANKS; ++i) t.ibuf[i] = reinterpret_cast<const long long*>(a->ibuf[i]);
    t.ibuf_out = reinterpret_cast<long long*>(a->ibuf_out);
    t.n_int = a->n_int;
    return t;
}

#define DISPATCH_WORLD(W, ...)         \
    switch (W) {                       \
        case 1: { constexpr int KW = 1; __VA_ARGS__; } break;   \
        case 2: { constexpr int KW = 2; __VA_ARGS__; } break;   \
        case 3: { constexpr int KW = 3; __VA_ARGS__; } break;   \
        case 4: { constexpr int KW = 4; __VA_ARGS__; } break;   \
        case 5: { constexpr int KW = 5; __VA_ARGS__; } break;   \
        case 6: { constexpr int KW = 6; __VA_ARGS__; } break;   \
        case 7: { constexpr int KW = 7; __VA_ARGS__; } break;   \
        case 8: { constexpr int KW = 8; __VA_ARGS__; } break;   \
        default: return (int)cudaErrorInvalidValue;      \
    }

int fl4h_agg_fused(const Fl4hPeerArgs* args, const float* wcur, float* m, float* v, int mode, float eta, float beta1,
                   float beta2, float tau, float server_lr, float momentum, int64_t numel, uint32_t epoch,
                   cudaStream_t stream) {
    if (numel & 3) return (int)cudaErrorInvalidValue;
    PeerTable t = make_table(args);
    EpiArgs ea{eta, beta1, beta2, tau, server_lr, momentum, mode};
    // slice = ceil(numel / world) rounded up to 4 elements
    int64_t slice = (numel + t.world - 1) / t.world;
    slice = (slice + 3) & ~int64_t(3);
    cudaError_t err = cudaSuccess;
    if (args->use_nvls) {
        if (t.mc_contrib == nullptr || t.mc_result == nullptr) return (int)cudaErrorInvalidValue;
        static int unroll = env_int("FL4H_NVLS_UNROLL", 4);
        float uniform_coef = args->uniform_coef;
        void* kargs[] = {&t, &wcur, &m, &v, &ea, &numel, &slice, &epoch, &uniform_coef};
        void* kernel = unroll == 8 ? (void*)agg_nvls_kernel<8> : (unroll == 2 ? (void*)agg_nvls_kernel<2> : (void*)agg_nvls_kernel<4>);
        const int grid = coop_grid(agg_nvls_kernel<4>);
        return (int)cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kThreads), kargs, 0, stream);
    }
    DISPATCH_WORLD(t.world, {
        auto kernel = agg_fused_kernel<KW>;
        const int grid = coop_grid(kernel);
        void* kargs[] = {&t, &wcur, &m, &v, &ea, &numel, &slice, &epoch};
        err = cudaLaunchCooperativeKernel((void*)kernel, dim3(grid), dim3(kThreads), kargs, 0, stream);
    });
    return (int)err;
}

int fl4h_bcast_fused(const Fl4hPeerArgs* args, int root, float* w, float* anchor, void* shadow,
                     const float* c_server, const float* c_local, float* cv_out, int64_t numel, uint32_t epoch,
                     cudaStream_t stream) {
    if (numel & 3) return (int)cudaErrorInvalidValue;
    PeerTable t = make_table(args);
    int64_t slice = (numel + t.world - 1) / t.world;
    slice = (slice + 3) & ~int64_t(3);
    __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(shadow);
    cudaError_t err = cudaSuccess;
    if (args->use_nvls) {
        if (t.mc_result == nullptr) return (int)cudaErrorInvalidValue;
        static int unroll = env_int("FL4H_NVLS_UNROLL", 4);
        void* kargs[] = {&t, &root, &w, &anchor, &sh, &c_server, &c_local, &cv_out, &numel, &epoch};
        void* kernel = unroll == 8 ? (void*)bcast_nvls_kernel<8> : (unroll == 2 ? (void*)bcast_nvls_kernel<2> : (void*)bcast_nvls_kernel<4>);
        const int grid = coop_grid(bcast_nvls_kernel<4>);
        return (int)cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kThreads), kargs, 0, stream);
    }
    DISPATCH_WORLD(t.world, {
        auto kernel = bcast_fused_kernel<KW>;
        const int grid = coop_grid(kernel);
        void* kargs[] = {&t, &root, &w, &anchor, &sh, &c_server, &c_local, &cv_out, &numel, &slice, &epoch};
        err = cudaLaunchCooperativeKernel((void*)kernel, dim3(grid), dim3(kThreads), kargs, 0, stream);
    });
    return (int)err;
}

}  // extern "C"
