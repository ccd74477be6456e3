// BatchNorm(+residual)(+ReLU) training forward / backward for SMALL activations as ONE thread-block cluster.
//
// Why: for the <= 2 MB activation tensors of a CIFAR-scale network, the grid-wide kernels in bn_act.cu cost 11-16 us
// each no matter how little data they touch -- their critical path is a chain of global-memory round trips (load,
// RED atomics + fence, grid-barrier arrive, poll, read totals, store).  A single cluster of 8-16 CTAs keeps the whole
// dependency on chip: every CTA pushes its partial sums into all peers' shared memory with DSMEM reductions
// (red.shared::cluster), and the only synchronisation is the hardware cluster barrier.  The tile each
// thread normalises stays in registers between the two phases (forward) so x is read exactly once.
//
// Layout / mapping are those of bn_act.cu: x is [M, C] with C contiguous (NHWC), C % 8 == 0, one thread owns 8
// channels of a row per iteration; CTA r of S handles rows [r * rows_per_cta, ...).

#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "bn_common.cuh"

namespace cg = cooperative_groups;
using fl4h_bn::load8f;
using fl4h_bn::Vec8;

namespace {

constexpr int kThreads = 512;
constexpr int kCacheRegs = 64;  // registers per thread spent on the tile kept between the two phases
// forward caches x (packed as stored: 4 regs / 8 bf16), backward caches the masked gradient and x
template <typename T> constexpr int fwd_cache() { return kCacheRegs / Vec8<T>::kRawRegs; }
template <typename T> constexpr int bwd_cache() { return kCacheRegs / (2 * Vec8<T>::kRawRegs); }

// smem layout (floats): [0, 2C) cluster-wide totals (pushed by all CTAs) | [2C, 2C + 2*RP*C) block-reduction scratch, later reused for
// the per-channel coefficients (<= 3C floats) | shift values [C] (forward only).
// Push variant: every CTA adds its per-channel partials into EVERY peer's `totals` array with asynchronous DSMEM
// reductions (red.shared::cluster) -- S pipelined fire-and-forget operations per value instead of S dependent remote
// loads on the critical path.  `totals` must have been zeroed cluster-wide before (barrier_arrive / barrier_wait).
__device__ __forceinline__ void block_partials_push(cg::cluster_group& cluster, const float (&a)[8], const float (&b)[8], int C,
                                                    int LP, int RP, int lane, int ty, float* totals, float* scratch) {
    if (ty < RP) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            scratch[(ty * 2 + 0) * C + lane * 8 + k] = a[k];
            scratch[(ty * 2 + 1) * C + lane * 8 + k] = b[k];
        }
    }
    __syncthreads();
    const unsigned S = cluster.num_blocks();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sa = 0.f, sb = 0.f;
        for (int t = 0; t < RP; ++t) {
            sa += scratch[(t * 2 + 0) * C + c];
            sb += scratch[(t * 2 + 1) * C + c];
        }
        for (unsigned r = 0; r < S; ++r) {
            float* remote = cluster.map_shared_rank(totals, r);
            atomicAdd(remote + c, sa);
            atomicAdd(remote + C + c, sb);
        }
    }
}

template <typename T, bool kRelu, bool kRes>
__global__ void __launch_bounds__(kThreads, 1)
bn_fwd_cluster_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, int64_t M, int C, int rows_per_cta,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                      float* running_var, int64_t* nbt, float momentum, float eps, float* mean_out, float* invstd_out) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float smem[];
    const int LP = C >> 3, RP = blockDim.x / LP;
    const int lane = threadIdx.x % LP, ty = threadIdx.x / LP;
    float* partial = smem;
    float* scratch = smem + 2 * C;
    float* kshift = scratch + 2 * RP * C;
    const unsigned rank = cluster.block_rank();

    for (int c = threadIdx.x; c < C; c += blockDim.x) kshift[c] = running_mean != nullptr ? running_mean[c] : 0.f;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) partial[c] = 0.f;   // cluster-wide totals land here
    __syncthreads();
    cluster.barrier_arrive();                              // "my totals are zeroed" -- waited for just before the push
    const int64_t r0 = (int64_t)rank * rows_per_cta;
    const int64_t r1 = (r0 + rows_per_cta < M) ? r0 + rows_per_cta : M;
    constexpr int kFwdCache = fwd_cache<T>();
    const bool cached = rows_per_cta <= kFwdCache * RP;
    typename Vec8<T>::Raw tile[kFwdCache];
    float s[8], q[8], shift[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; }
    // issue the tile loads before waiting for the shift values: both latencies overlap
    if (ty < RP && cached) {
#pragma unroll
        for (int it = 0; it < kFwdCache; ++it) {
            const int64_t r = r0 + ty + (int64_t)it * RP;
            if (r < r1) tile[it] = Vec8<T>::load_raw(x + r * C + lane * 8);
        }
    }
    __syncthreads();
    if (ty < RP) {
#pragma unroll
        for (int k = 0; k < 8; ++k) shift[k] = kshift[lane * 8 + k];
        if (cached) {
#pragma unroll
            for (int it = 0; it < kFwdCache; ++it) {
                const int64_t r = r0 + ty + (int64_t)it * RP;
                if (r < r1) {
                    float v[8];
                    Vec8<T>::unpack(tile[it], v);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float d = v[k] - shift[k];
                        s[k] += d;
                        q[k] = fmaf(d, d, q[k]);
                    }
                }
            }
        } else {
            for (int64_t r = r0 + ty; r < r1; r += RP) {
                float v[8];
                Vec8<T>::load(x + r * C + lane * 8, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float d = v[k] - shift[k];
                    s[k] += d;
                    q[k] = fmaf(d, d, q[k]);
                }
            }
        }
    }
    cluster.barrier_wait();
    block_partials_push(cluster, s, q, C, LP, RP, lane, ty, partial, scratch);
    cluster.sync();                                        // all pushes have landed in every CTA's totals
    const float inv_m = 1.f / (float)M;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sd = partial[c], sq = partial[C + c];
        const float k0 = kshift[c];
        const float md = sd * inv_m;
        const float mean = k0 + md;
        float var = fmaf(-md, md, sq * inv_m);
        var = var > 0.f ? var : 0.f;
        const float invstd = rsqrtf(var + eps);
        const float sc = (gamma != nullptr ? gamma[c] : 1.f) * invstd;
        scratch[c] = sc;
        scratch[C + c] = (beta != nullptr ? beta[c] : 0.f) - mean * sc;
        if (rank == 0) {
            mean_out[c] = mean;
            invstd_out[c] = invstd;
            if (running_mean != nullptr) {
                const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
                running_mean[c] = (1.f - momentum) * k0 + momentum * mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
            }
        }
    }
    __syncthreads();
    if (ty < RP) {
        float sc[8], sh[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { sc[k] = scratch[lane * 8 + k]; sh[k] = scratch[C + lane * 8 + k]; }
        if (cached) {
#pragma unroll
            for (int it = 0; it < kFwdCache; ++it) {
                const int64_t r = r0 + ty + (int64_t)it * RP;
                if (r < r1) {
                    const int64_t off = r * C + lane * 8;
                    float v[8], rr[8];
                    Vec8<T>::unpack(tile[it], v);
                    if (kRes) Vec8<T>::load(res + off, rr);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float o = fmaf(v[k], sc[k], sh[k]);
                        if (kRes) o += rr[k];
                        if (kRelu) o = o > 0.f ? o : 0.f;
                        v[k] = o;
                    }
                    Vec8<T>::store(y + off, v);
                }
            }
        } else {
            for (int64_t r = r0 + ty; r < r1; r += RP) {
                const int64_t off = r * C + lane * 8;
                float v[8], rr[8];
                Vec8<T>::load(x + off, v);
                if (kRes) Vec8<T>::load(res + off, rr);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float o = fmaf(v[k], sc[k], sh[k]);
                    if (kRes) o += rr[k];
                    if (kRelu) o = o > 0.f ? o : 0.f;
                    v[k] = o;
                }
                Vec8<T>::store(y + off, v);
            }
        }
    }
    if (rank == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
    cluster.sync();                                        // nobody exits while peers may still read its partials
}

template <typename T, bool kRelu, bool kRes>
__global__ void __launch_bounds__(kThreads, 1)
bn_bwd_cluster_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x, T* __restrict__ dx,
                      T* __restrict__ dres, int64_t M, int C, int rows_per_cta, const float* __restrict__ gamma,
                      const float* __restrict__ mean, const float* __restrict__ invstd, float* dgamma, float* dbeta) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float smem[];
    const int LP = C >> 3, RP = blockDim.x / LP;
    const int lane = threadIdx.x % LP, ty = threadIdx.x / LP;
    float* partial = smem;
    float* scratch = smem + 2 * C;
    const unsigned rank = cluster.block_rank();
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) partial[c] = 0.f;
    __syncthreads();
    cluster.barrier_arrive();
    const int64_t r0 = (int64_t)rank * rows_per_cta;
    const int64_t r1 = (r0 + rows_per_cta < M) ? r0 + rows_per_cta : M;
    constexpr int kBwdCache = bwd_cache<T>();
    const bool cached = rows_per_cta <= kBwdCache * RP;
    typename Vec8<T>::Raw gt[kBwdCache], xt[kBwdCache];    // masked upstream gradient (exact in T) and raw x of the cached tile
    float sg[8], sgx[8], mu[8], is[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sg[k] = 0.f; sgx[k] = 0.f; mu[k] = 0.f; is[k] = 0.f; }
    if (ty < RP) {
        load8f(mean + lane * 8, mu);
        load8f(invstd + lane * 8, is);
        if (cached) {
#pragma unroll
            for (int it = 0; it < kBwdCache; ++it) {
                const int64_t r = r0 + ty + (int64_t)it * RP;
                if (r < r1) {
                    const int64_t off = r * C + lane * 8;
                    float g[8], xv[8], yv[8];
                    Vec8<T>::load(dy + off, g);
                    xt[it] = Vec8<T>::load_raw(x + off);
                    Vec8<T>::unpack(xt[it], xv);
                    if (kRelu) Vec8<T>::load(y + off, yv);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float gi = (!kRelu || yv[k] > 0.f) ? g[k] : 0.f;
                        g[k] = gi;
                        sg[k] += gi;
                        sgx[k] = fmaf(gi, (xv[k] - mu[k]) * is[k], sgx[k]);
                    }
                    gt[it] = Vec8<T>::pack(g);
                }
            }
        } else {
            for (int64_t r = r0 + ty; r < r1; r += RP) {
                const int64_t off = r * C + lane * 8;
                float g[8], xv[8], yv[8];
                Vec8<T>::load(dy + off, g);
                Vec8<T>::load(x + off, xv);
                if (kRelu) Vec8<T>::load(y + off, yv);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gi = (!kRelu || yv[k] > 0.f) ? g[k] : 0.f;
                    sg[k] += gi;
                    sgx[k] = fmaf(gi, (xv[k] - mu[k]) * is[k], sgx[k]);
                }
            }
        }
    }
    cluster.barrier_wait();
    block_partials_push(cluster, sg, sgx, C, LP, RP, lane, ty, partial, scratch);
    cluster.sync();
    const float inv_m = 1.f / (float)M;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float a = partial[c], b = partial[C + c];
        scratch[c] = (gamma != nullptr ? gamma[c] : 1.f) * invstd[c];
        scratch[C + c] = a * inv_m;
        scratch[2 * C + c] = b * inv_m;
        if (rank == 0) {
            if (dbeta != nullptr) dbeta[c] = a;
            if (dgamma != nullptr) dgamma[c] = b;
        }
    }
    __syncthreads();
    if (ty < RP) {
        float a[8], mg[8], mgx[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a[k] = scratch[lane * 8 + k];
            mg[k] = scratch[C + lane * 8 + k];
            mgx[k] = scratch[2 * C + lane * 8 + k];
        }
        if (cached) {
#pragma unroll
            for (int it = 0; it < kBwdCache; ++it) {
                const int64_t r = r0 + ty + (int64_t)it * RP;
                if (r < r1) {
                    const int64_t off = r * C + lane * 8;
                    float g[8], xv[8];
                    Vec8<T>::unpack(gt[it], g);
                    Vec8<T>::unpack(xt[it], xv);
#pragma unroll
                    for (int k = 0; k < 8; ++k) xv[k] = a[k] * (g[k] - mg[k] - (xv[k] - mu[k]) * is[k] * mgx[k]);
                    Vec8<T>::store(dx + off, xv);
                    if (kRes) Vec8<T>::store_raw(dres + off, gt[it]);
                }
            }
        } else {
            for (int64_t r = r0 + ty; r < r1; r += RP) {
                const int64_t off = r * C + lane * 8;
                float g[8], xv[8], yv[8];
                Vec8<T>::load(dy + off, g);
                Vec8<T>::load(x + off, xv);
                if (kRelu) Vec8<T>::load(y + off, yv);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gi = (!kRelu || yv[k] > 0.f) ? g[k] : 0.f;
                    g[k] = gi;
                    xv[k] = a[k] * (gi - mg[k] - (xv[k] - mu[k]) * is[k] * mgx[k]);
                }
                Vec8<T>::store(dx + off, xv);
                if (kRes) Vec8<T>::store(dres + off, g);
            }
        }
    }
    cluster.sync();
}

struct ClusterPlan {
    int cluster_size;
    int rows_per_cta;
    size_t smem_bytes;
};

inline ClusterPlan plan(int64_t M, int C, int max_cluster, bool forward) {
    const int LP = C / 8, RP = kThreads / LP;
    int s = max_cluster;
    while (s > 1 && (int64_t)(s - 1) * RP >= M) s >>= 1;   // do not spawn CTAs without rows
    int64_t rows = (M + s - 1) / s;
    rows = (rows + RP - 1) / RP * RP;
    ClusterPlan p;
    p.cluster_size = (int)((M + rows - 1) / rows);
    if (p.cluster_size < 1) p.cluster_size = 1;
    // cluster sizes must be supported shapes: round up to a power of two (empty CTAs just contribute zeros)
    int pow2 = 1;
    while (pow2 < p.cluster_size) pow2 <<= 1;
    p.cluster_size = pow2;
    p.rows_per_cta = (int)rows;
    const size_t scratch = (size_t)2 * RP * C > (size_t)3 * C ? (size_t)2 * RP * C : (size_t)3 * C;
    p.smem_bytes = ((size_t)2 * C + scratch + (forward ? (size_t)C : 0)) * sizeof(float);
    return p;
}

template <typename Kernel, typename... Args>
cudaError_t launch_cluster(Kernel kernel, const ClusterPlan& p, cudaStream_t stream, Args... args) {
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    if (err != cudaSuccess) return err;
    if (p.cluster_size > 8) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (err != cudaSuccess) return err;
    }
    cudaLaunchConfig_t config = {};
    config.gridDim = dim3(p.cluster_size);
    config.blockDim = dim3(kThreads);
    config.dynamicSmemBytes = p.smem_bytes;
    config.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    config.attrs = attr;
    config.numAttrs = 1;
    return cudaLaunchKernelEx(&config, kernel, args...);
}

#define FL4H_BNC_DISPATCH(T, relu, res, CALL)                                                     \
    do {                                                                                           \
        if (relu) { if (res) { CALL(T, true, true); } else { CALL(T, true, false); } }            \
        else      { if (res) { CALL(T, false, true); } else { CALL(T, false, false); } }          \
    } while (0)

}  // namespace

extern "C" {

// 1 if the cluster path handles this problem (C multiple of 8, 8 <= C/8 <= 512 lanes, tensor small enough that a
// single cluster beats the grid-wide kernel).
int fl4h_bn_cluster_supported(int64_t M, int C, int64_t max_bytes, int elem_bytes) {
    if (C % 8 != 0 || C / 8 > kThreads || M < 1) return 0;
    return (M * C * elem_bytes <= max_bytes) ? 1 : 0;
}

int fl4h_bn_fwd_train_cluster(const void* x, const void* res, void* y, int64_t M, int C, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, int64_t* nbt, float momentum, float eps,
                              float* mean_out, float* invstd_out, int is_bf16, int relu, int max_cluster, cudaStream_t stream) {
    const ClusterPlan p = plan(M, C, max_cluster, true);
    const bool has_res = res != nullptr;
    cudaError_t err = cudaSuccess;
#define CALL_FWD(T, R, S)                                                                                              \
    err = launch_cluster(bn_fwd_cluster_kernel<T, R, S>, p, stream, (const T*)x, (const T*)res, (T*)y, M, C, p.rows_per_cta, \
                         gamma, beta, running_mean, running_var, nbt, momentum, eps, mean_out, invstd_out)
    if (is_bf16) FL4H_BNC_DISPATCH(__nv_bfloat16, relu, has_res, CALL_FWD);
    else FL4H_BNC_DISPATCH(float, relu, has_res, CALL_FWD);
#undef CALL_FWD
    if (err != cudaSuccess) (void)cudaGetLastError();
    return (int)err;
}

int fl4h_bn_bwd_cluster(const void* dy, const void* y, const void* x, int64_t M, int C, const float* gamma, const float* mean,
                        const float* invstd, void* dx, void* dres, float* dgamma, float* dbeta, int is_bf16, int relu,
                        int max_cluster, cudaStream_t stream) {
    const ClusterPlan p = plan(M, C, max_cluster, false);
    const bool has_res = dres != nullptr;
    cudaError_t err = cudaSuccess;
#define CALL_BWD(T, R, S)                                                                                             \
    err = launch_cluster(bn_bwd_cluster_kernel<T, R, S>, p, stream, (const T*)dy, (const T*)y, (const T*)x, (T*)dx, (T*)dres, \
                         M, C, p.rows_per_cta, gamma, mean, invstd, dgamma, dbeta)
    if (is_bf16) FL4H_BNC_DISPATCH(__nv_bfloat16, relu, has_res, CALL_BWD);
    else FL4H_BNC_DISPATCH(float, relu, has_res, CALL_BWD);
#undef CALL_BWD
    if (err != cudaSuccess) (void)cudaGetLastError();
    return (int)err;
}

}  // extern "C"
