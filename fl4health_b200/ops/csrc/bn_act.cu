// Fused BatchNorm(+residual add)(+ReLU) for channels-last activations on sm_100a.
//
// ResNet-style FL clients at CIFAR scale are launch-latency bound: ATen runs ~10 kernels per BN layer and step
// (collect stats, update running stats, transform, add, relu, num_batches_tracked += 1, and their backward twins).
// Here a training forward is 2 kernels and a backward is 2 kernels, whatever the tail ops:
//
//   fwd  K1  per-channel shifted sums (shift = running_mean, kills the E[x^2]-E[x]^2 cancellation) reduced through
//            smem -> fp32 RED atomics in L2; the last CTA to finish turns them into mean / invstd / scale / shift,
//            updates running_mean / running_var / num_batches_tracked and re-zeroes the workspace
//        K2  y = relu(x * scale + shift + residual)               (16-byte vectors of 8 channels)
//   bwd  K3  g = dy * (y > 0);  sum(g), sum(g * xhat) -> last CTA writes dgamma, dbeta and dx coefficients
//        K4  dx = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat));  dresidual = g
//
// On top of that, the default training path fuses each pair into ONE cooperative kernel (reduce -> grid barrier ->
// apply): measured on B200 the two-kernel chain costs ~15 us per layer and direction regardless of tensor size (it is
// a pure latency chain: load, smem reduce, RED + fence, election atomic, last-CTA epilogue, second launch), the fused
// kernel replaces the fence/election/second launch by one grid barrier and keeps the forward tile in registers.
// The two-kernel kernels remain as the fallback when a cooperative launch is not possible.
//
// Layout: x is [M, C] with C contiguous (NHWC storage), C % 8 == 0.  T is bf16 or fp32; statistics are fp32.

#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "bn_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;

using fl4h_bn::Vec8;
using fl4h_bn::load8f;

// Shared tail of both reduction kernels: fold the per-thread pairs (a, b) over the CTA's rows, RED them into
// acc[0..C) / acc[C..2C), and elect the last CTA.  Returns true in every thread of the last CTA.
__device__ __forceinline__ bool reduce_and_elect(const float (&a)[8], const float (&b)[8], int C, int LP, int RP, int lane,
                                                 int ty, float* acc, unsigned* counter, float* smem) {
    if (ty < RP) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            smem[(ty * 2 + 0) * C + lane * 8 + k] = a[k];
            smem[(ty * 2 + 1) * C + lane * 8 + k] = b[k];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sa = 0.f, sb = 0.f;
        for (int t = 0; t < RP; ++t) {
            sa += smem[(t * 2 + 0) * C + c];
            sb += smem[(t * 2 + 1) * C + c];
        }
        atomicAdd(acc + c, sa);
        atomicAdd(acc + C + c, sb);
    }
    __threadfence();
    __syncthreads();
    __shared__ bool is_last;
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
bn_fwd_stats_kernel(const T* __restrict__ x, int64_t M, int C, int rows_per_cta, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* running_mean, float* running_var, int64_t* nbt,
                    float momentum, float eps, float* mean_out, float* invstd_out, float* scale_shift, float* acc,
                    unsigned* counter) {
    extern __shared__ float smem[];
    const int LP = C >> 3, RP = blockDim.x / LP;
    const int lane = threadIdx.x % LP, ty = threadIdx.x / LP;
    float shift[8], s[8], q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; shift[k] = 0.f; }
    if (running_mean != nullptr && ty < RP) load8f(running_mean + lane * 8, shift);
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = (r0 + rows_per_cta < M) ? r0 + rows_per_cta : M;
    if (ty < RP) {
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
    if (!reduce_and_elect(s, q, C, LP, RP, lane, ty, acc, counter, smem)) return;
    const float inv_m = 1.f / (float)M;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sd = __ldcg(acc + c), sq = __ldcg(acc + C + c);
        acc[c] = 0.f;
        acc[C + c] = 0.f;
        const float k0 = running_mean != nullptr ? running_mean[c] : 0.f;
        const float md = sd * inv_m;
        const float mean = k0 + md;
        float var = fmaf(-md, md, sq * inv_m);
        var = var > 0.f ? var : 0.f;
        const float invstd = rsqrtf(var + eps);
        mean_out[c] = mean;
        invstd_out[c] = invstd;
        const float sc = (gamma != nullptr ? gamma[c] : 1.f) * invstd;
        scale_shift[c] = sc;
        scale_shift[C + c] = (beta != nullptr ? beta[c] : 0.f) - mean * sc;
        if (running_mean != nullptr) {
            const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
            running_mean[c] = (1.f - momentum) * k0 + momentum * mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
        }
    }
    if (threadIdx.x == 0) {
        *counter = 0u;
        if (nbt != nullptr) *nbt += 1;
    }
}

// mode 0: scale/shift precomputed in scale_shift[2][C];  mode 1 (eval): derive them from the running statistics.
template <typename T, bool kRelu, bool kRes>
__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, int64_t nvec, int C,
                const float* __restrict__ scale_shift, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ rmean, const float* __restrict__ rvar, float eps, int mode) {
    const int LP = C >> 3;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % LP) * 8;
        float sc[8], sh[8], v[8], r[8];
        if (mode == 0) {
            load8f(scale_shift + c0, sc);
            load8f(scale_shift + C + c0, sh);
        } else {
            float g[8], b[8], m[8], var[8];
            load8f(rmean + c0, m);
            load8f(rvar + c0, var);
#pragma unroll
            for (int k = 0; k < 8; ++k) { g[k] = 1.f; b[k] = 0.f; }
            if (gamma != nullptr) load8f(gamma + c0, g);
            if (beta != nullptr) load8f(beta + c0, b);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                sc[k] = g[k] * rsqrtf(var[k] + eps);
                sh[k] = b[k] - m[k] * sc[k];
            }
        }
        Vec8<T>::load(x + i * 8, v);
        if (kRes) Vec8<T>::load(res + i * 8, r);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float o = fmaf(v[k], sc[k], sh[k]);
            if (kRes) o += r[k];
            if (kRelu) o = o > 0.f ? o : 0.f;
            v[k] = o;
        }
        Vec8<T>::store(y + i * 8, v);
    }
}


// Training forward when the producing convolution already reduced the batch statistics in its epilogue
// (conv_tc.cu): `sums` = [sum | sum of squares] over all M pixels, per channel.  One streaming pass, no grid barrier:
// every thread derives the scale/shift of its 8 channels once (the grid stride is a multiple of C/8, so a thread keeps
// its channel group), CTA 0 does the bookkeeping (saved mean / invstd for the backward, running statistics, counter),
// and the last CTA to finish re-zeroes `sums` for the next step (election through `done`), which keeps the buffer
// self-resetting under CUDA-graph replay.
template <typename T, bool kRelu, bool kRes>
__global__ void __launch_bounds__(kThreads)
bn_apply_presum_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, int64_t nvec, int C, int64_t M,
                       float* sums, const float* __restrict__ gamma, const float* __restrict__ beta, float* rmean,
                       float* rvar, long long* nbt, float momentum, float eps, float* __restrict__ mean_out,
                       float* __restrict__ invstd_out, unsigned* done) {
    // programmatic dependent launch (see conv_tc.cu): started while the producing convolution drains; its statistics
    // and output are visible after the wait
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int LP = C >> 3;
    const float inv_m = 1.f / (float)M;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = (int)(first % LP) * 8;
    float sc[8], sh[8];
    {
        float s[8], q[8], g[8], b[8];
        load8f(sums + c0, s);
        load8f(sums + C + c0, q);
#pragma unroll
        for (int k = 0; k < 8; ++k) { g[k] = 1.f; b[k] = 0.f; }
        if (gamma != nullptr) load8f(gamma + c0, g);
        if (beta != nullptr) load8f(beta + c0, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float mean = s[k] * inv_m;
            const float var = fmaxf(fmaf(-mean, mean, q[k] * inv_m), 0.f);
            sc[k] = g[k] * rsqrtf(var + eps);
            sh[k] = fmaf(-mean, sc[k], b[k]);
        }
    }
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float mean = sums[c] * inv_m;
            const float var = fmaxf(fmaf(-mean, mean, sums[C + c] * inv_m), 0.f);
            mean_out[c] = mean;
            invstd_out[c] = rsqrtf(var + eps);
            if (rmean != nullptr) {
                const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
                rmean[c] = fmaf(momentum, mean - rmean[c], rmean[c]);
                rvar[c] = fmaf(momentum, unbiased - rvar[c], rvar[c]);
            }
        }
        if (threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
    }
    for (int64_t i = first; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float v[8], r[8];
        Vec8<T>::load(x + i * 8, v);
        if (kRes) Vec8<T>::load(res + i * 8, r);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float o = fmaf(v[k], sc[k], sh[k]);
            if (kRes) o += r[k];
            if (kRelu) o = o > 0.f ? o : 0.f;
            v[k] = o;
        }
        Vec8<T>::store(y + i * 8, v);
    }
    __shared__ bool is_last;
    __syncthreads();                                       // every thread of this CTA has consumed `sums`
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(done, 1u);
        is_last = prev == gridDim.x - 1;
        if (is_last) *done = 0;
    }
    __syncthreads();
    if (is_last)
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sums[c] = 0.f;
}

template <typename T, bool kRelu>
__global__ void __launch_bounds__(kThreads)
bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x, int64_t M, int C,
                     int rows_per_cta, const float* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ invstd, float* dgamma, float* dbeta, float* coef, float* acc,
                     unsigned* counter) {
    extern __shared__ float smem[];
    const int LP = C >> 3, RP = blockDim.x / LP;
    const int lane = threadIdx.x % LP, ty = threadIdx.x / LP;
    float sg[8], sgx[8], mu[8], is[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sg[k] = 0.f; sgx[k] = 0.f; mu[k] = 0.f; is[k] = 0.f; }
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = (r0 + rows_per_cta < M) ? r0 + rows_per_cta : M;
    if (ty < RP) {
        load8f(mean + lane * 8, mu);
        load8f(invstd + lane * 8, is);
        for (int64_t r = r0 + ty; r < r1; r += RP) {
            float g[8], xv[8], yv[8];
            const int64_t off = r * C + lane * 8;
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
    if (!reduce_and_elect(sg, sgx, C, LP, RP, lane, ty, acc, counter, smem)) return;
    const float inv_m = 1.f / (float)M;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float a = __ldcg(acc + c), b = __ldcg(acc + C + c);
        acc[c] = 0.f;
        acc[C + c] = 0.f;
        if (dbeta != nullptr) dbeta[c] = a;
        if (dgamma != nullptr) dgamma[c] = b;
        coef[c] = (gamma != nullptr ? gamma[c] : 1.f) * invstd[c];  // a
        coef[C + c] = a * inv_m;                                     // mean(g)
        coef[2 * C + c] = b * inv_m;                                 // mean(g * xhat)
    }
    if (threadIdx.x == 0) *counter = 0u;
}

template <typename T, bool kRelu, bool kRes>
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x, T* __restrict__ dx,
                    T* __restrict__ dres, int64_t nvec, int C, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ coef) {
    const int LP = C >> 3;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % LP) * 8;
        float mu[8], is[8], a[8], mg[8], mgx[8], g[8], xv[8], yv[8];
        load8f(mean + c0, mu);
        load8f(invstd + c0, is);
        load8f(coef + c0, a);
        load8f(coef + C + c0, mg);
        load8f(coef + 2 * C + c0, mgx);
        Vec8<T>::load(dy + i * 8, g);
        Vec8<T>::load(x + i * 8, xv);
        if (kRelu) Vec8<T>::load(y + i * 8, yv);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float gi = (!kRelu || yv[k] > 0.f) ? g[k] : 0.f;
            g[k] = gi;
            xv[k] = a[k] * (gi - mg[k] - (xv[k] - mu[k]) * is[k] * mgx[k]);
        }
        Vec8<T>::store(dx + i * 8, xv);
        if (kRes) Vec8<T>::store(dres + i * 8, g);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// fused cooperative kernels: reduce -> grid.sync() -> apply
// workspace: acc2 = two accumulator buffers of kAccStride floats used alternately (the buffer of the previous launch is
// re-zeroed by CTA 0 of the current one, so no extra barrier is needed for the reset); *parity selects the buffer.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAccStride = 4096;
constexpr int kCacheRows = 8;

__device__ __forceinline__ void red_partials(const float (&a)[8], const float (&b)[8], int C, int LP, int RP, int lane, int ty,
                                             float* acc, float* smem) {
    // Stage 1 (power-of-two LP < 32, i.e. C in {64, 128} ... ): the 32 / LP row-threads of a warp that own the same 8
    // channels combine through shuffles, so the shared-memory stage holds RP / (32 / LP) rows instead of RP and the
    // serial per-channel loop below shrinks by the same factor (32 -> 8 iterations at C = 64).
    const bool fold = (LP & (LP - 1)) == 0 && LP < 32 && (blockDim.x % 32) == 0;
    float a2[8], b2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a2[k] = a[k]; b2[k] = b[k]; }
    int group = 1;
    if (fold) {
        group = 32 / LP;
        for (int off = LP; off < 32; off <<= 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a2[k] += __shfl_xor_sync(0xffffffffu, a2[k], off);
                b2[k] += __shfl_xor_sync(0xffffffffu, b2[k], off);
            }
        }
    }
    const int rows = (RP + group - 1) / group;
    if (ty < RP && (ty % group) == 0) {
        const int row = ty / group;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            smem[(row * 2 + 0) * C + lane * 8 + k] = a2[k];
            smem[(row * 2 + 1) * C + lane * 8 + k] = b2[k];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sa = 0.f, sb = 0.f;
        for (int t = 0; t < rows; ++t) {
            sa += smem[(t * 2 + 0) * C + c];
            sb += smem[(t * 2 + 1) * C + c];
        }
        atomicAdd(acc + c, sa);
        atomicAdd(acc + C + c, sb);
    }
}

// Phase timestamps of CTA 0 of the fused forward kernel (SM clock), read back with fl4h_bn_debug_read:
//   [0] entry  [1] shift values staged  [2] tile loaded + per-thread sums  [3] block reduce + RED atomics issued
//   [4] grid barrier passed  [5] totals read, scale/shift ready  [6] normalised tile stored
__device__ long long g_bn_phase_clock[16];
#define BN_STAMP(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_bn_phase_clock[(slot)] = clock64(); } while (0)

template <typename T, bool kRelu, bool kRes>
__global__ void __launch_bounds__(kThreads)
bn_fwd_fused_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, int64_t M, int C, int rows_per_cta,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                    float* running_var, int64_t* nbt, float momentum, float eps, float* mean_out, float* invstd_out,
                    float* acc2, unsigned* parity) {
    cg::grid_group grid = cg::this_grid();
    BN_STAMP(0);
    extern __shared__ float smem[];            // [4096] reduce scratch (later scale|shift) + [C] shift values
    float* kshift = smem + 4096;
    const unsigned par = *reinterpret_cast<volatile unsigned*>(parity) & 1u;
    float* acc = acc2 + par * kAccStride;
    const int LP = C >> 3, RP = blockDim.x / LP;
    const int lane = threadIdx.x % LP, ty = threadIdx.x / LP;
    for (int c = threadIdx.x; c < C; c += blockDim.x) kshift[c] = running_mean != nullptr ? running_mean[c] : 0.f;
    float shift[8], s[8], q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; shift[k] = 0.f; }
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = (r0 + rows_per_cta < M) ? r0 + rows_per_cta : M;
    const bool cached = rows_per_cta <= kCacheRows * RP;   // uniform: the whole tile stays in registers
    float tile[kCacheRows][8];
    // The tile loads do not depend on the shift values: issue them BEFORE the barrier that publishes `kshift`, so the
    // running-mean fetch (a dependent ~1 us global round trip, phase timestamps) overlaps the tile's own L2 latency.
    if (ty < RP && cached) {
#pragma unroll
        for (int it = 0; it < kCacheRows; ++it) {
            const int64_t r = r0 + ty + (int64_t)it * RP;
            if (r < r1) Vec8<T>::load(x + r * C + lane * 8, tile[it]);
        }
    }
    __syncthreads();
    BN_STAMP(1);
    if (ty < RP) {
#pragma unroll
        for (int k = 0; k < 8; ++k) shift[k] = kshift[lane * 8 + k];
        if (cached) {
#pragma unroll
            for (int it = 0; it < kCacheRows; ++it) {
                const int64_t r = r0 + ty + (int64_t)it * RP;
                if (r < r1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float d = tile[it][k] - shift[k];
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
    BN_STAMP(2);
    red_partials(s, q, C, LP, RP, lane, ty, acc, smem);
    BN_STAMP(3);
    grid.sync();
    BN_STAMP(4);
    const float inv_m = 1.f / (float)M;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sd = __ldcg(acc + c), sq = __ldcg(acc + C + c);
        const float k0 = kshift[c];
        const float md = sd * inv_m;
        const float mean = k0 + md;
        float var = fmaf(-md, md, sq * inv_m);
        var = var > 0.f ? var : 0.f;
        const float invstd = rsqrtf(var + eps);
        const float sc = (gamma != nullptr ? gamma[c] : 1.f) * invstd;
        smem[c] = sc;
        smem[C + c] = (beta != nullptr ? beta[c] : 0.f) - mean * sc;
        if (blockIdx.x == 0) {
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
    BN_STAMP(5);
    if (ty < RP) {
        float sc[8], sh[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { sc[k] = smem[lane * 8 + k]; sh[k] = smem[C + lane * 8 + k]; }
        if (cached) {
#pragma unroll
            for (int it = 0; it < kCacheRows; ++it) {
                const int64_t r = r0 + ty + (int64_t)it * RP;
                if (r < r1) {
                    const int64_t off = r * C + lane * 8;
                    float rr[8];
                    if (kRes) Vec8<T>::load(res + off, rr);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float o = fmaf(tile[it][k], sc[k], sh[k]);
                        if (kRes) o += rr[k];
                        if (kRelu) o = o > 0.f ? o : 0.f;
                        tile[it][k] = o;
                    }
                    Vec8<T>::store(y + off, tile[it]);
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
    // Off the critical path (CTA 0 used to do this BEFORE loading its tile, delaying everyone at the grid barrier): re-zero
    // the WHOLE accumulator buffer the previous launch used -- it may have been a layer with more channels than this one.
    if (blockIdx.x == 0) {
        float4* other = reinterpret_cast<float4*>(acc2 + (par ^ 1u) * kAccStride);
        for (int c = threadIdx.x; c < kAccStride / 4; c += blockDim.x) other[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    BN_STAMP(6);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *parity = par ^ 1u;
        if (nbt != nullptr) *nbt += 1;
    }
}

template <typename T, bool kRelu, bool kRes>
__global__ void __launch_bounds__(kThreads)
bn_bwd_fused_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x, T* __restrict__ dx,
                    T* __restrict__ dres, int64_t M, int C, int rows_per_cta, const float* __restrict__ gamma,
                    const float* __restrict__ mean, const float* __restrict__ invstd, float* dgamma, float* dbeta,
                    float* acc2, unsigned* parity) {
    cg::grid_group grid = cg::this_grid();
    BN_STAMP(8);                               // backward: [8] entry [9] tile loaded + sums [10] block reduce + REDs [11] barrier
    extern __shared__ float smem[];            //           [12] totals staged [13] dx (and dres) stored
    // smem: max(4096, 3C) floats: reduce scratch, later a | mean(g) | mean(g xhat)
    const unsigned par = *reinterpret_cast<volatile unsigned*>(parity) & 1u;
    float* acc = acc2 + par * kAccStride;
    const int LP = C >> 3, RP = blockDim.x / LP;
    const int lane = threadIdx.x % LP, ty = threadIdx.x / LP;
    float sg[8], sgx[8], mu[8], is[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sg[k] = 0.f; sgx[k] = 0.f; mu[k] = 0.f; is[k] = 0.f; }
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = (r0 + rows_per_cta < M) ? r0 + rows_per_cta : M;
    // Common case (every ResNet-18 / CIFAR layer): the CTA's tile is <= kBwdCacheRows rows per thread.  dy and x stay in
    // registers PACKED as loaded (4 registers per 8 bf16) between the two phases and the ReLU mask shrinks to 8 bits per
    // row, so phase 2 reads nothing from memory and all phase-1 loads are in flight at once (the row loop below it issues
    // one row's loads per iteration: ~0.7 us of L2 latency per iteration, twice).
    constexpr int kBwdCacheRows = 8;
    constexpr bool kCanCache = Vec8<T>::kRawRegs == 4;
    const bool cached = kCanCache && rows_per_cta <= kBwdCacheRows * RP;
    typename Vec8<T>::Raw graw[kCanCache ? kBwdCacheRows : 1], xraw[kCanCache ? kBwdCacheRows : 1];
    uint32_t mask_lo = 0u, mask_hi = 0u;                   // bit (8 * (it % 4) + k) of lo (it < 4) / hi: y > 0
    if (ty < RP) {
        load8f(mean + lane * 8, mu);
        load8f(invstd + lane * 8, is);
        if (cached) {
            typename Vec8<T>::Raw yraw[kCanCache ? kBwdCacheRows : 1];
#pragma unroll
            for (int it = 0; it < kBwdCacheRows; ++it) {
                const int64_t r = r0 + ty + (int64_t)it * RP;
                if (r < r1) {
                    const int64_t off = r * C + lane * 8;
                    graw[it] = Vec8<T>::load_raw(dy + off);
                    xraw[it] = Vec8<T>::load_raw(x + off);
                    if (kRelu) yraw[it] = Vec8<T>::load_raw(y + off);
                }
            }
#pragma unroll
            for (int it = 0; it < kBwdCacheRows; ++it) {
                const int64_t r = r0 + ty + (int64_t)it * RP;
                if (r < r1) {
                    float g[8], xv[8], yv[8];
                    Vec8<T>::unpack(graw[it], g);
                    Vec8<T>::unpack(xraw[it], xv);
                    if (kRelu) Vec8<T>::unpack(yraw[it], yv);
                    uint32_t bits = 0u;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const bool on = !kRelu || yv[k] > 0.f;
                        bits |= (on ? 1u : 0u) << k;
                        const float gi = on ? g[k] : 0.f;
                        sg[k] += gi;
                        sgx[k] = fmaf(gi, (xv[k] - mu[k]) * is[k], sgx[k]);
                    }
                    if (it < 4) mask_lo |= bits << (8 * it); else mask_hi |= bits << (8 * (it - 4));
                }
            }
        } else {
            for (int64_t r = r0 + ty; r < r1; r += RP) {
                float g[8], xv[8], yv[8];
                const int64_t off = r * C + lane * 8;
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
    BN_STAMP(9);
    red_partials(sg, sgx, C, LP, RP, lane, ty, acc, smem);
    BN_STAMP(10);
    grid.sync();
    BN_STAMP(11);
    const float inv_m = 1.f / (float)M;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float a = __ldcg(acc + c), b = __ldcg(acc + C + c);
        smem[c] = (gamma != nullptr ? gamma[c] : 1.f) * invstd[c];
        smem[C + c] = a * inv_m;
        smem[2 * C + c] = b * inv_m;
        if (blockIdx.x == 0) {
            if (dbeta != nullptr) dbeta[c] = a;
            if (dgamma != nullptr) dgamma[c] = b;
        }
    }
    __syncthreads();
    BN_STAMP(12);
    if (ty < RP) {
        float a[8], mg[8], mgx[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a[k] = smem[lane * 8 + k];
            mg[k] = smem[C + lane * 8 + k];
            mgx[k] = smem[2 * C + lane * 8 + k];
        }
        if (cached) {
#pragma unroll
            for (int it = 0; it < kBwdCacheRows; ++it) {
                const int64_t r = r0 + ty + (int64_t)it * RP;
                if (r < r1) {
                    float g[8], xv[8];
                    Vec8<T>::unpack(graw[it], g);
                    Vec8<T>::unpack(xraw[it], xv);
                    const uint32_t bits = (it < 4 ? mask_lo >> (8 * it) : mask_hi >> (8 * (it - 4))) & 0xffu;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float gi = ((bits >> k) & 1u) ? g[k] : 0.f;
                        g[k] = gi;
                        xv[k] = a[k] * (gi - mg[k] - (xv[k] - mu[k]) * is[k] * mgx[k]);
                    }
                    const int64_t off = r * C + lane * 8;
                    Vec8<T>::store(dx + off, xv);
                    if (kRes) Vec8<T>::store(dres + off, g);
                }
            }
        } else {
            for (int64_t r = r0 + ty; r < r1; r += RP) {   // second read of the tile comes from L2
                float g[8], xv[8], yv[8];
                const int64_t off = r * C + lane * 8;
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
    BN_STAMP(13);
    if (blockIdx.x == 0) {  // see the forward kernel: zero the other parity's accumulators after the useful work
        float4* other = reinterpret_cast<float4*>(acc2 + (par ^ 1u) * kAccStride);
        for (int c = threadIdx.x; c < kAccStride / 4; c += blockDim.x) other[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *parity = par ^ 1u;
}

inline int reduce_grid(int64_t M, int C, int* rows_per_cta) {
    const int LP = C / 8, RP = kThreads / LP;
    int64_t rows = (int64_t)RP * 8;              // at least 8 rows per thread
    const int64_t even = (M + 147) / 148;          // ... and no more than one CTA per SM
    if (rows < even) rows = even;
    rows = (rows + RP - 1) / RP * RP;
    *rows_per_cta = (int)rows;
    return (int)((M + rows - 1) / rows);
}

inline int apply_grid(int64_t nvec) {
    int64_t g = (nvec + kThreads - 1) / kThreads;
    const int64_t cap = 148 * 8;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

#define FL4H_BN_DISPATCH(T, relu, res, ...)                                       \
    do {                                                                           \
        if (relu) { if (res) { __VA_ARGS__(T, true, true); } else { __VA_ARGS__(T, true, false); } } \
        else      { if (res) { __VA_ARGS__(T, false, true); } else { __VA_ARGS__(T, false, false); } } \
    } while (0)

}  // namespace

extern "C" {

// Copies the 16 phase timestamps (CTA 0, SM clock cycles) of the last fused forward launch to `out`.
int fl4h_bn_debug_read(long long* out) {
    return static_cast<int>(cudaMemcpyFromSymbol(out, g_bn_phase_clock, sizeof(long long) * 16));
}

int fl4h_bn_supported(int64_t M, int C) { return (C % 8 == 0 && C / 8 <= kThreads && M >= 1) ? 1 : 0; }

// workspace (persistent, zero-initialised by the caller once): acc [2*C] fp32, counter [1] u32; scratch scale_shift [2*C].
// `acc` layout: [0, 2*kAccStride) double-buffered accumulators of the fused kernels, [2*kAccStride, 3*kAccStride) the
// accumulator of the two-kernel fallback.  counter[0] = election counter (fallback), counter[1] = buffer parity (fused).
int fl4h_bn_fwd_train(const void* x, const void* res, void* y, int64_t M, int C, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, int64_t* nbt, float momentum, float eps, float* mean_out,
                      float* invstd_out, float* scale_shift, float* acc, unsigned* counter, int is_bf16, int relu,
                      int allow_fused, cudaStream_t stream) {
    int rows = 0;
    const int grid = reduce_grid(M, C, &rows);
    const size_t smem = (size_t)kThreads * 16 * sizeof(float);
    const int64_t nvec = M * C / 8;
    const bool has_res = res != nullptr;
    if (allow_fused) {
        unsigned* parity = counter + 1;
        void* kargs[] = {(void*)&x, (void*)&res, (void*)&y, (void*)&M, (void*)&C, (void*)&rows, (void*)&gamma, (void*)&beta,
                         (void*)&running_mean, (void*)&running_var, (void*)&nbt, (void*)&momentum, (void*)&eps,
                         (void*)&mean_out, (void*)&invstd_out, (void*)&acc, (void*)&parity};
        const size_t fsmem = (size_t)(4096 + C) * sizeof(float);
        const void* fn = nullptr;
#define PICK_FWD(T, R, S) fn = (const void*)bn_fwd_fused_kernel<T, R, S>
        if (is_bf16) FL4H_BN_DISPATCH(__nv_bfloat16, relu, has_res, PICK_FWD);
        else FL4H_BN_DISPATCH(float, relu, has_res, PICK_FWD);
#undef PICK_FWD
        cudaError_t err = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), kargs, fsmem, stream);
        if (err == cudaSuccess) return 0;
        (void)cudaGetLastError();  // not launchable cooperatively right now: use the two-kernel path below
    }
    acc += 2 * kAccStride;
#define LAUNCH_FWD(T, R, S)                                                                                          \
    bn_apply_kernel<T, R, S><<<apply_grid(nvec), kThreads, 0, stream>>>((const T*)x, (const T*)res, (T*)y, nvec, C, \
        scale_shift, nullptr, nullptr, nullptr, nullptr, eps, 0)
    if (is_bf16) {
        bn_fwd_stats_kernel<__nv_bfloat16><<<grid, kThreads, smem, stream>>>((const __nv_bfloat16*)x, M, C, rows, gamma,
            beta, running_mean, running_var, nbt, momentum, eps, mean_out, invstd_out, scale_shift, acc, counter);
        FL4H_BN_DISPATCH(__nv_bfloat16, relu, has_res, LAUNCH_FWD);
    } else {
        bn_fwd_stats_kernel<float><<<grid, kThreads, smem, stream>>>((const float*)x, M, C, rows, gamma, beta,
            running_mean, running_var, nbt, momentum, eps, mean_out, invstd_out, scale_shift, acc, counter);
        FL4H_BN_DISPATCH(float, relu, has_res, LAUNCH_FWD);
    }
#undef LAUNCH_FWD
    return (int)cudaGetLastError();
}

// Training forward from convolution-epilogue statistics (see bn_apply_presum_kernel).  `done` is a zero-initialised
// election counter owned by the caller (one per `sums` buffer).
int fl4h_bn_fwd_train_presum(const void* x, const void* res, void* y, int64_t M, int C, float* sums, const float* gamma,
                             const float* beta, float* running_mean, float* running_var, long long* nbt, float momentum,
                             float eps, float* mean_out, float* invstd_out, unsigned* done, int is_bf16, int relu,
                             cudaStream_t stream) {
    const int64_t nvec = M * C / 8;
    const bool has_res = res != nullptr;
    if ((C & 7) != 0 || kThreads % (C >> 3) != 0) return (int)cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(apply_grid(nvec));
    cfg.blockDim = dim3(kThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    const char* pdl = getenv("FL4H_PDL");
    cfg.numAttrs = (pdl != nullptr && pdl[0] == '0') ? 0 : 1;
    cudaError_t lerr = cudaSuccess;
#define LAUNCH_PRESUM(T, R, S)                                                                                         \
    lerr = cudaLaunchKernelEx(&cfg, bn_apply_presum_kernel<T, R, S>, (const T*)x, (const T*)res, (T*)y, nvec, C,       \
        M, sums, gamma, beta, running_mean, running_var, (long long*)nbt, momentum, eps, mean_out, invstd_out, done)
    if (is_bf16) FL4H_BN_DISPATCH(__nv_bfloat16, relu, has_res, LAUNCH_PRESUM);
    else FL4H_BN_DISPATCH(float, relu, has_res, LAUNCH_PRESUM);
#undef LAUNCH_PRESUM
    return (int)lerr;
}

int fl4h_bn_fwd_eval(const void* x, const void* res, void* y, int64_t M, int C, const float* gamma, const float* beta,
                     const float* running_mean, const float* running_var, float eps, int is_bf16, int relu,
                     cudaStream_t stream) {
    const int64_t nvec = M * C / 8;
    const bool has_res = res != nullptr;
#define LAUNCH_EVAL(T, R, S)                                                                                         \
    bn_apply_kernel<T, R, S><<<apply_grid(nvec), kThreads, 0, stream>>>((const T*)x, (const T*)res, (T*)y, nvec, C, \
        nullptr, gamma, beta, running_mean, running_var, eps, 1)
    if (is_bf16) FL4H_BN_DISPATCH(__nv_bfloat16, relu, has_res, LAUNCH_EVAL);
    else FL4H_BN_DISPATCH(float, relu, has_res, LAUNCH_EVAL);
#undef LAUNCH_EVAL
    return (int)cudaGetLastError();
}

int fl4h_bn_bwd(const void* dy, const void* y, const void* x, int64_t M, int C, const float* gamma, const float* mean,
                const float* invstd, void* dx, void* dres, float* dgamma, float* dbeta, float* coef, float* acc,
                unsigned* counter, int is_bf16, int relu, int allow_fused, cudaStream_t stream) {
    int rows = 0;
    const int grid = reduce_grid(M, C, &rows);
    const size_t smem = (size_t)kThreads * 16 * sizeof(float);
    const int64_t nvec = M * C / 8;
    const bool has_res = dres != nullptr;
    if (allow_fused) {
        unsigned* parity = counter + 1;
        void* kargs[] = {(void*)&dy, (void*)&y, (void*)&x, (void*)&dx, (void*)&dres, (void*)&M, (void*)&C, (void*)&rows,
                         (void*)&gamma, (void*)&mean, (void*)&invstd, (void*)&dgamma, (void*)&dbeta, (void*)&acc, (void*)&parity};
        const size_t fsmem = (size_t)(3 * C > 4096 ? 3 * C : 4096) * sizeof(float);
        const void* fn = nullptr;
#define PICK_BWD(T, R, S) fn = (const void*)bn_bwd_fused_kernel<T, R, S>
        if (is_bf16) FL4H_BN_DISPATCH(__nv_bfloat16, relu, has_res, PICK_BWD);
        else FL4H_BN_DISPATCH(float, relu, has_res, PICK_BWD);
#undef PICK_BWD
        cudaError_t err = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), kargs, fsmem, stream);
        if (err == cudaSuccess) return 0;
        (void)cudaGetLastError();
    }
    acc += 2 * kAccStride;
#define LAUNCH_BWD(T, R, S)                                                                                     \
    bn_bwd_apply_kernel<T, R, S><<<apply_grid(nvec), kThreads, 0, stream>>>((const T*)dy, (const T*)y, (const T*)x, \
        (T*)dx, (T*)dres, nvec, C, mean, invstd, coef)
    if (is_bf16) {
        if (relu) bn_bwd_reduce_kernel<__nv_bfloat16, true><<<grid, kThreads, smem, stream>>>((const __nv_bfloat16*)dy,
            (const __nv_bfloat16*)y, (const __nv_bfloat16*)x, M, C, rows, gamma, mean, invstd, dgamma, dbeta, coef, acc, counter);
        else bn_bwd_reduce_kernel<__nv_bfloat16, false><<<grid, kThreads, smem, stream>>>((const __nv_bfloat16*)dy,
            (const __nv_bfloat16*)y, (const __nv_bfloat16*)x, M, C, rows, gamma, mean, invstd, dgamma, dbeta, coef, acc, counter);
        FL4H_BN_DISPATCH(__nv_bfloat16, relu, has_res, LAUNCH_BWD);
    } else {
        if (relu) bn_bwd_reduce_kernel<float, true><<<grid, kThreads, smem, stream>>>((const float*)dy, (const float*)y,
            (const float*)x, M, C, rows, gamma, mean, invstd, dgamma, dbeta, coef, acc, counter);
        else bn_bwd_reduce_kernel<float, false><<<grid, kThreads, smem, stream>>>((const float*)dy, (const float*)y,
            (const float*)x, M, C, rows, gamma, mean, invstd, dgamma, dbeta, coef, acc, counter);
        FL4H_BN_DISPATCH(float, relu, has_res, LAUNCH_BWD);
    }
#undef LAUNCH_BWD
    return (int)cudaGetLastError();
}

}  // extern "C"
