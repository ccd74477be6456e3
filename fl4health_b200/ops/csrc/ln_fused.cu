// Residual add + dropout + LayerNorm, forward and backward, one pass each  (sm_100a).
//
//   pre = residual + dropout(y)          out = LayerNorm(pre) * gamma + beta
//
// The transformer block epilogue ``x = LN(x + dropout(sublayer(x)))`` (reference workload:
// examples/bert_finetuning_example/client.py, HF BertSelfOutput / BertOutput) is three memory-bound ATen passes forward
// (dropout + mask, add, layer_norm) and four backward, with fp32 intermediates under autocast.  Here: one read of y and
// residual, one write of `pre` (kept for the backward) and of `out`; the dropout mask is never stored -- it is a pure
// function of (stream seed, draw index, element index) and the backward regenerates it.  The stream position lives in
// device memory and is advanced by a follow-up launch, so a captured CUDA graph drops different units on every replay.
//
// One warp per row, 8 contiguous elements per lane per 256-column chunk (16-byte bf16 vectors); H = CHUNKS * 256.
// Statistics are two-pass over the register-resident row (exact), computed on the values as rounded to the storage type
// so that the backward, which re-reads `pre`, differentiates exactly the function the forward evaluated.

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kThreads = kWarpsPerCta * 32;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint64_t stream_key(uint64_t seed, uint64_t draw) {
    return mix64(seed ^ mix64(draw * 0xD1342543DE82EF95ull + 1ull));
}

// keep-mask of the 8 elements starting at flat element index `e` (a multiple of 8): bit i set = element kept.
// 16 random bits per element, threshold = p * 65536.
__device__ __forceinline__ uint32_t keep_mask8(uint64_t key, uint64_t e, uint32_t threshold) {
    const uint64_t r0 = mix64(key ^ (e >> 2)), r1 = mix64(key ^ ((e >> 2) + 1));
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m |= (uint32_t)(((r0 >> (16 * i)) & 0xFFFFu) >= threshold) << i;
        m |= (uint32_t)(((r1 >> (16 * i)) & 0xFFFFu) >= threshold) << (4 + i);
    }
    return m;
}

template <typename T>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 raw = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 raw;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = raw;
    }
    static __device__ __forceinline__ float round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};
template <>
struct Vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    static __device__ __forceinline__ float round(float x) { return x; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T, int CHUNKS>
__global__ void __launch_bounds__(kThreads)
ln_fwd_kernel(const T* __restrict__ y, const T* __restrict__ residual, const float* __restrict__ gamma,
              const float* __restrict__ beta, T* __restrict__ out, T* __restrict__ pre, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, int64_t rows, float eps, float p_drop, const uint64_t* __restrict__ rng_state,
              uint64_t* __restrict__ used_draw) {
    constexpr int H = CHUNKS * 256;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerCta;
    const bool drop = p_drop > 0.f && rng_state != nullptr;
    uint64_t key = 0;
    uint32_t threshold = 0;
    float keep_scale = 1.f;
    if (drop) {
        key = stream_key(rng_state[0], rng_state[1]);
        threshold = (uint32_t)(p_drop * 65536.f);
        keep_scale = 1.f / (1.f - p_drop);
        if (blockIdx.x == 0 && threadIdx.x == 0 && used_draw != nullptr) *used_draw = rng_state[1];
    }
    float g[CHUNKS][8], b[CHUNKS][8];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        Vec8<float>::load(gamma + c * 256 + lane * 8, g[c]);
        Vec8<float>::load(beta + c * 256 + lane * 8, b[c]);
    }
    for (int64_t row = warp; row < rows; row += n_warps) {
        float v[CHUNKS][8];
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int64_t e = row * H + c * 256 + lane * 8;
            Vec8<T>::load(y + e, v[c]);
            if (drop) {
                const uint32_t m = keep_mask8(key, (uint64_t)e, threshold);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[c][i] = ((m >> i) & 1u) ? v[c][i] * keep_scale : 0.f;
            }
            if (residual != nullptr) {
                float r[8];
                Vec8<T>::load(residual + e, r);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[c][i] += r[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[c][i] = Vec8<T>::round(v[c][i]);
                sum += v[c][i];
            }
            if (pre != nullptr) Vec8<T>::store(pre + e, v[c]);
        }
        const float mean = warp_sum(sum) * (1.f / H);
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = v[c][i] - mean;
                sq += d * d;
            }
        const float rstd = rsqrtf(warp_sum(sq) * (1.f / H) + eps);
        if (lane == 0) {
            if (mean_out != nullptr) mean_out[row] = mean;
            if (rstd_out != nullptr) rstd_out[row] = rstd;
        }
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = (v[c][i] - mean) * rstd * g[c][i] + b[c][i];
            Vec8<T>::store(out + row * H + c * 256 + lane * 8, o);
        }
    }
}

// dpre = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dout * gamma;   dy = dpre * keep / (1 - p)
// dgamma += dout * xhat, dbeta += dout: per-lane registers over the CTA's rows, one smem reduction, fp32 atomics.
template <typename T, int CHUNKS>
__global__ void __launch_bounds__(kThreads)
ln_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ pre, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const float* __restrict__ gamma, T* __restrict__ dpre, T* __restrict__ dy,
              float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, float p_drop,
              const uint64_t* __restrict__ rng_state, const uint64_t* __restrict__ used_draw) {
    constexpr int H = CHUNKS * 256;
    __shared__ float acc_g[kWarpsPerCta][H], acc_b[kWarpsPerCta][H];
    const int lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + warp_in_cta;
    const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerCta;
    const bool drop = p_drop > 0.f && rng_state != nullptr && dy != nullptr;
    uint64_t key = 0;
    uint32_t threshold = 0;
    float keep_scale = 1.f;
    if (drop) {
        key = stream_key(rng_state[0], *used_draw);
        threshold = (uint32_t)(p_drop * 65536.f);
        keep_scale = 1.f / (1.f - p_drop);
    }
    float gam[CHUNKS][8], dg[CHUNKS][8], db[CHUNKS][8];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        Vec8<float>::load(gamma + c * 256 + lane * 8, gam[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) dg[c][i] = db[c][i] = 0.f;
    }
    for (int64_t row = warp; row < rows; row += n_warps) {
        const float mean = mean_in[row], rstd = rstd_in[row];
        float xh[CHUNKS][8], go[CHUNKS][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int64_t e = row * H + c * 256 + lane * 8;
            float d[8], x[8];
            Vec8<T>::load(dout + e, d);
            Vec8<T>::load(pre + e, x);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                xh[c][i] = (x[i] - mean) * rstd;
                go[c][i] = d[i] * gam[c][i];
                s1 += go[c][i];
                s2 += go[c][i] * xh[c][i];
                dg[c][i] += d[i] * xh[c][i];
                db[c][i] += d[i];
            }
        }
        const float c1 = warp_sum(s1) * (1.f / H), c2 = warp_sum(s2) * (1.f / H);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int64_t e = row * H + c * 256 + lane * 8;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = rstd * (go[c][i] - c1 - xh[c][i] * c2);
            Vec8<T>::store(dpre + e, o);
            if (drop) {
                const uint32_t m = keep_mask8(key, (uint64_t)e, threshold);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = ((m >> i) & 1u) ? o[i] * keep_scale : 0.f;
                Vec8<T>::store(dy + e, o);
            }
        }
    }
    // CTA reduction of the parameter gradients: every warp parks its partials in shared memory, then all 128 threads
    // add the four copies of "their" columns and publish them (2 * H / 128 atomics per thread, spread over the CTA)
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc_g[warp_in_cta][c * 256 + lane * 8 + i] = dg[c][i];
            acc_b[warp_in_cta][c * 256 + lane * 8 + i] = db[c][i];
        }
    __syncthreads();
    for (int col = threadIdx.x; col < H; col += kThreads) {
        float tg = 0.f, tb = 0.f;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) {
            tg += acc_g[w][col];
            tb += acc_b[w][col];
        }
        atomicAdd(dgamma + col, tg);
        atomicAdd(dbeta + col, tb);
    }
}

__global__ void bump_draws_kernel(uint64_t* state) { state[1] += 1ull; }

int grid_for(int64_t rows) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t wanted = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t cap = (int64_t)sms * 8;  // 8 CTAs of 4 warps per SM keep 32 rows in flight per SM
    return (int)(wanted < cap ? (wanted > 0 ? wanted : 1) : cap);
}

}  // namespace

extern "C" {

// y / residual / out / pre: [rows, H] of bf16 (is_bf16) or fp32; gamma / beta / mean / rstd fp32.  residual, pre, mean,
// rstd may be null.  rng_state = {seed, draws} (device) or null for no dropout; used_draw receives the draw index used.
int fl4h_ln_fwd(const void* y, const void* residual, const float* gamma, const float* beta, void* out, void* pre,
                float* mean, float* rstd, int64_t rows, int hidden, float eps, float p_drop, uint64_t* rng_state,
                uint64_t* used_draw, int is_bf16, cudaStream_t stream) {
    if (hidden % 256 != 0 || hidden < 256 || hidden > 1024) return (int)cudaErrorInvalidValue;
    const int grid = grid_for(rows);
#define LN_FWD(T, C)                                                                                                  \
    ln_fwd_kernel<T, C><<<grid, kThreads, 0, stream>>>((const T*)y, (const T*)residual, gamma, beta, (T*)out, (T*)pre, \
                                                       mean, rstd, rows, eps, p_drop, rng_state, used_draw)
    const int chunks = hidden / 256;
    if (is_bf16) {
        switch (chunks) {
            case 1: LN_FWD(__nv_bfloat16, 1); break;
            case 2: LN_FWD(__nv_bfloat16, 2); break;
            case 3: LN_FWD(__nv_bfloat16, 3); break;
            default: LN_FWD(__nv_bfloat16, 4); break;
        }
    } else {
        switch (chunks) {
            case 1: LN_FWD(float, 1); break;
            case 2: LN_FWD(float, 2); break;
            case 3: LN_FWD(float, 3); break;
            default: LN_FWD(float, 4); break;
        }
    }
#undef LN_FWD
    if (p_drop > 0.f && rng_state != nullptr) bump_draws_kernel<<<1, 1, 0, stream>>>(rng_state);
    return (int)cudaGetLastError();
}

// dgamma / dbeta must be zero on entry (fp32 [H]).  dy may be null (no dropout: the gradient of y IS dpre).
int fl4h_ln_bwd(const void* dout, const void* pre, const float* mean, const float* rstd, const float* gamma, void* dpre,
                void* dy, float* dgamma, float* dbeta, int64_t rows, int hidden, float p_drop, const uint64_t* rng_state,
                const uint64_t* used_draw, int is_bf16, cudaStream_t stream) {
    if (hidden % 256 != 0 || hidden < 256 || hidden > 1024) return (int)cudaErrorInvalidValue;
    int grid = grid_for(rows);
    if (grid > 148) grid = 148;  // fewer, longer CTAs: each ends with 2 * H atomics
#define LN_BWD(T, C)                                                                                                   \
    ln_bwd_kernel<T, C><<<grid, kThreads, 0, stream>>>((const T*)dout, (const T*)pre, mean, rstd, gamma, (T*)dpre, (T*)dy, \
                                                       dgamma, dbeta, rows, p_drop, rng_state, used_draw)
    const int chunks = hidden / 256;
    if (is_bf16) {
        switch (chunks) {
            case 1: LN_BWD(__nv_bfloat16, 1); break;
            case 2: LN_BWD(__nv_bfloat16, 2); break;
            case 3: LN_BWD(__nv_bfloat16, 3); break;
            default: LN_BWD(__nv_bfloat16, 4); break;
        }
    } else {
        switch (chunks) {
            case 1: LN_BWD(float, 1); break;
            case 2: LN_BWD(float, 2); break;
            case 3: LN_BWD(float, 3); break;
            default: LN_BWD(float, 4); break;
        }
    }
#undef LN_BWD
    return (int)cudaGetLastError();
}

}  // extern "C"
