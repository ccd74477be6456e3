// Flat-arena fused elementwise / reduction kernels for sm_100a.
//
// Everything in the FL round that is not a GEMM is a streaming pass over the rank's flat parameter arena, so these
// kernels are written once, vectorized (128-bit), grid-strided over a grid sized from the SM count, and take their
// scalars from a small device-resident hyper-parameter block so that a captured CUDA graph can be replayed while the
// host changes lr / mu / step between replays.
//
// What the reference does for the same work (per-layer Python loops):
//   local SGD/AdamW ............... torch.optim foreach            (examples/basic_example/client.py:34)
//   FedProx penalty grad mu(w-w_t) . autograd over WeightDriftLoss  (fl4health/losses/weight_drift_loss.py:57-64)
//   SCAFFOLD g += c - c_i .......... per-param H2D + add            (fl4health/clients/scaffold_client.py:187-197)
//   FedAvg reduce(np.add) .......... NumPy on the server CPU        (fl4health/strategies/aggregate_utils.py:23-32)
//   FedOpt server update ........... NumPy per layer                (Flower FedAdam; fl4health/strategies/flash.py:125-170)
//
// C ABI (ctypes): every entry point returns cudaError_t as int and launches on the given stream.

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define FL4H_MAX_SRC 16

namespace {

constexpr int kThreads = 256;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int g_num_sms = 0;
inline int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

// grid for a streaming pass over n elements with `vec` elements per thread-iteration: enough CTAs for ~8 resident
// CTAs/SM (256 thr) but never more than the work.
inline int stream_grid(int64_t n, int vec) {
    int64_t blocks = ceil_div(n, (int64_t)kThreads * vec);
    int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ld4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&lo);
    packed.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = packed;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-level sum, result valid in thread 0
__device__ __forceinline__ float block_sum(float v) {
    __shared__ float warp_part[kThreads / 32];
    v = warp_sum(v);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_part[wid] = v;
    __syncthreads();
    float total = 0.f;
    if (wid == 0) {
        total = lane < (kThreads / 32) ? warp_part[lane] : 0.f;
        total = warp_sum(total);
    }
    __syncthreads();
    return total;
}

// ---------------------------------------------------------------------------------------------------------------
// Hyper-parameter block layout (floats), shared by the optimizer kernels.
//   [0] lr  [1] momentum  [2] dampening  [3] weight_decay  [4] mu (FedProx/Ditto drift weight)  [5] nesterov(0/1)
//   [6] beta1 [7] beta2 [8] eps [9] step (float, incremented by the kernel wrapper's tiny tick kernel)
//   [10] first_step flag for SGD momentum buffer initialisation (1 => buf = g)
//   [11] grad_scale (1/loss_scale or DP 1/batch)  [12] noise_std (DP)  [13] max_grad_norm (reserved)
// ---------------------------------------------------------------------------------------------------------------
enum { HP_LR = 0, HP_MOM, HP_DAMP, HP_WD, HP_MU, HP_NESTEROV, HP_B1, HP_B2, HP_EPS, HP_STEP, HP_FIRST, HP_GSCALE,
       HP_NOISE, HP_MAXNORM, HP_COUNT = 16 };

template <bool kHasAnchor, bool kHasCv, bool kHasShadow, bool kGradBf16>
__device__ __forceinline__ void sgd_update4(float4& w, float4 g, float4& m, const float4 a, const float4 c,
                                            const float lr, const float mom, const float damp, const float wd,
                                            const float mu, const bool nesterov, const bool first) {
    float* wp = &w.x;
    float* gp = &g.x;
    float* mp = &m.x;
    const float* ap = &a.x;
    const float* cp = &c.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float gi = gp[i];
        if (kHasCv) gi += cp[i];                       // SCAFFOLD: g += (c - c_i), precomputed at broadcast time
        // zero coefficients are skipped like torch.optim does: 0 * inf is NaN, and FedPM's scores are legitimately
        // +-inf after a Bayesian aggregate of 0 or 1 (sigmoid_inverse)
        if (kHasAnchor && mu != 0.f) gi += mu * (wp[i] - ap[i]);    // FedProx/Ditto/MR-MTL: analytic grad of mu/2 |w - w_t|^2
        if (wd != 0.f) gi += wd * wp[i];
        float buf = first ? gi : mom * mp[i] + (1.f - damp) * gi;
        mp[i] = buf;
        float upd = (mom != 0.f) ? (nesterov ? gi + mom * buf : buf) : gi;
        wp[i] -= lr * upd;
    }
}

template <bool kHasAnchor, bool kHasCv, bool kHasShadow, bool kGradBf16>
__global__ void __launch_bounds__(kThreads)
sgd_step_kernel(float* __restrict__ w, const void* __restrict__ grad, float* __restrict__ mbuf,
                const float* __restrict__ anchor, const float* __restrict__ cv, __nv_bfloat16* __restrict__ shadow,
                const float* __restrict__ hp, int64_t n) {
    const float gscale = hp[HP_GSCALE];
    const float lr = hp[HP_LR], mom = hp[HP_MOM], damp = hp[HP_DAMP], wd = hp[HP_WD], mu = hp[HP_MU];
    const bool nesterov = hp[HP_NESTEROV] != 0.f, first = hp[HP_FIRST] != 0.f;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        float4 wv = ld4(w + e);
        float4 gv;
        if (kGradBf16) {
            uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(grad) + e);
            __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
            __nv_bfloat162 hi = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
            float2 flo = __bfloat1622float2(lo), fhi = __bfloat1622float2(hi);
            gv = make_float4(flo.x, flo.y, fhi.x, fhi.y);
        } else {
            gv = ld4(reinterpret_cast<const float*>(grad) + e);
        }
        gv.x *= gscale; gv.y *= gscale; gv.z *= gscale; gv.w *= gscale;
        float4 mv = (mom != 0.f && !first) ? ld4(mbuf + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 av = kHasAnchor ? ld4(anchor + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 cvv = kHasCv ? ld4(cv + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        sgd_update4<kHasAnchor, kHasCv, kHasShadow, kGradBf16>(wv, gv, mv, av, cvv, lr, mom, damp, wd, mu, nesterov,
                                                               first);
        st4(w + e, wv);
        if (mom != 0.f) st4(mbuf + e, mv);
        if (kHasShadow) st_bf16x4(shadow + e, wv);
    }
}

template <bool kHasAnchor, bool kHasShadow, bool kGradBf16>
__global__ void __launch_bounds__(kThreads)
adamw_step_kernel(float* __restrict__ w, const void* __restrict__ grad, float* __restrict__ m1,
                  float* __restrict__ m2, const float* __restrict__ anchor, __nv_bfloat16* __restrict__ shadow,
                  const float* __restrict__ hp, int64_t n, int decoupled) {
    const float gscale = hp[HP_GSCALE];
    const float lr = hp[HP_LR], wd = hp[HP_WD], mu = hp[HP_MU];
    const float b1 = hp[HP_B1], b2 = hp[HP_B2], eps = hp[HP_EPS], step = hp[HP_STEP];
    const float bc1 = 1.f - __powf(b1, step);
    const float bc2 = 1.f - __powf(b2, step);
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        float4 wv = ld4(w + e);
        float4 gv;
        if (kGradBf16) {
            uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(grad) + e);
            __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
            __nv_bfloat162 hi = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
            float2 flo = __bfloat1622float2(lo), fhi = __bfloat1622float2(hi);
            gv = make_float4(flo.x, flo.y, fhi.x, fhi.y);
        } else {
            gv = ld4(reinterpret_cast<const float*>(grad) + e);
        }
        float4 mv = ld4(m1 + e), vv = ld4(m2 + e);
        float4 av = kHasAnchor ? ld4(anchor + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        float* wp = &wv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x; const float* ap = &av.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gi = gp[k] * gscale;
            if (kHasAnchor && mu != 0.f) gi += mu * (wp[k] - ap[k]);
            if (wd != 0.f) { if (decoupled) wp[k] *= (1.f - lr * wd); else gi += wd * wp[k]; }
            mp[k] = b1 * mp[k] + (1.f - b1) * gi;
            vp[k] = b2 * vp[k] + (1.f - b2) * gi * gi;
            float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
            wp[k] -= step_size * mp[k] / denom;
        }
        st4(w + e, wv); st4(m1 + e, mv); st4(m2 + e, vv);
        if (kHasShadow) st_bf16x4(shadow + e, wv);
    }
}

__global__ void tick_kernel(float* hp) {
    hp[HP_STEP] += 1.f;
}
__global__ void clear_first_kernel(float* hp) { hp[HP_FIRST] = 0.f; }

// ---------------------------------------------------------------------------------------------------------------
// Weighted K-way aggregation:  out = sum_k coef[k] * src[k]   (fixed k order => bit-deterministic)
// epilogue modes implement the strategy update on top of the mean in the same pass.
// ---------------------------------------------------------------------------------------------------------------
struct SrcPack {
    const float* src[FL4H_MAX_SRC];
    float coef[FL4H_MAX_SRC];
    int k;
};

enum { EPI_NONE = 0, EPI_FEDADAM = 1, EPI_FEDADAGRAD = 2, EPI_FEDYOGI = 3, EPI_SERVER_LR = 4, EPI_MOMENTUM = 5 };

// epilogue scalars: [0] eta  [1] beta1  [2] beta2  [3] tau  [4] bias-corrected eta flag  [5] step
struct EpiArgs {
    float eta, beta1, beta2, tau, server_lr, momentum;
    int mode;
};

__device__ __forceinline__ float epi_apply(int mode, const EpiArgs& ea, float avg, float& wcur, float& m, float& v) {
    // returns the new global weight.  wcur = current server weight (x_t); m,v = server optimizer moments.
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
        case EPI_SERVER_LR:  // SCAFFOLD weights: x <- x + eta_s (ybar - x)
            return wcur + ea.server_lr * (avg - wcur);
        case EPI_MOMENTUM: {  // FedAvgM on the mean delta: m <- beta m + avg ; w <- w + eta m   (avg is a delta)
            m = ea.momentum * m + avg;
            return wcur + ea.server_lr * m;
        }
        default:
            return avg;
    }
}

__global__ void __launch_bounds__(kThreads)
weighted_sum_kernel(float* __restrict__ out, SrcPack pack, float* __restrict__ wcur, float* __restrict__ m,
                    float* __restrict__ v, EpiArgs ea, int64_t n) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // issue all K loads before consuming them (memory-level parallelism), fixed order accumulate
        float4 vals[FL4H_MAX_SRC];
#pragma unroll
        for (int k = 0; k < FL4H_MAX_SRC; ++k)
            if (k < pack.k) vals[k] = ld4_stream(pack.src[k] + e);
#pragma unroll
        for (int k = 0; k < FL4H_MAX_SRC; ++k)
            if (k < pack.k) {
                const float c = pack.coef[k];
                acc.x = fmaf(c, vals[k].x, acc.x); acc.y = fmaf(c, vals[k].y, acc.y);
                acc.z = fmaf(c, vals[k].z, acc.z); acc.w = fmaf(c, vals[k].w, acc.w);
            }
        if (ea.mode != EPI_NONE) {
            float4 wv = ld4(wcur + e);
            float4 mv = m ? ld4(m + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 vv = v ? ld4(v + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            acc.x = epi_apply(ea.mode, ea, acc.x, wv.x, mv.x, vv.x);
            acc.y = epi_apply(ea.mode, ea, acc.y, wv.y, mv.y, vv.y);
            acc.z = epi_apply(ea.mode, ea, acc.z, wv.z, mv.z, vv.z);
            acc.w = epi_apply(ea.mode, ea, acc.w, wv.w, mv.w, vv.w);
            if (m) st4(m + e, mv);
            if (v) st4(v + e, vv);
        }
        st4(out + e, acc);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Broadcast unpack (receiver-side tail of parameter_exchange): one read of the incoming global buffer, up to four
// writes: w <- g ; anchor w_t <- g ; bf16 compute shadow <- g ; SCAFFOLD correction d <- c - c_i.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
bcast_unpack_kernel(const float* __restrict__ g, float* __restrict__ w, float* __restrict__ anchor,
                    __nv_bfloat16* __restrict__ shadow, const float* __restrict__ c_server,
                    const float* __restrict__ c_local, float* __restrict__ cv_out, int64_t n) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        float4 gv = ld4_stream(g + e);
        if (w) st4(w + e, gv);
        if (anchor) st4(anchor + e, gv);
        if (shadow) st_bf16x4(shadow + e, gv);
        if (cv_out) {
            float4 cs = ld4_stream(c_server + e), cl = ld4(c_local + e);
            st4(cv_out + e, make_float4(cs.x - cl.x, cs.y - cl.y, cs.z - cl.z, cs.w - cl.w));
        }
    }
}

// SCAFFOLD client round-end: c_i+ = c_i - c + (x - y)/(K*lr); delta_c = c_i+ - c_i ; c_i <- c_i+
__global__ void __launch_bounds__(kThreads)
scaffold_variate_kernel(const float* __restrict__ x_global, const float* __restrict__ y_local,
                        const float* __restrict__ c_server, float* __restrict__ c_local,
                        float* __restrict__ delta_c, float inv_k_lr, int64_t n) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        float4 x = ld4(x_global + e), y = ld4(y_local + e), c = ld4(c_server + e), ci = ld4(c_local + e);
        float4 nci = make_float4(ci.x - c.x + (x.x - y.x) * inv_k_lr, ci.y - c.y + (x.y - y.y) * inv_k_lr,
                                 ci.z - c.z + (x.z - y.z) * inv_k_lr, ci.w - c.w + (x.w - y.w) * inv_k_lr);
        st4(delta_c + e, make_float4(nci.x - ci.x, nci.y - ci.y, nci.z - ci.z, nci.w - ci.w));
        st4(c_local + e, nci);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Reductions: out[0] += sum((a-b)^2) (b may be null => sum a^2);  out[0] += dot(a, b);  two-level, one atomic/CTA.
// ---------------------------------------------------------------------------------------------------------------
template <int kMode>  // 0: sq diff, 1: dot, 2: apfl alpha-grad dot  <a-b, alpha*ga + (1-alpha)*gb>
__global__ void __launch_bounds__(kThreads)
reduce_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ ga,
              const float* __restrict__ gb, float alpha, float* __restrict__ out, int64_t n) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        float4 av = ld4(a + e);
        float4 bv = b ? ld4(b + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (kMode == 0) {
            float dx = av.x - bv.x, dy = av.y - bv.y, dz = av.z - bv.z, dw = av.w - bv.w;
            acc += dx * dx + dy * dy + dz * dz + dw * dw;
        } else if (kMode == 1) {
            acc += av.x * bv.x + av.y * bv.y + av.z * bv.z + av.w * bv.w;
        } else {
            float4 g1 = ld4(ga + e), g2 = ld4(gb + e);
            acc += (av.x - bv.x) * (alpha * g1.x + (1.f - alpha) * g2.x) +
                   (av.y - bv.y) * (alpha * g1.y + (1.f - alpha) * g2.y) +
                   (av.z - bv.z) * (alpha * g1.z + (1.f - alpha) * g2.z) +
                   (av.w - bv.w) * (alpha * g1.w + (1.f - alpha) * g2.w);
        }
    }
    float total = block_sum(acc);
    if (threadIdx.x == 0) atomicAdd(out, total);
}

// scale-clip for client-level DP: y = x * min(1, C / norm) with norm read from device memory; bit <- norm <= C
__global__ void __launch_bounds__(kThreads)
clip_scale_kernel(float* __restrict__ x, const float* __restrict__ sq_norm, float clip, float* __restrict__ bit,
                  int64_t n) {
    const float norm = sqrtf(*sq_norm);
    const float s = fminf(1.f, clip / (norm + 1e-12f));
    if (blockIdx.x == 0 && threadIdx.x == 0 && bit) *bit = norm <= clip ? 1.f : 0.f;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        float4 v = ld4(x + e);
        st4(x + e, make_float4(v.x * s, v.y * s, v.z * s, v.w * s));
    }
}

// Philox-free counter RNG (splitmix/xorshift hash -> Box-Muller) for DP noise inside streaming kernels.
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float2 gauss2(uint64_t seed, uint64_t idx) {
    uint64_t r = mix64(seed ^ mix64(idx));
    float u1 = ((uint32_t)(r >> 40) + 1.0f) * (1.0f / 16777217.0f);  // (0,1]
    float u2 = ((uint32_t)((r >> 8) & 0xFFFFFF)) * (1.0f / 16777216.0f);
    float rad = sqrtf(-2.f * __logf(u1));
    float s, c;
    __sincosf(6.28318530718f * u2, &s, &c);
    return make_float2(rad * c, rad * s);
}

// y += std * N(0,1)   (client-level DP noisy aggregate epilogue, instance-level DP grad noise)
__global__ void __launch_bounds__(kThreads)
add_gaussian_kernel(float* __restrict__ y, float stddev, uint64_t seed, int64_t n) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        float4 v = ld4(y + e);
        float2 a = gauss2(seed, (uint64_t)i * 2), b = gauss2(seed, (uint64_t)i * 2 + 1);
        st4(y + e, make_float4(v.x + stddev * a.x, v.y + stddev * a.y, v.z + stddev * b.x, v.w + stddev * b.y));
    }
}

// y = (y + std * N(0,1)) * scale with the stream taken from DEVICE memory: state[0] = seed, state[1] = draws so far.
// A captured CUDA graph replays the same launch arguments; reading the stream position from memory (and bumping it in
// a follow-up launch) gives every replay fresh noise -- DP-SGD's clip/noise/step inside one graph.
__global__ void __launch_bounds__(kThreads)
add_gaussian_state_kernel(float* __restrict__ y, float stddev, float scale, const uint64_t* __restrict__ state, int64_t n) {
    const uint64_t seed = mix64(state[0] ^ mix64(state[1] * 0xD1342543DE82EF95ull + 1ull));
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        float4 v = ld4(y + e);
        float2 a = gauss2(seed, (uint64_t)i * 2), b = gauss2(seed, (uint64_t)i * 2 + 1);
        st4(y + e, make_float4((v.x + stddev * a.x) * scale, (v.y + stddev * a.y) * scale, (v.z + stddev * b.x) * scale,
                               (v.w + stddev * b.y) * scale));
    }
}
__global__ void bump_draws_kernel(uint64_t* state) { state[1] += 1ull; }

// fp32 -> bf16 shadow refresh
__global__ void __launch_bounds__(kThreads)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e = i << 2;
        st_bf16x4(dst + e, ld4(src + e));
    }
}

// FedPM: bit-vote accumulate. masks are uint8 {0,1}; alpha += sum_k M_k ; beta += K - sum_k M_k ;
// theta = (alpha - 1) / (alpha + beta - 2)  (posterior mode)  -- fl4health/strategies/fedpm.py:128-154
struct MaskPack {
    const uint8_t* src[FL4H_MAX_SRC];
    int k;
};
__global__ void __launch_bounds__(kThreads)
fedpm_vote_kernel(MaskPack pack, float* __restrict__ alpha, float* __restrict__ beta, float* __restrict__ theta,
                  int bayesian, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < FL4H_MAX_SRC; ++k)
            if (k < pack.k) s += (float)pack.src[k][i];
        if (bayesian) {
            float a = alpha[i] + s, b = beta[i] + ((float)pack.k - s);
            alpha[i] = a; beta[i] = b;
            theta[i] = (a - 1.f) / (a + b - 2.f);
        } else {
            theta[i] = s / (float)pack.k;
        }
    }
}

// FedPM vote over BIT-PACKED masks (the cross-GPU form: every client ships n/32 words instead of n bytes and the K
// contributions are one all-gather).  pack: one ballot per 32 scores.  vote: a warp owns 32 consecutive words = 1024
// scores; in step t every lane reads word (base + t) of each client (one broadcast transaction) and lane l counts bit
// l, so the alpha / beta / theta accesses of a step are 32 consecutive floats.
template <typename T>
__global__ void __launch_bounds__(kThreads)
pack_mask_bits_kernel(const T* __restrict__ mask, uint32_t* __restrict__ words, int64_t n) {
    const int64_t padded = (n + 31) & ~int64_t(31);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += stride) {
        const bool set = i < n && mask[i] != T(0);
        const uint32_t word = __ballot_sync(0xffffffffu, set);
        if ((threadIdx.x & 31) == 0) words[i >> 5] = word;
    }
}

__global__ void __launch_bounds__(kThreads)
fedpm_vote_packed_kernel(const uint32_t* __restrict__ words, int k, int64_t n_words, float* __restrict__ alpha,
                         float* __restrict__ beta, float* __restrict__ theta, int bayesian, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t groups = (n_words + 31) >> 5;
    for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < groups; g += warps) {
        const int64_t base = g << 5;
        for (int t = 0; t < 32 && base + t < n_words; ++t) {
            const int64_t i = ((base + t) << 5) + lane;
            int votes = 0;
            for (int c = 0; c < k; ++c) votes += (int)((words[(int64_t)c * n_words + base + t] >> lane) & 1u);
            if (i >= n) continue;
            const float s = (float)votes;
            if (bayesian) {
                const float a = alpha[i] + s, b = beta[i] + ((float)k - s);
                alpha[i] = a; beta[i] = b;
                theta[i] = (a - 1.f) / (a + b - 2.f);
            } else {
                theta[i] = s / (float)k;
            }
        }
    }
}

}  // namespace

// ===================================================================================================================
// C ABI
// ===================================================================================================================
extern "C" {

int fl4h_num_sms() { return num_sms(); }

// `flags`: bit 0 = gradient is bf16; bit 1 = more ranges of the same parameter group follow, so the "first step" marker
// in the shared hyper-parameter block stays armed (a group over non-contiguous arena ranges is stepped range by range
// and every range must initialise its momentum as g, not (1 - dampening) g).
int fl4h_sgd_step(float* w, const void* grad, float* mbuf, const float* anchor, const float* cv, void* shadow,
                  float* hp, int64_t n, int flags, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    const int grad_is_bf16 = flags & 1;
    const int grid = stream_grid(n, 4);
    __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(shadow);
#define LAUNCH_SGD(A, C, S, G) \
    sgd_step_kernel<A, C, S, G><<<grid, kThreads, 0, stream>>>(w, grad, mbuf, anchor, cv, sh, hp, n)
    const int key = (anchor ? 8 : 0) | (cv ? 4 : 0) | (shadow ? 2 : 0) | (grad_is_bf16 ? 1 : 0);
    switch (key) {
        case 0: LAUNCH_SGD(false, false, false, false); break;
        case 1: LAUNCH_SGD(false, false, false, true); break;
        case 2: LAUNCH_SGD(false, false, true, false); break;
        case 3: LAUNCH_SGD(false, false, true, true); break;
        case 4: LAUNCH_SGD(false, true, false, false); break;
        case 5: LAUNCH_SGD(false, true, false, true); break;
        case 6: LAUNCH_SGD(false, true, true, false); break;
        case 7: LAUNCH_SGD(false, true, true, true); break;
        case 8: LAUNCH_SGD(true, false, false, false); break;
        case 9: LAUNCH_SGD(true, false, false, true); break;
        case 10: LAUNCH_SGD(true, false, true, false); break;
        case 11: LAUNCH_SGD(true, false, true, true); break;
        case 12: LAUNCH_SGD(true, true, false, false); break;
        case 13: LAUNCH_SGD(true, true, false, true); break;
        case 14: LAUNCH_SGD(true, true, true, false); break;
        default: LAUNCH_SGD(true, true, true, true); break;
    }
#undef LAUNCH_SGD
    if (!(flags & 2)) clear_first_kernel<<<1, 1, 0, stream>>>(hp);
    return (int)cudaGetLastError();
}

int fl4h_adamw_step(float* w, const void* grad, float* m1, float* m2, const float* anchor, void* shadow, float* hp,
                    int64_t n, int grad_is_bf16, int decoupled, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    const int grid = stream_grid(n, 4);
    __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(shadow);
    tick_kernel<<<1, 1, 0, stream>>>(hp);
#define LAUNCH_ADAM(A, S, G) \
    adamw_step_kernel<A, S, G><<<grid, kThreads, 0, stream>>>(w, grad, m1, m2, anchor, sh, hp, n, decoupled)
    const int key = (anchor ? 4 : 0) | (shadow ? 2 : 0) | (grad_is_bf16 ? 1 : 0);
    switch (key) {
        case 0: LAUNCH_ADAM(false, false, false); break;
        case 1: LAUNCH_ADAM(false, false, true); break;
        case 2: LAUNCH_ADAM(false, true, false); break;
        case 3: LAUNCH_ADAM(false, true, true); break;
        case 4: LAUNCH_ADAM(true, false, false); break;
        case 5: LAUNCH_ADAM(true, false, true); break;
        case 6: LAUNCH_ADAM(true, true, false); break;
        default: LAUNCH_ADAM(true, true, true); break;
    }
#undef LAUNCH_ADAM
    return (int)cudaGetLastError();
}

int fl4h_weighted_sum(float* out, const float* const* srcs, const float* coefs, int k, float* wcur, float* m,
                      float* v, int mode, float eta, float beta1, float beta2, float tau, float server_lr,
                      float momentum, int64_t n, cudaStream_t stream) {
    if (k < 1 || k > FL4H_MAX_SRC || (n & 3)) return (int)cudaErrorInvalidValue;
    SrcPack pack;
    pack.k = k;
    for (int i = 0; i < k; ++i) { pack.src[i] = srcs[i]; pack.coef[i] = coefs[i]; }
    for (int i = k; i < FL4H_MAX_SRC; ++i) { pack.src[i] = nullptr; pack.coef[i] = 0.f; }
    EpiArgs ea{eta, beta1, beta2, tau, server_lr, momentum, mode};
    weighted_sum_kernel<<<stream_grid(n, 4), kThreads, 0, stream>>>(out, pack, wcur, m, v, ea, n);
    return (int)cudaGetLastError();
}

int fl4h_bcast_unpack(const float* g, float* w, float* anchor, void* shadow, const float* c_server,
                      const float* c_local, float* cv_out, int64_t n, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    bcast_unpack_kernel<<<stream_grid(n, 4), kThreads, 0, stream>>>(
        g, w, anchor, reinterpret_cast<__nv_bfloat16*>(shadow), c_server, c_local, cv_out, n);
    return (int)cudaGetLastError();
}

int fl4h_scaffold_variate(const float* x_global, const float* y_local, const float* c_server, float* c_local,
                          float* delta_c, float inv_k_lr, int64_t n, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    scaffold_variate_kernel<<<stream_grid(n, 4), kThreads, 0, stream>>>(x_global, y_local, c_server, c_local,
                                                                        delta_c, inv_k_lr, n);
    return (int)cudaGetLastError();
}

// out must be pre-zeroed by the caller (or accumulate across calls)
int fl4h_sq_diff_sum(const float* a, const float* b, float* out, int64_t n, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    int grid = stream_grid(n, 4);
    if (grid > num_sms() * 4) grid = num_sms() * 4;
    reduce_kernel<0><<<grid, kThreads, 0, stream>>>(a, b, nullptr, nullptr, 0.f, out, n);
    return (int)cudaGetLastError();
}

int fl4h_dot(const float* a, const float* b, float* out, int64_t n, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    int grid = stream_grid(n, 4);
    if (grid > num_sms() * 4) grid = num_sms() * 4;
    reduce_kernel<1><<<grid, kThreads, 0, stream>>>(a, b, nullptr, nullptr, 0.f, out, n);
    return (int)cudaGetLastError();
}

int fl4h_apfl_alpha_grad(const float* w_local, const float* w_global, const float* g_local, const float* g_global,
                         float alpha, float* out, int64_t n, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    int grid = stream_grid(n, 4);
    if (grid > num_sms() * 4) grid = num_sms() * 4;
    reduce_kernel<2><<<grid, kThreads, 0, stream>>>(w_local, w_global, g_local, g_global, alpha, out, n);
    return (int)cudaGetLastError();
}

int fl4h_clip_scale(float* x, const float* sq_norm, float clip, float* bit, int64_t n, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    clip_scale_kernel<<<stream_grid(n, 4), kThreads, 0, stream>>>(x, sq_norm, clip, bit, n);
    return (int)cudaGetLastError();
}

int fl4h_add_gaussian(float* y, float stddev, uint64_t seed, int64_t n, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    add_gaussian_kernel<<<stream_grid(n, 4), kThreads, 0, stream>>>(y, stddev, seed, n);
    return (int)cudaGetLastError();
}

int fl4h_add_gaussian_state(float* y, float stddev, float scale, uint64_t* state, int64_t n, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    add_gaussian_state_kernel<<<stream_grid(n, 4), kThreads, 0, stream>>>(y, stddev, scale, state, n);
    bump_draws_kernel<<<1, 1, 0, stream>>>(state);
    return (int)cudaGetLastError();
}

int fl4h_cast_bf16(const float* src, void* dst, int64_t n, cudaStream_t stream) {
    if (n & 3) return (int)cudaErrorInvalidValue;
    cast_bf16_kernel<<<stream_grid(n, 4), kThreads, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
    return (int)cudaGetLastError();
}

int fl4h_fedpm_vote(const uint8_t* const* masks, int k, float* alpha, float* beta, float* theta, int bayesian,
                    int64_t n, cudaStream_t stream) {
    if (k < 1 || k > FL4H_MAX_SRC) return (int)cudaErrorInvalidValue;
    MaskPack pack;
    pack.k = k;
    for (int i = 0; i < FL4H_MAX_SRC; ++i) pack.src[i] = i < k ? masks[i] : nullptr;
    fedpm_vote_kernel<<<stream_grid(n, 1), kThreads, 0, stream>>>(pack, alpha, beta, theta, bayesian, n);
    return (int)cudaGetLastError();
}

// `kind`: 0 = uint8 masks, 1 = fp32 masks.  `words` holds ceil(n / 32) uint32.
int fl4h_pack_mask_bits(const void* mask, int kind, uint32_t* words, int64_t n, cudaStream_t stream) {
    const int64_t padded = (n + 31) & ~int64_t(31);
    if (kind == 0)
        pack_mask_bits_kernel<uint8_t><<<stream_grid(padded, 1), kThreads, 0, stream>>>((const uint8_t*)mask, words, n);
    else if (kind == 1)
        pack_mask_bits_kernel<float><<<stream_grid(padded, 1), kThreads, 0, stream>>>((const float*)mask, words, n);
    else
        return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}

// `words`: k rows of n_words (row c = client c's packed masks, as an all-gather leaves them).
int fl4h_fedpm_vote_packed(const uint32_t* words, int k, int64_t n_words, float* alpha, float* beta, float* theta,
                           int bayesian, int64_t n, cudaStream_t stream) {
    if (k < 1 || n_words * 32 < n) return (int)cudaErrorInvalidValue;
    fedpm_vote_packed_kernel<<<stream_grid(n_words * 32, 4), kThreads, 0, stream>>>(words, k, n_words, alpha, beta, theta,
                                                                                 bayesian, n);
    return (int)cudaGetLastError();
}

}  // extern "C"
