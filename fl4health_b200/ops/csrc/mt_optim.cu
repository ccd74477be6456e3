// Multi-tensor ("table") optimizer steps for sm_100a: gradients stay where autograd produced them (one tensor per
// parameter, bf16 or fp32), the fp32 master weights / moments / FedProx anchor / SCAFFOLD correction / bf16 compute
// shadow live in the flat arena.  ONE launch updates every parameter of a group:
//
//   * no 60+ AccumulateGrad "+=" kernels into a flat gradient buffer (autograd *assigns* fresh grads because .grad is
//     None at backward time);
//   * no fp32->bf16 weight casts in the forward (the kernel writes the bf16 shadow the model computes with) and no
//     bf16->fp32 gradient casts in the backward (bf16 grads are consumed directly);
//   * the pointer table travels in the kernel parameter block (CUDA 12.1+: up to 32 KB), so a captured CUDA graph
//     replays it with no host work and no side buffer.
//
// Work decomposition: fixed 4096-element chunks; chunk -> tensor by binary search over a prefix table that also
// sits in the parameter block (uniform per CTA => constant-bank broadcast).

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define FL4H_MT_MAX 512

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 4096;

struct MtTable {
    const void* grad[FL4H_MT_MAX];
    int64_t offset[FL4H_MT_MAX];      // element offset of the parameter inside the arena
    int32_t numel[FL4H_MT_MAX];
    int32_t chunk_prefix[FL4H_MT_MAX + 1];
    uint8_t grad_bf16[FL4H_MT_MAX];
    int32_t count;
};

enum { HP_LR = 0, HP_MOM, HP_DAMP, HP_WD, HP_MU, HP_NESTEROV, HP_B1, HP_B2, HP_EPS, HP_STEP, HP_FIRST, HP_GSCALE };

__device__ __forceinline__ int find_tensor(const MtTable& t, int chunk) {
    int lo = 0, hi = t.count;  // invariant: prefix[lo] <= chunk < prefix[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (t.chunk_prefix[mid] <= chunk) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float4 load_grad4(const void* base, bool is_bf16, int64_t e, int32_t numel, bool aligned) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (is_bf16) {
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base);
        if (aligned && e + 3 < numel) {
            uint2 raw = *reinterpret_cast<const uint2*>(p + e);
            float2 lo = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.x));
            float2 hi = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.y));
            g = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
            float* gp = &g.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) if (e + k < numel) gp[k] = __bfloat162float(p[e + k]);
        }
    } else {
        const float* p = reinterpret_cast<const float*>(base);
        if (aligned && e + 3 < numel) {
            g = *reinterpret_cast<const float4*>(p + e);
        } else {
            float* gp = &g.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) if (e + k < numel) gp[k] = p[e + k];
        }
    }
    return g;
}

__device__ __forceinline__ void store_shadow4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&lo);
    packed.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = packed;
}

template <bool kAdam>
__global__ void __launch_bounds__(kThreads)
mt_step_kernel(const __grid_constant__ MtTable t, float* __restrict__ w, float* __restrict__ m1,
               float* __restrict__ m2, const float* __restrict__ anchor, const float* __restrict__ cv,
               __nv_bfloat16* __restrict__ shadow, const float* __restrict__ hp, int decoupled) {
    const float gscale = hp[HP_GSCALE], lr = hp[HP_LR], wd = hp[HP_WD], mu = hp[HP_MU];
    // SGD
    const float mom = hp[HP_MOM], damp = hp[HP_DAMP];
    const bool nesterov = hp[HP_NESTEROV] != 0.f, first = hp[HP_FIRST] != 0.f;
    // Adam
    const float b1 = hp[HP_B1], b2 = hp[HP_B2], eps = hp[HP_EPS], step = hp[HP_STEP];
    float step_size = 0.f, inv_sqrt_bc2 = 0.f;
    if (kAdam) {
        step_size = lr / (1.f - __powf(b1, step));
        inv_sqrt_bc2 = rsqrtf(1.f - __powf(b2, step));
    }
    const int total_chunks = t.chunk_prefix[t.count];
    for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        const int ti = find_tensor(t, chunk);
        const int32_t numel = t.numel[ti];
        const int64_t base = t.offset[ti];
        const void* gptr = t.grad[ti];
        const bool is_bf16 = t.grad_bf16[ti] != 0;
        const bool aligned = (reinterpret_cast<uintptr_t>(gptr) & 15) == 0;
        const int64_t cs = (int64_t)(chunk - t.chunk_prefix[ti]) * kChunk;
#pragma unroll
        for (int it = 0; it < kChunk / (kThreads * 4); ++it) {
            const int64_t e = cs + ((int64_t)it * kThreads + threadIdx.x) * 4;
            if (e >= numel) break;
            float4 gv = load_grad4(gptr, is_bf16, e, numel, aligned);
            const int64_t a = base + e;  // arena entries are padded to 32 elements: full float4 access is in bounds
            float4 wv = *reinterpret_cast<const float4*>(w + a);
            float4 av = anchor ? *reinterpret_cast<const float4*>(anchor + a) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 cvv = cv ? *reinterpret_cast<const float4*>(cv + a) : make_float4(0.f, 0.f, 0.f, 0.f);
            float* wp = &wv.x; float* gp = &gv.x; const float* ap = &av.x; const float* cp = &cvv.x;
            if (kAdam) {
                float4 mv = *reinterpret_cast<const float4*>(m1 + a), vv = *reinterpret_cast<const float4*>(m2 + a);
                float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (e + k >= numel) continue;
                    float gi = gp[k] * gscale;
                    if (anchor && mu != 0.f) gi += mu * (wp[k] - ap[k]);
                    if (wd != 0.f) { if (decoupled) wp[k] *= (1.f - lr * wd); else gi += wd * wp[k]; }
                    mp[k] = b1 * mp[k] + (1.f - b1) * gi;
                    vp[k] = b2 * vp[k] + (1.f - b2) * gi * gi;
                    wp[k] -= step_size * mp[k] / (sqrtf(vp[k]) * inv_sqrt_bc2 + eps);
                }
                *reinterpret_cast<float4*>(m1 + a) = mv;
                *reinterpret_cast<float4*>(m2 + a) = vv;
            } else {
                float4 mv = (mom != 0.f && !first) ? *reinterpret_cast<const float4*>(m1 + a) : make_float4(0.f, 0.f, 0.f, 0.f);
                float* mp = &mv.x;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (e + k >= numel) continue;
                    float gi = gp[k] * gscale;
                    if (cv) gi += cp[k];
                    if (anchor && mu != 0.f) gi += mu * (wp[k] - ap[k]);
                    if (wd != 0.f) gi += wd * wp[k];
                    const float buf = first ? gi : mom * mp[k] + (1.f - damp) * gi;
                    mp[k] = buf;
                    wp[k] -= lr * ((mom != 0.f) ? (nesterov ? gi + mom * buf : buf) : gi);
                }
                if (mom != 0.f) *reinterpret_cast<float4*>(m1 + a) = mv;
            }
            *reinterpret_cast<float4*>(w + a) = wv;
            if (shadow) store_shadow4(shadow + a, wv);
        }
    }
}

__global__ void mt_post_kernel(float* hp, int adam) {
    if (adam) return;
    hp[HP_FIRST] = 0.f;
}
__global__ void mt_tick_kernel(float* hp) { hp[HP_STEP] += 1.f; }

}  // namespace

extern "C" {

int fl4h_mt_table_size() { return (int)sizeof(MtTable); }
int fl4h_mt_chunk() { return kChunk; }

// adam: 0 = SGD(momentum) ; 1 = Adam/AdamW.  `table` points to a host MtTable (copied into the parameter block).
int fl4h_mt_step(const void* table, int adam, float* w, float* m1, float* m2, const float* anchor, const float* cv,
                 void* shadow, float* hp, int decoupled, int tick, cudaStream_t stream) {
    const MtTable* t = reinterpret_cast<const MtTable*>(table);
    if (t->count < 1 || t->count > FL4H_MT_MAX) return (int)cudaErrorInvalidValue;
    const int total_chunks = t->chunk_prefix[t->count];
    int grid = total_chunks < 148 * 8 ? total_chunks : 148 * 8;
    if (grid < 1) grid = 1;
    __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(shadow);
    if (adam) {
        if (tick) mt_tick_kernel<<<1, 1, 0, stream>>>(hp);
        mt_step_kernel<true><<<grid, kThreads, 0, stream>>>(*t, w, m1, m2, anchor, cv, sh, hp, decoupled);
    } else {
        mt_step_kernel<false><<<grid, kThreads, 0, stream>>>(*t, w, m1, m2, anchor, cv, sh, hp, decoupled);
    }
    return (int)cudaGetLastError();
}

int fl4h_mt_clear_first(float* hp, cudaStream_t stream) {
    mt_post_kernel<<<1, 1, 0, stream>>>(hp, 0);
    return (int)cudaGetLastError();
}

}  // extern "C"
