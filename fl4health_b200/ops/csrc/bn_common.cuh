// Shared helpers of the BatchNorm kernels: 8-channel (16-byte for bf16) vector loads / stores in fp32 registers.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fl4h_bn {

template <typename T> struct Vec8;

template <> struct Vec8<__nv_bfloat16> {
    using Raw = uint4;                                   // 8 packed bf16: what a thread keeps in registers between phases
    static constexpr int kRawRegs = 4;
    static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
    static __device__ __forceinline__ void store_raw(__nv_bfloat16* p, const Raw& r) { *reinterpret_cast<uint4*>(p) = r; }
    static __device__ __forceinline__ void unpack(const Raw& raw, float (&v)[8]) {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
            v[2 * k] = f.x;
            v[2 * k + 1] = f.y;
        }
    }
    static __device__ __forceinline__ Raw pack(const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 raw = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
            v[2 * k] = f.x;
            v[2 * k + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

template <> struct Vec8<float> {
    struct Raw { float4 a, b; };
    static constexpr int kRawRegs = 8;
    static __device__ __forceinline__ Raw load_raw(const float* p) {
        Raw r;
        r.a = *reinterpret_cast<const float4*>(p);
        r.b = *reinterpret_cast<const float4*>(p + 4);
        return r;
    }
    static __device__ __forceinline__ void store_raw(float* p, const Raw& r) {
        *reinterpret_cast<float4*>(p) = r.a;
        *reinterpret_cast<float4*>(p + 4) = r.b;
    }
    static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
    }
    static __device__ __forceinline__ Raw pack(const float (&v)[8]) {
        Raw r;
        r.a = make_float4(v[0], v[1], v[2], v[3]);
        r.b = make_float4(v[4], v[5], v[6], v[7]);
        return r;
    }
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};

__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) { Vec8<float>::load(p, v); }


}  // namespace fl4h_bn
