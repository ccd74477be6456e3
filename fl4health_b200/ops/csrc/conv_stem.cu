// First-layer ("stem") 3x3 convolution for images with a handful of channels (RGB: Cin = 3), NHWC, stride 1, pad 1.
//
// With K = 27 the contraction is far too thin for the tensor cores (one 128-byte TMA row would hold ten pixels of
// channels), so the stem runs on the CUDA cores: 0.1 GFLOP per step for CIFAR batches, bound by writing the 64-channel
// output.  Two kernels (the input needs no gradient):
//   * stem_fwd   : 4 threads per output pixel, 16 output channels each, filters broadcast from shared memory; the
//                  epilogue reduces the BatchNorm statistics (sum, sum of squares of the stored values) like the
//                  tensor-core kernels do;
//   * stem_wgrad : every CTA accumulates dW[64][27] over a strip of pixels in registers (thread = one output channel x
//                  seven taps, the input halo tile staged in shared memory), adds it to an fp32 scratch with atomics;
//                  the last CTA converts the scratch to the gradient tensor and re-zeroes it (self-resetting: no memset
//                  node in the captured step).
// Reference hot op: the first nn.Conv2d of the CNNs in examples/models/cnn_model.py:16 / research/cifar10/model.py:38.

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kCout = 64, kTaps = 9, kCinMax = 4;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// x: [N, H, W, CIN]   w: [64, 3, 3, CIN]   y: [N, H, W, 64]   stats: [2][64] (nullable)
template <typename T, int CIN>
__global__ void __launch_bounds__(256)
stem_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y, float* stats, int N, int H, int W) {
    __shared__ float ws[kTaps * CIN][kCout];          // [k][co]: lanes of a pixel read 16 consecutive co
    __shared__ float red[2][kCout];
    for (int i = threadIdx.x; i < kCout * kTaps * CIN; i += blockDim.x) {
        const int co = i / (kTaps * CIN), k = i - co * (kTaps * CIN);
        ws[k][co] = to_f(w[i]);
    }
    if (threadIdx.x < 2 * kCout) red[threadIdx.x / kCout][threadIdx.x % kCout] = 0.f;
    __syncthreads();
    const int group = threadIdx.x & 3;                // 16-channel group of this thread
    const long long pixels = (long long)N * H * W;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 2) + (threadIdx.x >> 2);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    const bool live = p < pixels;
    if (live) {
        const int wq = (int)(p % W), hq = (int)((p / W) % H);
        const long long n = p / ((long long)W * H);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int hh = hq + r - 1;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int ww = wq + s - 1;
                if (ww < 0 || ww >= W) continue;
                const T* px = x + ((n * H + hh) * W + ww) * CIN;
#pragma unroll
                for (int c = 0; c < CIN; ++c) {
                    const float xv = to_f(px[c]);
                    const float4* wrow = reinterpret_cast<const float4*>(&ws[(r * 3 + s) * CIN + c][group * 16]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 wv = wrow[q];
                        acc[4 * q] = fmaf(xv, wv.x, acc[4 * q]);
                        acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
                    }
                }
            }
        }
        T* out = y + p * kCout + group * 16;
        if constexpr (sizeof(T) == 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                reinterpret_cast<float4*>(out)[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        } else {
            uint32_t packed[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
                packed[q] = *reinterpret_cast<uint32_t*>(&h);
                acc[2 * q] = __bfloat162float(h.x);   // statistics of the STORED values
                acc[2 * q + 1] = __bfloat162float(h.y);
            }
            reinterpret_cast<uint4*>(out)[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            reinterpret_cast<uint4*>(out)[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
    }
    if (stats != nullptr) {
        // lanes l, l^4, l^8, l^16 hold the same channel group for 8 different pixels: butterfly over the pixels
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float s = live ? acc[j] : 0.f, q = s * s;
#pragma unroll
            for (int m = 4; m < 32; m <<= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, m);
                q += __shfl_xor_sync(0xffffffffu, q, m);
            }
            if ((threadIdx.x & 31) < 4) {
                atomicAdd(&red[0][group * 16 + j], s);
                atomicAdd(&red[1][group * 16 + j], q);
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * kCout) atomicAdd(stats + threadIdx.x, red[threadIdx.x / kCout][threadIdx.x % kCout]);
    }
}

// dW[co][r][s][ci] += sum over this CTA's rows of pixels.  One CTA = `rows` image rows of one image; threads =
// 64 output channels x 4 tap groups; the input halo strip ((rows + 2) x (W + 2) x CIN, zero padded) sits in smem.
constexpr int kWgThreads = 256, kWgRows = 4, kWgMaxW = 64;

template <typename T, int CIN>
__global__ void __launch_bounds__(kWgThreads)
stem_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dw, float* scratch, unsigned* done, int N,
                  int H, int W) {
    __shared__ float halo[(kWgRows + 2) * (kWgMaxW + 2) * CIN];
    const int strips = H / kWgRows;
    const int n = blockIdx.x / strips, h0 = (blockIdx.x % strips) * kWgRows;
    const int pitch = (W + 2) * CIN;
    for (int i = threadIdx.x; i < (kWgRows + 2) * pitch; i += blockDim.x) {
        const int hh = h0 - 1 + i / pitch, rem = i % pitch, ww = rem / CIN - 1, c = rem % CIN;
        halo[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? to_f(x[(((long long)n * H + hh) * W + ww) * CIN + c]) : 0.f;
    }
    __syncthreads();
    constexpr int kK = kTaps * CIN;                    // 27 reduction-free outputs per channel
    constexpr int kPer = (kK + 3) / 4;                 // handled by 4 threads per channel
    const int co = threadIdx.x & 63, part = threadIdx.x >> 6;
    float acc[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) acc[j] = 0.f;
    for (int r = 0; r < kWgRows; ++r) {
        for (int wq = 0; wq < W; ++wq) {
            const float g = to_f(dy[(((long long)n * H + h0 + r) * W + wq) * kCout + co]);   // coalesced over co
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int k = part * kPer + j;         // k = (tap_r * 3 + tap_s) * CIN + c
                if (k < kK) {
                    const int tap = k / CIN, c = k - tap * CIN;
                    acc[j] = fmaf(g, halo[(r + tap / 3) * pitch + (wq + tap % 3) * CIN + c], acc[j]);   // warp-wide broadcast
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int k = part * kPer + j;
        if (k < kK) atomicAdd(scratch + co * kK + k, acc[j]);
    }
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(done, 1u) == gridDim.x - 1;
        if (is_last) *done = 0;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int i = threadIdx.x; i < kCout * kK; i += blockDim.x) {
            dw[i] = from_f<T>(__ldcg(scratch + i));
            scratch[i] = 0.f;
        }
    }
}

}  // namespace

extern "C" {

// dtype: 0 = fp32, 1 = bf16.  cin in {1, 3, 4}; cout must be 64.
int fl4h_conv_stem_fwd(const void* x, const void* w, void* y, float* stats, int dtype, int N, int H, int W, int cin, int cout,
                       cudaStream_t stream) {
    if (cout != kCout || cin < 1 || cin > kCinMax || cin == 2) return (int)cudaErrorInvalidValue;
    const long long pixels = (long long)N * H * W;
    const int grid = (int)((pixels + 63) / 64);
#define STEM_FWD(T, C) stem_fwd_kernel<T, C><<<grid, 256, 0, stream>>>((const T*)x, (const T*)w, (T*)y, stats, N, H, W)
    if (dtype == 0) { if (cin == 3) STEM_FWD(float, 3); else if (cin == 1) STEM_FWD(float, 1); else STEM_FWD(float, 4); }
    else { if (cin == 3) STEM_FWD(__nv_bfloat16, 3); else if (cin == 1) STEM_FWD(__nv_bfloat16, 1); else STEM_FWD(__nv_bfloat16, 4); }
#undef STEM_FWD
    return (int)cudaGetLastError();
}

// scratch: zero-initialised fp32 [64 * 9 * cin] + done: zero-initialised uint32, both owned by the caller (self-resetting).
int fl4h_conv_stem_wgrad(const void* x, const void* dy, void* dw, float* scratch, unsigned* done, int dtype, int N, int H, int W,
                         int cin, int cout, cudaStream_t stream) {
    if (cout != kCout || cin < 1 || cin > kCinMax || cin == 2 || H % kWgRows != 0 || W > kWgMaxW) return (int)cudaErrorInvalidValue;
    const int grid = N * (H / kWgRows);
#define STEM_WG(T, C) stem_wgrad_kernel<T, C><<<grid, kWgThreads, 0, stream>>>((const T*)x, (const T*)dy, (T*)dw, scratch, done, N, H, W)
    if (dtype == 0) { if (cin == 3) STEM_WG(float, 3); else if (cin == 1) STEM_WG(float, 1); else STEM_WG(float, 4); }
    else { if (cin == 3) STEM_WG(__nv_bfloat16, 3); else if (cin == 1) STEM_WG(__nv_bfloat16, 1); else STEM_WG(__nv_bfloat16, 4); }
#undef STEM_WG
    return (int)cudaGetLastError();
}

}  // extern "C"
