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
// One CTA = 2 image rows x 32 pixels (64 pixels, 4 threads each).  Thread `g` of a pixel owns channels 16 q + 4 g + e
// (q, e in 0..3): the four threads of a pixel then read 64 contiguous bytes of a filter row (conflict-free LDS.128) and
// write 64 contiguous bytes of the output row per store instruction (full sectors).  The 4 x 34 input halo tile is staged
// in shared memory once (every input value is used by 4 threads x up to 9 taps).
constexpr int kFwdRows = 2, kFwdCols = 32;

template <typename T, int CIN>
__global__ void __launch_bounds__(256)
stem_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y, float* stats, int N, int H, int W) {
    __shared__ __align__(16) float ws[kTaps * CIN][kCout];     // [k][co]
    __shared__ float halo[(kFwdRows + 2) * (kFwdCols + 2) * CIN];
    __shared__ float red[2][kCout];
    const int tiles_w = W / kFwdCols, tiles_h = H / kFwdRows;
    const int tw = blockIdx.x % tiles_w, th = (blockIdx.x / tiles_w) % tiles_h, n = blockIdx.x / (tiles_w * tiles_h);
    const int w0 = tw * kFwdCols, h0 = th * kFwdRows;
    for (int i = threadIdx.x; i < kCout * kTaps * CIN; i += blockDim.x) {   // co fastest: conflict-free smem stores
        const int k = i / kCout, co = i - k * kCout;
        ws[k][co] = to_f(w[co * (kTaps * CIN) + k]);
    }
    constexpr int pitch = (kFwdCols + 2) * CIN;
    for (int i = threadIdx.x; i < (kFwdRows + 2) * pitch; i += blockDim.x) {
        const int hh = h0 - 1 + i / pitch, rem = i % pitch, ww = w0 - 1 + rem / CIN, c = rem % CIN;
        halo[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? to_f(x[(((long long)n * H + hh) * W + ww) * CIN + c]) : 0.f;
    }
    if (threadIdx.x < 2 * kCout) red[threadIdx.x / kCout][threadIdx.x % kCout] = 0.f;
    __syncthreads();
    const int group = threadIdx.x & 3, pix = threadIdx.x >> 2;    // pixel of the tile: row pix / 32, column pix % 32
    const int pr = pix / kFwdCols, pc = pix % kFwdCols;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                const float xv = halo[(pr + r) * pitch + (pc + s) * CIN + c];
                const float* wrow = &ws[(r * 3 + s) * CIN + c][4 * group];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = *reinterpret_cast<const float4*>(wrow + 16 * q);
                    acc[4 * q] = fmaf(xv, wv.x, acc[4 * q]);
                    acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
                }
            }
        }
    }
    T* out = y + ((((long long)n * H + h0 + pr) * W + w0 + pc) * kCout) + 4 * group;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(out + 16 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        } else {
            __nv_bfloat162 lo = __floats2bfloat162_rn(acc[4 * q], acc[4 * q + 1]), hi = __floats2bfloat162_rn(acc[4 * q + 2], acc[4 * q + 3]);
            *reinterpret_cast<uint2*>(out + 16 * q) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            acc[4 * q] = __bfloat162float(lo.x); acc[4 * q + 1] = __bfloat162float(lo.y);   // statistics of the STORED values
            acc[4 * q + 2] = __bfloat162float(hi.x); acc[4 * q + 3] = __bfloat162float(hi.y);
        }
    }
    if (stats != nullptr) {
        // lanes l, l^4, l^8, l^16 hold the same channel set for 8 different pixels: butterfly over the pixels
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float s = acc[j], q = s * s;
#pragma unroll
            for (int m = 4; m < 32; m <<= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, m);
                q += __shfl_xor_sync(0xffffffffu, q, m);
            }
            if ((threadIdx.x & 31) < 4) {
                const int ch = 16 * (j >> 2) + 4 * group + (j & 3);
                atomicAdd(&red[0][ch], s);
                atomicAdd(&red[1][ch], q);
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * kCout) atomicAdd(stats + threadIdx.x, red[threadIdx.x / kCout][threadIdx.x % kCout]);
    }
}

// dW[co][r][s][ci] += sum over this CTA's rows of pixels.  One CTA = `rows` image rows of one image; threads =
// 64 output channels x 4 tap groups; the input halo strip ((rows + 2) x (W + 2) x CIN, zero padded) sits in smem.
constexpr int kWgThreads = 256, kWgRows = 8, kWgMaxW = 64;

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
    int off[kPer];                                     // halo offset of this thread's taps relative to the pixel
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        acc[j] = 0.f;
        const int k = min(part * kPer + j, kK - 1);    // k = (tap_r * 3 + tap_s) * CIN + c
        const int tap = k / CIN, c = k - tap * CIN;
        off[j] = (tap / 3) * pitch + (tap % 3) * CIN + c;
    }
    const T* dyp = dy + (((long long)n * H + h0) * W) * kCout + co;
    for (int r = 0; r < kWgRows; ++r) {
        const float* hrow = halo + r * pitch;
#pragma unroll 4
        for (int wq = 0; wq < W; ++wq) {
            const float g = to_f(dyp[((long long)r * W + wq) * kCout]);   // coalesced over co
#pragma unroll
            for (int j = 0; j < kPer; ++j) acc[j] = fmaf(g, hrow[wq * CIN + off[j]], acc[j]);   // warp-wide broadcast
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
    if (W % kFwdCols != 0 || H % kFwdRows != 0) return (int)cudaErrorInvalidValue;
    const int grid = N * (H / kFwdRows) * (W / kFwdCols);
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
