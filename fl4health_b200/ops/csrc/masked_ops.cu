// FedPM masked-parameter sampling for sm_100a (SURVEY hot-op L13).
//
//   forward : p = sigmoid(score); m ~ Bernoulli(p); out = m * frozen_weight          (1 kernel, counter-based RNG)
//   backward: d score = grad_out * frozen_weight * p * [p (1 - p)]                     (straight-through estimator of
//             fl4health/utils/functions.py:10-42 — backward of the sample is "p * grad" — chained with sigmoid')
//
// The reference runs sigmoid, torch.bernoulli, a multiply and the stock op as separate kernels per layer per forward
// (fl4health/model_bases/masked_layers/masked_linear.py:56-79).  The RNG stream is (device seed counter, element
// index): the counter lives in device memory and is ticked by a 1-thread kernel after each forward so that CUDA-graph
// replays draw fresh masks.

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t idx) {
    return (float)(mix64(seed ^ mix64(idx)) >> 40) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(kThreads)
masked_fwd_kernel(const float* __restrict__ scores, const float* __restrict__ frozen, float* __restrict__ out,
                  uint8_t* __restrict__ mask_out, const uint64_t* __restrict__ seed_state, uint64_t stream_id,
                  int64_t n) {
    const uint64_t seed = seed_state[0] * 0x9E3779B97F4A7C15ull + stream_id;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float p = 1.f / (1.f + __expf(-scores[i]));
        const bool keep = uniform01(seed, (uint64_t)i) < p;
        out[i] = keep ? frozen[i] : 0.f;
        if (mask_out) mask_out[i] = keep ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kThreads)
masked_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ frozen,
                  const float* __restrict__ grad_out, float* __restrict__ grad_scores, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float p = 1.f / (1.f + __expf(-scores[i]));
        grad_scores[i] = grad_out[i] * frozen[i] * p * (p * (1.f - p));
    }
}

__global__ void tick_seed_kernel(uint64_t* seed_state) { seed_state[0] += 1; }

inline int grid_for(int64_t n) {
    int64_t blocks = (n + kThreads - 1) / kThreads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace

extern "C" {

int fl4h_masked_fwd(const float* scores, const float* frozen, float* out, uint8_t* mask_out, uint64_t* seed_state,
                    uint64_t stream_id, int64_t n, int tick, cudaStream_t stream) {
    masked_fwd_kernel<<<grid_for(n), kThreads, 0, stream>>>(scores, frozen, out, mask_out, seed_state, stream_id, n);
    if (tick) tick_seed_kernel<<<1, 1, 0, stream>>>(seed_state);
    return (int)cudaGetLastError();
}

int fl4h_masked_bwd(const float* scores, const float* frozen, const float* grad_out, float* grad_scores, int64_t n,
                    cudaStream_t stream) {
    masked_bwd_kernel<<<grid_for(n), kThreads, 0, stream>>>(scores, frozen, grad_out, grad_scores, n);
    return (int)cudaGetLastError();
}

}  // extern "C"
