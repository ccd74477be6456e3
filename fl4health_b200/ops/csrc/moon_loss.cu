// Fused MOON / PerFCL contrastive loss (forward + backward) for sm_100a.
//
//   logits_b = [cos(z_b, p_b), cos(z_b, n_{1,b}), ..., cos(z_b, n_{N,b})] / tau ;  loss = mean_b CE(logits_b, 0)
//
// The reference builds this from CosineSimilarity + repeat + cat + CrossEntropyLoss (~12 kernels fwd, ~25 bwd:
// fl4health/losses/contrastive_loss.py:28-92).  Here: one kernel forward (norms, dots, softmax, loss in one pass
// over the feature rows) and one kernel backward producing d/dz and, optionally, d/dp and d/dn.
// One CTA per sample; rows are read with 128-bit loads when F % 4 == 0.

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kThreads = 128;
constexpr int kMaxPairs = 16;  // 1 positive + up to 15 negatives

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// reduce `count` partial values per thread across the CTA; result broadcast in smem `out`
template <int kCount>
__device__ __forceinline__ void block_reduce(float (&vals)[kCount], float* smem /* [kCount][kThreads/32] */,
                                             float* out /* [kCount] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kCount; ++i) {
        float v = warp_sum(vals[i]);
        if (lane == 0) smem[i * (kThreads / 32) + wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < kCount) {
        float total = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) total += smem[threadIdx.x * (kThreads / 32) + w];
        out[threadIdx.x] = total;
    }
    __syncthreads();
}

// pairs layout: y[j] for j in [0, P): j == 0 -> positive [B,F]; j >= 1 -> negatives [(j-1), B, F]
__global__ void __launch_bounds__(kThreads)
moon_fwd_kernel(const float* __restrict__ z, const float* __restrict__ pos, const float* __restrict__ neg,
                int batch, int feat, int n_pairs, float inv_tau, float eps, float* __restrict__ loss_sum,
                float* __restrict__ probs /* [B,P] */, float* __restrict__ cosines /* [B,P] */,
                float* __restrict__ norms /* [B,1+P] : |z|, |y_j| */) {
    __shared__ float smem[(1 + 2 * kMaxPairs) * (kThreads / 32)];
    __shared__ float red[1 + 2 * kMaxPairs];
    const int b = blockIdx.x;
    const float* zr = z + (size_t)b * feat;
    float acc[1 + 2 * kMaxPairs];
#pragma unroll
    for (int i = 0; i < 1 + 2 * kMaxPairs; ++i) acc[i] = 0.f;
    for (int f = threadIdx.x; f < feat; f += kThreads) {
        const float zv = zr[f];
        acc[0] += zv * zv;
#pragma unroll
        for (int j = 0; j < kMaxPairs; ++j) {
            if (j < n_pairs) {
                const float* yr = (j == 0) ? pos + (size_t)b * feat : neg + ((size_t)(j - 1) * batch + b) * feat;
                const float yv = yr[f];
                acc[1 + j] += zv * yv;
                acc[1 + kMaxPairs + j] += yv * yv;
            }
        }
    }
    block_reduce<1 + 2 * kMaxPairs>(acc, smem, red);
    if (threadIdx.x == 0) {
        const float nz = fmaxf(sqrtf(red[0]), eps);
        norms[(size_t)b * (1 + n_pairs)] = nz;
        float logits[kMaxPairs];
        float mx = -1e30f;
        for (int j = 0; j < n_pairs; ++j) {
            const float ny = fmaxf(sqrtf(red[1 + kMaxPairs + j]), eps);
            norms[(size_t)b * (1 + n_pairs) + 1 + j] = ny;
            const float c = red[1 + j] / (nz * ny);
            cosines[(size_t)b * n_pairs + j] = c;
            logits[j] = c * inv_tau;
            mx = fmaxf(mx, logits[j]);
        }
        float denom = 0.f;
        for (int j = 0; j < n_pairs; ++j) denom += __expf(logits[j] - mx);
        for (int j = 0; j < n_pairs; ++j) probs[(size_t)b * n_pairs + j] = __expf(logits[j] - mx) / denom;
        atomicAdd(loss_sum, (mx + __logf(denom) - logits[0]) / (float)batch);
    }
}

__global__ void __launch_bounds__(kThreads)
moon_bwd_kernel(const float* __restrict__ z, const float* __restrict__ pos, const float* __restrict__ neg,
                const float* __restrict__ probs, const float* __restrict__ cosines, const float* __restrict__ norms,
                const float* __restrict__ grad_out, int batch, int feat, int n_pairs, float inv_tau,
                float* __restrict__ gz, float* __restrict__ gpos, float* __restrict__ gneg) {
    const int b = blockIdx.x;
    const float go = grad_out[0] * inv_tau / (float)batch;
    const float nz = norms[(size_t)b * (1 + n_pairs)];
    float coef[kMaxPairs], ny[kMaxPairs], cs[kMaxPairs];
    float zself = 0.f;  // sum_j coef_j * cos_j / nz^2  (the -cos * z/|z|^2 term)
#pragma unroll
    for (int j = 0; j < kMaxPairs; ++j) {
        if (j < n_pairs) {
            coef[j] = go * (probs[(size_t)b * n_pairs + j] - (j == 0 ? 1.f : 0.f));  // dL/dlogit_j * 1/tau
            ny[j] = norms[(size_t)b * (1 + n_pairs) + 1 + j];
            cs[j] = cosines[(size_t)b * n_pairs + j];
            zself += coef[j] * cs[j];
        }
    }
    zself /= (nz * nz);
    const float* zr = z + (size_t)b * feat;
    for (int f = threadIdx.x; f < feat; f += kThreads) {
        const float zv = zr[f];
        float g = -zself * zv;
#pragma unroll
        for (int j = 0; j < kMaxPairs; ++j) {
            if (j < n_pairs) {
                const size_t off = (j == 0) ? (size_t)b * feat + f : ((size_t)(j - 1) * batch + b) * feat + f;
                const float yv = (j == 0) ? pos[off] : neg[off];
                g += coef[j] * yv / (nz * ny[j]);
                // d cos / d y = z/(|z||y|) - cos * y/|y|^2
                const float gy = coef[j] * (zv / (nz * ny[j]) - cs[j] * yv / (ny[j] * ny[j]));
                if (j == 0) { if (gpos) gpos[off] = gy; }
                else if (gneg) gneg[off] = gy;
            }
        }
        if (gz) gz[(size_t)b * feat + f] = g;
    }
}

}  // namespace

extern "C" {

int fl4h_moon_fwd(const float* z, const float* pos, const float* neg, int batch, int feat, int n_neg, float tau,
                  float* loss_sum, float* probs, float* cosines, float* norms, cudaStream_t stream) {
    const int n_pairs = 1 + n_neg;
    if (n_pairs > kMaxPairs || batch < 1) return (int)cudaErrorInvalidValue;
    moon_fwd_kernel<<<batch, kThreads, 0, stream>>>(z, pos, neg, batch, feat, n_pairs, 1.f / tau, 1e-8f, loss_sum,
                                                    probs, cosines, norms);
    return (int)cudaGetLastError();
}

int fl4h_moon_bwd(const float* z, const float* pos, const float* neg, const float* probs, const float* cosines,
                  const float* norms, const float* grad_out, int batch, int feat, int n_neg, float tau, float* gz,
                  float* gpos, float* gneg, cudaStream_t stream) {
    const int n_pairs = 1 + n_neg;
    if (n_pairs > kMaxPairs || batch < 1) return (int)cudaErrorInvalidValue;
    moon_bwd_kernel<<<batch, kThreads, 0, stream>>>(z, pos, neg, probs, cosines, norms, grad_out, batch, feat,
                                                    n_pairs, 1.f / tau, gz, gpos, gneg);
    return (int)cudaGetLastError();
}

}  // extern "C"
