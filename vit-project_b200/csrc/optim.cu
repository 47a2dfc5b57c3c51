// Multi-tensor optimisers, one launch for all parameter tensors.
//   AdamW: torch.optim.AdamW arithmetic (decoupled weight decay, bias-corrected, eps added after
//          the sqrt / bias_correction2_sqrt), reference NEW:1181 `AdamW(model.parameters(), lr)`.
//   SGD:   torch.optim.SGD with momentum + L2 weight decay (dampening 0, no nesterov), VIT:294-299.
// The optional device-side skip flag implements the reference's NaN/Inf batch guard
// (NEW:989-998: `continue` before backward/step) without a host synchronisation.
#include <cmath>

#include "common.cuh"

namespace hba {

constexpr int kOptMaxTensors = 1024;
constexpr int kOptChunk = 2048;  // elements per CTA

// finds the tensor that owns concatenated index `g` (prefix[] is the exclusive scan of sizes)
__device__ __forceinline__ int find_tensor(const int64_t* prefix, int n, int64_t g) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (prefix[mid] <= g) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
    adamw_multi_kernel(void* const* __restrict__ ptrs, const int64_t* __restrict__ sizes, int n,
                       int64_t total, float lr, float beta1, float beta2, float eps, float wd,
                       float step_size, float bc2_sqrt, const int* __restrict__ step_dev,
                       const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag != 0) return;
  __shared__ int64_t prefix[kOptMaxTensors];
  __shared__ float s_corr[2];
  if (threadIdx.x == 0) {
    int64_t acc = 0;
    for (int t = 0; t < n; ++t) prefix[t] = acc, acc += sizes[t];
    if (step_dev) {
      // step count kept on the device (CUDA-graph replays): same double-precision bias corrections
      // as the host path / torch.optim.AdamW
      const double st = (double)*step_dev;
      s_corr[0] = (float)((double)lr / (1.0 - pow((double)beta1, st)));
      s_corr[1] = (float)sqrt(1.0 - pow((double)beta2, st));
    }
  }
  __syncthreads();
  if (step_dev) step_size = s_corr[0], bc2_sqrt = s_corr[1];
  const int64_t base = (int64_t)blockIdx.x * kOptChunk;
  for (int k = threadIdx.x; k < kOptChunk; k += 256) {
    const int64_t g = base + k;
    if (g >= total) break;
    const int t = find_tensor(prefix, n, g);
    const int64_t off = g - prefix[t];
    float* p = static_cast<float*>(ptrs[4 * t + 0]) + off;
    const float grad = static_cast<const float*>(ptrs[4 * t + 1])[off];
    float* m = static_cast<float*>(ptrs[4 * t + 2]) + off;
    float* v = static_cast<float*>(ptrs[4 * t + 3]) + off;
    float pv = *p * (1.0f - lr * wd);
    const float mv = *m + (grad - *m) * (1.0f - beta1);          // exp_avg.lerp_(grad, 1 - beta1)
    const float vv = *v * beta2 + (1.0f - beta2) * grad * grad;  // mul_(beta2).addcmul_(g, g)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv += -step_size * (mv / denom);
    *p = pv, *m = mv, *v = vv;
  }
}

__global__ void __launch_bounds__(256)
    sgd_multi_kernel(void* const* __restrict__ ptrs, const int64_t* __restrict__ sizes, int n,
                     int64_t total, float lr, float momentum, float wd, int first_step,
                     const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag != 0) return;
  __shared__ int64_t prefix[kOptMaxTensors];
  if (threadIdx.x == 0) {
    int64_t acc = 0;
    for (int t = 0; t < n; ++t) prefix[t] = acc, acc += sizes[t];
  }
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kOptChunk;
  for (int k = threadIdx.x; k < kOptChunk; k += 256) {
    const int64_t g = base + k;
    if (g >= total) break;
    const int t = find_tensor(prefix, n, g);
    const int64_t off = g - prefix[t];
    float* p = static_cast<float*>(ptrs[3 * t + 0]) + off;
    float grad = static_cast<const float*>(ptrs[3 * t + 1])[off];
    float* buf = static_cast<float*>(ptrs[3 * t + 2]) + off;
    if (wd != 0.f) grad += wd * *p;
    if (momentum != 0.f) {
      const float b = first_step ? grad : (*buf * momentum + grad);
      *buf = b;
      grad = b;
    }
    *p += -lr * grad;
  }
}

// SGD step of the fully trained ViT-B/16 (VIT:294-299) over 86.6 M parameters, vectorised (every tensor's
// element count is a multiple of 4, every pointer 16-byte aligned), with the exclusive prefix of the sizes
// precomputed on the host, and with the bf16 GEMM operand of each weight matrix refreshed in the same pass
// (ptrs[4 t + 3], or null): the separate fp32 -> bf16 staging pass of 50 matrices per step disappears.
constexpr int kSgdVecPerCta = 2048;   // float4 per CTA

__global__ void __launch_bounds__(256)
    sgd_staged_kernel(void* const* __restrict__ ptrs, const int64_t* __restrict__ prefix4, int n,
                      int64_t total4, float lr, float momentum, float wd, int first_step,
                      const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag != 0) return;
  const int64_t base = (int64_t)blockIdx.x * kSgdVecPerCta;
  int t = 0;
  {  // tensor that owns the first float4 of this CTA (binary search in global memory, once per thread)
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(prefix4 + mid) <= base) lo = mid; else hi = mid - 1;
    }
    t = lo;
  }
  for (int k = threadIdx.x; k < kSgdVecPerCta; k += 256) {
    const int64_t gidx = base + k;
    if (gidx >= total4) break;
    while (t + 1 < n && __ldg(prefix4 + t + 1) <= gidx) ++t;
    const int64_t off = gidx - __ldg(prefix4 + t);
    float4* p = static_cast<float4*>(ptrs[4 * t + 0]) + off;
    const float4 g4 = static_cast<const float4*>(ptrs[4 * t + 1])[off];
    float4* buf = static_cast<float4*>(ptrs[4 * t + 2]) + off;
    float4 pv = *p, b4 = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : *buf;
    float g[4] = {g4.x, g4.y, g4.z, g4.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float grad = g[e];
      if (wd != 0.f) grad += wd * pp[e];
      if (momentum != 0.f) {
        const float b = first_step ? grad : (bb[e] * momentum + grad);
        bb[e] = b;
        grad = b;
      }
      pp[e] += -lr * grad;
    }
    *p = make_float4(pp[0], pp[1], pp[2], pp[3]);
    if (momentum != 0.f) *buf = make_float4(bb[0], bb[1], bb[2], bb[3]);
    uint2* w16 = static_cast<uint2*>(ptrs[4 * t + 3]);
    if (w16) w16[off] = make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
  }
}

}  // namespace hba

using namespace hba;

extern "C" int hba_adamw_multi(void* const* ptrs, const int64_t* sizes, int32_t n, int64_t total,
                               float lr, float beta1, float beta2, float eps, float weight_decay,
                               int64_t step, const int32_t* step_dev, const int32_t* skip_flag,
                               void* stream) {
  HBA_REQUIRE(ptrs && sizes && n > 0 && n <= kOptMaxTensors && total > 0 && (step >= 1 || step_dev),
              "hba_adamw_multi: bad arguments");
  const double hstep = step >= 1 ? (double)step : 1.0;
  const double bc1 = 1.0 - std::pow((double)beta1, hstep);
  const double bc2 = 1.0 - std::pow((double)beta2, hstep);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)std::sqrt(bc2);
  const unsigned grid = (unsigned)((total + kOptChunk - 1) / kOptChunk);
  adamw_multi_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ptrs, sizes, n, total, lr, beta1, beta2, eps, weight_decay, step_size, bc2_sqrt, step_dev, skip_flag);
  return check_launch("adamw_multi_kernel");
}

extern "C" int hba_sgd_multi(void* const* ptrs, const int64_t* sizes, int32_t n, int64_t total,
                             float lr, float momentum, float weight_decay, int32_t first_step,
                             const int32_t* skip_flag, void* stream) {
  HBA_REQUIRE(ptrs && sizes && n > 0 && n <= kOptMaxTensors && total > 0, "hba_sgd_multi: bad arguments");
  const unsigned grid = (unsigned)((total + kOptChunk - 1) / kOptChunk);
  sgd_multi_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ptrs, sizes, n, total, lr, momentum, weight_decay, first_step, skip_flag);
  return check_launch("sgd_multi_kernel");
}

extern "C" int hba_sgd_staged(void* const* ptrs, const int64_t* prefix4, int32_t n, int64_t total4,
                              float lr, float momentum, float weight_decay, int32_t first_step,
                              const int32_t* skip_flag, void* stream) {
  HBA_REQUIRE(ptrs && prefix4 && n > 0 && total4 > 0, "hba_sgd_staged: bad arguments");
  const unsigned grid = (unsigned)((total4 + kSgdVecPerCta - 1) / kSgdVecPerCta);
  sgd_staged_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ptrs, prefix4, n, total4, lr, momentum, weight_decay, first_step, skip_flag);
  return check_launch("sgd_staged_kernel");
}
