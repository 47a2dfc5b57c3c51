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
