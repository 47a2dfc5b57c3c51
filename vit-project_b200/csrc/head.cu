// Output heads: CLIP cosine logits (+ fused nn.MSELoss) forward / backward, and softmax
// cross-entropy forward + backward for the ViT-B/16 classifier.  Bandwidth / latency bound; warp
// shuffles + 128-bit loads.
//
// Reference: the tail of the un-vendored CLIP.forward (contract at NEW:298-300: [B,66] logits =
// exp(logit_scale) * cos(img, txt)), nn.MSELoss (BDRV:31 applied at NEW:994 / NEW:597) and
// nn.CrossEntropyLoss (VIT:291 applied at VIT:139).
#include "common.cuh"

namespace hba {

constexpr int kHeadThreads = 256;
constexpr int kHeadMaxPer = 8;  // E <= 2048
constexpr int kHeadMaxOther = 1024;

__device__ __forceinline__ float row_sumsq_warp(const float* __restrict__ row, int E, int lane) {
  float s = 0.f;
  for (int e = lane * 4; e < E; e += 128) {
    const float4 v = *reinterpret_cast<const float4*>(row + e);
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  return warp_sum(s);
}

// grid = (B, G): group g (one lock-stepped sweep condition) owns img [B, E], txt [C, E], pred [B, C].
// pred[b, c] = exp(ls) * <img_b, txt_c> / (|img_b| |txt_c|); one warp per class row.
// With `target` the same launch produces nn.MSELoss(reduction='mean') (BDRV:31, applied NEW:994 / NEW:597)
// and the bookkeeping the reference's loop does on the host around it (NEW:989-1004): every CTA leaves its
// row's squared error in `ws`, the last CTA to finish sums the B partials in a fixed order (deterministic),
// and writes loss, the non-finite flag (a NaN / Inf anywhere in pred or target makes the loss non-finite,
// so the three checks of NEW:932-935, 989-998 collapse into this one), the count of skipped batches and the
// running sum of loss * batch.
__global__ void __launch_bounds__(kHeadThreads)
    cos_mse_fwd_kernel(const float* __restrict__ img, const float* __restrict__ txt, int B, int C, int E,
                       const float* __restrict__ logit_scale, float* __restrict__ pred,
                       const float* __restrict__ target, int64_t target_gstride, float* __restrict__ loss,
                       int* __restrict__ bad_step, int* __restrict__ bad_total, double* __restrict__ total,
                       float* __restrict__ ws) {
  __shared__ float sErr[kHeadThreads / 32];
  __shared__ int sLast;
  const int b = blockIdx.x, g = blockIdx.y, G = gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* x = img + ((size_t)g * B + b) * E;
  const float* tg = txt + (size_t)g * C * E;
  float* prow = pred + ((size_t)g * B + b) * C;
  const float* trow = target ? target + (size_t)g * target_gstride + (size_t)b * C : nullptr;
  const float nx = sqrtf(row_sumsq_warp(x, E, lane));
  const float s = expf(*logit_scale);
  float err = 0.f;
  for (int c = warp; c < C; c += kHeadThreads / 32) {
    const float* y = tg + (size_t)c * E;
    float dot = 0.f, sq = 0.f;
    for (int e = lane * 4; e < E; e += 128) {
      const float4 a = *reinterpret_cast<const float4*>(x + e);
      const float4 v = *reinterpret_cast<const float4*>(y + e);
      dot += (a.x * v.x + a.y * v.y) + (a.z * v.z + a.w * v.w);
      sq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    dot = warp_sum(dot);
    sq = warp_sum(sq);
    const float p = s * (dot / nx / sqrtf(sq));
    if (lane == 0) {
      prow[c] = p;
      if (trow) {
        const float d = p - trow[c];
        err += d * d;
      }
    }
  }
  if (!target) return;
  if (lane == 0) sErr[warp] = err;
  __syncthreads();
  unsigned int* counter = reinterpret_cast<unsigned int*>(ws) + g;
  float* part = ws + G + (size_t)g * B;
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kHeadThreads / 32; ++w) t += sErr[w];
    part[b] = t;
    __threadfence();
    sLast = (atomicAdd(counter, 1u) == (unsigned int)(B - 1));
  }
  __syncthreads();
  if (sLast && warp == 0) {
    __threadfence();
    float t = 0.f;
    for (int i = lane; i < B; i += 32) t += *reinterpret_cast<volatile float*>(part + i);
    t = warp_sum(t);
    if (lane == 0) {
      const float l = t / (float)((size_t)B * C);
      loss[g] = l;
      const int bad = isfinite(l) ? 0 : 1;
      if (bad_step) bad_step[g] = bad;
      if (bad_total) bad_total[g] += bad;
      if (total && !(bad && bad_step)) total[g] += (double)l * (double)B;
      *counter = 0u;   // ready for the next launch (stream-ordered reuse of the workspace)
    }
  }
}

// single CTA, deterministic: loss = mean((pred - target)^2)  (the two-launch form behind hba_cos_head_fwd)
__global__ void __launch_bounds__(256)
    mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, int n,
               float* __restrict__ loss) {
  __shared__ float scratch[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float d = pred[i] - target[i];
    s += d * d;
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) *loss = s / n;
}

// One CTA per row of X (own side).  g_x = (d_u - u <u, d_u>) / |x| with u = x/|x| and
// d_u = s * sum_o dpred(own, o) * y_o / |y_o|.
// dpred(own, o) = dp[own * s_own + o * s_oth]  (or the fused MSE gradient when dp == nullptr)
// PER = ceil(E / kHeadThreads): elements of the row per thread (compile-time: the accumulators stay in
// registers; the 8-wide generic form spilled)
template <int PER>
__device__ __forceinline__ void cos_head_bwd_side(const float* __restrict__ X,
                                                  const float* __restrict__ Y, int n_other, int E,
                                                  float s, const float* __restrict__ dp,
                                                  const float* __restrict__ pred,
                                                  const float* __restrict__ target, float mse_coef,
                                                  int own, int s_own, int s_oth,
                                                  float* __restrict__ gX, float* sCoef,
                                                  float* scratch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < n_other; o += kHeadThreads / 32) {
    const float ny = sqrtf(row_sumsq_warp(Y + (size_t)o * E, E, lane));
    if (lane == 0) {
      const size_t idx = (size_t)own * s_own + (size_t)o * s_oth;
      const float d = dp ? dp[idx] : mse_coef * (pred[idx] - target[idx]);
      sCoef[o] = d / ny;
    }
  }
  __syncthreads();
  const float* x = X + (size_t)own * E;
  float acc[PER], xv[PER];
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int e = threadIdx.x + kHeadThreads * k;
    acc[k] = 0.f;
    xv[k] = (e < E) ? x[e] : 0.f;
    sq += xv[k] * xv[k];
  }
#pragma unroll 2
  for (int o = 0; o < n_other; ++o) {
    const float cf = sCoef[o];
    const float* y = Y + (size_t)o * E;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int e = threadIdx.x + kHeadThreads * k;
      if (e < E) acc[k] += cf * y[e];
    }
  }
  const float nx = sqrtf(block_sum(sq, scratch));
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    acc[k] *= s;
    xv[k] /= nx;
    dot += xv[k] * acc[k];
  }
  dot = block_sum(dot, scratch);
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int e = threadIdx.x + kHeadThreads * k;
    if (e < E) gX[(size_t)own * E + e] = (acc[k] - xv[k] * dot) / nx;
  }
}

// grid = (B + C, G); group g owns img [B, E], txt [C, E], d_pred / pred [B, C], target (+ g * target_gstride),
// d_loss[g] (optional upstream gradient of the fused loss; 1 when absent)
template <int PER>
__global__ void __launch_bounds__(kHeadThreads)
    cos_head_bwd_kernel(const float* __restrict__ img, const float* __restrict__ txt, int B, int C,
                        int E, const float* __restrict__ logit_scale,
                        const float* __restrict__ d_pred, const float* __restrict__ pred,
                        const float* __restrict__ target, int64_t target_gstride,
                        const float* __restrict__ d_loss, float* __restrict__ d_img,
                        float* __restrict__ d_txt) {
  __shared__ float sCoef[kHeadMaxOther];
  __shared__ float scratch[32];
  const int g = blockIdx.y;
  img += (size_t)g * B * E, txt += (size_t)g * C * E;
  if (d_pred) d_pred += (size_t)g * B * C;
  if (pred) pred += (size_t)g * B * C;
  if (target) target += (size_t)g * target_gstride;
  const float s = expf(*logit_scale);
  const float mse_coef = 2.0f / (float)(B * C) * (d_loss ? d_loss[g] : 1.0f);
  if ((int)blockIdx.x < B) {
    if (d_img)
      cos_head_bwd_side<PER>(img, txt, C, E, s, d_pred, pred, target, mse_coef, blockIdx.x, C, 1,
                             d_img + (size_t)g * B * E, sCoef, scratch);
  } else {
    if (d_txt)
      cos_head_bwd_side<PER>(txt, img, B, E, s, d_pred, pred, target, mse_coef, blockIdx.x - B, 1, C,
                             d_txt + (size_t)g * C * E, sCoef, scratch);
  }
}

// softmax cross-entropy, one warp per sample
__global__ void __launch_bounds__(256)
    softmax_ce_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                      int B, int C, float* __restrict__ loss_rows, float* __restrict__ d_logits,
                      int64_t ld_d, int* __restrict__ correct_rows) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* z = logits + (size_t)b * ld;
  float mx = -INFINITY;
  int arg = 0;
  for (int c = lane; c < C; c += 32)
    if (z[c] > mx) mx = z[c], arg = c;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) mx = om, arg = oa;
  }
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(z[c] - mx);
  sum = warp_sum(sum);
  const int y = (int)labels[b];
  if (lane == 0) {
    loss_rows[b] = logf(sum) + mx - z[y];
    if (correct_rows) correct_rows[b] = (arg == y) ? 1 : 0;
  }
  if (d_logits) {
    const float inv = 1.0f / sum, invB = 1.0f / B;
    for (int c = lane; c < C; c += 32)
      d_logits[(size_t)b * ld_d + c] = (expf(z[c] - mx) * inv - (c == y ? 1.f : 0.f)) * invB;
  }
}

__global__ void __launch_bounds__(256)
    ce_finalize_kernel(const float* __restrict__ loss_rows, const int* __restrict__ correct_rows, int B,
                       float* __restrict__ loss, int* __restrict__ correct) {
  __shared__ float scratch[32];
  __shared__ int iscratch[32];
  float s = 0.f;
  int k = 0;
  for (int i = threadIdx.x; i < B; i += 256) {
    s += loss_rows[i];
    if (correct_rows) k += correct_rows[i];
  }
  s = block_sum(s, scratch);
  k = block_sum(k, iscratch);
  if (threadIdx.x == 0) {
    *loss = s / B;
    if (correct) *correct = k;
  }
}

}  // namespace hba

using namespace hba;

extern "C" int hba_cos_mse_fwd(const float* img, const float* txt, int32_t B, int32_t C, int32_t E,
                               int32_t groups, const float* logit_scale, float* pred, const float* target,
                               int64_t target_group_stride, float* loss, int32_t* bad_step,
                               int32_t* bad_total, double* total, float* workspace, void* stream) {
  HBA_REQUIRE(img && txt && logit_scale && pred && B > 0 && C > 0 && groups > 0, "hba_cos_mse_fwd: bad arguments");
  HBA_REQUIRE(E > 0 && E % 4 == 0, "hba_cos_mse_fwd: E=%d must be a multiple of 4", E);
  HBA_REQUIRE(groups <= 65535, "hba_cos_mse_fwd: too many groups");
  if (target) HBA_REQUIRE(loss && workspace, "hba_cos_mse_fwd: target given without loss / workspace");
  else HBA_REQUIRE(!loss && !bad_step && !bad_total && !total, "hba_cos_mse_fwd: loss outputs requested without target");
  cos_mse_fwd_kernel<<<dim3(B, groups), kHeadThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      img, txt, B, C, E, logit_scale, pred, target, target_group_stride, loss, bad_step, bad_total, total,
      workspace);
  return check_launch("cos_mse_fwd_kernel");
}

extern "C" int hba_cos_head_fwd(const float* img, const float* txt, int32_t B, int32_t C, int32_t E,
                                const float* logit_scale, float* pred, const float* target,
                                float* loss, void* stream) {
  HBA_REQUIRE(img && txt && logit_scale && pred && B > 0 && C > 0, "hba_cos_head_fwd: bad arguments");
  HBA_REQUIRE(E > 0 && E % 4 == 0, "hba_cos_head_fwd: E=%d must be a multiple of 4", E);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cos_mse_fwd_kernel<<<dim3(B, 1), kHeadThreads, 0, s>>>(img, txt, B, C, E, logit_scale, pred, nullptr, 0,
                                                         nullptr, nullptr, nullptr, nullptr, nullptr);
  HBA_CHECK(check_launch("cos_mse_fwd_kernel"));
  if (loss) {
    HBA_REQUIRE(target != nullptr, "hba_cos_head_fwd: loss requested without target");
    mse_kernel<<<1, 256, 0, s>>>(pred, target, B * C, loss);
    HBA_CHECK(check_launch("mse_kernel"));
  }
  return HBA_OK;
}

static int launch_cos_head_bwd(const float* img, const float* txt, int B, int C, int E, int groups,
                               const float* logit_scale, const float* d_pred, const float* pred,
                               const float* target, int64_t target_gstride, const float* d_loss,
                               float* d_img, float* d_txt, cudaStream_t s) {
  const dim3 grid(B + C, groups);
  const int per = (E + kHeadThreads - 1) / kHeadThreads;
#define HBA_HEAD_BWD(P)                                                                                     \
  cos_head_bwd_kernel<P><<<grid, kHeadThreads, 0, s>>>(img, txt, B, C, E, logit_scale, d_pred, pred, target, \
                                                       target_gstride, d_loss, d_img, d_txt)
  if (per <= 1) HBA_HEAD_BWD(1);
  else if (per <= 2) HBA_HEAD_BWD(2);
  else if (per <= 3) HBA_HEAD_BWD(3);
  else if (per <= 4) HBA_HEAD_BWD(4);
  else HBA_HEAD_BWD(kHeadMaxPer);
#undef HBA_HEAD_BWD
  return check_launch("cos_head_bwd_kernel");
}

extern "C" int hba_cos_head_bwd(const float* img, const float* txt, int32_t B, int32_t C, int32_t E,
                                const float* logit_scale, const float* d_pred, const float* pred,
                                const float* target, float* d_img, float* d_txt, void* stream) {
  HBA_REQUIRE(img && txt && logit_scale && (d_img || d_txt) && B > 0 && C > 0, "hba_cos_head_bwd: bad arguments");
  HBA_REQUIRE(d_pred || (pred && target), "hba_cos_head_bwd: need d_pred or (pred, target)");
  HBA_REQUIRE(E % 4 == 0 && E <= kHeadThreads * kHeadMaxPer, "hba_cos_head_bwd: E=%d unsupported", E);
  HBA_REQUIRE(B <= kHeadMaxOther && C <= kHeadMaxOther, "hba_cos_head_bwd: B, C must be <= %d", kHeadMaxOther);
  return launch_cos_head_bwd(img, txt, B, C, E, 1, logit_scale, d_pred, pred, target, 0, nullptr, d_img, d_txt,
                             static_cast<cudaStream_t>(stream));
}

extern "C" int hba_cos_mse_bwd(const float* img, const float* txt, int32_t B, int32_t C, int32_t E,
                               int32_t groups, const float* logit_scale, const float* pred,
                               const float* target, int64_t target_group_stride, const float* d_loss,
                               float* d_img, float* d_txt, void* stream) {
  HBA_REQUIRE(img && txt && logit_scale && pred && target && (d_img || d_txt) && B > 0 && C > 0 && groups > 0,
              "hba_cos_mse_bwd: bad arguments");
  HBA_REQUIRE(E % 4 == 0 && E <= kHeadThreads * kHeadMaxPer, "hba_cos_mse_bwd: E=%d unsupported", E);
  HBA_REQUIRE(B <= kHeadMaxOther && C <= kHeadMaxOther && groups <= 65535,
              "hba_cos_mse_bwd: B, C must be <= %d", kHeadMaxOther);
  return launch_cos_head_bwd(img, txt, B, C, E, groups, logit_scale, nullptr, pred, target, target_group_stride,
                             d_loss, d_img, d_txt, static_cast<cudaStream_t>(stream));
}

extern "C" int hba_softmax_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* labels,
                                      int32_t B, int32_t C, float* loss, float* d_logits,
                                      int64_t ld_d, int32_t* correct_top1, float* workspace,
                                      void* stream) {
  HBA_REQUIRE(logits && labels && loss && workspace && B > 0 && C > 0, "hba_softmax_ce_fwd_bwd: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* d_rows = workspace;
  int* d_hits = reinterpret_cast<int*>(workspace + B);
  softmax_ce_kernel<<<(B + 7) / 8, 256, 0, s>>>(logits, ld, labels, B, C, d_rows, d_logits, ld_d, d_hits);
  HBA_CHECK(check_launch("softmax_ce_kernel"));
  ce_finalize_kernel<<<1, 256, 0, s>>>(d_rows, d_hits, B, loss, correct_top1);
  return check_launch("ce_finalize_kernel");
}
