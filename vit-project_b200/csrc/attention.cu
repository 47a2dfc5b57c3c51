// Softmax attention of the CLIP / ViT blocks (head_dim 64), reference semantics of
// F.multi_head_attention_forward -> scaled_dot_product_attention (torch/nn/functional.py:6682):
// softmax(q k^T / 8 [+ causal mask]) v, all statistics in fp32.
//
// This file holds the exact-fp32 CUDA-core path (used by the fp32 parity mode and for the CLS-row
// pruned forward/backward of the last vision block); the bf16 tensor-core path is attention_tc.cu.
#include <cstdlib>

#include "common.cuh"

namespace hba {

constexpr int kHd = 64;
constexpr int kAttnWarps = 8;
constexpr int kQPerWarp = 4;
constexpr int kMaxKeyChunks = 9;  // keys per lane: T <= 288
constexpr int kKStride = kHd + 1;

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

// grid = (B*H); the CTA stages K and V of one (sequence, head) in shared memory (fp32) and its
// 8 warps sweep the query rows 4 at a time.
template <typename T>
__global__ void __launch_bounds__(kAttnWarps * 32, 1)
    attention_fwd_kernel(const T* __restrict__ qkv, int64_t ld_qkv, int Tn, int H, int causal,
                         __nv_bfloat16* __restrict__ out, int64_t ld_out, int64_t lo_off,
                         float* __restrict__ out_f32, int64_t ld_of) {
  extern __shared__ __align__(16) float smem_f[];
  const int Tp = (Tn + 3) & ~3;
  float* sK = smem_f;                                    // [Tn][65], padded to 16 bytes
  float* sV = sK + (((size_t)Tn * kKStride + 3) & ~(size_t)3);  // [Tn][64]
  float* sQ = sV + (size_t)Tn * kHd;      // [warps][4][64]
  float* sP = sQ + kAttnWarps * kQPerWarp * kHd;  // [warps][4][Tp]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int d = H * kHd;
  const T* base = qkv + (int64_t)b * Tn * ld_qkv + h * kHd;
  for (int i = threadIdx.x; i < Tn * kHd; i += blockDim.x) {
    const int j = i >> 6, c = i & 63;
    sK[j * kKStride + c] = ld_as_float(base + (int64_t)j * ld_qkv + d + c);
    sV[j * kHd + c] = ld_as_float(base + (int64_t)j * ld_qkv + 2 * d + c);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q_s = sQ + warp * kQPerWarp * kHd;
  float* p_s = sP + (size_t)warp * kQPerWarp * Tp;
  const int nchunks = (Tn + 31) >> 5;
  for (int q0 = warp * kQPerWarp; q0 < Tn; q0 += kAttnWarps * kQPerWarp) {
    // stage 4 query rows (scaled by 1/sqrt(64)) for broadcast reads
    for (int i = lane; i < kQPerWarp * kHd; i += 32) {
      const int qi = q0 + (i >> 6);
      q_s[i] = (qi < Tn) ? 0.125f * ld_as_float(base + (int64_t)qi * ld_qkv + (i & 63)) : 0.f;
    }
    __syncwarp();
    float s[kQPerWarp][kMaxKeyChunks];
#pragma unroll
    for (int a = 0; a < kQPerWarp; ++a)
#pragma unroll
      for (int jj = 0; jj < kMaxKeyChunks; ++jj) s[a][jj] = 0.f;
#pragma unroll 2
    for (int c = 0; c < kHd; c += 4) {
      float4 qv[kQPerWarp];
#pragma unroll
      for (int a = 0; a < kQPerWarp; ++a) qv[a] = *reinterpret_cast<const float4*>(q_s + a * kHd + c);
#pragma unroll
      for (int jj = 0; jj < kMaxKeyChunks; ++jj) {
        if (jj < nchunks) {
          const int j = min(lane + 32 * jj, Tn - 1);
          const float* kr = sK + j * kKStride + c;
          const float k0 = kr[0], k1 = kr[1], k2 = kr[2], k3 = kr[3];
#pragma unroll
          for (int a = 0; a < kQPerWarp; ++a)
            s[a][jj] += qv[a].x * k0 + qv[a].y * k1 + qv[a].z * k2 + qv[a].w * k3;
        }
      }
    }
    // softmax per query row
#pragma unroll
    for (int a = 0; a < kQPerWarp; ++a) {
      const int qi = q0 + a;
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < kMaxKeyChunks; ++jj) {
        const int j = lane + 32 * jj;
        const bool ok = (jj < nchunks) && (j < Tn) && (!causal || j <= qi);
        s[a][jj] = ok ? s[a][jj] : -INFINITY;
        mx = fmaxf(mx, s[a][jj]);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int jj = 0; jj < kMaxKeyChunks; ++jj) {
        const float e = (s[a][jj] == -INFINITY) ? 0.f : expf(s[a][jj] - mx);
        s[a][jj] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int jj = 0; jj < kMaxKeyChunks; ++jj) {
        const int j = lane + 32 * jj;
        if (jj < nchunks && j < Tp) p_s[a * Tp + j] = (j < Tn) ? s[a][jj] * inv : 0.f;
      }
    }
    __syncwarp();
    // O = P V : lane owns output dims lane and lane + 32
    float o[kQPerWarp][2];
#pragma unroll
    for (int a = 0; a < kQPerWarp; ++a) o[a][0] = o[a][1] = 0.f;
    for (int j = 0; j < Tp; j += 4) {
      float4 pv[kQPerWarp];
#pragma unroll
      for (int a = 0; a < kQPerWarp; ++a) pv[a] = *reinterpret_cast<const float4*>(p_s + a * Tp + j);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jr = min(j + u, Tn - 1);  // p is 0 beyond Tn
        const float v0 = sV[jr * kHd + lane], v1 = sV[jr * kHd + lane + 32];
#pragma unroll
        for (int a = 0; a < kQPerWarp; ++a) {
          const float p = (u == 0) ? pv[a].x : (u == 1) ? pv[a].y : (u == 2) ? pv[a].z : pv[a].w;
          o[a][0] += p * v0;
          o[a][1] += p * v1;
        }
      }
    }
#pragma unroll
    for (int a = 0; a < kQPerWarp; ++a) {
      const int qi = q0 + a;
      if (qi < Tn) {
        const int64_t row = (int64_t)b * Tn + qi;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int c = h * kHd + lane + 32 * u;
          if (out) {
            __nv_bfloat16 hi, lo;
            split_bf16(o[a][u], hi, lo);
            out[row * ld_out + c] = hi;
            if (lo_off > 0) out[row * ld_out + lo_off + c] = lo;
          }
          if (out_f32) out_f32[row * ld_of + c] = o[a][u];
        }
      }
    }
    __syncwarp();
  }
}

// CLS-row-only forward: one warp per (sequence, head); K and V are streamed from global memory.
template <typename T>
__global__ void __launch_bounds__(128)
    attention_row0_fwd_kernel(const T* __restrict__ qkv, int64_t ld_qkv, int Bn, int Tn, int H,
                              __nv_bfloat16* __restrict__ out, int64_t ld_out, int64_t lo_off,
                              float* __restrict__ out_f32, int64_t ld_of) {
  __shared__ float sP[4][kMaxKeyChunks * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x * 4 + warp;
  if (bh >= Bn * H) return;
  const int b = bh / H, h = bh % H, d = H * kHd;
  const T* base = qkv + (int64_t)b * Tn * ld_qkv + h * kHd;
  const float q_lo = 0.125f * ld_as_float(base + lane), q_hi = 0.125f * ld_as_float(base + lane + 32);
  const int nchunks = (Tn + 31) >> 5;
  float s[kMaxKeyChunks];
  float mx = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < kMaxKeyChunks; ++jj) {
    s[jj] = -INFINITY;
    if (jj < nchunks) {
      // dot products of the 32 keys of this chunk, computed cooperatively (lane owns dims), then
      // redistributed so that lane l keeps key 32*jj + l
      for (int u = 0; u < 32; ++u) {
        const int j = 32 * jj + u;
        float part = 0.f;
        if (j < Tn) {
          const T* kr = base + (int64_t)j * ld_qkv + d;
          part = q_lo * ld_as_float(kr + lane) + q_hi * ld_as_float(kr + lane + 32);
        }
        part = warp_sum(part);
        if (lane == u && j < Tn) s[jj] = part;
      }
      mx = fmaxf(mx, s[jj]);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int jj = 0; jj < kMaxKeyChunks; ++jj) {
    const float e = (s[jj] == -INFINITY) ? 0.f : expf(s[jj] - mx);
    s[jj] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int jj = 0; jj < kMaxKeyChunks; ++jj)
    if (jj < nchunks) sP[warp][32 * jj + lane] = s[jj] * inv;
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
  for (int j = 0; j < Tn; ++j) {
    const float p = sP[warp][j];
    const T* vr = base + (int64_t)j * ld_qkv + 2 * d;
    o0 += p * ld_as_float(vr + lane);
    o1 += p * ld_as_float(vr + lane + 32);
  }
  const int c = h * kHd + lane;
  if (out) {
    __nv_bfloat16 hi, lo;
    split_bf16(o0, hi, lo);
    out[(int64_t)b * ld_out + c] = hi;
    if (lo_off > 0) out[(int64_t)b * ld_out + lo_off + c] = lo;
    split_bf16(o1, hi, lo);
    out[(int64_t)b * ld_out + c + 32] = hi;
    if (lo_off > 0) out[(int64_t)b * ld_out + lo_off + c + 32] = lo;
  }
  if (out_f32) {
    out_f32[(int64_t)b * ld_of + c] = o0;
    out_f32[(int64_t)b * ld_of + c + 32] = o1;
  }
}

// CLS-row-only backward: d_out [B, H*64] -> d_qkv [B*T, 3*H*64] (fully written)
//   p = softmax(q0 K^T / 8); dV_j = p_j dO; dp_j = dO.V_j; ds_j = p_j (dp_j - sum_k p_k dp_k);
//   dq0 = sum_j ds_j K_j / 8; dK_j = ds_j q0 / 8; dq_i = 0 for i > 0.
template <typename T>
__global__ void __launch_bounds__(128)
    attention_row0_bwd_kernel(const T* __restrict__ qkv, int64_t ld_qkv, int Bn, int Tn, int H,
                              const float* __restrict__ d_out, int64_t ld_do,
                              float* __restrict__ d_qkv, int64_t ld_dqkv) {
  __shared__ float sS[4][kMaxKeyChunks * 32];  // p, then ds
  __shared__ float sD[4][kMaxKeyChunks * 32];  // dp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x * 4 + warp;
  if (bh >= Bn * H) return;
  const int b = bh / H, h = bh % H, d = H * kHd;
  const T* base = qkv + (int64_t)b * Tn * ld_qkv + h * kHd;
  float* gbase = d_qkv + (int64_t)b * Tn * ld_dqkv + h * kHd;
  const float q_lo = ld_as_float(base + lane), q_hi = ld_as_float(base + lane + 32);
  const float do_lo = d_out[(int64_t)b * ld_do + h * kHd + lane];
  const float do_hi = d_out[(int64_t)b * ld_do + h * kHd + lane + 32];
  // pass 1: scores and dp for every key (lane owns dims; warp_sum per key)
  float mx = -INFINITY;
  for (int j = 0; j < Tn; ++j) {
    const T* kr = base + (int64_t)j * ld_qkv + d;
    const T* vr = base + (int64_t)j * ld_qkv + 2 * d;
    float sc = 0.125f * (q_lo * ld_as_float(kr + lane) + q_hi * ld_as_float(kr + lane + 32));
    float dp = do_lo * ld_as_float(vr + lane) + do_hi * ld_as_float(vr + lane + 32);
    sc = warp_sum(sc);
    dp = warp_sum(dp);
    if (lane == 0) sS[warp][j] = sc, sD[warp][j] = dp;
    mx = fmaxf(mx, sc);
  }
  __syncwarp();
  float sum = 0.f;
  for (int j = lane; j < Tn; j += 32) sum += expf(sS[warp][j] - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  float dot = 0.f;
  for (int j = lane; j < Tn; j += 32) {
    const float p = expf(sS[warp][j] - mx) * inv;
    sS[warp][j] = p;
    dot += p * sD[warp][j];
  }
  dot = warp_sum(dot);
  __syncwarp();
  // pass 2: gradients
  float dq_lo = 0.f, dq_hi = 0.f;
  for (int j = 0; j < Tn; ++j) {
    const float p = sS[warp][j];
    const float ds = p * (sD[warp][j] - dot);
    const T* kr = base + (int64_t)j * ld_qkv + d;
    dq_lo += ds * ld_as_float(kr + lane);
    dq_hi += ds * ld_as_float(kr + lane + 32);
    float* gr = gbase + (int64_t)j * ld_dqkv;
    gr[d + lane] = 0.125f * ds * q_lo;
    gr[d + lane + 32] = 0.125f * ds * q_hi;
    gr[2 * d + lane] = p * do_lo;
    gr[2 * d + lane + 32] = p * do_hi;
    if (j > 0) gr[lane] = 0.f, gr[lane + 32] = 0.f;
  }
  gbase[lane] = 0.125f * dq_lo;
  gbase[lane + 32] = 0.125f * dq_hi;
}

int attention_tc_launch(const __nv_bfloat16* qkv, int64_t ld_qkv, int B, int T, int H, int causal,
                        __nv_bfloat16* out, int64_t ld_out, float* out_f32, int64_t ld_of,
                        cudaStream_t stream);

template <typename T>
static int attention_fwd_dispatch(const T* qkv, int64_t ld_qkv, int B, int Tn, int H, int causal,
                                  int first_row_only, __nv_bfloat16* out, int64_t ld_out,
                                  int64_t lo_off, float* out_f32, int64_t ld_of, cudaStream_t s) {
  if (first_row_only) {
    attention_row0_fwd_kernel<T><<<(B * H + 3) / 4, 128, 0, s>>>(qkv, ld_qkv, B, Tn, H, out, ld_out,
                                                               lo_off, out_f32, ld_of);
    return check_launch("attention_row0_fwd_kernel");
  }
  const int Tp = (Tn + 3) & ~3;
  const size_t smem = sizeof(float) * ((((size_t)Tn * kKStride + 3) & ~(size_t)3) + (size_t)Tn * kHd +
                                       kAttnWarps * kQPerWarp * kHd +
                                       (size_t)kAttnWarps * kQPerWarp * Tp);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<T>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("attention_fwd_kernel: cannot reserve %zu bytes of shared memory: %s", smem,
                cudaGetErrorString(e));
      return HBA_ERR_CUDA;
    }
    configured = smem;
  }
  attention_fwd_kernel<T><<<B * H, kAttnWarps * 32, smem, s>>>(qkv, ld_qkv, Tn, H, causal, out,
                                                              ld_out, lo_off, out_f32, ld_of);
  return check_launch("attention_fwd_kernel");
}

}  // namespace hba

using namespace hba;

extern "C" int hba_attention_fwd(const void* qkv, int32_t qkv_dtype, int64_t ld_qkv, int32_t B,
                                 int32_t T, int32_t H, int32_t causal, int32_t first_row_only,
                                 void* out, int64_t ld_out, int64_t lo_off, float* out_f32,
                                 int64_t ld_of, void* stream) {
  HBA_REQUIRE(qkv && (out || out_f32) && B > 0 && T > 0 && H > 0, "hba_attention_fwd: bad arguments");
  HBA_REQUIRE(T <= 32 * kMaxKeyChunks, "hba_attention_fwd: T=%d exceeds %d", T, 32 * kMaxKeyChunks);
  HBA_REQUIRE(!(first_row_only && causal), "hba_attention_fwd: first_row_only with causal is unsupported");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (qkv_dtype == HBA_DT_F32)
    return attention_fwd_dispatch<float>(static_cast<const float*>(qkv), ld_qkv, B, T, H, causal,
                                         first_row_only, static_cast<__nv_bfloat16*>(out), ld_out,
                                         lo_off, out_f32, ld_of, s);
  if (qkv_dtype == HBA_DT_BF16 && !first_row_only && lo_off == 0 && T <= 257 && ld_qkv % 8 == 0 &&
      (!out || ld_out % 8 == 0) && (!out_f32 || ld_of % 4 == 0) && getenv("HBA_ATTN_SIMT") == nullptr)
    return attention_tc_launch(static_cast<const __nv_bfloat16*>(qkv), ld_qkv, B, T, H, causal,
                               static_cast<__nv_bfloat16*>(out), ld_out, out_f32, ld_of, s);
  if (qkv_dtype == HBA_DT_BF16)
    return attention_fwd_dispatch<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(qkv), ld_qkv, B,
                                                 T, H, causal, first_row_only,
                                                 static_cast<__nv_bfloat16*>(out), ld_out, lo_off,
                                                 out_f32, ld_of, s);
  set_error("hba_attention_fwd: unknown dtype %d", qkv_dtype);
  return HBA_ERR_ARG;
}

extern "C" int hba_attention_bwd_row0(const void* qkv, int32_t qkv_dtype, int64_t ld_qkv, int32_t B,
                                      int32_t T, int32_t H, const float* d_out, int64_t ld_do,
                                      float* d_qkv, int64_t ld_dqkv, void* stream) {
  HBA_REQUIRE(qkv && d_out && d_qkv && B > 0 && T > 0 && H > 0, "hba_attention_bwd_row0: bad arguments");
  HBA_REQUIRE(T <= 32 * kMaxKeyChunks, "hba_attention_bwd_row0: T=%d exceeds %d", T, 32 * kMaxKeyChunks);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (qkv_dtype == HBA_DT_F32)
    attention_row0_bwd_kernel<float><<<(B * H + 3) / 4, 128, 0, s>>>(
        static_cast<const float*>(qkv), ld_qkv, B, T, H, d_out, ld_do, d_qkv, ld_dqkv);
  else if (qkv_dtype == HBA_DT_BF16)
    attention_row0_bwd_kernel<__nv_bfloat16><<<(B * H + 3) / 4, 128, 0, s>>>(
        static_cast<const __nv_bfloat16*>(qkv), ld_qkv, B, T, H, d_out, ld_do, d_qkv, ld_dqkv);
  else {
    set_error("hba_attention_bwd_row0: unknown dtype %d", qkv_dtype);
    return HBA_ERR_ARG;
  }
  return check_launch("attention_row0_bwd_kernel");
}
