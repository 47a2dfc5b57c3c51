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

// ---- CLS-row-only kernels (last vision block, SURVEY 7 "row pruning") -------------------------
// One 256-thread CTA per (sequence, head).  Scores: one thread per key, reading its 64-element K row
// with 16-byte loads; block-wide max / sum; P.V and the dq accumulation use a (32 key-groups x 8
// eight-dim chunks) thread grid, so that 8 consecutive threads read one contiguous K / V row.
constexpr int kR0Threads = 256;
constexpr int kR0MaxKeys = 32 * kMaxKeyChunks;  // 288

// 8 consecutive elements (16-byte aligned for bf16, 32-byte for fp32) as floats
template <typename T>
__device__ __forceinline__ void load8(const T* p, float* v);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __bfloat1622float2(h[e]);
    v[2 * e] = f.x, v[2 * e + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ float dot64(const T* row, const float* sv) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float v[8];
    load8(row + 8 * c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += v[e] * sv[8 * c + e];
  }
  return acc;
}
__device__ __forceinline__ float block_max256(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kR0Threads / 32; ++w) r = fmaxf(r, red[w]);
  return r;
}
__device__ __forceinline__ float block_sum256(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kR0Threads / 32; ++w) r += red[w];
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kR0Threads)
    attention_row0_fwd_kernel(const T* __restrict__ qkv, int64_t ld_qkv, int Bn, int Tn, int H,
                              __nv_bfloat16* __restrict__ out, int64_t ld_out, int64_t lo_off,
                              float* __restrict__ out_f32, int64_t ld_of) {
  __shared__ __align__(16) float sq[kHd];
  __shared__ float sp[kR0MaxKeys];
  __shared__ float red[kR0Threads / 32];
  __shared__ float so[32][kHd + 1];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / H, h = blockIdx.x % H, d = H * kHd;
  const T* base = qkv + (int64_t)b * Tn * ld_qkv + h * kHd;
  if (tid < kHd) sq[tid] = 0.125f * ld_as_float(base + tid);
  __syncthreads();
  float s[2];
  float mx = -INFINITY;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = tid + kR0Threads * u;
    s[u] = -INFINITY;
    if (j < Tn) s[u] = dot64(base + (int64_t)j * ld_qkv + d, sq);
    mx = fmaxf(mx, s[u]);
  }
  mx = block_max256(mx, red);
  float sum = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = tid + kR0Threads * u;
    if (j < Tn) {
      s[u] = expf(s[u] - mx);
      sum += s[u];
    }
  }
  sum = block_sum256(sum, red);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = tid + kR0Threads * u;
    if (j < Tn) sp[j] = s[u] * inv;
  }
  __syncthreads();
  const int grp = tid >> 3, c = tid & 7;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int j = grp; j < Tn; j += 32) {
    float v[8];
    load8(base + (int64_t)j * ld_qkv + 2 * d + 8 * c, v);
    const float pj = sp[j];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += pj * v[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) so[grp][8 * c + e] = acc[e];
  __syncthreads();
  if (tid < kHd) {
    float o = 0.f;
#pragma unroll
    for (int gq = 0; gq < 32; ++gq) o += so[gq][tid];
    const int col = h * kHd + tid;
    if (out) {
      __nv_bfloat16 hi, lo;
      split_bf16(o, hi, lo);
      out[(int64_t)b * ld_out + col] = hi;
      if (lo_off > 0) out[(int64_t)b * ld_out + lo_off + col] = lo;
    }
    if (out_f32) out_f32[(int64_t)b * ld_of + col] = o;
  }
}

// CLS-row-only backward: d_out [B, H*64] -> d_qkv [B*T, 3*H*64] (fully written)
//   p = softmax(q0 K^T / 8); dV_j = p_j dO; dp_j = dO.V_j; ds_j = p_j (dp_j - sum_k p_k dp_k);
//   dq0 = sum_j ds_j K_j / 8; dK_j = ds_j q0 / 8; dq_i = 0 for i > 0.
template <typename T>
__global__ void __launch_bounds__(kR0Threads)
    attention_row0_bwd_kernel(const T* __restrict__ qkv, int64_t ld_qkv, int Bn, int Tn, int H,
                              const float* __restrict__ d_out, int64_t ld_do,
                              float* __restrict__ d_qkv, int64_t ld_dqkv) {
  __shared__ __align__(16) float sq[kHd];    // q0 (unscaled)
  __shared__ __align__(16) float sdo[kHd];   // dO
  __shared__ float sp[kR0MaxKeys];           // p
  __shared__ float sds[kR0MaxKeys];          // ds
  __shared__ float red[kR0Threads / 32];
  __shared__ float so[32][kHd + 1];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / H, h = blockIdx.x % H, d = H * kHd;
  const T* base = qkv + (int64_t)b * Tn * ld_qkv + h * kHd;
  float* gbase = d_qkv + (int64_t)b * Tn * ld_dqkv + h * kHd;
  if (tid < kHd) {
    sq[tid] = ld_as_float(base + tid);
    sdo[tid] = d_out[(int64_t)b * ld_do + h * kHd + tid];
  }
  __syncthreads();
  float sc[2], dp[2];
  float mx = -INFINITY;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = tid + kR0Threads * u;
    sc[u] = -INFINITY, dp[u] = 0.f;
    if (j < Tn) {
      sc[u] = 0.125f * dot64(base + (int64_t)j * ld_qkv + d, sq);
      dp[u] = dot64(base + (int64_t)j * ld_qkv + 2 * d, sdo);
    }
    mx = fmaxf(mx, sc[u]);
  }
  mx = block_max256(mx, red);
  float sum = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = tid + kR0Threads * u;
    if (j < Tn) {
      sc[u] = expf(sc[u] - mx);
      sum += sc[u];
    }
  }
  sum = block_sum256(sum, red);
  const float inv = 1.0f / sum;
  float dot = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = tid + kR0Threads * u;
    if (j < Tn) {
      sc[u] *= inv;
      dot += sc[u] * dp[u];
    }
  }
  dot = block_sum256(dot, red);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = tid + kR0Threads * u;
    if (j < Tn) {
      sp[j] = sc[u];
      sds[j] = sc[u] * (dp[u] - dot);
    }
  }
  __syncthreads();
  // dq0 = 0.125 * sum_j ds_j K_j: (32 key groups) x (8 dim chunks), reduced through shared memory
  const int grp = tid >> 3, c = tid & 7;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int j = grp; j < Tn; j += 32) {
    float v[8];
    load8(base + (int64_t)j * ld_qkv + d + 8 * c, v);
    const float dsj = sds[j];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += dsj * v[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) so[grp][8 * c + e] = acc[e];
  // dK_j = 0.125 ds_j q0, dV_j = p_j dO, dq_j = 0 (j > 0): 16 threads per key row segment (float4 each)
  const int c4 = tid & 15;
  const float4 q4 = *reinterpret_cast<const float4*>(sq + 4 * c4);
  const float4 do4 = *reinterpret_cast<const float4*>(sdo + 4 * c4);
  for (int j = tid >> 4; j < Tn; j += kR0Threads / 16) {
    const float dsj = 0.125f * sds[j], pj = sp[j];
    float* gr = gbase + (int64_t)j * ld_dqkv + 4 * c4;
    if (j > 0) *reinterpret_cast<float4*>(gr) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(gr + d) = make_float4(dsj * q4.x, dsj * q4.y, dsj * q4.z, dsj * q4.w);
    *reinterpret_cast<float4*>(gr + 2 * d) = make_float4(pj * do4.x, pj * do4.y, pj * do4.z, pj * do4.w);
  }
  __syncthreads();
  if (tid < kHd) {
    float o = 0.f;
#pragma unroll
    for (int gq = 0; gq < 32; ++gq) o += so[gq][tid];
    gbase[tid] = 0.125f * o;
  }
}

int attention_tc_launch(const __nv_bfloat16* qkv, int64_t ld_qkv, int B, int T, int H, int causal,
                        __nv_bfloat16* out, int64_t ld_out, float* out_f32, int64_t ld_of, float* lse,
                        cudaStream_t stream);
int attention_bwd_tc_launch(const __nv_bfloat16* qkv, int64_t ld_qkv, int B, int T, int H, int causal,
                            const __nv_bfloat16* o, int64_t ld_o, const __nv_bfloat16* d_out, int64_t ld_do,
                            const float* lse, __nv_bfloat16* d_qkv, int64_t ld_dqkv, cudaStream_t stream);

template <typename T>
static int attention_fwd_dispatch(const T* qkv, int64_t ld_qkv, int B, int Tn, int H, int causal,
                                  int first_row_only, __nv_bfloat16* out, int64_t ld_out,
                                  int64_t lo_off, float* out_f32, int64_t ld_of, cudaStream_t s) {
  if (first_row_only) {
    attention_row0_fwd_kernel<T><<<B * H, kR0Threads, 0, s>>>(qkv, ld_qkv, B, Tn, H, out, ld_out, lo_off,
                                                            out_f32, ld_of);
    return check_launch("attention_row0_fwd_kernel");
  }
  const int Tp = (Tn + 3) & ~3;
  const size_t smem = sizeof(float) * ((((size_t)Tn * kKStride + 3) & ~(size_t)3) + (size_t)Tn * kHd +
                                       kAttnWarps * kQPerWarp * kHd +
                                       (size_t)kAttnWarps * kQPerWarp * Tp);
  static SmemAttr attr;
  HBA_CHECK(ensure_dyn_smem(attention_fwd_kernel<T>, smem, attr, "attention_fwd_kernel"));
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
  if (first_row_only)
    HBA_REQUIRE(ld_qkv % 8 == 0 && ((uintptr_t)qkv & 15) == 0,
                "hba_attention_fwd: first_row_only needs 16-byte aligned qkv rows (ld %% 8 == 0)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (qkv_dtype == HBA_DT_F32)
    return attention_fwd_dispatch<float>(static_cast<const float*>(qkv), ld_qkv, B, T, H, causal,
                                         first_row_only, static_cast<__nv_bfloat16*>(out), ld_out,
                                         lo_off, out_f32, ld_of, s);
  if (qkv_dtype == HBA_DT_BF16 && !first_row_only && lo_off == 0 && T <= 257 && ld_qkv % 8 == 0 &&
      (!out || ld_out % 8 == 0) && (!out_f32 || ld_of % 4 == 0) && getenv("HBA_ATTN_SIMT") == nullptr)
    return attention_tc_launch(static_cast<const __nv_bfloat16*>(qkv), ld_qkv, B, T, H, causal,
                               static_cast<__nv_bfloat16*>(out), ld_out, out_f32, ld_of, nullptr, s);
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
  HBA_REQUIRE(ld_qkv % 8 == 0 && ((uintptr_t)qkv & 15) == 0 && ld_dqkv % 4 == 0 && ((uintptr_t)d_qkv & 15) == 0,
              "hba_attention_bwd_row0: qkv / d_qkv rows must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (qkv_dtype == HBA_DT_F32)
    attention_row0_bwd_kernel<float><<<B * H, kR0Threads, 0, s>>>(
        static_cast<const float*>(qkv), ld_qkv, B, T, H, d_out, ld_do, d_qkv, ld_dqkv);
  else if (qkv_dtype == HBA_DT_BF16)
    attention_row0_bwd_kernel<__nv_bfloat16><<<B * H, kR0Threads, 0, s>>>(
        static_cast<const __nv_bfloat16*>(qkv), ld_qkv, B, T, H, d_out, ld_do, d_qkv, ld_dqkv);
  else {
    set_error("hba_attention_bwd_row0: unknown dtype %d", qkv_dtype);
    return HBA_ERR_ARG;
  }
  return check_launch("attention_row0_bwd_kernel");
}

extern "C" int hba_attention_fwd_lse(const void* qkv, int64_t ld_qkv, int32_t B, int32_t T, int32_t H,
                                     int32_t causal, void* out, int64_t ld_out, float* lse, void* stream) {
  HBA_REQUIRE(qkv && out && lse && B > 0 && T > 0 && H > 0, "hba_attention_fwd_lse: bad arguments");
  HBA_REQUIRE(T <= 257 && ld_qkv % 8 == 0 && ld_out % 8 == 0, "hba_attention_fwd_lse: T=%d (max 257) / leading dimensions unsupported", T);
  return attention_tc_launch(static_cast<const __nv_bfloat16*>(qkv), ld_qkv, B, T, H, causal,
                             static_cast<__nv_bfloat16*>(out), ld_out, nullptr, 0, lse,
                             static_cast<cudaStream_t>(stream));
}

extern "C" int hba_attention_bwd_lse(const void* qkv, int64_t ld_qkv, int32_t B, int32_t T, int32_t H,
                                     int32_t causal, const void* o, int64_t ld_o, const void* d_out,
                                     int64_t ld_do, const float* lse, void* d_qkv, int64_t ld_dqkv,
                                     void* stream) {
  HBA_REQUIRE(qkv && o && d_out && lse && d_qkv && B > 0 && T > 0 && H > 0, "hba_attention_bwd_lse: bad arguments");
  HBA_REQUIRE(T <= 256, "hba_attention_bwd_lse: T=%d exceeds 256", T);
  HBA_REQUIRE(ld_qkv % 8 == 0 && ld_o % 8 == 0 && ld_do % 8 == 0 && ld_dqkv % 8 == 0,
              "hba_attention_bwd_lse: leading dimensions must be multiples of 8");
  return attention_bwd_tc_launch(static_cast<const __nv_bfloat16*>(qkv), ld_qkv, B, T, H, causal,
                                 static_cast<const __nv_bfloat16*>(o), ld_o,
                                 static_cast<const __nv_bfloat16*>(d_out), ld_do, lse,
                                 static_cast<__nv_bfloat16*>(d_qkv), ld_dqkv, static_cast<cudaStream_t>(stream));
}
