// DoRA merge and its backward (reference: DoRALayer.weight, NEW:447-463, and the autograd mirror
// of those lines).  Bandwidth-bound: D is read once from HBM, every output written once.
//
//   V = D + scale * Bm @ A          D,V [in,out]; Bm [in,r]; A [r,out]
//   n_j = ||V[:,j]||_2 + eps        (eps added AFTER the sqrt, NEW:455)
//   Wt[i,j] = V[i,j] / n_j * m[j]   (the reference returns Wt^T as `.weight`)
//
// A CTA owns 8 consecutive output columns (one 32-byte sector per D row) and all `in` rows, so the
// column norm is a CTA-local reduction and D is read exactly once (single pass, V kept in
// registers).  The rank-r update costs in*out*r FMAs (33.5 M at 1024^2 x 32): negligible.
#include "common.cuh"

namespace hba {

constexpr int kDoraThreads = 256;
constexpr int kDoraCols = 8;
constexpr int kDoraMaxRows = 8;  // in_f <= 2048
constexpr int kDoraMaxRank = 64;

// v[k][j] for rows i = tid + 256 k; returns per-thread partial column sums of squares in ss[8]
__device__ __forceinline__ void dora_compute_v(const float* __restrict__ D,
                                               const float* __restrict__ Bm, const float (*sA)[kDoraCols],
                                               int in_f, int out_f, int r, float scale, int c0,
                                               float v[kDoraMaxRows][kDoraCols], float ss[kDoraCols]) {
#pragma unroll
  for (int j = 0; j < kDoraCols; ++j) ss[j] = 0.f;
#pragma unroll
  for (int k = 0; k < kDoraMaxRows; ++k) {
    const int i = threadIdx.x + kDoraThreads * k;
    if (i < in_f) {
      float acc[kDoraCols];
#pragma unroll
      for (int j = 0; j < kDoraCols; ++j) acc[j] = 0.f;
      const float* brow = Bm + (size_t)i * r;
      for (int kk = 0; kk < r; kk += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(brow + kk));
#pragma unroll
        for (int j = 0; j < kDoraCols; ++j)
          acc[j] += b.x * sA[kk][j] + b.y * sA[kk + 1][j] + b.z * sA[kk + 2][j] + b.w * sA[kk + 3][j];
      }
      const float4 d0 = __ldg(reinterpret_cast<const float4*>(D + (size_t)i * out_f + c0));
      const float4 d1 = __ldg(reinterpret_cast<const float4*>(D + (size_t)i * out_f + c0 + 4));
      v[k][0] = d0.x + acc[0] * scale, v[k][1] = d0.y + acc[1] * scale;
      v[k][2] = d0.z + acc[2] * scale, v[k][3] = d0.w + acc[3] * scale;
      v[k][4] = d1.x + acc[4] * scale, v[k][5] = d1.y + acc[5] * scale;
      v[k][6] = d1.z + acc[6] * scale, v[k][7] = d1.w + acc[7] * scale;
#pragma unroll
      for (int j = 0; j < kDoraCols; ++j) ss[j] += v[k][j] * v[k][j];
    } else {
#pragma unroll
      for (int j = 0; j < kDoraCols; ++j) v[k][j] = 0.f;
    }
  }
}

// block-wide sums of 8 values, result broadcast to all threads through smem
__device__ __forceinline__ void block_sum8(float x[kDoraCols], float (*red)[kDoraCols]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < kDoraCols; ++j) x[j] = warp_sum(x[j]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < kDoraCols; ++j) red[warp][j] = x[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kDoraCols; ++j) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kDoraThreads / 32; ++w) t += red[w][j];
    x[j] = t;
  }
}

__global__ void __launch_bounds__(kDoraThreads)
    dora_merge_fwd_kernel(const float* __restrict__ D, const float* __restrict__ A,
                          const float* __restrict__ Bm, const float* __restrict__ m, int in_f,
                          int out_f, int r, float scale, float eps, float* __restrict__ w_t_f32,
                          __nv_bfloat16* __restrict__ w_bf16, int64_t ld_w, int64_t w_lo_off,
                          __nv_bfloat16* __restrict__ wt_bf16, int64_t ld_wt, int64_t wt_lo_off,
                          float* __restrict__ norm_out) {
  __shared__ float sA[kDoraMaxRank][kDoraCols];
  __shared__ float red[kDoraThreads / 32][kDoraCols];
  const int c0 = blockIdx.x * kDoraCols;
  for (int i = threadIdx.x; i < r * kDoraCols; i += kDoraThreads)
    sA[i / kDoraCols][i % kDoraCols] = A[(size_t)(i / kDoraCols) * out_f + c0 + (i % kDoraCols)];
  __syncthreads();
  float v[kDoraMaxRows][kDoraCols], ss[kDoraCols];
  dora_compute_v(D, Bm, sA, in_f, out_f, r, scale, c0, v, ss);
  block_sum8(ss, red);
  float nrm[kDoraCols], mj[kDoraCols];
#pragma unroll
  for (int j = 0; j < kDoraCols; ++j) {
    nrm[j] = sqrtf(ss[j]) + eps;
    mj[j] = __ldg(m + c0 + j);
  }
  if (norm_out && threadIdx.x < kDoraCols) norm_out[c0 + threadIdx.x] = nrm[threadIdx.x];
#pragma unroll
  for (int k = 0; k < kDoraMaxRows; ++k) {
    const int i = threadIdx.x + kDoraThreads * k;
    if (i < in_f) {
      float w[kDoraCols];
#pragma unroll
      for (int j = 0; j < kDoraCols; ++j) w[j] = v[k][j] / nrm[j] * mj[j];
      if (w_t_f32) {
        float* o = w_t_f32 + (size_t)i * out_f + c0;
        *reinterpret_cast<float4*>(o) = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(w[4], w[5], w[6], w[7]);
      }
      if (wt_bf16) {
        __nv_bfloat16* o = wt_bf16 + (size_t)i * ld_wt + c0;
        *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]),
                                                  pack_bf16x2(w[4], w[5]), pack_bf16x2(w[6], w[7]));
        if (wt_lo_off > 0) {
          float l[kDoraCols];
#pragma unroll
          for (int j = 0; j < kDoraCols; ++j) l[j] = w[j] - __bfloat162float(__float2bfloat16_rn(w[j]));
          *reinterpret_cast<uint4*>(o + wt_lo_off) =
              make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]),
                         pack_bf16x2(l[6], l[7]));
        }
      }
      if (w_bf16) {
#pragma unroll
        for (int j = 0; j < kDoraCols; ++j) {
          __nv_bfloat16 h, l;
          split_bf16(w[j], h, l);
          w_bf16[(size_t)(c0 + j) * ld_w + i] = h;
          if (w_lo_off > 0) w_bf16[(size_t)(c0 + j) * ld_w + w_lo_off + i] = l;
        }
      }
    }
  }
}

// backward, phase 1 (column-owning CTAs): dm, dA and dV (written to the workspace)
__global__ void __launch_bounds__(kDoraThreads)
    dora_merge_bwd_cols_kernel(const float* __restrict__ G, int64_t ld_g,
                               const float* __restrict__ D, const float* __restrict__ A,
                               const float* __restrict__ Bm, const float* __restrict__ m, int in_f,
                               int out_f, int r, float scale, float eps, float* __restrict__ dm,
                               float* __restrict__ dA, float* __restrict__ dV) {
  __shared__ float sA[kDoraMaxRank][kDoraCols];
  __shared__ float red[kDoraThreads / 32][kDoraCols];
  extern __shared__ __align__(16) float sDV[];  // [in_f][8]
  const int c0 = blockIdx.x * kDoraCols;
  for (int i = threadIdx.x; i < r * kDoraCols; i += kDoraThreads)
    sA[i / kDoraCols][i % kDoraCols] = A[(size_t)(i / kDoraCols) * out_f + c0 + (i % kDoraCols)];
  __syncthreads();
  float v[kDoraMaxRows][kDoraCols], ss[kDoraCols];
  dora_compute_v(D, Bm, sA, in_f, out_f, r, scale, c0, v, ss);
  block_sum8(ss, red);
  float g[kDoraMaxRows][kDoraCols], c[kDoraCols];
#pragma unroll
  for (int j = 0; j < kDoraCols; ++j) c[j] = 0.f;
#pragma unroll
  for (int k = 0; k < kDoraMaxRows; ++k) {
    const int i = threadIdx.x + kDoraThreads * k;
#pragma unroll
    for (int j = 0; j < kDoraCols; ++j) {
      g[k][j] = (i < in_f) ? __ldg(G + (size_t)(c0 + j) * ld_g + i) : 0.f;
      c[j] += g[k][j] * v[k][j];
    }
  }
  block_sum8(c, red);
  float coef[kDoraCols], back[kDoraCols];
#pragma unroll
  for (int j = 0; j < kDoraCols; ++j) {
    const float vn = sqrtf(ss[j]);
    const float n = vn + eps;
    const float mj = __ldg(m + c0 + j);
    coef[j] = mj / n;
    back[j] = (vn > 0.f) ? c[j] / (n * vn) : 0.f;
    if (threadIdx.x == j) dm[c0 + j] = c[j] / n;
  }
#pragma unroll
  for (int k = 0; k < kDoraMaxRows; ++k) {
    const int i = threadIdx.x + kDoraThreads * k;
    if (i < in_f) {
      float dv[kDoraCols];
#pragma unroll
      for (int j = 0; j < kDoraCols; ++j) dv[j] = coef[j] * (g[k][j] - v[k][j] * back[j]);
      *reinterpret_cast<float4*>(sDV + (size_t)i * kDoraCols) = make_float4(dv[0], dv[1], dv[2], dv[3]);
      *reinterpret_cast<float4*>(sDV + (size_t)i * kDoraCols + 4) = make_float4(dv[4], dv[5], dv[6], dv[7]);
      float* o = dV + (size_t)i * out_f + c0;
      *reinterpret_cast<float4*>(o) = make_float4(dv[0], dv[1], dv[2], dv[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(dv[4], dv[5], dv[6], dv[7]);
    }
  }
  __syncthreads();
  // dA[k, c0 + j] = scale * sum_i Bm[i,k] dV[i,j].  Lane = rank index k (two per lane when r > 32), the
  // 8 warps split the rows: every Bm row (one 128-byte line at r = 32) is read exactly once, coalesced,
  // 16 rows in flight per warp; dV[i, j] is a shared-memory broadcast.  Cross-warp sum through smem.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float accA[2][kDoraCols];
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
    for (int j = 0; j < kDoraCols; ++j) accA[h2][j] = 0.f;
  const int rows_per_warp = (in_f + 7) / 8;
  const int i_begin = warp * rows_per_warp, i_end = min(in_f, i_begin + rows_per_warp);
  for (int i0 = i_begin; i0 < i_end; i0 += 16) {
    float b0[16], b1[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int i = i0 + u;
      b0[u] = (i < i_end && lane < r) ? __ldg(Bm + (size_t)i * r + lane) : 0.f;
      b1[u] = (i < i_end && lane + 32 < r) ? __ldg(Bm + (size_t)i * r + lane + 32) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int i = min(i0 + u, in_f - 1);
      const float4 d0 = *reinterpret_cast<const float4*>(sDV + (size_t)i * kDoraCols);
      const float4 d1 = *reinterpret_cast<const float4*>(sDV + (size_t)i * kDoraCols + 4);
      const float dv[kDoraCols] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int j = 0; j < kDoraCols; ++j) {
        accA[0][j] += b0[u] * dv[j];
        accA[1][j] += b1[u] * dv[j];
      }
    }
  }
  __syncthreads();  // every warp is done reading sDV: reuse it for the per-warp partial sums
  float* part = sDV;  // [8 warps][64 k][8 j] floats = 16 KB (the launcher reserves at least that)
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
    for (int j = 0; j < kDoraCols; ++j)
      part[((size_t)warp * 64 + lane + 32 * h2) * kDoraCols + j] = accA[h2][j];
  __syncthreads();
  for (int t = threadIdx.x; t < r * kDoraCols; t += kDoraThreads) {
    const int k = t / kDoraCols, j = t % kDoraCols;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kDoraThreads / 32; ++w) acc += part[((size_t)w * 64 + k) * kDoraCols + j];
    dA[(size_t)k * out_f + c0 + j] = scale * acc;
  }
}

// backward, phase 2 (row-owning warps): dB[i,k] = scale * sum_j dV[i,j] A[k,j].
// The CTA stages 32 rank rows of A (all out_f columns, <= 2048) in shared memory once; a warp owns
// one row i, its lanes split the columns (coalesced reads of dV, 32 loads in flight), every lane keeps
// 32 partial sums (one per rank index) and the warp reduces them with shuffles at the end.
__global__ void __launch_bounds__(256)
    dora_merge_bwd_rows_kernel(const float* __restrict__ dV, const float* __restrict__ A, int in_f,
                               int out_f, int r, float scale, float* __restrict__ dB) {
  extern __shared__ __align__(16) float sAr[];  // [32][out_f + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  const int pitch = out_f + 1;
  for (int k0 = 0; k0 < r; k0 += 32) {
    __syncthreads();
    for (int t = threadIdx.x; t < 32 * (out_f / 4); t += 256) {
      const int k = t / (out_f / 4), j4 = (t % (out_f / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < r) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(k0 + k) * out_f + j4));
      float* dst = sAr + (size_t)k * pitch + j4;
      dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;
    }
    __syncthreads();
    if (i < in_f) {
      float acc[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[k] = 0.f;
      for (int c0 = 0; c0 < out_f; c0 += 256) {
        float dv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = c0 + lane + 32 * u;
          dv[u] = (j < out_f) ? __ldg(dV + (size_t)i * out_f + j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = min(c0 + lane + 32 * u, out_f - 1);
#pragma unroll
          for (int k = 0; k < 32; ++k) acc[k] += dv[u] * sAr[(size_t)k * pitch + j];
        }
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float t = warp_sum(acc[k]);
        if (lane == k && k0 + k < r) dB[(size_t)i * r + k0 + k] = scale * t;
      }
    }
  }
}

}  // namespace hba

using namespace hba;

static int dora_check(const char* who, int in_f, int out_f, int r) {
  HBA_REQUIRE(in_f > 0 && in_f <= kDoraThreads * kDoraMaxRows, "%s: in_features=%d unsupported (max %d)", who, in_f, kDoraThreads * kDoraMaxRows);
  HBA_REQUIRE(out_f > 0 && out_f % kDoraCols == 0, "%s: out_features=%d must be a multiple of %d", who, out_f, kDoraCols);
  HBA_REQUIRE(r > 0 && r <= kDoraMaxRank && r % 4 == 0, "%s: rank=%d must be a multiple of 4 and <= %d", who, r, kDoraMaxRank);
  return HBA_OK;
}

extern "C" int hba_dora_merge_fwd(const float* D, const float* A, const float* Bm, const float* m,
                                  int32_t in_f, int32_t out_f, int32_t r, float scale, float eps,
                                  float* w_t_f32, void* w_bf16, int64_t ld_w, int64_t w_lo_off,
                                  void* wt_bf16, int64_t ld_wt, int64_t wt_lo_off, float* norm_out,
                                  void* stream) {
  HBA_REQUIRE(D && A && Bm && m, "hba_dora_merge_fwd: null input");
  HBA_REQUIRE(w_t_f32 || w_bf16 || wt_bf16, "hba_dora_merge_fwd: no output requested");
  HBA_CHECK(dora_check("hba_dora_merge_fwd", in_f, out_f, r));
  HBA_REQUIRE(!wt_bf16 || (ld_wt % 8 == 0 && wt_lo_off % 8 == 0 && ((uintptr_t)wt_bf16 & 15) == 0), "hba_dora_merge_fwd: wt_bf16 alignment");
  dora_merge_fwd_kernel<<<out_f / kDoraCols, kDoraThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      D, A, Bm, m, in_f, out_f, r, scale, eps, w_t_f32, static_cast<__nv_bfloat16*>(w_bf16), ld_w,
      w_lo_off, static_cast<__nv_bfloat16*>(wt_bf16), ld_wt, wt_lo_off, norm_out);
  return check_launch("dora_merge_fwd_kernel");
}

extern "C" int hba_dora_merge_bwd(const float* G, int64_t ld_g, const float* D, const float* A,
                                  const float* Bm, const float* m, int32_t in_f, int32_t out_f,
                                  int32_t r, float scale, float eps, float* dm, float* dA, float* dB,
                                  float* workspace, void* stream) {
  HBA_REQUIRE(G && D && A && Bm && m && dm && dA && dB && workspace, "hba_dora_merge_bwd: null pointer");
  HBA_CHECK(dora_check("hba_dora_merge_bwd", in_f, out_f, r));
  HBA_REQUIRE(out_f <= 1792, "hba_dora_merge_bwd: out_features=%d exceeds the shared-memory staging (max 1792)", out_f);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // [in_f][8] dV tile, later reused for the [8 warps][64][8] partial sums of dA
  const size_t smem = (size_t)(in_f > 512 ? in_f : 512) * kDoraCols * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 - 4096 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(dora_merge_bwd_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("hba_dora_merge_bwd: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return HBA_ERR_CUDA;
    }
    configured = smem;
  }
  dora_merge_bwd_cols_kernel<<<out_f / kDoraCols, kDoraThreads, smem, s>>>(
      G, ld_g, D, A, Bm, m, in_f, out_f, r, scale, eps, dm, dA, workspace);
  HBA_CHECK(check_launch("dora_merge_bwd_cols_kernel"));
  const size_t smem_rows = (size_t)32 * (out_f + 1) * sizeof(float);
  static size_t configured_rows = 0;
  if (smem_rows > configured_rows) {
    cudaError_t e = cudaFuncSetAttribute(dora_merge_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("hba_dora_merge_bwd: cannot reserve %zu bytes of shared memory: %s", smem_rows, cudaGetErrorString(e));
      return HBA_ERR_CUDA;
    }
    configured_rows = smem_rows;
  }
  dora_merge_bwd_rows_kernel<<<(in_f + 7) / 8, 256, smem_rows, s>>>(workspace, A, in_f, out_f, r, scale, dB);
  return check_launch("dora_merge_bwd_rows_kernel");
}
