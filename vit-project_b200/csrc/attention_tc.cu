// Fused softmax attention on the tcgen05 tensor pipe (bf16 mode), head_dim 64, T <= 272.
// Reference semantics: F.multi_head_attention_forward -> SDPA (torch/nn/functional.py:6682).
//
// One CTA per (sequence, head).  Q/K/V tiles are fetched by TMA straight out of the packed
// [B*T, 3*H*64] bf16 QKV activation (SWIZZLE_128B boxes of 64 columns), S = Q K^T accumulates in
// TMEM (128 lanes x NK columns), the four softmax warps (one thread per query row) read S with
// tcgen05.ld, write P as bf16 into shared memory in the K-major SWIZZLE_128B layout, and
// O = P V runs as a second UMMA whose B operand is the V tile as loaded (MN-major descriptor, no
// transpose).  T = 257 = 2*128 + 1: the two 128-row tiles go through the tensor cores, the
// left-over rows (T mod 128 <= 8) are computed by two CUDA-core warps from the same shared-memory
// K/V tiles while the tensor pipe works.
//
//   warp 0      TMA producer (one lane)        warps 2..5  softmax + O epilogue (thread = row)
//   warp 1      UMMA issuer (one lane), TMEM    warps 6..7  left-over rows (CUDA cores)
#include "common.cuh"

namespace hba {

constexpr int kTcHd = 64;
constexpr int kTcMaxKeys = 272;
constexpr int kTcRowBytes = 128;               // 64 bf16
constexpr int kTcQBytes = 128 * kTcRowBytes;   // 16 KB per Q tile
constexpr int kTcKVBytes = kTcMaxKeys * kTcRowBytes;        // 34 KB
constexpr int kTcPChunks = (kTcMaxKeys + 63) / 64;          // 5
constexpr int kTcPBytes = kTcPChunks * 128 * kTcRowBytes;   // 80 KB
constexpr int kTcThreads = 256;
constexpr int kTcSmemBytes = 2 * kTcQBytes + 2 * kTcKVBytes + kTcPBytes + 1024 + 4096;
constexpr int kTcTmemCols = 512;
constexpr int kTcOCol = 320;  // O accumulator columns [320, 384)

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// instruction descriptor: bf16 A (K-major) x bf16 B (K-major or MN-major), fp32 accumulate
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of element (row, col) inside a [rows x 64] bf16 SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int row, int col) {
  return (uint32_t)row * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4) +
         (((uint32_t)col & 7u) << 1);
}

struct AttnTcArgs {
  int T, H, causal;
  int n_tiles;     // 128-row query tiles on the tensor cores
  int n_left;      // left-over query rows (CUDA cores), rows [128*n_tiles, T)
  int NK;          // keys padded to a multiple of 16
  __nv_bfloat16* out;
  int64_t ld_out;
  float* out_f32;
  int64_t ld_of;
  float scale_log2;  // log2(e) / sqrt(64)
  const __nv_bfloat16* qkv;
  int64_t ld_qkv;
};

__global__ void __launch_bounds__(kTcThreads, 1)
    attention_tc_kernel(const __grid_constant__ CUtensorMap tma128,
                        const __grid_constant__ CUtensorMap tma16, const AttnTcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                       // 2 tiles
  uint8_t* sK = sQ + 2 * kTcQBytes;
  uint8_t* sV = sK + kTcKVBytes;            // 34816 = 34 * 1024: stays 1024-aligned
  uint8_t* sP = sV + kTcKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kTcPBytes);
  uint64_t* bar_kv = bars;        // K, V and Q tile 0 landed
  uint64_t* bar_q1 = bars + 1;    // Q tile 1 landed
  uint64_t* bar_s = bars + 2;     // S = Q K^T complete           (per tile, phase = tile & 1)
  uint64_t* bar_p = bars + 3;     // P written by the 128 softmax threads
  uint64_t* bar_o = bars + 4;     // O = P V complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* sPleft = reinterpret_cast<float*>(bars + 16);  // [2 warps][272] probabilities

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / g.H, h = blockIdx.x % g.H;
  const int d = g.H * kTcHd;
  const int row0 = b * g.T;  // first row of this sequence in the [B*T, 3d] activation

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma128);
    tma_prefetch_desc(&tma16);
    mbar_init(bar_kv, 1);
    mbar_init(bar_q1, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTcTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int full = g.NK / 128, tail = (g.NK % 128) / 16;
      const uint32_t kv_bytes = (uint32_t)g.NK * kTcRowBytes;
      mbar_arrive_expect_tx(bar_kv, 2 * kv_bytes + (g.n_tiles > 0 ? kTcQBytes : 0));
      if (g.n_tiles > 0) tma_load_2d(sQ, &tma128, bar_kv, h * kTcHd, row0);
      for (int i = 0; i < full; ++i) {
        tma_load_2d(sK + i * kTcQBytes, &tma128, bar_kv, d + h * kTcHd, row0 + i * 128);
        tma_load_2d(sV + i * kTcQBytes, &tma128, bar_kv, 2 * d + h * kTcHd, row0 + i * 128);
      }
      for (int i = 0; i < tail; ++i) {
        const int r = full * 128 + i * 16;
        tma_load_2d(sK + r * kTcRowBytes, &tma16, bar_kv, d + h * kTcHd, row0 + r);
        tma_load_2d(sV + r * kTcRowBytes, &tma16, bar_kv, 2 * d + h * kTcHd, row0 + r);
      }
      if (g.n_tiles > 1) {
        mbar_arrive_expect_tx(bar_q1, kTcQBytes);
        tma_load_2d(sQ + kTcQBytes, &tma128, bar_q1, h * kTcHd, row0 + 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && g.n_tiles > 0) {
      const int n1 = g.NK > 256 ? 256 : g.NK, n2 = g.NK - n1;
      const uint32_t idesc_s1 = idesc_bf16(128, n1, false);
      const uint32_t idesc_s2 = idesc_bf16(128, n2 > 0 ? n2 : 16, false);
      const uint32_t idesc_o = idesc_bf16(128, kTcHd, true);
      const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK));
      const uint64_t k_desc2 = make_smem_desc_sw128(smem_u32(sK + 256 * kTcRowBytes));
      const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV));
      const uint32_t tS = tmem_base, tO = tmem_base + kTcOCol;
      auto issue_s = [&](int tile) {
        const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ + (tile & 1) * kTcQBytes));
#pragma unroll
        for (int k = 0; k < kTcHd / 16; ++k) {
          umma_bf16(tS, q_desc + 2 * k, k_desc + 2 * k, idesc_s1, k > 0 ? 1u : 0u);
          if (n2 > 0) umma_bf16(tS + 256, q_desc + 2 * k, k_desc2 + 2 * k, idesc_s2, k > 0 ? 1u : 0u);
        }
        umma_commit(bar_s);
      };
      mbar_wait(bar_kv, 0);
      tc_fence_after();
      issue_s(0);
      for (int tile = 0; tile < g.n_tiles; ++tile) {
        mbar_wait(bar_p, tile & 1);  // P(tile) in smem, S(tile) fully consumed
        tc_fence_after();
        for (int kk = 0; kk < g.NK / 16; ++kk) {
          const uint64_t p_desc =
              make_smem_desc_sw128(smem_u32(sP + (kk >> 2) * kTcQBytes)) + 2 * (kk & 3);
          umma_bf16(tO, p_desc, v_desc + (uint64_t)(kk * 16 * kTcRowBytes >> 4), idesc_o,
                    kk > 0 ? 1u : 0u);
        }
        umma_commit(bar_o);
        if (tile + 1 < g.n_tiles) {
          if (tile + 1 == 1) {
            mbar_wait(bar_q1, 0);
            tc_fence_after();
          }
          issue_s(tile + 1);
        }
      }
    }
  } else if (warp < 6) {
    // ---- softmax + epilogue: thread = query row (TMEM lane) ----
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int tile = 0; tile < g.n_tiles; ++tile) {
      const int qi = tile * 128 + r;  // query index inside the sequence
      const int last_key = g.causal ? min(qi, g.T - 1) : g.T - 1;
      mbar_wait(bar_s, tile & 1);
      tc_fence_after();
      float mx = -INFINITY;
      for (int c = 0; c < g.NK; c += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(lane_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c + j <= last_key) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      const float mxs = mx * g.scale_log2;
      float sum = 0.f;
      for (int c = 0; c < g.NK; c += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(lane_addr + c, v);
        tmem_ld_wait();
        float p[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          p[j] = (c + j <= last_key) ? exp2f(__uint_as_float(v[j]) * g.scale_log2 - mxs) : 0.f;
          sum += p[j];
        }
        uint8_t* chunk = sP + (c >> 6) * kTcQBytes;
        const int cc = c & 63;
        *reinterpret_cast<uint4*>(chunk + sw128_off(r, cc)) =
            make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]),
                       pack_bf16x2(p[6], p[7]));
        *reinterpret_cast<uint4*>(chunk + sw128_off(r, cc + 8)) =
            make_uint4(pack_bf16x2(p[8], p[9]), pack_bf16x2(p[10], p[11]),
                       pack_bf16x2(p[12], p[13]), pack_bf16x2(p[14], p[15]));
      }
      tc_fence_before();
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core
      mbar_arrive(bar_p);
      // ---- O epilogue ----
      mbar_wait(bar_o, tile & 1);
      tc_fence_after();
      const float inv = 1.0f / sum;
      uint32_t o[64];
      tmem_ld_32x32b_x32(lane_addr + kTcOCol, o);
      tmem_ld_32x32b_x32(lane_addr + kTcOCol + 32, o + 32);
      tmem_ld_wait();
      if (qi < g.T) {
        const int64_t grow = (int64_t)row0 + qi;
        if (g.out) {
          __nv_bfloat16* dst = g.out + grow * g.ld_out + h * kTcHd;
#pragma unroll
          for (int j = 0; j < 64; j += 8)
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(
                pack_bf16x2(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv),
                pack_bf16x2(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv),
                pack_bf16x2(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv),
                pack_bf16x2(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv));
        }
        if (g.out_f32) {
          float* dst = g.out_f32 + grow * g.ld_of + h * kTcHd;
#pragma unroll
          for (int j = 0; j < 64; j += 4)
            *reinterpret_cast<float4*>(dst + j) =
                make_float4(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv,
                            __uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
        }
      }
      tc_fence_before();
    }
  } else {
    // ---- left-over query rows on CUDA cores, K/V read from the swizzled smem tiles ----
    const int w = warp - 6;
    if (g.n_left > 0) {
      mbar_wait(bar_kv, 0);
      float* pbuf = sPleft + w * kTcMaxKeys;
      float* qbuf = sPleft + 2 * kTcMaxKeys + w * kTcHd;
      for (int lr = w; lr < g.n_left; lr += 2) {
        const int qi = g.n_tiles * 128 + lr;
        const int last_key = g.causal ? min(qi, g.T - 1) : g.T - 1;
        const __nv_bfloat16* qrow = g.qkv + ((int64_t)row0 + qi) * g.ld_qkv + h * kTcHd;
        qbuf[lane] = __bfloat162float(qrow[lane]);
        qbuf[lane + 32] = __bfloat162float(qrow[lane + 32]);
        __syncwarp();
        float s[(kTcMaxKeys + 31) / 32];
        float mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < (kTcMaxKeys + 31) / 32; ++jj) {
          const int j = lane + 32 * jj;
          s[jj] = -INFINITY;
          if (j <= last_key) {
            float dot = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint4 kv = *reinterpret_cast<const uint4*>(sK + j * kTcRowBytes + ((u ^ (j & 7)) << 4));
              const __nv_bfloat162* kp = reinterpret_cast<const __nv_bfloat162*>(&kv);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 kf = __bfloat1622float2(kp[e]);
                dot += qbuf[u * 8 + 2 * e] * kf.x + qbuf[u * 8 + 2 * e + 1] * kf.y;
              }
            }
            s[jj] = dot;
            mx = fmaxf(mx, dot);
          }
        }
        mx = warp_max(mx);
        const float mxs = mx * g.scale_log2;
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < (kTcMaxKeys + 31) / 32; ++jj) {
          const int j = lane + 32 * jj;
          const float p = (j <= last_key) ? exp2f(s[jj] * g.scale_log2 - mxs) : 0.f;
          sum += p;
          if (j < kTcMaxKeys) pbuf[j] = p;
        }
        sum = warp_sum(sum);
        __syncwarp();
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j <= last_key; ++j) {
          const float p = pbuf[j];
          const __nv_bfloat162 vv =
              *reinterpret_cast<const __nv_bfloat162*>(sV + sw128_off(j, 2 * lane));
          const float2 vf = __bfloat1622float2(vv);
          a0 += p * vf.x;
          a1 += p * vf.y;
        }
        const float inv = 1.0f / sum;
        const int64_t grow = (int64_t)row0 + qi;
        if (g.out)
          *reinterpret_cast<__nv_bfloat162*>(g.out + grow * g.ld_out + h * kTcHd + 2 * lane) =
              __floats2bfloat162_rn(a0 * inv, a1 * inv);
        if (g.out_f32)
          *reinterpret_cast<float2*>(g.out_f32 + grow * g.ld_of + h * kTcHd + 2 * lane) =
              make_float2(a0 * inv, a1 * inv);
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTcTmemCols);
}

// host launcher, called from hba_attention_fwd (attention.cu) for bf16 activations
int attention_tc_launch(const __nv_bfloat16* qkv, int64_t ld_qkv, int B, int T, int H, int causal,
                        __nv_bfloat16* out, int64_t ld_out, float* out_f32, int64_t ld_of,
                        cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaFuncSetAttribute(attention_tc_kernel): %s", cudaGetErrorString(e));
      return HBA_ERR_CUDA;
    }
    attr_set = true;
  }
  AttnTcArgs g;
  g.T = T, g.H = H, g.causal = causal;
  const int rem = T % 128;
  if (T > 128 && rem > 0 && rem <= 8) {
    g.n_tiles = T / 128, g.n_left = rem;
  } else {
    g.n_tiles = (T + 127) / 128, g.n_left = 0;
  }
  if (g.n_tiles > 2) {
    set_error("attention_tc: T=%d needs more than two query tiles", T);
    return HBA_ERR_ARG;
  }
  g.NK = (T + 15) / 16 * 16;
  g.out = out, g.ld_out = ld_out, g.out_f32 = out_f32, g.ld_of = ld_of;
  g.scale_log2 = 1.4426950408889634f * 0.125f;
  g.qkv = qkv, g.ld_qkv = ld_qkv;
  CUtensorMap t128, t16;
  const uint64_t rows = (uint64_t)B * T, cols = (uint64_t)3 * H * kTcHd;
  HBA_CHECK(make_tma_2d_bf16(&t128, qkv, rows, cols, ld_qkv, 128, 64));
  HBA_CHECK(make_tma_2d_bf16(&t16, qkv, rows, cols, ld_qkv, 16, 64));
  attention_tc_kernel<<<B * H, kTcThreads, kTcSmemBytes, stream>>>(t128, t16, g);
  return check_launch("attention_tc_kernel");
}

}  // namespace hba
