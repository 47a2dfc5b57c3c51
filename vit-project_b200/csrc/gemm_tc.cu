// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] . B[N,K]^T), bf16 in, fp32 acc.
//
// One persistent CTA per SM, 6 warps, warp-specialised:
//   warp 0 (one lane)  TMA producer: A box 128x64 and B box BNx64 (SWIZZLE_128B) into a smem ring
//   warp 1 (one lane)  tcgen05.mma issuer: 128 x BN x 16 UMMAs, accumulators in TMEM, 2 accumulator
//                      stages so that the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2..5         epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> fused
//                      bias / QuickGELU / GELU / activation-gradient / residual -> global
// "fp32 mode" (nsplit == 3) runs three bf16 passes per k-block over hi/lo split operands
// (hi.hi + lo.hi + hi.lo), which reproduces fp32 products to ~2^-16 on the bf16 tensor pipe.
//
// Replaces F.linear at torch/nn/functional.py:6244 (in_proj), :6690 (out_proj) and the CLIP / timm
// MLP linears reached through the reference's CLIPHBA.forward (NEW:298) and VIT:138-140.
#include "common.cuh"

namespace hba {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;  // two warps per TMEM lane quarter, each takes half of the columns
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + 1024 /*align*/ + 256 /*bars*/;
};

struct GemmArgs {
  int M, N, K;
  int nsplit, a_lo_off, b_lo_off;
  float alpha;
  const float* bias;
  const float* residual;
  int ldr;
  int act;
  const void* aux;
  int ld_aux, aux_dtype;
  void* pre_out;
  int ld_pre, pre_dtype;
  float* out_f32;
  int ld_f32;
  __nv_bfloat16* out_bf16;
  int ld_bf16, out_lo_off;
  int transpose_out;
  int a_mn, b_mn;  // operand stored MN-major: [K rows, M (resp. N) columns]
};

// x * sigmoid(1.702 x) on the SFU fast paths (ex2.approx + rcp.approx, ~2 ulp): the epilogue of the
// c_fc GEMM evaluates this 33.7 M times per layer and must stay under the main loop's time
__device__ __forceinline__ float quickgelu(float v) {
  return __fdividef(v, 1.0f + __expf(-1.702f * v));
}
__device__ __forceinline__ float quickgelu_grad(float a) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-1.702f * a));
  return s * (1.0f + 1.702f * a * (1.0f - s));
}
__device__ __forceinline__ float gelu_erf(float v) {
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_erf_grad(float a) {
  return 0.5f * (1.0f + erff(a * 0.70710678118654752f)) +
         a * 0.39894228040143268f * expf(-0.5f * a * a);
}

// epilogue for one thread: row `row`, 32 consecutive columns starting at `col`
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& g, const uint32_t* r, int row,
                                               int col) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * g.alpha;
  const bool full = (col + 32 <= g.N);
  if (g.bias) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col + j));
        v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col + j < g.N) v[j] += __ldg(g.bias + col + j);
    }
  }
  if (g.pre_out) {
    if (g.pre_dtype == HBA_DT_F32) {
      float* p = static_cast<float*>(g.pre_out) + (size_t)row * g.ld_pre + col;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col + j < g.N) p[j] = v[j];
      }
    } else {
      __nv_bfloat16* p = static_cast<__nv_bfloat16*>(g.pre_out) + (size_t)row * g.ld_pre + col;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          *reinterpret_cast<uint4*>(p + j) =
              make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                         pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col + j < g.N) p[j] = __float2bfloat16_rn(v[j]);
      }
    }
  }
  if (g.act == HBA_ACT_QUICKGELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = quickgelu(v[j]);
  } else if (g.act == HBA_ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (g.act == HBA_ACT_QUICKGELU_GRAD || g.act == HBA_ACT_GELU_ERF_GRAD) {
    float a[32];
    if (g.aux_dtype == HBA_DT_F32) {
      const float* p = static_cast<const float*>(g.aux) + (size_t)row * g.ld_aux + col;
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = (col + j < g.N) ? __ldg(p + j) : 0.f;
    } else {
      const __nv_bfloat16* p =
          static_cast<const __nv_bfloat16*>(g.aux) + (size_t)row * g.ld_aux + col;
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = (col + j < g.N) ? __bfloat162float(p[j]) : 0.f;
    }
    if (g.act == HBA_ACT_QUICKGELU_GRAD) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= quickgelu_grad(a[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= gelu_erf_grad(a[j]);
    }
  }
  if (g.residual) {
    const float* p = g.residual + (size_t)row * g.ldr + col;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p + j));
        v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col + j < g.N) v[j] += __ldg(p + j);
    }
  }
  if (!g.transpose_out) {
    if (g.out_f32) {
      float* p = g.out_f32 + (size_t)row * g.ld_f32 + col;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col + j < g.N) p[j] = v[j];
      }
    }
    if (g.out_bf16) {
      __nv_bfloat16* p = g.out_bf16 + (size_t)row * g.ld_bf16 + col;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          *reinterpret_cast<uint4*>(p + j) =
              make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                         pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
        if (g.out_lo_off > 0) {
          float l[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) l[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            *reinterpret_cast<uint4*>(p + g.out_lo_off + j) =
                make_uint4(pack_bf16x2(l[j], l[j + 1]), pack_bf16x2(l[j + 2], l[j + 3]),
                           pack_bf16x2(l[j + 4], l[j + 5]), pack_bf16x2(l[j + 6], l[j + 7]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col + j < g.N) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
            p[j] = h;
            if (g.out_lo_off > 0)
              p[g.out_lo_off + j] = __float2bfloat16_rn(v[j] - __bfloat162float(h));
          }
      }
    }
  } else {
    // transposed store: for a fixed column the 32 lanes of the warp write 32 consecutive rows
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (col + j < g.N) {
        if (g.out_f32) g.out_f32[(size_t)(col + j) * g.ld_f32 + row] = v[j];
        if (g.out_bf16) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
          g.out_bf16[(size_t)(col + j) * g.ld_bf16 + row] = h;
          if (g.out_lo_off > 0)
            g.out_bf16[(size_t)(col + j) * g.ld_bf16 + g.out_lo_off + row] =
                __float2bfloat16_rn(v[j] - __bfloat162float(h));
        }
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a,
                   const __grid_constant__ CUtensorMap tma_b, const GemmArgs g) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + kStages * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + kStages * Cfg::kBBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m_tiles = (g.M + BM - 1) / BM;
  const int num_n_tiles = (g.N + BN - 1) / BN;
  const int total_tiles = num_m_tiles * num_n_tiles;
  const int kblocks = (g.K + BK - 1) / BK;  // a ragged last block is zero-filled by TMA
  const int kiters = kblocks * g.nsplit;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile % num_m_tiles) * BM;
        const int n0 = (tile / num_m_tiles) * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          for (int s = 0; s < g.nsplit; ++s) {
            const int a_col = kb * BK + (s == 1 ? g.a_lo_off : 0);
            const int b_col = kb * BK + (s == 2 ? g.b_lo_off : 0);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::kBBytes);
            if (!g.a_mn) {
              tma_load_2d(sA + stage * Cfg::kABytes, &tma_a, &full_bar[stage], a_col, m0);
            } else {  // [K, M] storage: BM/64 boxes of 64 k-rows x 64 m-columns, 8 KB apart (LBO)
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d(sA + stage * Cfg::kABytes + j * 8192, &tma_a, &full_bar[stage],
                            m0 + 64 * j + (s == 1 ? g.a_lo_off : 0), kb * BK);
            }
            if (!g.b_mn) {
              tma_load_2d(sB + stage * Cfg::kBBytes, &tma_b, &full_bar[stage], b_col, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(sB + stage * Cfg::kBBytes + j * 8192, &tma_b, &full_bar[stage],
                            n0 + 64 * j + (s == 2 ? g.b_lo_off : 0), kb * BK);
            }
            if (++stage == kStages) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, BN) | (g.a_mn ? (1u << 15) : 0u) |
                             (g.b_mn ? (1u << 16) : 0u);
      // K-major: +32 B per 16-element k-step inside the 128-byte swizzle row (start address += 2);
      // MN-major: 16 k-rows of 128 B further down (start address += 128)
      const uint32_t a_step = g.a_mn ? 128u : 2u, b_step = g.b_mn ? 128u : 2u;
      const uint64_t mn_lbo = (uint64_t)(8192 >> 4) << 16;  // next 64-wide MN block of an MN-major tile
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA + stage * Cfg::kABytes));
          uint64_t b_desc = make_smem_desc_sw128(smem_u32(sB + stage * Cfg::kBBytes));
          if (g.a_mn) a_desc = (a_desc & ~((uint64_t)0x3FFF << 16)) | mn_lbo;
          if (g.b_mn) b_desc = (b_desc & ~((uint64_t)0x3FFF << 16)) | mn_lbo;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem, a_desc + a_step * k, b_desc + b_step * k, idesc,
                      (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (it == kiters - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which half of the tile's columns this warp drains
    constexpr int kChunksPerWarp = BN / 32 / (kEpiWarps / 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile % num_m_tiles) * BM;
      const int n0 = (tile / num_m_tiles) * BN;
      const int row = m0 + q * 32 + lane;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = half * kChunksPerWarp; c < (half + 1) * kChunksPerWarp; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 32;
        tmem_ld_32x32b_x32(taddr, r);
        tmem_ld_wait();
        const int col = n0 + c * 32;
        if (row < g.M && col < g.N) epilogue_chunk(g, r, row, col);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int BN>
static int launch_gemm(const hba_gemm_params* p, const GemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaFuncSetAttribute(gemm_tc_kernel<%d>): %s", BN, cudaGetErrorString(e));
      return HBA_ERR_CUDA;
    }
    attr_set = true;
  }
  const uint64_t a_cols = (uint64_t)p->K + (p->nsplit == 3 ? (uint64_t)p->a_lo_off : 0);
  const uint64_t b_cols = (uint64_t)p->K + (p->nsplit == 3 ? (uint64_t)p->b_lo_off : 0);
  CUtensorMap ta, tb;
  if (!p->a_mn_major) {
    HBA_CHECK(make_tma_2d_bf16(&ta, p->A, p->M, a_cols, p->lda, BM, BK));
  } else {
    const uint64_t cols = (uint64_t)p->M + (p->nsplit == 3 ? (uint64_t)p->a_lo_off : 0);
    HBA_CHECK(make_tma_2d_bf16(&ta, p->A, p->K, cols, p->lda, 64, 64));
  }
  if (!p->b_mn_major) {
    HBA_CHECK(make_tma_2d_bf16(&tb, p->B, p->N, b_cols, p->ldb, BN, BK));
  } else {
    const uint64_t cols = (uint64_t)p->N + (p->nsplit == 3 ? (uint64_t)p->b_lo_off : 0);
    HBA_CHECK(make_tma_2d_bf16(&tb, p->B, p->K, cols, p->ldb, 64, 64));
  }
  const int tiles = ((p->M + BM - 1) / BM) * ((p->N + BN - 1) / BN);
  int ctas = num_sms();
  if (p->max_ctas > 0 && p->max_ctas < ctas) ctas = p->max_ctas;
  if (tiles < ctas) ctas = tiles;
  gemm_tc_kernel<BN><<<ctas, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, g);
  return check_launch("gemm_tc_kernel");
}

}  // namespace hba

extern "C" int hba_gemm_bf16(const hba_gemm_params* p, void* stream) {
  using namespace hba;
  HBA_REQUIRE(p != nullptr, "hba_gemm_bf16: null params");
  HBA_REQUIRE(p->A && p->B, "hba_gemm_bf16: null operand");
  HBA_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0, "hba_gemm_bf16: empty problem M=%d N=%d K=%d",
              p->M, p->N, p->K);
  // a ragged last k-block is zero-filled by TMA; a split K-major operand must then keep its hi part
  // zero padded up to the lo part (checked below through the lo offsets)
  HBA_REQUIRE(p->lda % 8 == 0 && p->ldb % 8 == 0, "hba_gemm_bf16: lda/ldb must be multiples of 8");
  HBA_REQUIRE(p->nsplit == 1 || p->nsplit == 3, "hba_gemm_bf16: nsplit must be 1 or 3");
  if (p->nsplit == 3)
    HBA_REQUIRE(p->a_lo_off >= (p->a_mn_major ? p->M : (p->K + BK - 1) / BK * BK) &&
                    p->b_lo_off >= (p->b_mn_major ? p->N : (p->K + BK - 1) / BK * BK) && p->a_lo_off % 8 == 0 &&
                    p->b_lo_off % 8 == 0,
                "hba_gemm_bf16: lo offsets must lie behind the hi part and be multiples of 8");
  HBA_REQUIRE(p->out_f32 || p->out_bf16 || p->pre_out, "hba_gemm_bf16: no output");
  HBA_REQUIRE(p->act >= HBA_ACT_NONE && p->act <= HBA_ACT_GELU_ERF_GRAD, "hba_gemm_bf16: bad act");
  if (p->act == HBA_ACT_QUICKGELU_GRAD || p->act == HBA_ACT_GELU_ERF_GRAD)
    HBA_REQUIRE(p->aux != nullptr, "hba_gemm_bf16: activation gradient needs aux");
  if (!p->transpose_out) {
    HBA_REQUIRE(!p->out_f32 || (p->ld_f32 % 4 == 0 && ((uintptr_t)p->out_f32 & 15) == 0),
                "hba_gemm_bf16: out_f32 must be 16-byte aligned with ld %% 4 == 0");
    HBA_REQUIRE(!p->out_bf16 || (p->ld_bf16 % 8 == 0 && p->out_lo_off % 8 == 0 &&
                                 ((uintptr_t)p->out_bf16 & 15) == 0),
                "hba_gemm_bf16: out_bf16 must be 16-byte aligned with ld %% 8 == 0");
  }
  HBA_REQUIRE(!p->residual || (p->ldr % 4 == 0 && ((uintptr_t)p->residual & 15) == 0),
              "hba_gemm_bf16: residual must be 16-byte aligned with ld %% 4 == 0");
  HBA_REQUIRE(!p->bias || ((uintptr_t)p->bias & 15) == 0, "hba_gemm_bf16: bias must be 16-byte aligned");
  if (p->pre_out)
    HBA_REQUIRE(((uintptr_t)p->pre_out & 15) == 0 && p->ld_pre % 8 == 0,
                "hba_gemm_bf16: pre_out must be 16-byte aligned with ld %% 8 == 0");
  GemmArgs g;
  g.M = p->M, g.N = p->N, g.K = p->K;
  g.nsplit = p->nsplit, g.a_lo_off = p->a_lo_off, g.b_lo_off = p->b_lo_off;
  g.alpha = p->alpha;
  g.bias = p->bias, g.residual = p->residual, g.ldr = p->ldr;
  g.act = p->act, g.aux = p->aux, g.ld_aux = p->ld_aux, g.aux_dtype = p->aux_dtype;
  g.pre_out = p->pre_out, g.ld_pre = p->ld_pre, g.pre_dtype = p->pre_dtype;
  g.out_f32 = p->out_f32, g.ld_f32 = p->ld_f32;
  g.out_bf16 = static_cast<__nv_bfloat16*>(p->out_bf16), g.ld_bf16 = p->ld_bf16;
  g.out_lo_off = p->out_lo_off, g.transpose_out = p->transpose_out;
  g.a_mn = p->a_mn_major, g.b_mn = p->b_mn_major;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->N >= 256) return launch_gemm<256>(p, g, s);
  return launch_gemm<128>(p, g, s);
}
