// RSA evaluation tail on the GPU (reference: behavioral_RSA, NEW:625-652, i.e. numpy.corrcoef in
// float64 + scipy.stats.spearmanr = rankdata('average') + Pearson on the ranks):
//   hba_rdm_f64       E [N, Dm] fp32 -> RDM = 1 - corrcoef(E) (f64, diag 0) and its row-major upper
//                     triangle (k=1) packed as a vector of N(N-1)/2 doubles
//   hba_rank_avg_f64  ranks with ties averaged, bit-exact w.r.t. scipy.stats.rankdata(x,'average')
//                     n <= 2048: single-CTA bitonic sort in shared memory
//                     larger n : 8-pass LSD radix sort (8-bit digits, warp match_any ranking)
//   hba_pearson_f64   two-pass (means, then centred sums) deterministic float64 reduction
// Bound: HBM / L2 bandwidth (40 bytes per pair end to end), no tensor-core work.
#include "common.cuh"

namespace hba {

// ------------------------------------------------------------------------------------------
// RDM: 32x32 pair tiles; each CTA centres its 64 rows in float64 (numpy.cov promotes to f64
// before subtracting the mean) and forms cov/(sd_i sd_j) exactly in numpy's order of operations:
// c = X X^T / (Dm - 1); c /= sd_i; c /= sd_j; clip to [-1, 1].
constexpr int kRdmTile = 32;
constexpr int kRdmChunk = 64;

__global__ void __launch_bounds__(256)
    rdm_kernel(const float* __restrict__ E, int N, int Dm, double* __restrict__ rdm,
               double* __restrict__ tri) {
  __shared__ double sX[2][kRdmTile][kRdmChunk + 1];
  __shared__ double sMean[2][kRdmTile];
  __shared__ double sSd[2][kRdmTile];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;  // upper tiles only; the lower half of rdm is written by symmetry
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rr = warp; rr < 2 * kRdmTile; rr += 8) {
    const int side = rr / kRdmTile, r = rr % kRdmTile;
    const int row = (side == 0 ? bi : bj) * kRdmTile + r;
    double s = 0.0;
    if (row < N)
      for (int k = lane; k < Dm; k += 32) s += (double)E[(size_t)row * Dm + k];
    s = warp_sum(s);
    const double mean = s / Dm;
    double sq = 0.0;
    if (row < N)
      for (int k = lane; k < Dm; k += 32) {
        const double c = (double)E[(size_t)row * Dm + k] - mean;
        sq += c * c;
      }
    sq = warp_sum(sq);
    if (lane == 0) sMean[side][r] = mean, sSd[side][r] = sqrt(sq / (Dm - 1));
  }
  double dot[4] = {0.0, 0.0, 0.0, 0.0};  // pairs p = threadIdx.x + 256 u
  for (int k0 = 0; k0 < Dm; k0 += kRdmChunk) {
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * kRdmTile * kRdmChunk; t += 256) {
      const int side = t / (kRdmTile * kRdmChunk), r = (t / kRdmChunk) % kRdmTile, k = t % kRdmChunk;
      const int row = (side == 0 ? bi : bj) * kRdmTile + r;
      sX[side][r][k] = (row < N && k0 + k < Dm)
                           ? (double)E[(size_t)row * Dm + k0 + k] - sMean[side][r]
                           : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = threadIdx.x + 256 * u;
      const int a = p / kRdmTile, b = p % kRdmTile;
      double acc = dot[u];
#pragma unroll 8
      for (int k = 0; k < kRdmChunk; ++k) acc += sX[0][a][k] * sX[1][b][k];
      dot[u] = acc;
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int p = threadIdx.x + 256 * u;
    const int a = p / kRdmTile, b = p % kRdmTile;
    const int i = bi * kRdmTile + a, j = bj * kRdmTile + b;
    if (i >= N || j >= N || j < i) continue;
    double c = dot[u] / (Dm - 1);
    c /= sSd[0][a];
    c /= sSd[1][b];
    c = fmin(fmax(c, -1.0), 1.0);
    const double v = (i == j) ? 0.0 : 1.0 - c;
    if (rdm) {
      rdm[(size_t)i * N + j] = v;
      rdm[(size_t)j * N + i] = v;
    }
    if (tri && j > i) tri[(size_t)i * N - (size_t)i * (i + 1) / 2 + (j - i - 1)] = v;
  }
}

// ------------------------------------------------------------------------------------------
// double -> uint64 whose unsigned order equals the numeric order (-0.0 canonicalised to +0.0)
__device__ __forceinline__ unsigned long long f64_to_key(double x) {
  x = x + 0.0;
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// single CTA: bitonic sort of up to 2048 (key, index) pairs in shared memory, tie-run averaging
constexpr int kSmallN = 2048;
__global__ void __launch_bounds__(1024)
    rank_small_kernel(const double* __restrict__ x, int n, double* __restrict__ ranks) {
  __shared__ unsigned long long sk[kSmallN];
  __shared__ unsigned short si[kSmallN];
  for (int i = threadIdx.x; i < kSmallN; i += 1024) {
    sk[i] = (i < n) ? f64_to_key(x[i]) : ~0ull;  // padding sorts last (index breaks the tie)
    si[i] = (unsigned short)i;
  }
  __syncthreads();
  for (int k = 2; k <= kSmallN; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < kSmallN; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = ((i & k) == 0);
          const unsigned long long a = sk[i], b = sk[l];
          const unsigned short ia = si[i], ib = si[l];
          const bool gt = (a > b) || (a == b && ia > ib);
          if (gt == up) {
            sk[i] = b, sk[l] = a;
            si[i] = ib, si[l] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  // positions [0, n) now hold the real elements (padding has the largest key or index)
  for (int p = threadIdx.x; p < n; p += 1024) {
    if (p == 0 || sk[p] != sk[p - 1]) {
      int e = p;
      while (e + 1 < n && sk[e + 1] == sk[p]) ++e;
      const double r = 0.5 * (double)(p + e) + 1.0;
      for (int q = p; q <= e; ++q) ranks[si[q]] = r;
    }
  }
}

// ---- radix sort (large n) ----
constexpr int kRadixThreads = 256;
constexpr int kRadixWarps = kRadixThreads / 32;
constexpr int kRadixIters = 16;                                  // per warp
constexpr int kRadixTile = kRadixWarps * kRadixIters * 32;       // 4096 elements per CTA

__global__ void radix_init_kernel(const double* __restrict__ x, int64_t n,
                                  unsigned long long* __restrict__ keys,
                                  unsigned int* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = f64_to_key(x[i]);
    vals[i] = (unsigned int)i;
  }
}

// per-warp digit counts of this CTA's tile; element order = (warp, iteration, lane)
__device__ __forceinline__ void radix_count(const unsigned long long* __restrict__ keys, int64_t n,
                                            int shift, unsigned int (*warp_hist)[256]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = lane; b < 256; b += 32) warp_hist[warp][b] = 0;
  __syncwarp();
  const int64_t wbase = (int64_t)blockIdx.x * kRadixTile + (int64_t)warp * kRadixIters * 32;
  for (int it = 0; it < kRadixIters; ++it) {
    const int64_t idx = wbase + it * 32 + lane;
    const bool valid = idx < n;
    const unsigned int digit = valid ? (unsigned int)((keys[idx] >> shift) & 0xff) : 0x100u;
    const unsigned int peers = __match_any_sync(0xffffffffu, digit);
    if (valid && lane == (__ffs(peers) - 1)) warp_hist[warp][digit] += __popc(peers);
    __syncwarp();
  }
}

// hist[bin * num_ctas + cta] = count of `bin` in the tile of `cta`
__global__ void __launch_bounds__(kRadixThreads)
    radix_hist_kernel(const unsigned long long* __restrict__ keys, int64_t n, int shift,
                      unsigned int* __restrict__ hist) {
  __shared__ unsigned int warp_hist[kRadixWarps][256];
  radix_count(keys, n, shift, warp_hist);
  __syncthreads();
  const int b = threadIdx.x;
  unsigned int t = 0;
#pragma unroll
  for (int w = 0; w < kRadixWarps; ++w) t += warp_hist[w][b];
  hist[(size_t)b * gridDim.x + blockIdx.x] = t;
}

// exclusive scan over hist (bin-major); single CTA
__global__ void __launch_bounds__(1024)
    radix_scan_kernel(unsigned int* __restrict__ hist, int64_t len) {
  __shared__ unsigned int sums[1024];
  const int64_t per = (len + 1023) / 1024;
  const int64_t lo = threadIdx.x * per, hi = min(lo + per, len);
  unsigned int s = 0;
  for (int64_t i = lo; i < hi; ++i) s += hist[i];
  sums[threadIdx.x] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const unsigned int v = (threadIdx.x >= off) ? sums[threadIdx.x - off] : 0;
    __syncthreads();
    sums[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned int run = sums[threadIdx.x] - s;  // exclusive prefix of this thread's chunk
  for (int64_t i = lo; i < hi; ++i) {
    const unsigned int c = hist[i];
    hist[i] = run;
    run += c;
  }
}

__global__ void __launch_bounds__(kRadixThreads)
    radix_scatter_kernel(const unsigned long long* __restrict__ keys,
                         const unsigned int* __restrict__ vals, int64_t n, int shift,
                         const unsigned int* __restrict__ hist,
                         unsigned long long* __restrict__ keys_out,
                         unsigned int* __restrict__ vals_out) {
  __shared__ unsigned int warp_hist[kRadixWarps][256];
  radix_count(keys, n, shift, warp_hist);
  __syncthreads();
  {
    const int b = threadIdx.x;
    unsigned int run = hist[(size_t)b * gridDim.x + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kRadixWarps; ++w) {
      const unsigned int c = warp_hist[w][b];
      warp_hist[w][b] = run;
      run += c;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wbase = (int64_t)blockIdx.x * kRadixTile + (int64_t)warp * kRadixIters * 32;
  for (int it = 0; it < kRadixIters; ++it) {
    const int64_t idx = wbase + it * 32 + lane;
    const bool valid = idx < n;
    unsigned long long key = 0;
    unsigned int digit = 0x100u;
    if (valid) {
      key = keys[idx];
      digit = (unsigned int)((key >> shift) & 0xff);
    }
    const unsigned int peers = __match_any_sync(0xffffffffu, digit);
    const unsigned int rank = __popc(peers & ((1u << lane) - 1u));
    unsigned int base = 0;
    if (valid) base = warp_hist[warp][digit];
    __syncwarp();
    if (valid) {
      const unsigned int pos = base + rank;
      keys_out[pos] = key;
      vals_out[pos] = vals[idx];
      if (lane == (__ffs(peers) - 1)) warp_hist[warp][digit] = base + __popc(peers);
    }
    __syncwarp();
  }
}

// after the sort: each tie-run head assigns the averaged 1-based rank to every member of its run
__global__ void rank_runs_kernel(const unsigned long long* __restrict__ keys,
                                 const unsigned int* __restrict__ vals, int64_t n,
                                 double* __restrict__ ranks) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[p];
    if (p == 0 || keys[p - 1] != k) {
      int64_t e = p;
      while (e + 1 < n && keys[e + 1] == k) ++e;
      const double r = 0.5 * (double)(p + e) + 1.0;
      for (int64_t q = p; q <= e; ++q) ranks[vals[q]] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Pearson correlation in float64: pass 1 partial sums -> means; pass 2 centred partial sums;
// finalize in a fixed order (deterministic).  ws: [0,1024) sum a, [1024,2048) sum b,
// [2048,3072) Sab, [3072,4096) Saa, [4096,5120) Sbb.
constexpr int kPearsonBlocks = 1024;

__global__ void __launch_bounds__(256)
    pearson_sums_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                        double* __restrict__ ws) {
  __shared__ double scratch[32];
  double sa = 0.0, sb = 0.0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    sa += a[i];
    sb += b[i];
  }
  sa = block_sum(sa, scratch);
  sb = block_sum(sb, scratch);
  if (threadIdx.x == 0) ws[blockIdx.x] = sa, ws[kPearsonBlocks + blockIdx.x] = sb;
}

__global__ void __launch_bounds__(256)
    pearson_centered_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                            int nblocks, double* __restrict__ ws) {
  __shared__ double scratch[32];
  __shared__ double means[2];
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int i = 0; i < nblocks; ++i) sa += ws[i], sb += ws[kPearsonBlocks + i];
    means[0] = sa / (double)n, means[1] = sb / (double)n;
  }
  __syncthreads();
  const double ma = means[0], mb = means[1];
  double sab = 0.0, saa = 0.0, sbb = 0.0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const double da = a[i] - ma, db = b[i] - mb;
    sab += da * db, saa += da * da, sbb += db * db;
  }
  sab = block_sum(sab, scratch);
  saa = block_sum(saa, scratch);
  sbb = block_sum(sbb, scratch);
  if (threadIdx.x == 0) {
    ws[2 * kPearsonBlocks + blockIdx.x] = sab;
    ws[3 * kPearsonBlocks + blockIdx.x] = saa;
    ws[4 * kPearsonBlocks + blockIdx.x] = sbb;
  }
}

__global__ void pearson_final_kernel(const double* __restrict__ ws, int nblocks,
                                     double* __restrict__ rho) {
  double sab = 0.0, saa = 0.0, sbb = 0.0;
  for (int i = 0; i < nblocks; ++i) {
    sab += ws[2 * kPearsonBlocks + i];
    saa += ws[3 * kPearsonBlocks + i];
    sbb += ws[4 * kPearsonBlocks + i];
  }
  // numpy.corrcoef: c / sqrt(d_a) / sqrt(d_b), clipped to [-1, 1]
  double r = sab / sqrt(saa);
  r /= sqrt(sbb);
  *rho = fmin(fmax(r, -1.0), 1.0);
}

static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

}  // namespace hba

using namespace hba;

extern "C" int hba_rdm_f64(const float* E, int32_t N, int32_t Dm, double* rdm, double* tri,
                           void* stream) {
  HBA_REQUIRE(E && (rdm || tri) && N > 1, "hba_rdm_f64: bad arguments");
  HBA_REQUIRE(Dm > 1, "hba_rdm_f64: Dm=%d must be > 1", Dm);
  const int nb = (N + kRdmTile - 1) / kRdmTile;
  rdm_kernel<<<dim3(nb, nb), 256, 0, static_cast<cudaStream_t>(stream)>>>(E, N, Dm, rdm, tri);
  return check_launch("rdm_kernel");
}

extern "C" int64_t hba_rank_workspace_bytes(int64_t n) {
  if (n <= kSmallN) return 256;
  const int64_t ctas = (n + kRadixTile - 1) / kRadixTile;
  return 2 * align256(n * 8) + 2 * align256(n * 4) + align256(256 * ctas * 4) + 256;
}

extern "C" int hba_rank_avg_f64(const double* x, int64_t n, double* ranks, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  HBA_REQUIRE(x && ranks && n > 0, "hba_rank_avg_f64: bad arguments");
  HBA_REQUIRE(n < (1ll << 32), "hba_rank_avg_f64: n too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n <= kSmallN) {
    rank_small_kernel<<<1, 1024, 0, s>>>(x, (int)n, ranks);
    return check_launch("rank_small_kernel");
  }
  HBA_REQUIRE(workspace && workspace_bytes >= hba_rank_workspace_bytes(n) &&
                  ((uintptr_t)workspace & 255) == 0,
              "hba_rank_avg_f64: workspace too small or not 256-byte aligned (need %lld bytes)",
              (long long)hba_rank_workspace_bytes(n));
  const int ctas = (int)((n + kRadixTile - 1) / kRadixTile);
  char* w = static_cast<char*>(workspace);
  unsigned long long* k0 = reinterpret_cast<unsigned long long*>(w);
  unsigned long long* k1 = reinterpret_cast<unsigned long long*>(w + align256(n * 8));
  unsigned int* v0 = reinterpret_cast<unsigned int*>(w + 2 * align256(n * 8));
  unsigned int* v1 = reinterpret_cast<unsigned int*>(w + 2 * align256(n * 8) + align256(n * 4));
  unsigned int* hist = reinterpret_cast<unsigned int*>(w + 2 * align256(n * 8) + 2 * align256(n * 4));
  int g = (int)((n + 255) / 256);
  if (g > num_sms() * 8) g = num_sms() * 8;
  radix_init_kernel<<<g, 256, 0, s>>>(x, n, k0, v0);
  HBA_CHECK(check_launch("radix_init_kernel"));
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = pass * 8;
    radix_hist_kernel<<<ctas, kRadixThreads, 0, s>>>(k0, n, shift, hist);
    radix_scan_kernel<<<1, 1024, 0, s>>>(hist, (int64_t)256 * ctas);
    radix_scatter_kernel<<<ctas, kRadixThreads, 0, s>>>(k0, v0, n, shift, hist, k1, v1);
    HBA_CHECK(check_launch("radix pass"));
    unsigned long long* tk = k0; k0 = k1; k1 = tk;
    unsigned int* tv = v0; v0 = v1; v1 = tv;
  }
  rank_runs_kernel<<<g, 256, 0, s>>>(k0, v0, n, ranks);
  return check_launch("rank_runs_kernel");
}

extern "C" int hba_pearson_f64(const double* a, const double* b, int64_t n, double* rho_out,
                               double* workspace, void* stream) {
  HBA_REQUIRE(a && b && rho_out && workspace && n > 1, "hba_pearson_f64: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int nblocks = (int)((n + 255) / 256);
  if (nblocks > kPearsonBlocks) nblocks = kPearsonBlocks;
  pearson_sums_kernel<<<nblocks, 256, 0, s>>>(a, b, n, workspace);
  pearson_centered_kernel<<<nblocks, 256, 0, s>>>(a, b, n, nblocks, workspace);
  pearson_final_kernel<<<1, 1, 0, s>>>(workspace, nblocks, rho_out);
  return check_launch("pearson kernels");
}
