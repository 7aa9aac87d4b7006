// Skinny FP64 GEMMs of the rSVD range finder on sm_100a: FP64 tensor-core (DMMA) tiles fed by TMA.
//
//   K1  gemm_an :  Y (M x N) = A (M x K) * X (K x N)          replaces Eigen `A * Omega`, `A * Q`   (reference src/rSVD.cpp:59,66)
//   K2  gemm_at :  Z (M x N) = A^T, A (K x M), times Q (K x N) replaces Eigen `A.transpose() * Q`    (src/rSVD.cpp:63)
//   K3  gemm_at with transposed store: B (N x M) = Q^T A      replaces Eigen `Q.transpose() * A`     (src/rSVD.cpp:89)
//
// All matrices are column-major FP64 (Eigen::MatrixXd layout, include/rSVD.hpp:9).  N = l <= 128 per launch.
//
// Structure (both kernels): persistent CTAs, 1 TMA producer warp + 8 DMMA consumer warps, an mbarrier full/empty ring of
// `stages` slabs of BK = 16 reduction indices.  Every smem row is 16 doubles = 128 bytes and is written by TMA with
// SWIZZLE_128B, so that the per-lane 16-byte fragment loads below are bank-conflict free:
//   * K1: A slab = 8 boxes {16 rows(m) x 16 k}, one per consumer warp; smem row = k, 16-byte chunk c holds rows (2c,2c+1)
//         at chunk position c ^ (k & 7).
//   * K2: A slab = one box {16 k(m) x 128 cols(j)}; smem row = j, chunk c holds k = (2c,2c+1) at position c ^ (j & 7).
//   * X / Q slab = one box {16 k x 8*NB cols}; smem row = column n, chunk c holds k = (2c,2c+1) at position c ^ (n & 7).
// The MMA is mma.sync.m16n8k4.f64 (SASS DMMA).  Index permutations (free, because the MMA does not care which matrix row
// sits in which fragment slot) make one ld.shared.v2.f64 deliver fragment values for two MMAs:
//   k-slot t of MMA#1 <-> k = kk+2t, of MMA#2 <-> k = kk+2t+1;  n-slot x <-> column 8j + perm(x), perm(x) = (x>>1)|((x&1)<<2);
//   K1: m-slot g <-> row 2g, m-slot g+8 <-> row 2g+1;  K2: m-slot g <-> column perm(g), m-slot g+8 <-> column 8+perm(g).
// Work is cut into (tile, k-split) units; with nsplit > 1 each unit stores a partial tile to a workspace and a second
// kernel sums the partials in a fixed order (bitwise reproducible; no atomics).
#include "dev_once.cuh"
#include "gemm_dmma.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace rsvdb {

namespace {

constexpr int BM = 128;    // tile extent along the non-reduced dimension of A (K1: rows, K2: columns)
constexpr int BK = 16;     // reduction slab per stage = one 128-byte swizzle row
constexpr int NCONS = 8;   // consumer warps; each owns 16 of the BM tile rows
constexpr int NTHREADS = (NCONS + 1) * 32;
constexpr uint32_t A_BYTES = BM * BK * 8;

struct GemmParams {
  double* out;            // S == 1: final output; S > 1: workspace, partial s at out + s * split_stride
  long long ld_out;
  long long split_stride;
  int M, N, K;            // tile dimension, l, reduction length
  int ntiles, nsplit, kchunk;
  int transpose_out;      // K2 only: store element (j, c) at out[c + j * ld_out]
  int vec_ok;             // K1 only: 16-byte stores allowed
  int stages;
  int sh0, sh1;           // SPLIT only: 1 = the even- / odd-column class starts at 8 mod 16 and is copied with cp.async
  const double* Araw;     // SPLIT only: A and its leading dimension for the cp.async classes
  long long lda;
};

// SPLIT = A is not TMA-addressable as one tensor: its leading dimension is odd (column starts alternate between 16-byte
// aligned and 8 mod 16) or its base is only 8-byte aligned.  The columns of one parity are 2*lda apart -- a legal TMA stride --
// so A is split into its even and its odd columns.  A class whose columns start 16-byte aligned gets its own tensor map; a
// class that starts at 8 mod 16 cannot be read by TMA at all (every box would start at an address that is 8 mod 16, which
// the unit rejects: measured, "illegal instruction"), so the producer warp copies it with 8-byte cp.async (LDGSTS) into the
// same swizzled layout and signals the stage's mbarrier with cp.async.mbarrier.arrive.  K1: the 16 reduction indices of a
// stage land as smem lines [k even | k odd]; the consumers pair line L with X row k(L) = L < 8 ? 2L : 2(L-8)+1.  K2: the
// 128 tile columns land as smem rows [even | odd]; only the epilogue's row -> column map changes.  Same bytes, same DMMA work.
template <int NB> struct SmemCfg {
  static constexpr uint32_t X_BYTES = NB * 8 * BK * 8;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + X_BYTES;
};

__device__ __forceinline__ const double2* frag_ptr(const uint8_t* base, uint32_t off) {
  return reinterpret_cast<const double2*>(base + off);
}

// ------------------------------------------------------------------------------------------------------------------
// K1: Y = A * X
// ------------------------------------------------------------------------------------------------------------------
template <int NB, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_an(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmX, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t STAGE = SmemCfg<NB>::STAGE_BYTES;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], SPLIT ? 33 : 1); mbar_init(&empty[s], NCONS); }   // SPLIT: + one cp.async arrival per producer lane
    fence_mbar_init();
  }
  __syncthreads();

  const int nunits = p.ntiles * p.nsplit;
  int stage = 0; uint32_t phase = 0;

  if (warp == NCONS) {
    // ---------------- TMA producer ----------------
    if (SPLIT) {
      // all 32 lanes: lane 0 issues the TMA loads (X, and the A class(es) that TMA can address), every lane copies its share
      // of the other class(es) with cp.async and arrives on the stage barrier when its copies have landed
      if (lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmX); }
      const uint32_t tma_bytes = SmemCfg<NB>::X_BYTES + (p.sh0 ? 0u : A_BYTES / 2) + (p.sh1 ? 0u : A_BYTES / 2);
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int tile = u / p.nsplit, split = u - tile * p.nsplit;
        const int k0 = split * p.kchunk, k1 = min(p.K, k0 + p.kchunk);
        const int m0 = tile * BM;
        for (int k = k0; k < k1; k += BK) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * STAGE;
          if (lane == 0) {
            mbar_expect_tx(&full[stage], tma_bytes);
#pragma unroll
            for (int b = 0; b < NCONS; ++b) {
              if (!p.sh0) tma_load_2d(sa + b * (16 * BK * 8), &tmA, &full[stage], m0 + 16 * b, k >> 1);            // lines 0-7: k even
              if (!p.sh1) tma_load_2d(sa + b * (16 * BK * 8) + 1024, &tmA1, &full[stage], m0 + 16 * b, k >> 1);    // lines 8-15: k odd
            }
            tma_load_2d(sa + A_BYTES, &tmX, &full[stage], k, 0);
          }
#pragma unroll
          for (int cls = 0; cls < 2; ++cls) {
            if (cls ? p.sh1 : p.sh0) {
              // 8 boxes x 8 lines x 16 rows = 1024 elements per class; element e = lane + 32 q: row i = lane & 15, line kk = (lane >> 4) + 2 (q & 3), box b = q >> 2
              const int li = lane & 15, lj = lane >> 4;
              const double* base = p.Araw + (size_t)(k + cls + 2 * lj) * p.lda + m0 + li;
              const size_t step = (size_t)4 * p.lda;
#pragma unroll
              for (int q = 0; q < 32; ++q) {
                const int kk = lj + 2 * (q & 3), b = q >> 2;
                const bool ok = (k + 2 * kk + cls) < p.K && (m0 + 16 * b + li) < p.M;
                const double* src = ok ? base + (q & 3) * step + 16 * b : p.Araw;
                cp_async_8(sa + b * (16 * BK * 8) + cls * 1024 + kk * 128 + (((li >> 1) ^ kk) << 4) + (li & 1) * 8, src, ok ? 8u : 0u);
              }
            }
          }
          cp_async_mbar_arrive_noinc(&full[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    } else
    if (lane == 0) {
      tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmX);
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int tile = u / p.nsplit, split = u - tile * p.nsplit;
        const int k0 = split * p.kchunk, k1 = min(p.K, k0 + p.kchunk);
        const int m0 = tile * BM;
        for (int k = k0; k < k1; k += BK) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE);
          uint8_t* sa = smem + (size_t)stage * STAGE;
#pragma unroll
          for (int b = 0; b < NCONS; ++b) tma_load_2d(sa + b * (16 * BK * 8), &tmA, &full[stage], m0 + 16 * b, k);
          tma_load_2d(sa + A_BYTES, &tmX, &full[stage], k, 0);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ---------------- DMMA consumers ----------------
    const int g = lane >> 2, t = lane & 3;
    const int pg = (g >> 1) | ((g & 1) << 2);
    const uint32_t offA0 = warp * (16 * BK * 8) + (2 * t) * 128 + ((g ^ (2 * t)) << 4);
    const uint32_t offA1 = warp * (16 * BK * 8) + (2 * t + 1) * 128 + ((g ^ (2 * t + 1)) << 4);
    const uint32_t offB0 = A_BYTES + pg * 128 + ((t ^ pg) << 4);
    const uint32_t offB1 = A_BYTES + pg * 128 + (((4 + t) ^ pg) << 4);
    const uint32_t offBs0 = A_BYTES + pg * 128 + (((2 * t) ^ pg) << 4);       // SPLIT: X rows k = 4t, 4t+1
    const uint32_t offBs1 = A_BYTES + pg * 128 + (((2 * t + 1) ^ pg) << 4);   //        X rows k = 4t+2, 4t+3

    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int tile = u / p.nsplit, split = u - tile * p.nsplit;
      const int k0 = split * p.kchunk, k1 = min(p.K, k0 + p.kchunk);
      double c[NB][4];
#pragma unroll
      for (int j = 0; j < NB; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.0;

      for (int k = k0; k < k1; k += BK) {
        mbar_wait(&full[stage], phase);
        const uint8_t* ss = smem + (size_t)stage * STAGE;
        if (SPLIT) {
          // line 2t <-> k = 4t, line 2t+1 <-> k = 4t+2, line 8+2t <-> k = 4t+1, line 8+2t+1 <-> k = 4t+3
          const double2 a00 = *frag_ptr(ss, offA0), a10 = *frag_ptr(ss, offA1);
          const double2 a01 = *frag_ptr(ss, offA0 + 1024), a11 = *frag_ptr(ss, offA1 + 1024);
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const double2 blo = *frag_ptr(ss, offBs0 + j * 1024);   // column 8j + perm(g), k = (4t, 4t+1)
            const double2 bhi = *frag_ptr(ss, offBs1 + j * 1024);   //                      k = (4t+2, 4t+3)
            dmma_16x8x4(c[j], a00.x, a00.y, blo.x);
            dmma_16x8x4(c[j], a10.x, a10.y, bhi.x);
            dmma_16x8x4(c[j], a01.x, a01.y, blo.y);
            dmma_16x8x4(c[j], a11.x, a11.y, bhi.y);
          }
        } else
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double2 a0 = *frag_ptr(ss, offA0 + h * 1024);   // rows (2g, 2g+1), k = 8h + 2t
          const double2 a1 = *frag_ptr(ss, offA1 + h * 1024);   // rows (2g, 2g+1), k = 8h + 2t + 1
          const uint32_t ob = h ? offB1 : offB0;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const double2 b = *frag_ptr(ss, ob + j * 1024);     // column 8j + perm(g), k = 8h + (2t, 2t+1)
            dmma_16x8x4(c[j], a0.x, a0.y, b.x);
            dmma_16x8x4(c[j], a1.x, a1.y, b.y);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }

      // epilogue: thread owns rows (r, r+1), r = m0 + 16*warp + 2g, columns 8j + t and 8j + t + 4
      double* out = p.out + (size_t)split * p.split_stride;
      const int r = tile * BM + 16 * warp + 2 * g;
      if (r < p.M) {
        const bool pair = (r + 1 < p.M);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int n = 8 * j + t + 4 * e;
            if (n < p.N) {
              double* dst = out + (size_t)n * p.ld_out + r;
              if (pair && p.vec_ok) {
                *reinterpret_cast<double2*>(dst) = make_double2(c[j][e], c[j][2 + e]);
              } else {
                dst[0] = c[j][e];
                if (pair) dst[1] = c[j][2 + e];
              }
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K2 / K3: Z = A^T * Q   (A is K x M column-major: the reduction runs down the contiguous dimension)
// ------------------------------------------------------------------------------------------------------------------
template <int NB, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_at(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmQ, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t STAGE = SmemCfg<NB>::STAGE_BYTES;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], SPLIT ? 33 : 1); mbar_init(&empty[s], NCONS); }   // SPLIT: + one cp.async arrival per producer lane
    fence_mbar_init();
  }
  __syncthreads();

  const int nunits = p.ntiles * p.nsplit;
  int stage = 0; uint32_t phase = 0;

  if (warp == NCONS) {
    if (SPLIT) {
      if (lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmQ); }
      const uint32_t tma_bytes = SmemCfg<NB>::X_BYTES + (p.sh0 ? 0u : A_BYTES / 2) + (p.sh1 ? 0u : A_BYTES / 2);
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int tile = u / p.nsplit, split = u - tile * p.nsplit;
        const int k0 = split * p.kchunk, k1 = min(p.K, k0 + p.kchunk);
        const int j0 = tile * BM;
        for (int k = k0; k < k1; k += BK) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * STAGE;
          if (lane == 0) {
            mbar_expect_tx(&full[stage], tma_bytes);
            if (!p.sh0) tma_load_2d(sa, &tmA, &full[stage], k, j0 >> 1);                 // smem rows 0-63: even tile columns
            if (!p.sh1) tma_load_2d(sa + 64 * 128, &tmA1, &full[stage], k, j0 >> 1);     // smem rows 64-127: odd tile columns
            tma_load_2d(sa + A_BYTES, &tmQ, &full[stage], k, 0);
          }
#pragma unroll
          for (int cls = 0; cls < 2; ++cls) {
            if (cls ? p.sh1 : p.sh0) {
              // 64 columns x 16 reduction indices per class; element e = lane + 32 q: k offset i = lane & 15, column jj = (lane >> 4) + 2 q
              const int li = lane & 15, lj = lane >> 4;
              const double* base = p.Araw + (size_t)(j0 + cls + 2 * lj) * p.lda + k + li;
              const size_t step = (size_t)4 * p.lda;
              const bool kok = (k + li) < p.K;
#pragma unroll
              for (int q = 0; q < 32; ++q) {
                const int jj = lj + 2 * q, R = 64 * cls + jj;
                const bool ok = kok && (j0 + 2 * jj + cls) < p.M;
                const double* src = ok ? base + q * step : p.Araw;
                cp_async_8(sa + R * 128 + (((li >> 1) ^ (R & 7)) << 4) + (li & 1) * 8, src, ok ? 8u : 0u);
              }
            }
          }
          cp_async_mbar_arrive_noinc(&full[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    } else
    if (lane == 0) {
      tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmQ);
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int tile = u / p.nsplit, split = u - tile * p.nsplit;
        const int k0 = split * p.kchunk, k1 = min(p.K, k0 + p.kchunk);
        const int j0 = tile * BM;
        for (int k = k0; k < k1; k += BK) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE);
          uint8_t* sa = smem + (size_t)stage * STAGE;
          tma_load_2d(sa, &tmA, &full[stage], k, j0);
          tma_load_2d(sa + A_BYTES, &tmQ, &full[stage], k, 0);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int g = lane >> 2, t = lane & 3;
    const int pg = (g >> 1) | ((g & 1) << 2);
    // A slab row = tile column (16*warp + perm(g)) and (16*warp + 8 + perm(g)); both have (row & 7) == perm(g)
    const uint32_t offAlo0 = (16 * warp + pg) * 128 + ((t ^ pg) << 4);
    const uint32_t offAlo1 = (16 * warp + pg) * 128 + (((4 + t) ^ pg) << 4);
    const uint32_t offB0 = A_BYTES + pg * 128 + ((t ^ pg) << 4);
    const uint32_t offB1 = A_BYTES + pg * 128 + (((4 + t) ^ pg) << 4);

    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int tile = u / p.nsplit, split = u - tile * p.nsplit;
      const int k0 = split * p.kchunk, k1 = min(p.K, k0 + p.kchunk);
      double c[NB][4];
#pragma unroll
      for (int j = 0; j < NB; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.0;

      for (int k = k0; k < k1; k += BK) {
        mbar_wait(&full[stage], phase);
        const uint8_t* ss = smem + (size_t)stage * STAGE;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t oa = h ? offAlo1 : offAlo0;
          const double2 alo = *frag_ptr(ss, oa);              // column 16w + perm(g),     k = 8h + (2t, 2t+1)
          const double2 ahi = *frag_ptr(ss, oa + 8 * 128);    // column 16w + 8 + perm(g), k = 8h + (2t, 2t+1)
          const uint32_t ob = h ? offB1 : offB0;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const double2 b = *frag_ptr(ss, ob + j * 1024);
            dmma_16x8x4(c[j], alo.x, ahi.x, b.x);
            dmma_16x8x4(c[j], alo.y, ahi.y, b.y);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }

      // epilogue: thread owns tile columns jlo = 16w + perm(g), jhi = jlo + 8 and output columns 8j + t, 8j + t + 4
      double* out = p.out + (size_t)split * p.split_stride;
      const int rlo = 16 * warp + pg, rhi = rlo + 8;                      // smem rows of this thread's two tile columns
      const int jlo = tile * BM + (SPLIT ? ((rlo & 63) << 1) + (rlo >> 6) : rlo);
      const int jhi = tile * BM + (SPLIT ? ((rhi & 63) << 1) + (rhi >> 6) : rhi);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = 8 * j + t + 4 * e;
          if (n < p.N) {
            if (p.transpose_out) {
              if (jlo < p.M) out[(size_t)jlo * p.ld_out + n] = c[j][e];
              if (jhi < p.M) out[(size_t)jhi * p.ld_out + n] = c[j][2 + e];
            } else {
              if (jlo < p.M) out[(size_t)n * p.ld_out + jlo] = c[j][e];
              if (jhi < p.M) out[(size_t)n * p.ld_out + jhi] = c[j][2 + e];
            }
          }
        }
      }
    }
  }
}

// Sum nsplit partial (rows x cols) tiles in a fixed order.  Partials are column-major with leading dimension ld_ws.
// transpose == 0: out[r + c*ld_out]; transpose == 1: out[c + r*ld_out] (staged through shared memory so both sides coalesce).
// gridDim.z > 1: slab z sums the partials [z * group, min(nsplit, (z + 1) * group)) into out + z * out_z_stride (first stage of
// the two-stage reduction used when the output has too few tiles to occupy the GPU).
__global__ void k_reduce_splits(const double* __restrict__ ws, long long split_stride, int nsplit, long long ld_ws,
                                double* __restrict__ out, long long ld_out, int rows, int cols, int transpose,
                                int group, long long out_z_stride) {
  __shared__ double tile[32][33];
  if (gridDim.z > 1) {
    const int k0 = blockIdx.z * group;
    ws += (size_t)k0 * split_stride; out += (size_t)blockIdx.z * out_z_stride;
    nsplit = min(group, nsplit - k0);
  }
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  // the four elements of a thread and eight partials of each are loaded together (32 independent loads in flight: with few
  // output tiles -- a Gram matrix -- the kernel is latency-bound); the additions keep the fixed order k = 0, 1, 2, ...
  const int r = r0 + tx;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  bool on[4];
  const double* src[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = c0 + ty + 8 * q;
    on[q] = r < rows && c < cols;
    src[q] = ws + (size_t)(on[q] ? c : 0) * ld_ws + (on[q] ? r : 0);
  }
  for (int k = 0; k < nsplit; k += 8) {
    double v[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int u = 0; u < 8; ++u) v[q][u] = (on[q] && k + u < nsplit) ? src[q][(size_t)(k + u) * split_stride] : 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int u = 0; u < 8; ++u) if (k + u < nsplit) s[q] += v[q][u];
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int cc = ty + 8 * q, c = c0 + cc;
    if (on[q] && !transpose) out[(size_t)c * ld_out + r] = s[q];
    tile[cc][tx] = s[q];
  }
  if (transpose) {
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
      const int r = r0 + rr, c = c0 + tx;
      if (r < rows && c < cols) out[(size_t)r * ld_out + c] = tile[tx][rr];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Generic FP64 GEMM on CUDA cores for the small dense products off the streaming path (l x l factors, API helpers,
// operands that TMA cannot address).  C (m x n) = alpha * op(A) * op(B) + beta * C, column-major.
// ------------------------------------------------------------------------------------------------------------------
constexpr int GT = 64, GKT = 16;
__global__ void __launch_bounds__(256)
k_gemm_generic(int ta, int tb, int m, int n, int k, double alpha, const double* __restrict__ A, long long lda,
               const double* __restrict__ B, long long ldb, double beta, double* __restrict__ C, long long ldc) {
  __shared__ double As[GKT][GT + 1];
  __shared__ double Bs[GKT][GT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < k; k0 += GKT) {
    for (int e = threadIdx.x; e < GT * GKT; e += 256) {
      int i, kk;
      if (!ta) { i = e % GT; kk = e / GT; } else { kk = e % GKT; i = e / GKT; }
      const int gi = m0 + i, gk = k0 + kk;
      As[kk][i] = (gi < m && gk < k) ? (ta ? A[(size_t)gi * lda + gk] : A[(size_t)gk * lda + gi]) : 0.0;
      int j, kb;
      if (!tb) { kb = e % GKT; j = e / GKT; } else { j = e % GT; kb = e / GT; }
      const int gj = n0 + j, gkb = k0 + kb;
      Bs[kb][j] = (gj < n && gkb < k) ? (tb ? B[(size_t)gkb * ldb + gj] : B[(size_t)gj * ldb + gkb]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GKT; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][tx + 16 * i]; b[i] = Bs[kk][ty + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = m0 + tx + 16 * i, gj = n0 + ty + 16 * j;
      if (gi < m && gj < n) {
        double* dst = C + (size_t)gj * ldc + gi;
        *dst = (beta == 0.0) ? alpha * acc[i][j] : alpha * acc[i][j] + beta * (*dst);
      }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D FP64 column-major matrix (dim0 = rows, contiguous; dim1 = cols, stride ld), box {16 rows, box_cols}, SWIZZLE_128B.
bool make_map(CUtensorMap* map, const double* base, long long rows, long long cols, long long ld, int box_cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {16, (cuuint32_t)box_cols};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool tma_addressable(const double* p, long long ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 1) == 0; }

// Cut the reduction into nsplit slabs so that (tiles x splits) fills the SMs in whole waves.  Cost model in units of one
// k-step of one CTA: waves(S) * ceil(ksteps / S) + (S > 1 ? (S + 1) * reduce_units : 0), where reduce_units is the time
// to stream one partial output through HBM (the fixed-order reduction reads S partials and writes one result).
void choose_split(int ntiles, int ksteps, int nsm, long long out_elems, int nb, int* nsplit, int* kchunk) {
  const double kstep_us = (double)BM * (8.0 * nb) * BK * 2.0 / (37.1e12 / nsm) * 1e6;   // one CTA, one slab, at the DMMA rate
  const double reduce_units = ((double)out_elems * 8.0 / 6.5e12 * 1e6) / kstep_us;
  int best = 1; double best_t = 1e300;
  // at most 48 slabs, except when the output has so few tiles (a Gram matrix: one) that 48 would leave SMs idle
  const int smax = std::max(1, std::min(std::max(48, nsm / std::max(1, ntiles)), ksteps / 8));
  for (int s = 1; s <= smax; ++s) {
    const long long units = (long long)ntiles * s;
    const long long waves = (units + nsm - 1) / nsm;
    const int per = (ksteps + s - 1) / s;
    double t = (double)waves * per + (s > 1 ? (s + 1) * reduce_units + 4.0 : 0.0);
    if (t < best_t * 0.995) { best_t = t; best = s; }
  }
  const int per = (ksteps + best - 1) / best;
  *kchunk = per * BK;
  *nsplit = (ksteps + per - 1) / per;
}

template <int NB, bool SPLIT> cudaError_t launch_an(const CUtensorMap& tA, const CUtensorMap& tA1, const CUtensorMap& tX, GemmParams p, int grid, cudaStream_t st) {
  constexpr uint32_t STAGE = SmemCfg<NB>::STAGE_BYTES;
  int stages = std::min(8, (int)((220 * 1024 - 1024) / STAGE));
  p.stages = stages;
  const size_t smem = (size_t)stages * STAGE + 1024 + 2 * stages * sizeof(uint64_t);
  static DevOnce attr_set;
  if (!attr_set.get()) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_an<NB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set.set();
  }
  k_gemm_an<NB, SPLIT><<<grid, NTHREADS, smem, st>>>(tA, tA1, tX, p);
  return cudaGetLastError();
}
template <int NB, bool SPLIT> cudaError_t launch_at(const CUtensorMap& tA, const CUtensorMap& tA1, const CUtensorMap& tQ, GemmParams p, int grid, cudaStream_t st) {
  constexpr uint32_t STAGE = SmemCfg<NB>::STAGE_BYTES;
  int stages = std::min(8, (int)((220 * 1024 - 1024) / STAGE));
  p.stages = stages;
  const size_t smem = (size_t)stages * STAGE + 1024 + 2 * stages * sizeof(uint64_t);
  static DevOnce attr_set;
  if (!attr_set.get()) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_at<NB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set.set();
  }
  k_gemm_at<NB, SPLIT><<<grid, NTHREADS, smem, st>>>(tA, tA1, tQ, p);
  return cudaGetLastError();
}
#define RSVDB_NB_SWITCH(FN, SPL, ...)                                                                                  \
  switch (NB) {                                                                                                        \
    case 1: e = FN<1, SPL>(__VA_ARGS__); break;   case 2: e = FN<2, SPL>(__VA_ARGS__); break;                             \
    case 3: e = FN<3, SPL>(__VA_ARGS__); break;   case 4: e = FN<4, SPL>(__VA_ARGS__); break;                             \
    case 5: e = FN<5, SPL>(__VA_ARGS__); break;   case 6: e = FN<6, SPL>(__VA_ARGS__); break;                             \
    case 7: e = FN<7, SPL>(__VA_ARGS__); break;   case 8: e = FN<8, SPL>(__VA_ARGS__); break;                             \
    case 9: e = FN<9, SPL>(__VA_ARGS__); break;   case 10: e = FN<10, SPL>(__VA_ARGS__); break;                           \
    case 11: e = FN<11, SPL>(__VA_ARGS__); break; case 12: e = FN<12, SPL>(__VA_ARGS__); break;                           \
    case 13: e = FN<13, SPL>(__VA_ARGS__); break; case 14: e = FN<14, SPL>(__VA_ARGS__); break;                           \
    case 15: e = FN<15, SPL>(__VA_ARGS__); break; default: e = FN<16, SPL>(__VA_ARGS__); break;                           \
  }

// A that one tensor map cannot describe: one map per column parity where that class starts 16-byte aligned; sh = 1 marks a
// class that starts at 8 mod 16 and is copied with cp.async instead (its map slot gets a valid dummy).
// rows x cols matrix, column stride lda; box {16, box_cols}.  Returns false when there are fewer than two columns.
bool make_split_maps(CUtensorMap* m0, CUtensorMap* m1, int* sh0, int* sh1, const double* A, long long rows, long long cols, long long lda, int box_cols) {
  if (cols < 2) return false;
  const double* a0 = A; const double* a1 = A + lda;
  *sh0 = (int)((reinterpret_cast<uintptr_t>(a0) >> 3) & 1); *sh1 = (int)((reinterpret_cast<uintptr_t>(a1) >> 3) & 1);
  bool ok = true;
  if (!*sh0) ok = ok && make_map(m0, a0, rows, (cols + 1) / 2, 2 * lda, box_cols);
  if (!*sh1) ok = ok && make_map(m1, a1, rows, cols / 2, 2 * lda, box_cols);
  if (*sh0 && !*sh1) *m0 = *m1;
  if (*sh1 && !*sh0) *m1 = *m0;
  if (*sh0 && *sh1) {                        // both classes by cp.async: any valid descriptor (over the aligned element after A)
    ok = make_map(m0, A + 1, rows > 1 ? rows - 1 : 1, 1, 2 * lda, box_cols); *m1 = *m0;
  }
  return ok;
}

// the skinny operand (K x N) is small: when IT is not TMA-addressable it is repacked into an even-ld scratch
__global__ void k_repack(const double* __restrict__ src, long long lds, double* __restrict__ dst, long long ldd, long long rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) for (int k = blockIdx.y; k < cols; k += gridDim.y) dst[(size_t)k * ldd + i] = src[(size_t)k * lds + i];
}

// Fixed-order sum of the split-K partials.  An output of a few tiles (a Gram matrix: 16) with many slabs would leave one thread
// chasing ~150 dependent loads on 16 SMs; it is reduced in two stages instead (groups of ~sqrt(nsplit) slabs over gridDim.z, then
// the group sums): same result on every run and every rank, ~4x shorter.  `scratch` holds reduce_groups() partial outputs.
int reduce_groups(long long rows, int cols, int nsplit) {
  const long long tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
  if (tiles > 64 || nsplit < 24) return 1;
  int nz = 1; while (nz * nz < nsplit) ++nz;
  return nz;
}
cudaError_t reduce_partials(cudaStream_t st, const double* part, long long split_stride, int nsplit, long long ld_ws, double* out,
                            long long ld_out, int rows, int cols, int transpose, double* scratch, int* launches) {
  dim3 rg((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  const int nz = reduce_groups(rows, cols, nsplit);
  if (nz > 1) {
    const int group = (nsplit + nz - 1) / nz, nzz = (nsplit + group - 1) / group;
    const long long zs = (long long)rows * cols;
    k_reduce_splits<<<dim3(rg.x, rg.y, (unsigned)nzz), dim3(32, 8), 0, st>>>(part, split_stride, nsplit, ld_ws, scratch, rows, rows, cols, 0, group, zs);
    k_reduce_splits<<<rg, dim3(32, 8), 0, st>>>(scratch, zs, nzz, rows, out, ld_out, rows, cols, transpose, 0, 0);
    if (launches) *launches += 2;
  } else {
    k_reduce_splits<<<rg, dim3(32, 8), 0, st>>>(part, split_stride, nsplit, ld_ws, out, ld_out, rows, cols, transpose, 0, 0);
    if (launches) ++*launches;
  }
  return cudaGetLastError();
}

static long long g_generic_fallbacks = 0;   // products that took the CUDA-core kernel (operands TMA cannot describe at all)
static long long g_split_products = 0;      // products whose A needed the two-map (odd lda / 8-byte aligned base) path
void note_generic_fallback() { ++g_generic_fallbacks; }

}  // namespace

long long generic_fallback_count() { return g_generic_fallbacks; }

cudaError_t gemm_generic(cudaStream_t st, int ta, int tb, int m, int n, int k, double alpha, const double* A, long long lda,
                         const double* B, long long ldb, double beta, double* C, long long ldc) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  dim3 grid((m + GT - 1) / GT, (n + GT - 1) / GT);
  k_gemm_generic<<<grid, 256, 0, st>>>(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  return cudaGetLastError();
}

long long split_product_count() { return g_split_products; }

cudaError_t gemm_an(GemmWorkspace& ws, cudaStream_t st, int nsm, const double* A, long long M, long long K, long long lda,
                    const double* X, long long ldx, int N, double* Y, long long ldy, int* launches) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (K <= 0) {
    for (int n = 0; n < N; ++n) { cudaError_t e = cudaMemsetAsync(Y + (size_t)n * ldy, 0, (size_t)M * 8, st); if (e != cudaSuccess) return e; }
    return cudaSuccess;
  }
  if (M >= (1LL << 31) || K >= (1LL << 31)) return cudaErrorInvalidValue;     // explicit error: dimensions are 32-bit inside the kernels
  const bool splitA = !tma_addressable(A, lda);
  const bool repackX = !tma_addressable(X, ldx);
  CUtensorMap tA, tA1; int sh0 = 0, sh1 = 0;
  if (splitA ? !make_split_maps(&tA, &tA1, &sh0, &sh1, A, M, K, lda, BK / 2) : !make_map(&tA, A, M, K, lda, BK)) {
    if (!splitA) return cudaErrorInvalidValue;
    note_generic_fallback();                                                   // K == 1: nothing to stream
    if (launches) ++*launches;
    return gemm_generic(st, 0, 0, (int)M, N, (int)K, 1.0, A, lda, X, ldx, 0.0, Y, ldy);
  }
  if (!splitA) tA1 = tA; else ++g_split_products;
  const int ntiles = (int)((M + BM - 1) / BM), ksteps = (int)((K + BK - 1) / BK);
  int nsplit, kchunk; choose_split(ntiles, ksteps, nsm, M * std::min(N, 128), (std::min(N, 128) + 7) / 8, &nsplit, &kchunk);
  const long long ldw = (M + 1) & ~1LL, ldxr = (K + 1) & ~1LL;
  const size_t part_bytes = nsplit > 1 ? (((size_t)nsplit * ldw * std::min(N, 128) * 8 + 255) & ~size_t(255)) : 0;
  const size_t repack_bytes = repackX ? (((size_t)ldxr * N * 8 + 255) & ~size_t(255)) : 0;
  const int rgroups = nsplit > 1 ? reduce_groups(M, std::min(N, 128), nsplit) : 1;
  const size_t scratch_bytes = rgroups > 1 ? (size_t)(rgroups + 1) * M * std::min(N, 128) * 8 : 0;
  if (part_bytes || repackX) { cudaError_t e = ws.reserve(part_bytes + repack_bytes + scratch_bytes); if (e != cudaSuccess) return e; }
  double* scratch = reinterpret_cast<double*>(reinterpret_cast<char*>(ws.ptr) + part_bytes + repack_bytes);
  if (repackX) {
    double* Xr = reinterpret_cast<double*>(reinterpret_cast<char*>(ws.ptr) + part_bytes);
    k_repack<<<dim3((unsigned)((K + 255) / 256), (unsigned)std::min(N, 128)), 256, 0, st>>>(X, ldx, Xr, ldxr, K, N);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    X = Xr; ldx = ldxr;
  }
  for (int n0 = 0; n0 < N; n0 += 128) {
    const int nc = std::min(128, N - n0), NB = (nc + 7) / 8;
    CUtensorMap tX;
    if (!make_map(&tX, X + (size_t)n0 * ldx, K, nc, ldx, NB * 8)) return cudaErrorInvalidValue;
    GemmParams p{};
    p.M = (int)M; p.N = nc; p.K = (int)K; p.ntiles = ntiles; p.nsplit = nsplit; p.kchunk = kchunk; p.transpose_out = 0; p.sh0 = sh0; p.sh1 = sh1; p.Araw = A; p.lda = lda;
    double* yout = Y + (size_t)n0 * ldy;
    if (nsplit == 1) {
      p.out = yout; p.ld_out = ldy; p.split_stride = 0;
      p.vec_ok = ((ldy & 1) == 0 && (reinterpret_cast<uintptr_t>(yout) & 15) == 0) ? 1 : 0;
    } else {
      p.out = ws.ptr; p.ld_out = ldw; p.split_stride = ldw * nc; p.vec_ok = 1;
    }
    const int grid = std::min(nsm, ntiles * nsplit);
    cudaError_t e;
    if (splitA) { RSVDB_NB_SWITCH(launch_an, true, tA, tA1, tX, p, grid, st) } else { RSVDB_NB_SWITCH(launch_an, false, tA, tA1, tX, p, grid, st) }
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    if (nsplit > 1) {
      e = reduce_partials(st, ws.ptr, p.split_stride, nsplit, p.ld_out, yout, ldy, (int)M, nc, 0, scratch, launches);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

cudaError_t gemm_at(GemmWorkspace& ws, cudaStream_t st, int nsm, const double* A, long long K, long long M, long long lda,
                    const double* Q, long long ldq, int N, double* Z, long long ldz, int transpose_out, int* launches) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (K <= 0) {
    if (!transpose_out) { for (int n = 0; n < N; ++n) { cudaError_t e = cudaMemsetAsync(Z + (size_t)n * ldz, 0, (size_t)M * 8, st); if (e != cudaSuccess) return e; } }
    else { for (long long j = 0; j < M; ++j) { cudaError_t e = cudaMemsetAsync(Z + (size_t)j * ldz, 0, (size_t)N * 8, st); if (e != cudaSuccess) return e; } }
    return cudaSuccess;
  }
  if (M >= (1LL << 31) || K >= (1LL << 31)) return cudaErrorInvalidValue;
  const bool splitA = !tma_addressable(A, lda);
  const bool repackQ = !tma_addressable(Q, ldq);
  CUtensorMap tA, tA1; int sh0 = 0, sh1 = 0;
  if (splitA ? !make_split_maps(&tA, &tA1, &sh0, &sh1, A, K, M, lda, BM / 2) : !make_map(&tA, A, K, M, lda, BM)) {
    if (!splitA) return cudaErrorInvalidValue;
    note_generic_fallback();                                                   // a single column of A
    if (launches) ++*launches;
    if (!transpose_out) return gemm_generic(st, 1, 0, (int)M, N, (int)K, 1.0, A, lda, Q, ldq, 0.0, Z, ldz);
    return gemm_generic(st, 1, 0, N, (int)M, (int)K, 1.0, Q, ldq, A, lda, 0.0, Z, ldz);
  }
  if (!splitA) tA1 = tA; else ++g_split_products;
  const int ntiles = (int)((M + BM - 1) / BM), ksteps = (int)((K + BK - 1) / BK);
  int nsplit, kchunk; choose_split(ntiles, ksteps, nsm, M * std::min(N, 128), (std::min(N, 128) + 7) / 8, &nsplit, &kchunk);
  const long long ldqr = (K + 1) & ~1LL;
  const size_t part_bytes = nsplit > 1 ? (((size_t)nsplit * M * std::min(N, 128) * 8 + 255) & ~size_t(255)) : 0;
  const size_t repack_bytes = repackQ ? (((size_t)ldqr * N * 8 + 255) & ~size_t(255)) : 0;
  const int rgroups = nsplit > 1 ? reduce_groups(M, std::min(N, 128), nsplit) : 1;
  const size_t scratch_bytes = rgroups > 1 ? (size_t)(rgroups + 1) * M * std::min(N, 128) * 8 : 0;
  if (part_bytes || repackQ) { cudaError_t e = ws.reserve(part_bytes + repack_bytes + scratch_bytes); if (e != cudaSuccess) return e; }
  double* scratch = reinterpret_cast<double*>(reinterpret_cast<char*>(ws.ptr) + part_bytes + repack_bytes);
  if (repackQ) {
    double* Qr = reinterpret_cast<double*>(reinterpret_cast<char*>(ws.ptr) + part_bytes);
    k_repack<<<dim3((unsigned)((K + 255) / 256), (unsigned)std::min(N, 128)), 256, 0, st>>>(Q, ldq, Qr, ldqr, K, N);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    Q = Qr; ldq = ldqr;
  }
  for (int n0 = 0; n0 < N; n0 += 128) {
    const int nc = std::min(128, N - n0), NB = (nc + 7) / 8;
    CUtensorMap tQ;
    if (!make_map(&tQ, Q + (size_t)n0 * ldq, K, nc, ldq, NB * 8)) return cudaErrorInvalidValue;
    GemmParams p{};
    p.M = (int)M; p.N = nc; p.K = (int)K; p.ntiles = ntiles; p.nsplit = nsplit; p.kchunk = kchunk; p.vec_ok = 0; p.sh0 = sh0; p.sh1 = sh1; p.Araw = A; p.lda = lda;
    double* zout = transpose_out ? Z + n0 : Z + (size_t)n0 * ldz;
    if (nsplit == 1) {
      p.out = zout; p.ld_out = ldz; p.split_stride = 0; p.transpose_out = transpose_out;
    } else {
      p.out = ws.ptr; p.ld_out = M; p.split_stride = M * nc; p.transpose_out = 0;
    }
    const int grid = std::min(nsm, ntiles * nsplit);
    cudaError_t e;
    if (splitA) { RSVDB_NB_SWITCH(launch_at, true, tA, tA1, tQ, p, grid, st) } else { RSVDB_NB_SWITCH(launch_at, false, tA, tA1, tQ, p, grid, st) }
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    if (nsplit > 1) {
      e = reduce_partials(st, ws.ptr, p.split_stride, nsplit, p.ld_out, zout, ldz, (int)M, nc, transpose_out, scratch, launches);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

}  // namespace rsvdb
