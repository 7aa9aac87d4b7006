// Thread-block-cluster TSQR nodes (included by tsqr.cu, inside namespace rsvdb::{anonymous}).
//
// An upper tree level that factors 256-row blocks shrinks the R stack by only 256/l per level (2.56 at l = 100), and each
// level costs one full chain of l dependent reflector steps.  A cluster node spreads CL = 4 CTAs over 1024 rows: every
// CTA keeps its 256 rows in its own shared memory, and the three reductions of the blocked Householder step -- the
// per-reflector column products, the Gram matrix for T, and W = V^T C of the trailing update -- are summed across the
// cluster through distributed shared memory (remote st.shared::cluster for the 16-double reflector payload, remote
// ld.shared::cluster for the 8 x 8 partials) around hardware cluster barriers.  Fan-in becomes 1024/l, i.e. the tree over
// a 25000-row shard has 3 levels instead of 7.  Sums are taken in CTA-rank order on every CTA, so all CTAs (and all
// GPUs) hold bit-identical T, tau and W.
#pragma once

constexpr int CL = 4;                               // CTAs per cluster node of the upper tree levels
constexpr int CL_ROWS = CL * 256;
constexpr int CL0 = 8;                              // CTAs per cluster node at level 0 of panels of <= 16 such nodes (round 2): 2048 rows per node,
constexpr int CL0_ROWS = CL0 * 256;                 // i.e. one tree level fewer for the 20000- / 25000-row panels of the 8-GPU run

__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  __syncwarp();   // the .aligned forms need the whole warp converged
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned map_cluster(const void* p, unsigned rank) {
  unsigned la = static_cast<unsigned>(__cvta_generic_to_shared(p)), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  return ra;
}
__device__ __forceinline__ void st_cluster(double* p, unsigned rank, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(map_cluster(p, rank)), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_cluster(const double* p, unsigned rank) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(map_cluster(p, rank)) : "memory");
  return v;
}

// Remote store + completion signal in one operation: the 8-byte value lands in CTA `rank`'s copy of *p and 8 bytes are
// credited to that CTA's copy of *bar (complete_tx).  Lets the panel teams of a cluster exchange their per-reflector partial
// sums without a hardware cluster barrier (which every thread of every CTA would have to join).
__device__ __forceinline__ void st_async_cluster(double* p, unsigned rank, double v, uint64_t* bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(map_cluster(p, rank)), "l"(__double_as_longlong(v)), "r"(map_cluster(bar, rank)) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// clean reflector value V[grow][r] of the panel at column / node row c0; grow = node row index of local row `lrow`
__device__ __forceinline__ double clean_v_cl(const double* S, const double* Vtop, int c0, int grow, int lrow, int r) {
  constexpr int LDS = 256 + 4;
  if (grow < c0) return 0.0;
  if (grow < c0 + 8) return Vtop[r * 8 + (grow - c0)];
  return S[(size_t)(c0 + r) * LDS + lrow];
}

// shared-memory scratch of a cluster node (doubles): see the carve-up in the kernels
//   red 64 | drow 8 | xch 2*16*CL | Gtot 64 | xw 16*64 (aliases the per-warp Gram partials 8*64 and the CTA Gram partial 64)
template <int CLT> struct Cls { static constexpr int RED = 0, DROW = 64, XCH = 72, GTOT = XCH + 2 * 16 * CLT, XW = GTOT + 64, TOTAL = XW + 16 * 64; };

// All CTAs of the cluster factor the panel [c0, c0+pb) together.  Every thread of every CTA must call this (non-row
// warps only take part in the cluster barriers).  Outputs as panel_factor_la; T / tau are identical in all CTAs.
// Per reflector the 8 column products (+ the diagonal row, from rank 0) of every CTA go to every CTA with st.async and are
// awaited on a local mbarrier by the row team only (xbar[2], alternating per reflector; xphase tracks their parities): ncu
// showed barrier.cluster per reflector -- 2048 threads arriving 100 times per node -- as the top stall of the round-1 kernel.
template <int CLT>
__device__ __forceinline__ void panel_factor_cl(double* S, double* Vtop, double* Tsm, double* tau_s, double* Tglob, double* sc,
                                                int c0, int pb, int warp, int lane, unsigned rank, uint64_t* xbar, uint32_t& xphase) {
  constexpr int LDS = 256 + 4;
  double* red = sc + Cls<CLT>::RED; double* drow = sc + Cls<CLT>::DROW; double* xch = sc + Cls<CLT>::XCH; double* Gtot = sc + Cls<CLT>::GTOT;
  double* Gs = sc + Cls<CLT>::XW;                // [8 warps][64] during the Gram step
  double* Gcta = sc + Cls<CLT>::XW + 8 * 64;     // [64] this CTA's Gram partial (read remotely)
  const bool roww = warp < BQ_ROW_WARPS;
  const int i = roww ? 32 * warp + lane : 0;
  const int gi = (int)rank * 256 + i;            // node row index
  double a[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = (c < pb && roww) ? S[(size_t)(c0 + c) * LDS + i] : 0.0;
  double tau[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    tau[r] = 0.0;
    if (r < pb) {
      const int d = c0 + r;
      const int b = r & 1;
      if (roww) {
        double p[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) p[c] = (c >= r && gi > d) ? a[r] * a[c] : 0.0;
        {
          int colr; const double tcol = reduce8_transposed(p, lane, &colr);
          if ((lane & 3) == 0) red[colr * 8 + warp] = tcol;
        }
        if (gi == d) {
#pragma unroll
          for (int c = 0; c < 8; ++c) drow[c] = a[c];
        }
        bar_rows();
        if (warp == 0 && lane == 0) mbar_expect_tx(&xbar[b], CLT * 16 * 8);   // this round: 16 doubles from each CTA of the cluster
        if (warp == 0 && lane < 16) {            // CTA partial (8 column products) + the diagonal row (from rank 0) to every CTA
          double v;
          if (lane < 8) { v = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += red[lane * 8 + w];
          } else v = (rank == 0) ? drow[lane - 8] : 0.0;
          double* slot = xch + b * 16 * CLT + lane * CLT + rank;
#pragma unroll
          for (unsigned tr = 0; tr < CLT; ++tr) st_async_cluster(slot, tr, v, &xbar[b]);
        }
        mbar_wait_cluster(&xbar[b], (xphase >> b) & 1u);
        xphase ^= (1u << b);
      }
      if (!roww) continue;
      // 16 slots x CLT rank partials: each lane sums CLT/4 double2 of one slot, one butterfly step completes the
      // slot (lanes 2s, 2s+1), broadcasts hand it to every row thread; same order on every CTA of the cluster.
      double tot[8], dr[8];
      {
        const double2* q = reinterpret_cast<const double2*>(xch + b * 16 * CLT) + lane * (CLT / 4);
        double s2 = 0.0;
#pragma unroll
        for (int h = 0; h < CLT / 4; ++h) { const double2 qq = q[h]; s2 += qq.x + qq.y; }
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          tot[c] = (c >= r) ? __shfl_sync(0xffffffffu, s2, 2 * c) : 0.0;
          dr[c] = (c >= r) ? __shfl_sync(0xffffffffu, s2, 2 * (8 + c)) : 0.0;     // only rank 0 sends non-zero
        }
      }
      const double tail = tot[r], x0 = dr[r];
      double beta, scale;
      if (tail <= DBL_MIN) { tau[r] = 0.0; beta = x0; scale = 0.0; }
      else {
        const double n2 = fma(x0, x0, tail);
        const double inrm = rsqrt(n2);
        const double nrm = n2 * inrm;
        const double ax = fabs(x0);
        beta = (x0 >= 0.0) ? -nrm : nrm;
        tau[r] = fma(ax, inrm, 1.0);
        const double rc = __drcp_rn(ax + nrm);
        scale = (x0 >= 0.0) ? rc : -rc;
      }
      const double v = (gi > d) ? a[r] * scale : (gi == d ? 1.0 : 0.0);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c > r && c < pb) {
          const double w = tau[r] * fma(scale, tot[c], dr[c]);
          a[c] = fma(-w, v, a[c]);
        }
      }
      if (gi > d) a[r] = v; else if (gi == d) a[r] = beta;
    }
  }
  if (roww) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < pb) S[(size_t)(c0 + c) * LDS + i] = a[c];
      if (gi >= c0 && gi < c0 + 8) {
        const int d = c0 + c;
        Vtop[c * 8 + (gi - c0)] = (c < pb) ? ((gi > d) ? a[c] : (gi == d ? 1.0 : 0.0)) : 0.0;
      }
    }
    __syncwarp();
    const int g = lane >> 2, t = lane & 3;
    double acc[2] = {0.0, 0.0};
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 4) {
      const int lrow = 32 * warp + k0 + t;
      const double x = (g < pb) ? clean_v_cl(S, Vtop, c0, (int)rank * 256 + lrow, lrow, g) : 0.0;
      dmma884(acc, x, x);
    }
    Gs[warp * 64 + g + 8 * (2 * t)] = acc[0];
    Gs[warp * 64 + g + 8 * (2 * t + 1)] = acc[1];
    bar_rows();
    if (threadIdx.x < 64) {
      double s2 = 0.0;
#pragma unroll
      for (int w = 0; w < BQ_ROW_WARPS; ++w) s2 += Gs[w * 64 + threadIdx.x];
      Gcta[threadIdx.x] = s2;
    }
  }
  cluster_sync_all();
  if (threadIdx.x < 64) {
    double s2 = 0.0;
#pragma unroll
    for (unsigned sr = 0; sr < CLT; ++sr) s2 += ld_cluster(Gcta + threadIdx.x, sr);
    Gtot[threadIdx.x] = s2;
  }
  cluster_sync_all();                               // remote reads of Gcta are done before anyone reuses the xw region
  if (threadIdx.x < 8) {
    const int srow = threadIdx.x;
    double trow[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      double val = 0.0;
      if (c == srow) val = tau[c];
      else if (c > srow) {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r < c && r >= srow) acc = fma(trow[r], Gtot[r + 8 * c], acc);
        val = -tau[c] * acc;
      }
      trow[c] = val;
      Tsm[srow + 8 * c] = val;
      if (Tglob) Tglob[srow + 8 * c] = val;
    }
    double tv = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c == srow) tv = tau[c];
    if (srow < pb) tau_s[c0 + srow] = tv;
  }
}

// Block reflector on the columns [col_begin, col_end) of this CTA's rows, W = V^T C summed over the cluster.
// Reflectors come either from S in place (Vp == nullptr: factor kernel; + Vtop for the panel's first 8 node rows) or from
// a staged clean panel Vp (apply kernel).  One group of 8 columns per warp (<= 16 groups).  Every thread must call it.
template <int CLT>
__device__ __forceinline__ void block_reflect_cl(double* S, const double* Vp, const double* Vtop, const double* Tsm, double* xw, int c0,
                                                 int col_begin, int col_end, int warp, int lane, unsigned rank, bool transposeT,
                                                 bool second_barrier) {
  constexpr int LDS = 256 + 4;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = col_begin + 8 * warp;
  const bool have = n0 < col_end;
  const int row_lo = (rank == 0) ? c0 : 0;       // first local row that the panel touches (c0 is a multiple of 8)
  const int roff = (int)rank * 256;
  double w0 = 0.0, w1 = 0.0;
  const int colB = n0 + perm8(g);
  const bool bval = have && colB < col_end;
  if (have) {
    double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    const double* cb = S + (size_t)(bval ? colB : col_begin) * LDS + t;
    const int nk = (256 - row_lo) >> 2;
    for (int k = 0; k < nk; k += 4) {
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = k + u; const bool on = kk < nk; const int lrow = row_lo + 4 * kk + t;
        av[u] = on ? (Vp ? Vp[(size_t)g * LDS + lrow] : clean_v_cl(S, Vtop, c0, roff + lrow, lrow, g)) : 0.0;
        bv[u] = (on && bval) ? cb[row_lo + 4 * kk] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) dmma884(acc[u], av[u], bv[u]);
    }
    w0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
    w1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
    xw[warp * 64 + g + 8 * (2 * t)] = w0;
    xw[warp * 64 + g + 8 * (2 * t + 1)] = w1;
  }
  cluster_sync_all();
  if (have) {
    w0 = 0.0; w1 = 0.0;
#pragma unroll
    for (unsigned sr = 0; sr < CLT; ++sr) {
      w0 += ld_cluster(xw + warp * 64 + g + 8 * (2 * t), sr);
      w1 += ld_cluster(xw + warp * 64 + g + 8 * (2 * t + 1), sr);
    }
  }
  if (second_barrier) cluster_sync_all();           // callers that rewrite xw before another cluster barrier need this
  if (!have) return;
  const double bw0 = frag_c_to_b(w0, w1, t, g), bw1 = frag_c_to_b(w0, w1, t + 4, g);
  const double t0 = transposeT ? Tsm[t + 8 * g] : Tsm[g + 8 * t];
  const double t1 = transposeT ? Tsm[(t + 4) + 8 * g] : Tsm[g + 8 * (t + 4)];
  double x[2] = {0.0, 0.0};
  dmma884(x, t0, bw0);
  dmma884(x, t1, bw1);
  const double bx0 = -frag_c_to_b(x[0], x[1], t, g), bx1 = -frag_c_to_b(x[0], x[1], t + 4, g);
  const int col0 = n0 + perm8(2 * t), col1 = n0 + perm8(2 * t + 1);
  const bool v0 = col0 < col_end, v1 = col1 < col_end;
  double* p0 = S + (size_t)(v0 ? col0 : col_begin) * LDS + g;
  double* p1 = S + (size_t)(v1 ? col1 : col_begin) * LDS + g;
  const int nblk = (256 - row_lo) >> 3;
  for (int b0 = 0; b0 < nblk; b0 += 4) {
    double c[4][2], x0[4], x1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int bb = b0 + u; const bool on = bb < nblk; const int lr = row_lo + 8 * bb;
      c[u][0] = (on && v0) ? p0[lr] : 0.0; c[u][1] = (on && v1) ? p1[lr] : 0.0;
      x0[u] = on ? (Vp ? Vp[(size_t)t * LDS + lr + g] : clean_v_cl(S, Vtop, c0, roff + lr + g, lr + g, t)) : 0.0;
      x1[u] = on ? (Vp ? Vp[(size_t)(t + 4) * LDS + lr + g] : clean_v_cl(S, Vtop, c0, roff + lr + g, lr + g, t + 4)) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) dmma884(c[u], x0[u], bx0);
#pragma unroll
    for (int u = 0; u < 4; ++u) dmma884(c[u], x1[u], bx1);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int bb = b0 + u; const bool on = bb < nblk; const int lr = row_lo + 8 * bb;
      if (on && v0) p0[lr] = c[u][0];
      if (on && v1) p1[lr] = c[u][1];
    }
  }
}

// Factor one 1024-row node per cluster.  Same outputs as k_house_factor_la: reflectors overwrite Y, tau / T / R per node.
template <int CLT>
__global__ void __cluster_dims__(CLT, 1, 1) __launch_bounds__(BQ_THREADS, 1)
k_node_factor_cl(double* __restrict__ Y, long long ldy, long long rows, int l, double* __restrict__ tau_g,
                 double* __restrict__ Rstack, long long ldr, double* __restrict__ Tg) {
  constexpr int LDS = 256 + 4;
  extern __shared__ double sm[];
  double* S = sm;
  double* Vtop = S + (size_t)l * LDS;              // 64
  double* Tsm = Vtop + 64;                         // 64
  double* sc = Tsm + 64;                           // CLS_TOTAL
  double* tau_s = sc + Cls<CLT>::TOTAL;
  uint64_t* xbar = reinterpret_cast<uint64_t*>(tau_s + l);     // 2 mbarriers of the per-reflector exchange
  uint32_t xphase = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned rank = cluster_rank();
  const int node = blockIdx.x / CLT;
  const long long r0 = (long long)node * (CLT * 256) + (long long)rank * 256;
  const int nrows = (int)max(0LL, min(256LL, rows - r0));
  const int npanels = (l + 7) / 8;

  for (int k = warp; k < l; k += BQ_WARPS) {
    const double* src = Y + (size_t)k * ldy + r0;
    for (int i = lane; i < 256; i += 32) S[(size_t)k * LDS + i] = (i < nrows) ? src[i] : 0.0;
  }
  if (threadIdx.x == 0) { mbar_init(&xbar[0], 1); mbar_init(&xbar[1], 1); fence_mbar_init(); }
  __syncthreads();
  cluster_sync_all();                               // every CTA's mbarriers exist before any peer signals them
  double* Tblock = (rank == 0) ? Tg + (size_t)node * npanels * 64 : nullptr;
  for (int p = 0; p < npanels; ++p) {
    const int c0 = 8 * p, pb = min(8, l - c0);
    panel_factor_cl<CLT>(S, Vtop, Tsm, tau_s, Tblock ? Tblock + (size_t)p * 64 : nullptr, sc, c0, pb, warp, lane, rank, xbar, xphase);
    __syncthreads();
    if (c0 + pb < l) block_reflect_cl<CLT>(S, nullptr, Vtop, Tsm, sc + Cls<CLT>::XW, c0, c0 + pb, l, warp, lane, rank, true, false);
    __syncthreads();
  }
  for (int k = warp; k < l; k += BQ_WARPS) {
    double* dst = Y + (size_t)k * ldy + r0;
    for (int i = lane; i < 256; i += 32) {
      const double val = S[(size_t)k * LDS + i];
      if (i < nrows) dst[i] = val;
      if (rank == 0 && i < l) Rstack[(size_t)k * ldr + (size_t)node * l + i] = (i <= k) ? val : 0.0;
    }
  }
  if (rank == 0) for (int j = threadIdx.x; j < l; j += BQ_THREADS) tau_g[(size_t)node * l + j] = tau_s[j];
  cluster_sync_all();                               // no CTA may exit while a peer can still read its shared memory
}

// Form Q for cluster nodes: Q_node = H_0 ... H_{l-1} [C; 0], C = rows [node*l, (node+1)*l) of Ctop (or the identity).
template <int CLT>
__global__ void __cluster_dims__(CLT, 1, 1) __launch_bounds__(BQ_THREADS, 1)
k_node_apply_cl(const ApplyTable tab, int l) {
  constexpr int LDS = 256 + 4;
  constexpr int VPT = 8 * 256 / BQ_THREADS;
  extern __shared__ double sm[];
  double* S = sm;
  double* Vp = S + (size_t)l * LDS;
  double* Tsm = Vp + (size_t)8 * LDS;
  double* xw = Tsm + 64;                            // 16 * 64
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned rank = cluster_rank();
  int li = 0;
  while (li + 1 < tab.n && (int)blockIdx.x >= tab.lv[li + 1].first_block) ++li;
  const double* V = tab.lv[li].V; const long long ldv = tab.lv[li].ldv; const long long rows = tab.lv[li].rows;
  const double* Tg = tab.lv[li].Tg; const double* Ctop = tab.lv[li].Ctop; const long long ldc = tab.lv[li].ldc;
  double* Q = tab.lv[li].Q; const long long ldq = tab.lv[li].ldq;
  const int node = ((int)blockIdx.x - tab.lv[li].first_block) / CLT;
  const long long r0 = (long long)node * (CLT * 256) + (long long)rank * 256;
  const int nrows = (int)max(0LL, min(256LL, rows - r0));
  const int npanels = (l + 7) / 8;
  const int roff = (int)rank * 256;

  for (int k = warp; k < l; k += BQ_WARPS) {
    for (int i = lane; i < 256; i += 32) {
      double c = 0.0;
      if (rank == 0 && i < l) c = Ctop ? Ctop[(size_t)k * ldc + (size_t)node * l + i] : (i == k ? 1.0 : 0.0);
      S[(size_t)k * LDS + i] = c;
    }
  }
  const double* Tblock = Tg + (size_t)node * npanels * 64;
  double vreg[VPT];
  auto fetch = [&](int p) {
    const int c0 = 8 * p, pb = min(8, l - c0);
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      const int e = threadIdx.x + q * BQ_THREADS; const int c = e / 256, i = e % 256;
      const int d = c0 + c, gi = roff + i;
      double v = 0.0;
      if (c < pb) { if (gi > d) v = (i < nrows) ? V[(size_t)d * ldv + r0 + i] : 0.0; else if (gi == d) v = 1.0; }
      vreg[q] = v;
    }
  };
  fetch(npanels - 1);
  for (int p = npanels - 1; p >= 0; --p) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      const int e = threadIdx.x + q * BQ_THREADS; const int c = e / 256, i = e % 256;
      Vp[(size_t)c * LDS + i] = vreg[q];
    }
    if (threadIdx.x < 64) Tsm[threadIdx.x] = Tblock[(size_t)p * 64 + threadIdx.x];
    if (p > 0) fetch(p - 1);
    __syncthreads();
    block_reflect_cl<CLT>(S, Vp, nullptr, Tsm, xw, 8 * p, 0, l, warp, lane, rank, false, true);
  }
  __syncthreads();
  for (int k = warp; k < l; k += BQ_WARPS) {
    double* dst = Q + (size_t)k * ldq + r0;
    for (int i = lane; i < nrows; i += 32) dst[i] = S[(size_t)k * LDS + i];
  }
  cluster_sync_all();
}

inline size_t cl_factor_smem(int l, int clt = CL) { return ((size_t)l * 260 + 128 + (72 + 32 * clt + 64 + 16 * 64) + (size_t)l + 2) * sizeof(double); }
inline size_t cl_apply_smem(int l) { return ((size_t)(l + 8) * 260 + 64 + (size_t)((l + 7) / 8) * 64) * sizeof(double); }
