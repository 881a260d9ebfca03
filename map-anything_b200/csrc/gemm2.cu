// CTA-pair (tcgen05 cta_group::2) variant of the persistent bf16 GEMM / implicit 3x3 conv:  Y = epilogue(X . W^T)
//
// A cluster of two CTAs (the two SMs of a TPC) owns one 256 x BN output tile.  CTA r loads A rows [m0 + 128 r, +128) and
// B rows [n0 + BN/2 * r, + BN/2) of every K block; ONE thread of the leader CTA issues tcgen05.mma.cta_group::2 with
// M = 256, N = BN: each SM's tensor core multiplies its own 128 A rows with all BN columns, reading the other half of
// B from the peer SM's shared memory.  Per MMA an SM therefore reads 128x16 A + BN/2 x16 B instead of 128x16 + BNx16:
// at BN = 256 that is 8 KB instead of 12 KB per 128 tensor-core cycles, which is what keeps the 1-CTA kernel (gemm.cu)
// below the tensor pipe's rate (shared-memory operand bandwidth, B300_MICROARCH "tcgen05 floor").
//
//   warp 0 (both CTAs) : TMA producer; transaction bytes of both CTAs are credited to the LEADER's full barrier
//   warp 1 (leader)    : MMA issuer; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs
//   warp 2 (both)      : TMEM allocator (cta_group::2)
//   warps 4..11 (both) : epilogue for the CTA's own 128 rows; "accumulator drained" arrives on the leader's barrier
//
// Stream-K tail (in-place fp32 residual GEMMs, i.e. the reduce-add epilogue, K >= 2048): the tiles of the last, partially
// filled wave are not handed out whole.  Their K blocks form one flattened range that is cut into equal contiguous pieces,
// one per cluster, so every cluster works until (nearly) the same K block: 10960 x 1024 x 4096 on 74 clusters is 172 tiles =
// 2 full waves + 24 tiles; whole tiles cost 3 x 64 K blocks per cluster, stream-K 2 x 64 + 21.  A piece that covers part of
// a tile's K range is an ordinary work unit (tile, k0, k1): its partial product leaves through the same bulk tensor
// REDUCE-ADD as a whole tile (x += gamma * acc; the unit that holds K block 0 also adds gamma * bias), so no workspace and
// no fix-up pass exist.  The fp32 adds of one element's two or three partial products reach the L2 in arbitrary order: such
// a GEMM is reproducible to fp32 rounding of the sum (~1e-7 relative), not bit for bit (ma_set_stream_k(0) restores that).
#include "gemm_common.cuh"

namespace ma {

// Work units of one cluster: whole tiles t = cluster, cluster + W, ... below sk_lo, then the cluster's piece of the
// flattened (tile, K block) range of the tiles [sk_lo, total_tiles).  Every warp role walks the same sequence.
struct UnitWalk {
  int t, W, sk_lo, kblocks;
  int pos, end;  // stream-K piece, in K blocks from the start of tile sk_lo
  __device__ UnitWalk(int cluster, int W_, int total_tiles, int sk_lo_, int kblocks_) : t(cluster), W(W_), sk_lo(sk_lo_), kblocks(kblocks_) {
    const int total = (total_tiles - sk_lo) * kblocks;
    const int piece = (total + W - 1) / W;
    pos = cluster * piece;
    end = pos + piece < total ? pos + piece : total;
  }
  __device__ bool next(int& tile, int& k0, int& k1) {
    if (t < sk_lo) {
      tile = t; k0 = 0; k1 = kblocks;
      t += W;
      return true;
    }
    if (pos >= end) return false;
    const int j = pos / kblocks;
    tile = sk_lo + j;
    k0 = pos - j * kblocks;
    k1 = k0 + (end - pos) < kblocks ? k0 + (end - pos) : kblocks;
    pos += k1 - k0;
    return true;
  }
};

constexpr int G2_THREADS = 384;
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_SK_MIN_KBLOCKS = 32;  // stream-K only for K >= 2048 (fc2 of the transformer blocks)
constexpr int G2_SK_MIN_PIECE = 8;     // ... and pieces of at least 8 K blocks

template <int BN>
struct Gemm2Cfg {
  static constexpr int STAGES = BN == 256 ? 6 : 8;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;   // this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;  // this CTA's half of the N rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 1024;  // padded: the epilogue stages behind it stay 1024-byte aligned
  static constexpr int EPI_BYTES = G2_EPI_WARPS * EPI_STAGE_BYTES;  // fused head: reused as hw[8][128] + partial sums
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ CUtensorMap tmap_out, const ma_gemm_epilogue ep, const int M, const int N, const int K,
                      const ConvGeom cg, const int epi_mode, const int sk_lo) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bar_empty = bar_full + STAGES;
  uint64_t* bar_tfull = bar_empty + STAGES;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  uint8_t* sEpi = smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES;  // 8 x 4 KB TMA-store stages (1024-byte aligned)

  const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
  const int lane = lane_id();
  const uint32_t cta = cluster_ctarank();  // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bar_full[s], 1);   // leader's arrive.expect_tx; bytes from both CTAs
      mbar_init(&bar_empty[s], 1);  // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_tfull[a], 1);                   // multicast tcgen05.commit
      mbar_init(&bar_tempty[a], 2 * G2_EPI_WARPS);   // epilogue warps of BOTH CTAs (used on the leader only)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote signal; TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // prologue above overlaps the previous kernel's tail; no global memory access before this point

  const int tiles_per_img = cg.tiles_x * cg.tiles_y;
  const int rows_tiles = cg.mode ? (M / (cg.H * cg.W)) * tiles_per_img : (M + GEMM_BM - 1) / GEMM_BM;  // 128-row tiles
  const int tiles_m = (rows_tiles + 1) / 2;                                                             // 256-row tiles
  const int tiles_n = (N + BN - 1) / BN;
  const int total_tiles = tiles_m * tiles_n;
  const int kblocks = cg.mode ? 9 * cg.cblocks : (K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      UnitWalk walk(cluster_id, num_clusters, total_tiles, sk_lo, kblocks);
      int t, k0, k1;
      while (walk.next(t, k0, k1)) {
        int tm2, tn;
        raster_tile(t, tiles_n, num_clusters, tm2, tn);
        const int tm = tm2 * 2 + static_cast<int>(cta);  // this CTA's 128-row tile
        const int n0 = tn * BN + static_cast<int>(cta) * (BN / 2);
        if (cg.mode) {
          // a row tile past the end (odd tile count) reads image index n: fully out of bounds -> zero fill
          const int img = tm / tiles_per_img;
          const int r = tm - img * tiles_per_img;
          const int y0 = (r / cg.tiles_x) * cg.bh, x0 = (r % cg.tiles_x) * cg.bw;
          const uint32_t bytes = 2u * (static_cast<uint32_t>(cg.bw * cg.bh * 128) + Cfg::B_BYTES);
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            for (int cb = 0; cb < cg.cblocks; ++cb) {
              mbar_wait(&bar_empty[stage], phase ^ 1);
              if (cta == 0) mbar_arrive_expect_tx(&bar_full[stage], bytes);
              tma_load_4d_2sm(sA + stage * Cfg::A_BYTES, &tmap_x, &bar_full[stage], cb * GEMM_BK, x0 + dx, y0 + dy, img);
              tma_load_2d_2sm(sB + stage * Cfg::B_BYTES, &tmap_w, &bar_full[stage], tap * cg.C + cb * GEMM_BK, n0);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          const int m0 = tm * GEMM_BM;
          for (int kb = k0; kb < k1; ++kb) {
            mbar_wait(&bar_empty[stage], phase ^ 1);
            if (cta == 0) mbar_arrive_expect_tx(&bar_full[stage], 2u * Cfg::STAGE_BYTES);
            tma_load_2d_2sm(sA + stage * Cfg::A_BYTES, &tmap_x, &bar_full[stage], kb * GEMM_BK, m0);
            tma_load_2d_2sm(sB + stage * Cfg::B_BYTES, &tmap_w, &bar_full[stage], kb * GEMM_BK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (cta == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      UnitWalk walk(cluster_id, num_clusters, total_tiles, sk_lo, kblocks);
      int t, k0, k1;
      for (int it = 0; walk.next(t, k0, k1); ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait_cluster(&bar_tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16_ss_2sm(d_tmem, adesc, bdesc, idesc, (kb != k0 || k != 0) ? 1u : 0u);
          }
          umma_commit_2sm(&bar_empty[stage], 0x3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&bar_tfull[acc], 0x3);
      }
    }
  } else if (warp >= 4) {
    uint8_t* epi_stage = sEpi + (warp - 4) * EPI_STAGE_BYTES;
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t leader_tempty0 = mapa_shared(smem_u32(&bar_tempty[0]), 0);
    // fused narrow head (BN == 128 only): the epilogue stages are free (no TMA epilogue in this mode): stage 0 .. 3 hold
    // the head weights hw[8][128] fp32 (4 KB), stages 4 .. 7 the partial sums of the upper column half, [tile parity][row][8]
    const bool head = BN == 128 && ep.head_out != nullptr;
    float* hw = reinterpret_cast<float*>(sEpi);
    float* hpart = reinterpret_cast<float*>(sEpi + 4 * EPI_STAGE_BYTES);  // [2][128][8] fp32 = 8 KB
    if (head) {
      for (int i = threadIdx.x - 128; i < 8 * BN; i += G2_EPI_WARPS * 32) hw[i] = __ldg(ep.head_w + i);
      named_bar_sync(1, G2_EPI_WARPS * 32);
    }
    UnitWalk walk(cluster_id, num_clusters, total_tiles, sk_lo, kblocks);
    int t, k0, k1;
    for (int it = 0; walk.next(t, k0, k1); ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int tm2, tn;
      raster_tile(t, tiles_n, num_clusters, tm2, tn);
      const int tm = tm2 * 2 + static_cast<int>(cta);
      const int n0 = tn * BN;
      int m;
      bool row_ok;
      if (cg.mode) {
        const int img = tm / tiles_per_img;
        const int rr = tm - img * tiles_per_img;
        const int r = quarter * 32 + lane;
        const int y = (rr / cg.tiles_x) * cg.bh + r / cg.bw, x = (rr % cg.tiles_x) * cg.bw + r % cg.bw;
        row_ok = tm < rows_tiles && r < cg.bw * cg.bh && y < cg.H && x < cg.W;
        m = (img * cg.H + y) * cg.W + x;
      } else {
        m = tm * GEMM_BM + quarter * 32 + lane;
        row_ok = m < M;
      }
      mbar_wait(&bar_tfull[acc], acc_phase);
      tc_fence_after();
      float hsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 32) {
        const int col0 = n0 + half * (BN / 2) + c;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * (BN / 2) + c;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        if (head) {
          epilogue_head_chunk(ep, v, col0, BN, hw, hsum);
        } else if (epi_mode) {  // warp-uniform: asynchronous bulk tensor store / reduce-add of the 32 x 32 chunk
          // a stream-K unit that does not hold K block 0 contributes gamma * partial product only (no bias)
          if (col0 < N) epilogue_tma_chunk(&tmap_out, epi_mode, ep, v, tm * GEMM_BM + quarter * 32, col0, epi_stage, lane, k0 == 0);
        } else if (row_ok && col0 < N) {
          epilogue_store_chunk(ep, v, m, col0, N);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(leader_tempty0 + acc * 8);  // the accumulator is drained: the next tile's MMAs may start
      if (head) {
        // the two warps that own the same 32 rows (column halves 0 / 1) meet at a 64-thread named barrier: the upper half
        // hands its 8 partial sums over through shared memory (double-buffered by tile parity), the lower half adds the
        // head bias and writes 8 fp32 per pixel (consecutive rows -> 1 KB contiguous per warp)
        float* mine = hpart + ((it & 1) * 128 + quarter * 32 + lane) * 8;
        if (half == 1) {
          *reinterpret_cast<float4*>(mine) = make_float4(hsum[0], hsum[1], hsum[2], hsum[3]);
          *reinterpret_cast<float4*>(mine + 4) = make_float4(hsum[4], hsum[5], hsum[6], hsum[7]);
        }
        named_bar_sync(2 + quarter, 64);
        if (half == 0 && row_ok) {
          const float4 a = *reinterpret_cast<const float4*>(mine), b = *reinterpret_cast<const float4*>(mine + 4);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.head_bias)), b1 = __ldg(reinterpret_cast<const float4*>(ep.head_bias) + 1);
          float4* o = reinterpret_cast<float4*>(ep.head_out + static_cast<size_t>(m) * 8);
          o[0] = make_float4(hsum[0] + a.x + b0.x, hsum[1] + a.y + b0.y, hsum[2] + a.z + b0.z, hsum[3] + a.w + b0.w);
          o[1] = make_float4(hsum[4] + b.x + b1.x, hsum[5] + b.y + b1.y, hsum[6] + b.z + b1.z, hsum[7] + b.w + b1.w);
        }
      }
    }
    if (epi_mode && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's smem / TMEM stay alive until the leader's last MMA and all remote arrives are done
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_gemm2(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tout, int epi_mode,
                        const ma_gemm_epilogue& ep, int M, int N, int K, cudaStream_t stream, const ConvGeom& cg) {
  using Cfg = Gemm2Cfg<BN>;
  static bool configured_dev[MA_MAX_DEVICES] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    MA_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_2cta_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int rows_tiles = cg.mode ? (M / (cg.H * cg.W)) * cg.tiles_x * cg.tiles_y : (M + GEMM_BM - 1) / GEMM_BM;
  const int tiles = ((rows_tiles + 1) / 2) * ((N + BN - 1) / BN);
  const int max_clusters = device_sm_count() / 2;
  int clusters = tiles < max_clusters ? tiles : max_clusters;
  // stream-K over the last partial wave: only where a partial product can leave as a reduce-add (epi_mode 2, no activation)
  // and the K range is long enough for a piece to amortise its own epilogue
  int sk_lo = tiles;
  const int kblocks = (K + GEMM_BK - 1) / GEMM_BK;
  if (stream_k_enabled() && epi_mode == 2 && !cg.mode && ep.act == MA_ACT_NONE && kblocks >= G2_SK_MIN_KBLOCKS &&
      tiles % max_clusters != 0) {
    const int tail = tiles % max_clusters;
    if ((tail * kblocks + max_clusters - 1) / max_clusters >= G2_SK_MIN_PIECE) {
      sk_lo = tiles - tail;
      clusters = max_clusters;
    }
  }
  MA_CHECK_CUDA(launch_kernel(gemm_bf16_2cta_kernel<BN>, dim3(2 * clusters), dim3(G2_THREADS), Cfg::SMEM_BYTES, stream, pdl_enabled(), tx,
                              tw, tout, ep, M, N, K, cg, epi_mode, sk_lo));
  return MA_OK;
}

// Entry used by gemm.cu's dispatchers. bn2 in {128, 256}.
int launch_gemm_2cta(int bn2, const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tout, int epi_mode,
                     const ma_gemm_epilogue& ep, int M, int N, int K, cudaStream_t stream, const ConvGeom& cg) {
  if (bn2 == 256) return launch_gemm2<256>(tx, tw, tout, epi_mode, ep, M, N, K, stream, cg);
  return launch_gemm2<128>(tx, tw, tout, epi_mode, ep, M, N, K, stream, cg);
}

}  // namespace ma
