// Persistent, warp-specialised bf16 GEMM for sm_100a:  Y = epilogue(X . W^T)
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B swizzle, STAGES-deep mbarrier ring)
//   warp 1      : MMA issuer    (one thread, tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16)
//   warp 2      : TMEM allocator / deallocator
//   warps 4..11 : epilogue      (tcgen05.ld TMEM -> registers -> fused bias/act/scale/residual -> global)
//
// Two fp32 accumulators (2*BN TMEM columns) are double-buffered so the epilogue of tile i overlaps the
// MMA main loop of tile i+1.  Tiles are scheduled statically, M fastest, so CTAs running at the same
// time share the same weight tile in L2.  Ragged M / N / K edges are handled by TMA out-of-bounds
// zero fill on the load side and predicated stores on the store side.
#include <stdlib.h>

#include "gemm_common.cuh"

namespace ma {

constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARPS = 8;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 1024;  // barrier block, padded so that the epilogue stages stay 1024-byte aligned
  static constexpr int EPI_BYTES = GEMM_EPI_WARPS * EPI_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_BYTES + 1024;  // +1024: manual alignment slack
  static constexpr int TMEM_COLS = 2 * BN;                                   // 128 / 256 / 512: powers of two
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_out, const ma_gemm_epilogue ep, const int M, const int N, const int K,
                      const ConvGeom cg, const int epi_mode) {
  using Cfg = GemmCfg<BN>;
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

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_tfull[a], 1);
      mbar_init(&bar_tempty[a], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // prologue above overlaps the previous kernel's tail; no global memory access before this point

  const int tiles_per_img = cg.tiles_x * cg.tiles_y;
  const int tiles_m = cg.mode ? (M / (cg.H * cg.W)) * tiles_per_img : (M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int total_tiles = tiles_m * tiles_n;
  const int kblocks = cg.mode ? 9 * cg.cblocks : (K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int tm, tn;
        raster_tile(t, tiles_n, static_cast<int>(gridDim.x), tm, tn);
        const int n0 = tn * BN;
        if (cg.mode) {
          const int img = tm / tiles_per_img;
          const int r = tm - img * tiles_per_img;
          const int y0 = (r / cg.tiles_x) * cg.bh, x0 = (r % cg.tiles_x) * cg.bw;
          const uint32_t bytes = static_cast<uint32_t>(cg.bw * cg.bh * 128) + Cfg::B_BYTES;
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            for (int cb = 0; cb < cg.cblocks; ++cb) {
              mbar_wait(&bar_empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&bar_full[stage], bytes);
              tma_load_4d(sA + stage * Cfg::A_BYTES, &tmap_x, &bar_full[stage], cb * GEMM_BK, x0 + dx, y0 + dy, img);
              tma_load_2d(sB + stage * Cfg::B_BYTES, &tmap_w, &bar_full[stage], tap * cg.C + cb * GEMM_BK, n0);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          const int m0 = tm * GEMM_BM;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&bar_empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&bar_full[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sA + stage * Cfg::A_BYTES, &tmap_x, &bar_full[stage], kb * GEMM_BK, m0);
            tma_load_2d(sB + stage * Cfg::B_BYTES, &tmap_w, &bar_full[stage], kb * GEMM_BK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bar_tfull[acc]);
      }
    }
  } else if (warp >= 4) {
    uint8_t* epi_stage = sEpi + (warp - 4) * EPI_STAGE_BYTES;
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;   // which half of the BN columns this warp drains
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int tm, tn;
      raster_tile(t, tiles_n, static_cast<int>(gridDim.x), tm, tn);
      const int n0 = tn * BN;
      // output row of this thread's accumulator lane: tile row r = quarter*32 + lane
      int m;
      bool row_ok;
      if (cg.mode) {
        const int img = tm / tiles_per_img;
        const int rr = tm - img * tiles_per_img;
        const int r = quarter * 32 + lane;
        const int y = (rr / cg.tiles_x) * cg.bh + r / cg.bw, x = (rr % cg.tiles_x) * cg.bw + r % cg.bw;
        row_ok = r < cg.bw * cg.bh && y < cg.H && x < cg.W;
        m = (img * cg.H + y) * cg.W + x;
      } else {
        m = tm * GEMM_BM + quarter * 32 + lane;
        row_ok = m < M;
      }
      mbar_wait(&bar_tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 32) {
        const int col0 = n0 + half * (BN / 2) + c;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * (BN / 2) + c;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        if (epi_mode) {  // warp-uniform: asynchronous bulk tensor store / reduce-add of the 32 x 32 chunk
          if (col0 < N) epilogue_tma_chunk(&tmap_out, epi_mode, ep, v, tm * GEMM_BM + quarter * 32, col0, epi_stage, lane);
        } else if (row_ok && col0 < N) {
          epilogue_store_chunk(ep, v, m, col0, N);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[acc]);
    }
    if (epi_mode && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_gemm(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tout, int epi_mode,
                       const ma_gemm_epilogue& ep, int M, int N, int K, cudaStream_t stream, const ConvGeom& cg = ConvGeom{}) {
  using Cfg = GemmCfg<BN>;
  static bool configured_dev[MA_MAX_DEVICES] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    MA_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles_m = cg.mode ? (M / (cg.H * cg.W)) * cg.tiles_x * cg.tiles_y : (M + GEMM_BM - 1) / GEMM_BM;
  const int tiles = tiles_m * ((N + BN - 1) / BN);
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  MA_CHECK_CUDA(launch_kernel(gemm_bf16_tcgen05_kernel<BN>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, pdl_enabled(), tx, tw,
                              tout, ep, M, N, K, cg, epi_mode));
  return MA_OK;
}

int launch_gemm_2cta(int bn2, const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tout, int epi_mode,
                     const ma_gemm_epilogue& ep, int M, int N, int K, cudaStream_t stream, const ConvGeom& cg);  // gemm2.cu

// Tile configuration codes accepted as `block_n`: 64 / 128 / 256 = one CTA per 128 x block_n tile (gemm.cu);
// MA_GEMM_2CTA + 128 / 256 = CTA pair per 256 x bn tile (gemm2.cu).
constexpr int MA_GEMM_2CTA = 2000;

constexpr double G2_COST_256 = 0.82, G2_COST_128 = 0.56;

static int pick_block_n(int M, int N, int tiles_m_override = 0) {
  const int sms = device_sm_count();
  const int tiles_m = tiles_m_override ? tiles_m_override : (M + GEMM_BM - 1) / GEMM_BM;
  int best = 128;
  double best_cost = 1e30;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const int tiles = tiles_m * ((N + bn - 1) / bn);
    const int waves = (tiles + sms - 1) / sms;
    // measured on B200 (tools/bench_kernels.py): a 128-wide tile runs at ~75% and a 64-wide tile at ~43% of
    // the 256-wide tile's MMA rate (smem operand bandwidth), so per-tile cost is not proportional to BN.
    const double cost = waves * (bn == 256 ? 1.0 : (bn == 128 ? 0.66 : 0.58));
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  // CTA pairs: 256 x bn tiles over sms/2 clusters.  Per-tile cost relative to a 1-CTA 128x256 tile (measured on B200,
  // tools/bench_kernels.py): the pair tile does twice the work at the full tensor rate (operand traffic per SM is 2/3).
  static const bool use_pairs = [] {
    const char* e = getenv("MA_GEMM_2CTA");
    return e == nullptr || e[0] != '0';
  }();
  const int pcands[2] = {256, 128};
  for (int i = 0; use_pairs && i < 2; ++i) {
    const int bn = pcands[i];
    const int tiles = ((tiles_m + 1) / 2) * ((N + bn - 1) / bn);
    const int clusters = sms / 2;
    const int waves = (tiles + clusters - 1) / clusters;
    const double cost = waves * (bn == 256 ? G2_COST_256 : G2_COST_128);
    if (cost < best_cost) { best_cost = cost; best = MA_GEMM_2CTA + bn; }
  }
  return best;
}

thread_local int g_last_block = 0;

static bool valid_block_code(int bn) {
  return bn == 64 || bn == 128 || bn == 256 || bn == MA_GEMM_2CTA + 128 || bn == MA_GEMM_2CTA + 256;
}

}  // namespace ma

static int check_epilogue(const ma_gemm_epilogue* epi, int N, const char* who) {
  if (N % 32 == 0) {
    const int64_t oalign = epi->out_dtype == MA_F32 ? 4 : 8;
    MA_REQUIRE(epi->ldo % oalign == 0 && (reinterpret_cast<uintptr_t>(epi->out) & 15) == 0,
               "%s: output not 16-byte aligned (ldo=%lld)", who, (long long)epi->ldo);
    if (epi->residual) {
      const int64_t ralign = epi->residual_dtype == MA_F32 ? 4 : 8;
      MA_REQUIRE(epi->ldr % ralign == 0 && (reinterpret_cast<uintptr_t>(epi->residual) & 15) == 0,
                 "%s: residual not 16-byte aligned", who);
    }
    if (epi->out_relu)
      MA_REQUIRE(epi->ldo_relu % 8 == 0 && (reinterpret_cast<uintptr_t>(epi->out_relu) & 15) == 0,
                 "%s: out_relu not 16-byte aligned", who);
    if (epi->bias) MA_REQUIRE((reinterpret_cast<uintptr_t>(epi->bias) & 15) == 0, "%s: bias not 16-byte aligned", who);
    if (epi->colscale)
      MA_REQUIRE((reinterpret_cast<uintptr_t>(epi->colscale) & 15) == 0, "%s: colscale not 16-byte aligned", who);
  } else {
    // A ragged last N chunk takes the scalar store path, full chunks before it still use vector stores.
    const int64_t oalign = epi->out_dtype == MA_F32 ? 4 : 8;
    MA_REQUIRE(N <= 32 || (epi->ldo % oalign == 0 && (reinterpret_cast<uintptr_t>(epi->out) & 15) == 0),
               "%s: output not 16-byte aligned (ldo=%lld)", who, (long long)epi->ldo);
  }
  return MA_OK;
}

extern "C" int ma_conv3x3_bf16(const void* x, int n, int H, int W, int C, const void* w, int64_t ldw, int Cout,
                               const ma_gemm_epilogue* epi, int block_n, void* stream) {
  using namespace ma;
  MA_REQUIRE(x && w && epi && (epi->out || epi->head_out), "ma_conv3x3_bf16: null pointer");
  MA_REQUIRE(n > 0 && H > 0 && W > 0 && C > 0 && Cout > 0, "ma_conv3x3_bf16: bad shape");
  const bool head = epi->head_out != nullptr;
  if (head) {  // fused narrow head: the tile values feed 8 dot products instead of being stored
    MA_REQUIRE(Cout == 128 && (block_n == 0 || block_n == MA_GEMM_2CTA + 128),
               "ma_conv3x3_bf16: the fused head needs Cout == 128 and the CTA-pair 256x128 tile (block_n 0 / 2128)");
    MA_REQUIRE(epi->head_w && epi->head_bias && (reinterpret_cast<uintptr_t>(epi->head_w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(epi->head_bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(epi->head_out) & 15) == 0,
               "ma_conv3x3_bf16: fused head weights / bias / output missing or not 16-byte aligned");
    MA_REQUIRE(!epi->residual && !epi->colscale && !epi->out_relu && epi->flags == 0,
               "ma_conv3x3_bf16: the fused head takes bias + activation only");
    block_n = MA_GEMM_2CTA + 128;
  }
  MA_REQUIRE(C % 8 == 0 && ldw % 8 == 0 && ldw >= 9 * (int64_t)C, "ma_conv3x3_bf16: C / ldw must be multiples of 8, ldw >= 9C");
  MA_REQUIRE((int64_t)n * H * W < (1ll << 31), "ma_conv3x3_bf16: too many pixels");
  MA_REQUIRE(epi->rows_per_group_in == 0, "ma_conv3x3_bf16: row remapping is not supported");
  int rc = head ? MA_OK : check_epilogue(epi, Cout, "ma_conv3x3_bf16");
  if (rc != MA_OK) return rc;

  // pixel box of one M tile: maximise useful rows / 128 over all bw x bh <= 128 boxes
  ConvGeom cg;
  cg.mode = 1; cg.H = H; cg.W = W; cg.C = C;
  double best = -1.0;
  for (int bw = 1; bw <= 128 && bw <= W; ++bw) {
    int bh = 128 / bw;
    if (bh > H) bh = H;
    const int tx = (W + bw - 1) / bw, ty = (H + bh - 1) / bh;
    const double eff = (double)W * H / ((double)tx * ty * 128.0);
    if (eff > best + 1e-9) { best = eff; cg.bw = bw; cg.bh = bh; cg.tiles_x = tx; cg.tiles_y = ty; }
  }
  cg.cblocks = (C + GEMM_BK - 1) / GEMM_BK;
  const int M = n * H * W;
  const int tiles_m = n * cg.tiles_x * cg.tiles_y;
  int bn = block_n ? block_n : pick_block_n(M, Cout, tiles_m);
  MA_REQUIRE(valid_block_code(bn), "ma_conv3x3_bf16: bad block_n %d", block_n);
  g_last_block = bn;
  const bool pair = bn > MA_GEMM_2CTA;
  if (pair) bn -= MA_GEMM_2CTA;

  CUtensorMap tx, tw;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {GEMM_BK, (uint32_t)cg.bw, (uint32_t)cg.bh, 1};
    rc = make_tmap_bf16(&tx, x, 4, dims, strides, box);
    if (rc != MA_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * C, (uint64_t)Cout};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)(pair ? bn / 2 : bn)};
    rc = make_tmap_bf16(&tw, w, 2, dims, strides, box);
    if (rc != MA_OK) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int K = 9 * C;
  if (pair) return launch_gemm_2cta(bn, tx, tw, tw, 0, *epi, M, Cout, K, s, cg);
  switch (bn) {
    case 256: return launch_gemm<256>(tx, tw, tw, 0, *epi, M, Cout, K, s, cg);
    case 128: return launch_gemm<128>(tx, tw, tw, 0, *epi, M, Cout, K, s, cg);
    default: return launch_gemm<64>(tx, tw, tw, 0, *epi, M, Cout, K, s, cg);
  }
}

extern "C" int ma_last_gemm_block(void) { return ma::g_last_block; }

extern "C" int ma_gemm_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, int M, int N, int K,
                            const ma_gemm_epilogue* epi, int block_n, void* stream) {
  using namespace ma;
  MA_REQUIRE(x && w && epi && epi->out, "ma_gemm_bf16: null pointer");
  MA_REQUIRE(epi->head_out == nullptr, "ma_gemm_bf16: the fused narrow head is implemented for ma_conv3x3_bf16 only");
  MA_REQUIRE(M > 0 && N > 0 && K > 0, "ma_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  MA_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "ma_gemm_bf16: K/ldx/ldw must be multiples of 8 (K=%d ldx=%lld ldw=%lld)",
             K, (long long)ldx, (long long)ldw);
  MA_REQUIRE(ldx >= K && ldw >= K, "ma_gemm_bf16: leading dimension smaller than K");
  {
    int rc = check_epilogue(epi, N, "ma_gemm_bf16");
    if (rc != MA_OK) return rc;
  }
  int bn = block_n ? block_n : pick_block_n(M, N);
  MA_REQUIRE(valid_block_code(bn), "ma_gemm_bf16: bad block_n %d", block_n);
  g_last_block = bn;
  const bool pair = bn > MA_GEMM_2CTA;
  if (pair) bn -= MA_GEMM_2CTA;

  CUtensorMap tx, tw;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)ldx * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int rc = make_tmap_bf16(&tx, x, 2, dims, strides, box);
    if (rc != MA_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)(pair ? bn / 2 : bn)};
    int rc = make_tmap_bf16(&tw, w, 2, dims, strides, box);
    if (rc != MA_OK) return rc;
  }
  // TMA epilogue (bulk tensor store / reduce-add) whenever the output rows are the GEMM rows
  static const bool tma_epi = [] {
    const char* e = getenv("MA_GEMM_TMA_EPILOGUE");
    return e == nullptr || e[0] != '0';
  }();
  const int epi_mode = tma_epi ? tma_epilogue_mode(*epi, N, false) : 0;
  CUtensorMap tout = tw;
  if (epi_mode) {
    const bool f32 = epi->out_dtype == MA_F32;
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)epi->ldo * (f32 ? 4 : 2)};
    uint32_t box[2] = {32, 32};
    int rc = make_tmap(&tout, epi->out, f32 ? MA_F32 : MA_BF16, 2, dims, strides, box, f32 ? 128 : 64);
    if (rc != MA_OK) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (pair) return launch_gemm_2cta(bn, tx, tw, tout, epi_mode, *epi, M, N, K, s, ConvGeom{});
  switch (bn) {
    case 256: return launch_gemm<256>(tx, tw, tout, epi_mode, *epi, M, N, K, s);
    case 128: return launch_gemm<128>(tx, tw, tout, epi_mode, *epi, M, N, K, s);
    default: return launch_gemm<64>(tx, tw, tout, epi_mode, *epi, M, N, K, s);
  }
}
