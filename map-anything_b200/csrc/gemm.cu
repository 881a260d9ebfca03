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
#include "host_common.h"
#include "ptx.cuh"

namespace ma {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARPS = 8;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
  static constexpr int TMEM_COLS = 2 * BN;                                   // 128 / 256 / 512: powers of two
};

// 32 consecutive output columns of one row: fused epilogue + store.
__device__ __forceinline__ void epilogue_store_chunk(const ma_gemm_epilogue& ep, const uint32_t (&acc)[32], int m,
                                                     int col0, int N) {
  int out_row = m;
  if (ep.rows_per_group_in > 0) {
    int g = m / ep.rows_per_group_in;
    out_row = g * ep.rows_per_group_out + ep.row_offset_out + (m - g * ep.rows_per_group_in);
  }
  const int res_row = ep.residual_row_mod > 0 ? (m % ep.residual_row_mod) : out_row;
  const int nvalid = min(32, N - col0);

  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);

  if (nvalid == 32) {
    if (ep.bias) {
      const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b = __ldg(b4 + j);
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
    const bool act_late = (ep.flags & MA_GEMM_ACT_AFTER_RESIDUAL) != 0;
    const bool relu_early = (ep.flags & MA_GEMM_RELU_OUT_BEFORE_RESIDUAL) != 0;
    if (!act_late) {
      if (ep.act == MA_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
      } else if (ep.act == MA_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
      }
    }
    if (ep.out_relu && relu_early) {
      uint4* o4 =
          reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(ep.out_relu) + (size_t)out_row * ep.ldo_relu + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o4[j] = make_uint4(pack_bf16x2(fmaxf(v[8 * j], 0.f), fmaxf(v[8 * j + 1], 0.f)),
                           pack_bf16x2(fmaxf(v[8 * j + 2], 0.f), fmaxf(v[8 * j + 3], 0.f)),
                           pack_bf16x2(fmaxf(v[8 * j + 4], 0.f), fmaxf(v[8 * j + 5], 0.f)),
                           pack_bf16x2(fmaxf(v[8 * j + 6], 0.f), fmaxf(v[8 * j + 7], 0.f)));
    }
    if (ep.colscale) {
      const float4* s4 = reinterpret_cast<const float4*>(ep.colscale + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 s = __ldg(s4 + j);
        v[4 * j + 0] *= s.x; v[4 * j + 1] *= s.y; v[4 * j + 2] *= s.z; v[4 * j + 3] *= s.w;
      }
    }
    if (ep.residual) {
      if (ep.residual_dtype == MA_F32) {
        const float4* r4 =
            reinterpret_cast<const float4*>(static_cast<const float*>(ep.residual) + (size_t)res_row * ep.ldr + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 r = r4[j];
          v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
      } else {
        const uint4* r4 = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(ep.residual) +
                                                         (size_t)res_row * ep.ldr + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 r = r4[j];
          v[8 * j + 0] += bf16_lo(r.x); v[8 * j + 1] += bf16_hi(r.x);
          v[8 * j + 2] += bf16_lo(r.y); v[8 * j + 3] += bf16_hi(r.y);
          v[8 * j + 4] += bf16_lo(r.z); v[8 * j + 5] += bf16_hi(r.z);
          v[8 * j + 6] += bf16_lo(r.w); v[8 * j + 7] += bf16_hi(r.w);
        }
      }
    }
    if (act_late) {
      if (ep.act == MA_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
      } else if (ep.act == MA_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
      }
    }
    if (ep.out_dtype == MA_F32) {
      float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(ep.out) + (size_t)out_row * ep.ldo + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(ep.out) + (size_t)out_row * ep.ldo + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
    }
    if (ep.out_relu && !relu_early) {
      uint4* o4 =
          reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(ep.out_relu) + (size_t)out_row * ep.ldo_relu + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o4[j] = make_uint4(pack_bf16x2(fmaxf(v[8 * j], 0.f), fmaxf(v[8 * j + 1], 0.f)),
                           pack_bf16x2(fmaxf(v[8 * j + 2], 0.f), fmaxf(v[8 * j + 3], 0.f)),
                           pack_bf16x2(fmaxf(v[8 * j + 4], 0.f), fmaxf(v[8 * j + 5], 0.f)),
                           pack_bf16x2(fmaxf(v[8 * j + 6], 0.f), fmaxf(v[8 * j + 7], 0.f)));
    }
  } else {
    // ragged N edge: scalar path (fully unrolled + predicated so v[] stays in registers)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j >= nvalid) continue;
      const int n = col0 + j;
      float x = v[j];
      const bool act_late = (ep.flags & MA_GEMM_ACT_AFTER_RESIDUAL) != 0;
      const bool relu_early = (ep.flags & MA_GEMM_RELU_OUT_BEFORE_RESIDUAL) != 0;
      if (ep.bias) x += __ldg(ep.bias + n);
      if (!act_late) {
        if (ep.act == MA_ACT_GELU) x = gelu_erf(x);
        else if (ep.act == MA_ACT_RELU) x = fmaxf(x, 0.0f);
      }
      if (ep.out_relu && relu_early)
        static_cast<__nv_bfloat16*>(ep.out_relu)[(size_t)out_row * ep.ldo_relu + n] = __float2bfloat16(fmaxf(x, 0.f));
      if (ep.colscale) x *= __ldg(ep.colscale + n);
      if (ep.residual) {
        if (ep.residual_dtype == MA_F32) x += static_cast<const float*>(ep.residual)[(size_t)res_row * ep.ldr + n];
        else x += __bfloat162float(static_cast<const __nv_bfloat16*>(ep.residual)[(size_t)res_row * ep.ldr + n]);
      }
      if (act_late) {
        if (ep.act == MA_ACT_GELU) x = gelu_erf(x);
        else if (ep.act == MA_ACT_RELU) x = fmaxf(x, 0.0f);
      }
      if (ep.out_dtype == MA_F32) static_cast<float*>(ep.out)[(size_t)out_row * ep.ldo + n] = x;
      else static_cast<__nv_bfloat16*>(ep.out)[(size_t)out_row * ep.ldo + n] = __float2bfloat16(x);
      if (ep.out_relu && !relu_early)
        static_cast<__nv_bfloat16*>(ep.out_relu)[(size_t)out_row * ep.ldo_relu + n] = __float2bfloat16(fmaxf(x, 0.f));
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         const ma_gemm_epilogue ep, const int M, const int N, const int K) {
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

  const int tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int total_tiles = tiles_m * tiles_n;
  const int kblocks = (K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m0 = (t % tiles_m) * GEMM_BM;
        const int n0 = (t / tiles_m) * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&bar_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bar_full[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmap_x, &bar_full[stage], kb * GEMM_BK, m0);
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmap_w, &bar_full[stage], kb * GEMM_BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
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
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;   // which half of the BN columns this warp drains
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (t % tiles_m) * GEMM_BM;
      const int n0 = (t / tiles_m) * BN;
      mbar_wait(&bar_tfull[acc], acc_phase);
      tc_fence_after();
      const int m = m0 + quarter * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 32) {
        const int col0 = n0 + half * (BN / 2) + c;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * (BN / 2) + c;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        if (m < M && col0 < N) epilogue_store_chunk(ep, v, m, col0, N);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_gemm(const CUtensorMap& tx, const CUtensorMap& tw, const ma_gemm_epilogue& ep, int M, int N, int K,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    MA_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + BN - 1) / BN);
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  gemm_bf16_tcgen05_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tx, tw, ep, M, N, K);
  MA_CHECK_CUDA(cudaGetLastError());
  return MA_OK;
}

static int pick_block_n(int M, int N) {
  const int sms = device_sm_count();
  const int tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
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
  return best;
}

}  // namespace ma

extern "C" int ma_gemm_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, int M, int N, int K,
                            const ma_gemm_epilogue* epi, int block_n, void* stream) {
  using namespace ma;
  MA_REQUIRE(x && w && epi && epi->out, "ma_gemm_bf16: null pointer");
  MA_REQUIRE(M > 0 && N > 0 && K > 0, "ma_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  MA_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "ma_gemm_bf16: K/ldx/ldw must be multiples of 8 (K=%d ldx=%lld ldw=%lld)",
             K, (long long)ldx, (long long)ldw);
  MA_REQUIRE(ldx >= K && ldw >= K, "ma_gemm_bf16: leading dimension smaller than K");
  if (N % 32 == 0) {
    const int64_t oalign = epi->out_dtype == MA_F32 ? 4 : 8;
    MA_REQUIRE(epi->ldo % oalign == 0 && (reinterpret_cast<uintptr_t>(epi->out) & 15) == 0,
               "ma_gemm_bf16: output not 16-byte aligned (ldo=%lld)", (long long)epi->ldo);
    if (epi->residual) {
      const int64_t ralign = epi->residual_dtype == MA_F32 ? 4 : 8;
      MA_REQUIRE(epi->ldr % ralign == 0 && (reinterpret_cast<uintptr_t>(epi->residual) & 15) == 0,
                 "ma_gemm_bf16: residual not 16-byte aligned");
    }
    if (epi->out_relu)
      MA_REQUIRE(epi->ldo_relu % 8 == 0 && (reinterpret_cast<uintptr_t>(epi->out_relu) & 15) == 0,
                 "ma_gemm_bf16: out_relu not 16-byte aligned");
    if (epi->bias) MA_REQUIRE((reinterpret_cast<uintptr_t>(epi->bias) & 15) == 0, "ma_gemm_bf16: bias not 16-byte aligned");
    if (epi->colscale)
      MA_REQUIRE((reinterpret_cast<uintptr_t>(epi->colscale) & 15) == 0, "ma_gemm_bf16: colscale not 16-byte aligned");
  }
  // A ragged last N chunk takes the scalar store path, full chunks before it still use vector stores.
  if (N % 32 != 0) {
    const int64_t oalign = epi->out_dtype == MA_F32 ? 4 : 8;
    MA_REQUIRE(N <= 32 || (epi->ldo % oalign == 0 && (reinterpret_cast<uintptr_t>(epi->out) & 15) == 0),
               "ma_gemm_bf16: output not 16-byte aligned (ldo=%lld)", (long long)epi->ldo);
  }
  int bn = block_n ? block_n : pick_block_n(M, N);
  MA_REQUIRE(bn == 64 || bn == 128 || bn == 256, "ma_gemm_bf16: block_n must be 0/64/128/256, got %d", block_n);

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
    uint32_t box[2] = {GEMM_BK, (uint32_t)bn};
    int rc = make_tmap_bf16(&tw, w, 2, dims, strides, box);
    if (rc != MA_OK) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256: return launch_gemm<256>(tx, tw, *epi, M, N, K, s);
    case 128: return launch_gemm<128>(tx, tw, *epi, M, N, K, s);
    default: return launch_gemm<64>(tx, tw, *epi, M, N, K, s);
  }
}
