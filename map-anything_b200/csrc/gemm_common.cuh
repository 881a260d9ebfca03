// Shared by the 1-CTA (gemm.cu) and 2-CTA (gemm2.cu) tcgen05 GEMM kernels: implicit-conv geometry and the fused epilogue.
#pragma once
#include "host_common.h"
#include "ptx.cuh"

namespace ma {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;

// Implicit-GEMM geometry of a 3x3 / stride 1 / pad 1 convolution over an NHWC tensor (mode 1).  The A operand of the
// tile (bh x bw output pixels of one image) and of tap (ky,kx) is the input window shifted by (ky-1, kx-1): ONE 4-D TMA
// box {64 channels, bw, bh, 1} whose out-of-bounds pixels (the zero padding) and channels are zero-filled by the TMA
// unit -- no im2col matrix is ever written.  K runs over 9 taps x ceil(C/64) channel blocks.
// Tile raster of the persistent GEMMs: tile t of tiles_m x tiles_n -> (row tile, column tile).  Concurrently running
// workers (CTAs / CTA pairs) take consecutive t.  M-fastest order made every worker of a wave stream a DIFFERENT X row
// tile against the same W panel, so X was re-read once per column panel -- from DRAM when X + residual exceed the L2 (ncu,
// 10960 x 1024 x 4096: 260 MB read for 143 MB algorithmic).  Banded order: `band` row tiles x all column tiles are in flight
// together, column-fastest, so the tiles_n workers sharing an X row tile run at the same time (one DRAM read, L2 hits for
// the rest) while the W panels (a few MB) stay L2 resident.
__device__ __forceinline__ void raster_tile(int t, int tiles_n, int workers, int& tm, int& tn) {
  int band = workers / tiles_n;
  band = band < 1 ? 1 : band;
  const int per_band = band * tiles_n;
  const int b = t / per_band;
  const int r = t - b * per_band;  // the last band is shorter, but every earlier one is full: no special case
  tm = b * band + r / tiles_n;
  tn = r - (r / tiles_n) * tiles_n;
}

struct ConvGeom {
  int mode;  // 0 = plain GEMM
  int H, W, C;
  int bw, bh;            // pixel box of one M tile (bw * bh <= 128)
  int tiles_x, tiles_y;  // per image
  int cblocks;           // ceil(C / 64)
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
        if (ep.out_dtype == MA_F32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
        }
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
        if (ep.out_dtype == MA_F32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
        }
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


// Fused narrow head: this thread's 32 accumulator columns [col0, col0 + 32) of its row get bias + activation and are folded
// into 8 running dot products with the head weights (shared memory copy hw[8][N], read as broadcast float4).
__device__ __forceinline__ void epilogue_head_chunk(const ma_gemm_epilogue& ep, const uint32_t (&acc)[32], int col0, int N,
                                                    const float* __restrict__ hw, float (&h)[8]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  if (ep.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  if (ep.act == MA_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if (ep.act == MA_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const float4* w4 = reinterpret_cast<const float4*>(hw + n * N + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = w4[j];
      h[n] = fmaf(v[4 * j], w.x, h[n]); h[n] = fmaf(v[4 * j + 1], w.y, h[n]);
      h[n] = fmaf(v[4 * j + 2], w.z, h[n]); h[n] = fmaf(v[4 * j + 3], w.w, h[n]);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// TMA epilogue (plain GEMMs whose output rows are the GEMM rows): the warp's 32 x 32 accumulator chunk gets
// bias / activation / column scale applied per thread (thread = row), is written to a warp-private, hardware-swizzled
// shared-memory stage and leaves the SM as ONE asynchronous bulk tensor store -- or, for the in-place fp32 residual of a
// transformer block (x += gamma * (acc + b)), as a bulk tensor REDUCE-ADD executed at the L2, so the residual stream is
// never read by the SM.  No strided 16-byte stores, no exposed global-load latency in the epilogue warps, ragged
// M handled by TMA clipping.  mode: 1 = store, 2 = reduce-add.  All 32 lanes call; lane 0 issues the TMA.
// ----------------------------------------------------------------------------------------------------------------
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;

__device__ __forceinline__ void epilogue_tma_chunk(const CUtensorMap* tmap_out, int mode, const ma_gemm_epilogue& ep,
                                                   const uint32_t (&acc)[32], int row0, int col0, uint8_t* stage, int lane,
                                                   bool with_bias = true) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  if (ep.bias && with_bias) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  if (ep.act == MA_ACT_GELU) {
    if (ep.out_dtype == MA_F32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
    }
  } else if (ep.act == MA_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (ep.colscale) {
    const float4* s4 = reinterpret_cast<const float4*>(ep.colscale + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 s = __ldg(s4 + j);
      v[4 * j] *= s.x; v[4 * j + 1] *= s.y; v[4 * j + 2] *= s.z; v[4 * j + 3] *= s.w;
    }
  }
  // the previous bulk store of this warp must have finished reading the stage
  if (lane == 0) tma_store_wait_read<0>();
  __syncwarp();
  if (ep.out_dtype == MA_F32) {
    // rows of 128 B, SWIZZLE_128B: 16-byte chunk q of row r lives at r*128 + ((q ^ (r & 7)) << 4)
    uint4* row = reinterpret_cast<uint4*>(stage + lane * 128);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      row[q ^ (lane & 7)] = make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                                       __float_as_uint(v[4 * q + 3]));
  } else {
    // rows of 64 B, SWIZZLE_64B: chunk q of row r lives at r*64 + ((q ^ ((r >> 1) & 3)) << 4)
    uint4* row = reinterpret_cast<uint4*>(stage + lane * 64);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      row[q ^ ((lane >> 1) & 3)] = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                              pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    if (mode == 2) tma_reduce_add_2d(tmap_out, stage, col0, row0);
    else tma_store_2d(tmap_out, stage, col0, row0);
    tma_store_commit();
  }
}

// Host side: can this launch use the TMA epilogue, and in which mode (0 = no)?
inline int tma_epilogue_mode(const ma_gemm_epilogue& ep, int N, bool conv) {
  if (conv || ep.rows_per_group_in != 0 || ep.residual_row_mod != 0 || ep.out_relu != nullptr || N % 32 != 0) return 0;
  if ((ep.flags & (MA_GEMM_ACT_AFTER_RESIDUAL | MA_GEMM_RELU_OUT_BEFORE_RESIDUAL)) != 0) return 0;
  const int64_t esz = ep.out_dtype == MA_F32 ? 4 : 2;
  if ((ep.ldo * esz) % 16 != 0 || (reinterpret_cast<uintptr_t>(ep.out) & 15) != 0) return 0;
  if (ep.residual == nullptr) return 1;
  if (ep.residual == ep.out && ep.residual_dtype == MA_F32 && ep.out_dtype == MA_F32 && ep.ldr == ep.ldo) return 2;
  return 0;
}

}  // namespace ma
