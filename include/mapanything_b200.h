/*
 * mapanything_b200 -- C ABI of the B200-native (sm_100a) kernels behind the MapAnything
 * feed-forward inference hot path.
 *
 * The reference (etola/map-anything) is pure Python over PyTorch library kernels; it has no FFI of
 * its own.  The boundary a maintainer binds is therefore the Python surface
 * `MapAnything.forward / MapAnything.infer` (reference mapanything/models/mapanything/model.py:1477,
 * :1964); each entry point below replaces the PyTorch op(s) that a stage of that path dispatches
 * and cites the reference call site it stands in for.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - all functions return MA_OK (0) or a negative ma_status; the message for the calling thread is
 *    available from ma_last_error();
 *  - no C++ exceptions, no torch types, no ownership transfer: the caller owns every buffer;
 *  - bf16 tensors are `__nv_bfloat16` (uint16 storage), row-major, innermost dimension contiguous.
 */
#ifndef MAPANYTHING_B200_H_
#define MAPANYTHING_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MA_ABI_VERSION 1

typedef enum ma_status {
  MA_OK = 0,
  MA_ERR_INVALID = -1, /* bad argument (shape / alignment / dtype) */
  MA_ERR_CUDA = -2     /* CUDA runtime / driver error */
} ma_status;

typedef enum ma_dtype { MA_BF16 = 0, MA_F32 = 1 } ma_dtype;
typedef enum ma_act { MA_ACT_NONE = 0, MA_ACT_GELU = 1, MA_ACT_RELU = 2 } ma_act;

/* ---- library ------------------------------------------------------------------------------ */
const char* ma_last_error(void);
int ma_abi_version(void);
int ma_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- GEMM: Y = epilogue(X . W^T) ------------------------------------------------------------
 * Replaces every nn.Linear / 1x1 conv / im2col'ed conv on the path (cuBLASLt / cuDNN calls in the
 * reference: DINOv2 qkv/proj/fc1/fc2 mapanything/models/external/dinov2/layers/attention.py:53-70,
 * mlp.py:16-40; patch embed patch_embed.py:65-87; the uniception info-sharing / DPT / pose-head
 * layers invoked at model.py:1532-1542, :1302-1338, :1449-1457).
 *
 * X: [M,K] bf16 (row stride ldx), W: [N,K] bf16 (row stride ldw), fp32 accumulate in TMEM
 * (tcgen05.mma), operands staged by TMA.  K and the strides must be multiples of 8 elements.
 * Epilogue, applied per element in this order:
 *    v = acc + bias[n];  v = act(v);  v *= colscale[n];  v += residual[res_row, n]
 *    out[out_row, n] = v;  out_relu[out_row, n] = max(v, 0)   (optional second output, bf16)
 * Row mapping (token assembly without copy kernels): when rows_per_group_in > 0,
 *    out_row = (m / rows_per_group_in) * rows_per_group_out + row_offset_out + m % rows_per_group_in
 * else out_row = m.  res_row = m % residual_row_mod when residual_row_mod > 0, else out_row. */
typedef struct ma_gemm_epilogue {
  void* out;
  int64_t ldo;
  int32_t out_dtype; /* ma_dtype */
  int32_t act;       /* ma_act */
  const float* bias;     /* [N] or NULL */
  const float* colscale; /* [N] or NULL (LayerScale gamma) */
  const void* residual;  /* or NULL */
  int64_t ldr;
  int32_t residual_dtype; /* ma_dtype */
  int32_t residual_row_mod;
  void* out_relu; /* bf16 or NULL */
  int64_t ldo_relu;
  int32_t rows_per_group_in;
  int32_t rows_per_group_out;
  int32_t row_offset_out;
  int32_t reserved;
} ma_gemm_epilogue;

/* block_n: 0 = choose automatically, else one of 64 / 128 / 256. */
int ma_gemm_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, int M, int N, int K,
                 const ma_gemm_epilogue* epi, int block_n, void* stream);

/* ---- Attention forward (head_dim 64, non-causal) ---------------------------------------------
 * Replaces the attention of the DINOv2 blocks (reference dinov2/layers/attention.py:53-90,
 * F.scaled_dot_product_attention / xformers there) and of the info-sharing frame / global blocks
 * (uniception, invoked at model.py:1532-1542; analog vggt/layers/attention.py:46-76).
 *
 * q/k/v are token-major bf16 matrices (typically three column slices of one fused-qkv GEMM output):
 * head h of q occupies columns [q_col0 + 64 h, q_col0 + 64 h + 64).  Sequence s uses query rows
 * [s*q_seq_stride, s*q_seq_stride + q_len) and key/value rows [s*kv_seq_stride, ... + kv_len).
 * out[row, o_col0 + 64 h ...] = softmax(q k^T * softmax_scale) v, bf16.  QK^T and PV run on tcgen05
 * with accumulators in TMEM; softmax statistics and the output accumulator are fp32. */
int ma_attention_fwd(const void* q, int64_t ldq, int64_t q_rows, int q_col0, const void* k, int64_t ldk,
                     int64_t kv_rows, int k_col0, const void* v, int64_t ldv, int v_col0, void* out, int64_t ldo,
                     int o_col0, int num_seqs, int num_heads, int q_len, int kv_len, int64_t q_seq_stride,
                     int64_t kv_seq_stride, float softmax_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAPANYTHING_B200_H_ */
