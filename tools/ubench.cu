// Pipe micro-benchmarks behind the attention design decisions: how fast can an SM
//   * read tensor memory (tcgen05.ld 32x32b.x32), with 1 / 4 / 8 resident warps,
//   * evaluate ex2.approx in f32, f16x2 and bf16x2 form,
//   * issue fma.rn.f32 vs the packed fma.rn.f32x2.
// Every number is "per clock per SM", from %clock64 around an unrolled loop, median over the CTAs of a full-chip grid.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---------------------------------------------------------------------------------------------- tensor-memory reads
template <int COLS>
__global__ void ldtm_kernel(long long* cycles, uint32_t* sink, int iters, int active_warps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < active_warps) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 128; c += 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(base + c)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      }
    }
    t1 = clock64();
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(COLS) : "memory");
}

// ---------------------------------------------------------------------------------------------- ex2 forms
// mode 0: ex2.approx.ftz.f32   (1 value / instruction)
// mode 1: ex2.approx.ftz.f16x2 (2 values / instruction)
// mode 2: ex2.approx.ftz.bf16x2
// mode 3: fma.rn.f32           (1 FMA / instruction)
// mode 4: fma.rn.f32x2         (2 FMA / instruction)
// mode 5: add.rn.f32x2
template <int MODE>
__global__ void pipe_kernel(long long* cycles, uint32_t* sink, int iters) {
  constexpr int ILP = 16;
  uint32_t x[ILP];
  uint64_t y[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    x[i] = (MODE == 0 || MODE == 3) ? __float_as_uint(-0.001f * (threadIdx.x + i + 1)) : 0xB800B800u + i;  // small negatives
    y[i] = (static_cast<uint64_t>(__float_as_uint(0.5f + i)) << 32) | __float_as_uint(0.25f + threadIdx.x);
  }
  const uint64_t ca = (static_cast<uint64_t>(__float_as_uint(0.999f)) << 32) | __float_as_uint(1.001f);
  const uint64_t cb = (static_cast<uint64_t>(__float_as_uint(1e-3f)) << 32) | __float_as_uint(-1e-3f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(0x3f7fbe77u), "r"(0x3a83126fu));
      if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(ca), "l"(cb));
      if (MODE == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(cb));
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc ^= x[i] ^ static_cast<uint32_t>(y[i]) ^ static_cast<uint32_t>(y[i] >> 32);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
}

static double median_cycles(long long* d_cycles, int n) {
  std::vector<long long> h(n);
  CK(cudaMemcpy(h.data(), d_cycles, n * sizeof(long long), cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  return static_cast<double>(h[n / 2]);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d_cycles;
  uint32_t* d_sink;
  CK(cudaMalloc(&d_cycles, 4096 * sizeof(long long)));
  CK(cudaMalloc(&d_sink, 64));
  const int iters = 2000;
  printf("{\"sms\": %d", sms);
  // tensor-memory reads: one CTA per SM with 1, 2, 4 active warps (512 columns), two CTAs per SM with 4 warps each (256 columns)
  for (int aw : {1, 2, 4}) {
    ldtm_kernel<512><<<sms, 128>>>(d_cycles, d_sink, iters, aw);
    CK(cudaDeviceSynchronize());
    const double cyc = median_cycles(d_cycles, sms);
    printf(", \"ldtm_B_per_clk_sm_%dwarps\": %.1f", aw, aw * 4.0 * 4096.0 * iters / cyc);
  }
  ldtm_kernel<256><<<2 * sms, 128>>>(d_cycles, d_sink, iters, 4);
  CK(cudaDeviceSynchronize());
  printf(", \"ldtm_B_per_clk_sm_2ctas_8warps\": %.1f", 2 * 4 * 4.0 * 4096.0 * iters / median_cycles(d_cycles, 2 * sms));
  // pipes: 8 and 16 warps per SM (256 / 512 threads, one CTA per SM)
  const char* names[6] = {"ex2_f32", "ex2_f16x2", "ex2_bf16x2", "fma_f32", "fma_f32x2", "add_f32x2"};
  const int per_instr[6] = {1, 2, 2, 1, 2, 2};
  for (int mode = 0; mode < 6; ++mode) {
    for (int threads : {256, 512}) {
      switch (mode) {
        case 0: pipe_kernel<0><<<sms, threads>>>(d_cycles, d_sink, iters); break;
        case 1: pipe_kernel<1><<<sms, threads>>>(d_cycles, d_sink, iters); break;
        case 2: pipe_kernel<2><<<sms, threads>>>(d_cycles, d_sink, iters); break;
        case 3: pipe_kernel<3><<<sms, threads>>>(d_cycles, d_sink, iters); break;
        case 4: pipe_kernel<4><<<sms, threads>>>(d_cycles, d_sink, iters); break;
        default: pipe_kernel<5><<<sms, threads>>>(d_cycles, d_sink, iters); break;
      }
      CK(cudaDeviceSynchronize());
      const double cyc = median_cycles(d_cycles, sms);
      printf(", \"%s_values_per_clk_sm_%dthr\": %.2f", names[mode], threads, 16.0 * per_instr[mode] * threads * iters / cyc);
    }
  }
  printf("}\n");
  return 0;
}
