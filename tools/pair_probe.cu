// Semantics + throughput probe for tcgen05.mma.cta_group::2 (CTA pairs) on sm_100a.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/pair_probe tools/pair_probe.cu
// (1) semantics: M = 256 over two CTAs (128 A rows each), B split by N (N/2 rows of the K-major B tile in each CTA),
//     D = 128 lanes x N columns in each CTA's tensor memory; multicast commit; a remote mbarrier arrive (the relay the
//     window-GEMM kernel uses to tell the leader that the peer's operands have landed)
// (2) cycles per MMA for N = 32, 64, 128, 256 against the single-CTA instruction
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../baryon_painter_b200/csrc/bp_tc.cuh"

using namespace bp::tc;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

struct PArgs {
  int n, reps, pair;
  int fill_interval;   // > 0: a second thread streams 16 KB bulk copies into shared memory, one every so many cycles
  const uint4* fill_src;
  float* out;          // [2][128][n]
  long long* cyc;      // [grid]
};

// A: no-swizzle K-major, 128 rows x 16 k: core matrices of 8 rows x 16 B, SBO = 128 B, LBO = 128 rows x 16 B
// B: the same layout with `nb` rows
__device__ void fill_operand(__half* dst, int rows, int row0, int mulr, int mulk, int mod, int sub) {
  for (int i = threadIdx.x; i < rows * 16; i += blockDim.x) {
    const int r = i / 16, k = i % 16;
    const int half = k / 8, e = k % 8;
    const float v = (float)(((row0 + r) * mulr + k * mulk) % mod - sub);
    dst[(size_t)half * rows * 8 + (size_t)r * 8 + e] = __float2half(v);
  }
}

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) pair_kernel(PArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_done, bar_relay, bar_fill[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank = 0;
  if constexpr (PAIR) rank = cluster_ctarank();
  const int nb = PAIR ? a.n / 2 : a.n;        // B rows held by this CTA
  __half* A = reinterpret_cast<__half*>(smem);
  __half* B = reinterpret_cast<__half*>(smem + 16 * 1024);
  fill_operand(A, 128, (int)rank * 128, 3, 1, 7, 3);
  fill_operand(B, nb, (int)rank * nb, 5, 2, 5, 2);
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    mbar_init(&bar_relay, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_fill[i], 1);
    stop = 0;
    fence_barrier_init();
  }
  fence_proxy_async();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    if constexpr (PAIR) { tmem_alloc2(&slot, 512); tmem_relinquish2(); }
    else { tmem_alloc(&slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if constexpr (PAIR) { if (rank == 1 && tid == 0) mbar_arrive_remote(&bar_relay, 0); }      // "my operands are in place"
  if (warp == 1 && lane == 0 && a.fill_interval > 0) {
    // background fill traffic: what the patch / weight producers of the window-GEMM kernel write while MMAs run
    long long next = clock64();
    int i = 0;
    for (; !stop && i < 100000; ++i) {
      const int sl = i & 3;
      if (i >= 4) mbar_wait(&bar_fill[sl], ((i >> 2) - 1) & 1);
      mbar_arrive_expect_tx(&bar_fill[sl], 16384);
      bulk_g2s(smem + 64 * 1024 + sl * 16384, a.fill_src + (size_t)((blockIdx.x * 7 + i) & 63) * 1024, 16384, &bar_fill[sl]);
      next += a.fill_interval;
      while (clock64() < next) {
      }
    }
    for (int j = (i > 4 ? i - 4 : 0); j < i; ++j) mbar_wait(&bar_fill[j & 3], (j >> 2) & 1);
    if (rank == 0) a.cyc[148 + blockIdx.x] = i;
  }
  if (rank == 0 && tid == 0) {
    if constexpr (PAIR) mbar_wait_cluster(&bar_relay, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16_m(0, a.n, PAIR ? 256 : 128);
    const uint64_t da = make_smem_desc(smem_u32(A), 128 * 16, 128);
    const uint64_t db = make_smem_desc(smem_u32(B), (uint32_t)nb * 16, 128);
    if constexpr (PAIR) umma_f16_2(tm, da, db, idesc, 0); else umma_f16(tm, da, db, idesc, 0);
    const long long t0 = clock64();
    for (int r = 0; r < a.reps; ++r) {
      if constexpr (PAIR) umma_f16_2(tm + (uint32_t)a.n, da, db, idesc, r > 0);
      else umma_f16(tm + (uint32_t)a.n, da, db, idesc, r > 0);
    }
    if constexpr (PAIR) umma_commit_mc(&bar_done, 3); else umma_commit(&bar_done);
    mbar_wait(&bar_done, 0);
    a.cyc[blockIdx.x] = clock64() - t0;
    stop = 1;
  }
  if constexpr (PAIR) {
    // the peer's filler stops when the leader is done
    if (rank == 1 && tid == 0) { mbar_wait(&bar_done, 0); stop = 1; }
  }
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  if (blockIdx.x < 2) {
    uint32_t v[16];
    for (int c = 0; c < a.n; c += 16) {
      tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a.n + c), v);
      tmem_ld_wait();
      for (int e = 0; e < 16; ++e) a.out[((size_t)rank * 128 + warp * 32 + lane) * a.n + c + e] = __uint_as_float(v[e]);
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc2(tm, 512); else tmem_dealloc(tm, 512);
  }
}

static float ref_d(int row, int n) {
  float s = 0.f;
  for (int k = 0; k < 16; ++k) s += (float)((row * 3 + k) % 7 - 3) * (float)((n * 5 + k * 2) % 5 - 2);
  return s;
}

int main() {
  float* d_out;
  long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(float) * 2 * 128 * 256));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 296));
  uint4* d_fill;
  CK(cudaMalloc(&d_fill, 64 * 16384));
  CK(cudaMemset(d_fill, 0, 64 * 16384));
  CK(cudaFuncSetAttribute(pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  CK(cudaFuncSetAttribute(pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  for (int fill : {0, 1024, 512, 384, 256})
  for (int pair = 0; pair < 2; ++pair)
    for (int n : {64, 128}) {
      for (int grid : {148}) {
        PArgs a{n, 2048, pair, fill, d_fill, d_out, d_cyc};
        CK(cudaMemset(d_out, 0, sizeof(float) * 2 * 128 * 256));
        CK(cudaMemset(d_cyc, 0, sizeof(long long) * 296));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 128 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (pair) CK(cudaLaunchKernelEx(&cfg, pair_kernel<true>, a));
        else CK(cudaLaunchKernelEx(&cfg, pair_kernel<false>, a));
        CK(cudaDeviceSynchronize());
        std::vector<float> h(2 * 128 * n);
        std::vector<long long> c(296);
        CK(cudaMemcpy(h.data(), d_out, sizeof(float) * h.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(c.data(), d_cyc, sizeof(long long) * 296, cudaMemcpyDeviceToHost));
        int bad = 0, rows = pair ? 256 : 128;
        for (int r = 0; r < rows; ++r)
          for (int j = 0; j < n; ++j) {
            // D of the last MMA pass (accumulating passes all add the same product: reps x ref), first accumulator
            // holds one product
            const float want = ref_d(r, j) * 2048.f;
            if (h[(size_t)r * n + j] != want && bad++ < 4)
              printf("   mismatch row %d col %d: got %g want %g\n", r, j, h[(size_t)r * n + j], want);
          }
        double mx = 0;
        for (int i = 0; i < grid; ++i) mx = c[i] > mx ? (double)c[i] : mx;
        const double fill_rate = mx > 0 ? (double)c[148] * 16384.0 / mx : 0.0;
        printf("%s N=%3d grid=%3d  %s  %.1f cycles per MMA, fill %.1f B/clk per SM (asked %.1f)\n",
               pair ? "cta_group::2 M=256" : "cta_group::1 M=128", n, grid, bad ? "WRONG" : "values ok", mx / 2048.0, fill_rate,
               fill ? 16384.0 / fill : 0.0);
      }
    }
  return 0;
}
