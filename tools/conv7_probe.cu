// Does the conv7 formulation's operand geometry cost more than the 48-cycle N = 64 instruction floor?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/conv7_probe tools/conv7_probe.cu
// One CTA per SM issues the MMA sequence of a conv7 region (W_BLOCK, G = 4, Jy = 2, N = 64: 8 tap lines x 12
// 32-byte slices, two M-tiles) with the kernel's descriptors: A = 128B-swizzled patch, 8-row groups `sbo` bytes apart,
// start address walking over (line, slice); B = 2 KB no-swizzle tiles of a 24 KB stage.  Variants isolate each part.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../baryon_painter_b200/csrc/bp_tc.cuh"
using namespace bp::tc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

struct CArgs {
  int n, reps;
  uint32_t a_sbo;        // bytes between 8-row groups of A
  int walk_a, walk_b;    // walk the start addresses like the kernel (else: fixed)
  int two_tiles;         // alternate two M-tiles / accumulators
  int two_warps;         // two issuing warps (alternate M-tiles), as the kernel
  int hoist;             // slot-table reads of 8 slots issued ahead of their MMAs (else: read, issue, read, issue)
  long long* cyc;
  uint32_t ao[96], bo[96];   // per-slot operand offsets (16-byte units), from the parameter bank as in the kernel
};

__device__ __forceinline__ void issue_seq(const CArgs& a, uint32_t tm, uint32_t abase, uint32_t bbase, int half, int nhalf) {
  const uint32_t idesc = make_idesc_f16(0, a.n);
  const uint64_t da_t = make_smem_desc_sw(0, 128);
  const uint32_t a_hi = ((uint32_t)(da_t >> 32) & ~0x3FFFu) | (a.a_sbo >> 4);
  const uint64_t db_t = make_smem_desc(0, (uint32_t)a.n * 16u, 128u);
  const uint32_t b_hi = (uint32_t)(db_t >> 32);
  const uint32_t a_lo0 = (uint32_t)da_t + (abase >> 4), b_lo0 = (uint32_t)db_t + (bbase >> 4);
  const uint32_t el = elect_one() ? 1u : 0u;
  const int ntile = a.two_tiles ? 2 : 1;
  if (a.hoist) {
    for (int rep = 0; rep < a.reps; ++rep) {
      for (int s0 = 0; s0 < 96; s0 += 8) {
        uint32_t ao[8], bo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { ao[j] = a.ao[s0 + j]; bo[j] = a.bo[s0 + j]; }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          for (int mt = half; mt < ntile; mt += nhalf)
            umma_f16_lohi(tm + (uint32_t)(mt * a.n), a_lo0 + ao[j] + (uint32_t)mt * 64u, a_hi, b_lo0 + bo[j], b_hi, idesc,
                          (uint32_t)(rep > 0), el);
      }
    }
    return;
  }
  for (int rep = 0; rep < a.reps; ++rep) {
#pragma unroll 2
    for (int sl = 0; sl < 96; ++sl) {
      const uint32_t ao = a.ao[sl], bo = a.bo[sl];
      for (int mt = half; mt < ntile; mt += nhalf)
        umma_f16_lohi(tm + (uint32_t)(mt * a.n), a_lo0 + ao + (uint32_t)mt * 64u, a_hi, b_lo0 + bo, b_hi, idesc, (uint32_t)(rep > 0), el);
    }
  }
}

__global__ void __launch_bounds__(128, 1) conv7_kernel(CArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, a.two_warps ? 2 : 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const long long t0 = clock64();
  if (warp == 2 || (warp == 3 && a.two_warps)) {
    issue_seq(a, tm, smem_u32(smem), smem_u32(smem + 100 * 1024), a.two_warps ? warp - 2 : 0, a.two_warps ? 2 : 1);
    umma_commit_pred(&bar, elect_one() ? 1u : 0u);
  }
  mbar_wait(&bar, 0);
  if (tid == 64) a.cyc[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

int main() {
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 148));
  CK(cudaFuncSetAttribute(conv7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  struct V { const char* name; uint32_t sbo; int wa, wb, tt, tw, ho; };
  const V vs[] = {
      {"fixed A/B, SBO 1024, one tile, one warp", 1024, 0, 0, 0, 0},
      {"fixed A/B, SBO 4608 (block mode)", 4608, 0, 0, 0, 0},
      {"walk A (line, slice), SBO 4608", 4608, 1, 0, 0, 0},
      {"walk A, SBO 1024", 1024, 1, 0, 0, 0},
      {"walk B, SBO 4608", 4608, 0, 1, 0, 0},
      {"walk A + B, SBO 4608", 4608, 1, 1, 0, 0},
      {"walk A + B, two tiles, one warp", 4608, 1, 1, 1, 0},
      {"walk A + B, two tiles, two warps (the kernel)", 4608, 1, 1, 1, 1},
      {"walk A + B, two tiles, two warps, SBO 1024", 1024, 1, 1, 1, 1},
      {"hoisted table reads: one tile, one warp", 4608, 1, 1, 0, 0, 1},
      {"hoisted table reads: two tiles, one warp", 4608, 1, 1, 1, 0, 1},
      {"hoisted table reads: two tiles, two warps", 4608, 1, 1, 1, 1, 1},
  };
  for (int n : {64, 128})
    for (const V& v : vs) {
      CArgs a{n, 20, v.sbo, v.wa, v.wb, v.tt, v.tw, v.ho, d_cyc};
      for (int sl = 0; sl < 96; ++sl) {
        const int line = sl / 12, slice = sl % 12;
        a.ao[sl] = v.wa ? (uint32_t)(line * 18 * 8 + slice * 2) : 0u;
        a.bo[sl] = v.wb ? (uint32_t)((sl % 12) * (n * 2)) : 0u;
      }
      CK(cudaMemset(d_cyc, 0, sizeof(long long) * 148));
      conv7_kernel<<<148, 128, 160 * 1024>>>(a);
      CK(cudaDeviceSynchronize());
      std::vector<long long> c(148);
      CK(cudaMemcpy(c.data(), d_cyc, sizeof(long long) * 148, cudaMemcpyDeviceToHost));
      double mx = 0;
      for (long long x : c) mx = x > mx ? (double)x : mx;
      const double mmas = 96.0 * 20 * (v.tt ? 2 : 1);
      printf("N=%3d  %-52s %.1f cycles per MMA\n", n, v.name, mx / mmas);
    }
  return 0;
}
