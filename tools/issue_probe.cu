// Where do the cycles of a tcgen05.mma issue loop go?  Adds the ingredients of the real kernel's MMA-issuer loop
// one at a time (commit per stage, full/empty barrier ping-pong with a producer warp, fence after each wait,
// descriptors fetched from shared memory, bystander warps polling a barrier) and reports cycles per MMA.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/issue_probe tools/issue_probe.cu
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../baryon_painter_b200/csrc/bp_tc.cuh"

using namespace bp::tc;

struct PArgs {
  int n, stages_total, per_stage, flags, uniform;
  long long* out;
};

__global__ void __launch_bounds__(256, 1) issue_kernel(PArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[4], empty[4], never, done;
  __shared__ uint32_t slot;
  __shared__ uint64_t tmpl[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid < 64) tmpl[tid] = make_smem_desc_sw(0, 128) + (uint64_t)((tid % 9) * 8 + (tid % 4) * 2);
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&never, 1); mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int F = a.flags;
  if (warp == 1 && (F & 2)) {
    if (lane == 0)
      for (int it = 0; it < a.stages_total; ++it) {
        const uint32_t st = it & 3, par = (it >> 2) & 1;
        mbar_wait(&empty[st], par ^ 1u);
        mbar_arrive(&full[st]);
      }
  } else if (warp == 2) {
    const uint32_t idesc = make_idesc_f16(0, a.n);
    const uint32_t a16 = smem_u32(smem) >> 4, b16 = smem_u32(smem + 96 * 1024) >> 4;
    const uint64_t db_t = make_smem_desc(0, (uint32_t)a.n * 16u, 128u);
    long long t0 = clock64();
    if (!a.uniform) {
      for (int it = 0; it < a.stages_total; ++it) {
        const uint32_t st = it & 3, par = (it >> 2) & 1;
        if (F & 2) mbar_wait(&full[st], par);
        if (F & 4) tc_fence_after();
        if (lane == 0) {
          const uint64_t db0 = db_t + (uint64_t)(b16 + st * 64);
          for (int q = 0; q < a.per_stage; q += 2) {
            const uint64_t da = ((F & 8) ? tmpl[(it * 4 + q) & 63] : make_smem_desc_sw(0, 128) + (uint64_t)(q * 2)) + a16;
            const uint64_t db = db0 + (uint64_t)(q * 8);
            umma_f16(tm, da, db, idesc, 1);
            umma_f16(tm + a.n, da + 128, db, idesc, 1);
          }
          if (F & 1) umma_commit(&empty[st]);
        }
        __syncwarp();
      }
    } else {
      // warp-uniform issue: every lane runs the address arithmetic, one elected lane issues
      const uint32_t el = elect_one() ? 1u : 0u;
      for (int it = 0; it < a.stages_total; ++it) {
        const uint32_t st = it & 3, par = (it >> 2) & 1;
        if (F & 2) mbar_wait(&full[st], par);
        if (F & 4) tc_fence_after();
        const uint64_t db0 = db_t + (uint64_t)(b16 + st * 64);
        for (int q = 0; q < a.per_stage; q += 2) {
          const uint64_t da = make_smem_desc_sw(0, 128) + (uint64_t)(q * 2 + ((it * 5) & 31) * 8) + a16;
          const uint64_t db = db0 + (uint64_t)(q * 8);
          umma_f16_pred(tm, da, db, idesc, 1, el);
          umma_f16_pred(tm + a.n, da + 128, db, idesc, 1, el);
        }
        if (F & 1) umma_commit_pred(&empty[st], el);
        __syncwarp();
      }
    }
    long long t_issue = clock64();
    if (lane == 0) umma_commit(&done);
    mbar_wait(&done, 0);
    long long t1 = clock64();
    if (lane == 0) {
      a.out[blockIdx.x] = t1 - t0;
      a.out[148 + blockIdx.x] = t_issue - t0;
      mbar_arrive(&never);
    }
  } else if (warp >= 4 && (F & 16)) {
    mbar_wait(&never, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  CK(cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  long long* d_out;
  CK(cudaMalloc(&d_out, 2 * 148 * sizeof(long long)));
  // bit 2 (ping-pong) needs bit 1 (the commit is what releases the producer)
  const char* names[] = {"bare", "+commit/stage", "+commit+pingpong", "+fence", "+commit+pp+fence",
                         "+LDS desc", "all five", "all + pollers"};
  const int flagsets[] = {0, 1, 3, 4, 7, 8, 15, 31};
  for (int uniform = 0; uniform < 2; ++uniform)
    for (int n : {32, 128})
      for (int per_stage : {8, 16})
        for (int fi = 0; fi < 8; ++fi) {
          if (uniform && (flagsets[fi] & 8)) continue;
          PArgs a;
          a.n = n; a.per_stage = per_stage; a.stages_total = 2048 / per_stage; a.flags = flagsets[fi]; a.out = d_out;
          a.uniform = uniform;
          long long h[2 * 148];
          issue_kernel<<<148, 256, 160 * 1024>>>(a);
          CK(cudaDeviceSynchronize());
          issue_kernel<<<148, 256, 160 * 1024>>>(a);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
          long long mx = 0, mi = 0;
          for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; mi = h[148 + i] > mi ? h[148 + i] : mi; }
          printf("%s N=%3d mma/stage=%2d %-24s %7.1f cyc/mma (issue %6.1f)\n", uniform ? "uniform" : "lane0  ", n, per_stage,
                 names[fi], (double)mx / 2048, (double)mi / 2048);
          fflush(stdout);
        }
  return 0;
}
