// Microbenchmark / semantics probe for tcgen05.mma shared-memory operand layouts on sm_100a.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_probe tools/umma_probe.cu
// (1) cycles per MMA for A/B operand layouts {none(SBO128), none(LBO128), SW32, SW64, SW128} x N
// (2) does a 128B-swizzled K-major A operand tolerate a start address shifted by whole 128-byte rows
//     (the shifted-window trick), and with which descriptor base_offset?
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../baryon_painter_b200/csrc/bp_tc.cuh"

using namespace bp::tc;

__host__ __device__ inline uint64_t desc_full(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout,
                                              uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= (uint64_t)(layout & 7u) << 61;
  return d;
}

struct TArgs {
  int n, reps;
  uint32_t a_lbo, a_sbo, a_layout, a_kstep;   // a_kstep: start address advance per K=16 step (bytes)
  uint32_t b_lbo, b_sbo, b_layout, b_kstep;
  int ksteps;                                  // distinct K steps cycled through
  int nacc;                                    // independent TMEM accumulators cycled through
  uint32_t a_off;                              // extra byte offset of every A start address
  long long* out;
};

__global__ void __launch_bounds__(128, 1) throughput_kernel(TArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(0, a.n);
    const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 64 * 1024);
    uint64_t da[4], db[4];
    uint32_t dt[4];
    for (int i = 0; i < 4; ++i) {
      da[i] = desc_full(abase + a.a_off + (i % a.ksteps) * a.a_kstep, a.a_lbo, a.a_sbo, a.a_layout, 0);
      db[i] = desc_full(bbase + (i % a.ksteps) * a.b_kstep, a.b_lbo, a.b_sbo, a.b_layout, 0);
      dt[i] = tm + (a.nacc > 1 ? (uint32_t)((i % a.nacc) * a.n) : 0u);
    }
    // first pass initialises every accumulator
    for (int i = 0; i < 4; ++i) umma_f16(dt[i], da[i], db[i], idesc, 0);
    long long t0 = clock64();
    for (int r = 0; r < a.reps; r += 4) {
      umma_f16(dt[0], da[0], db[0], idesc, 1);
      umma_f16(dt[1], da[1], db[1], idesc, 1);
      umma_f16(dt[2], da[2], db[2], idesc, 1);
      umma_f16(dt[3], da[3], db[3], idesc, 1);
    }
    long long t_issue = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    a.out[blockIdx.x] = t1 - t0;
    a.out[148 + blockIdx.x] = t_issue - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

// ---- semantics probe ---------------------------------------------------------------------------------
struct SArgs {
  uint32_t a_off;        // byte offset of the A start address from the 1024-aligned patch base
  uint32_t base_off;     // descriptor base_offset field
  uint32_t a_layout, a_lbo, a_sbo;
  float* out;            // [128][16]
};

// patch: 160 "pixels" x 64 channels (128 B rows), value(p, c) = p + c/64, stored with the 128B swizzle
// (16-byte chunk index XOR (p % 8)) -- what a TMA SWIZZLE_128B copy (or a pre-swizzled global tensor
// copied linearly) would produce.  B = no-swizzle identity: D[m][n] = A[m][n].
__global__ void __launch_bounds__(128, 1) semantics_kernel(SArgs a, int swz_bits) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __half* P = reinterpret_cast<__half*>(smem);
  const int row_bytes = 16 << swz_bits;                 // 128 (SW128), 64 (SW64), 32 (SW32)
  const int chunks = row_bytes / 16;
  for (int i = tid; i < 200 * chunks * 8; i += 128) {
    const int p = i / (chunks * 8), rem = i % (chunks * 8), c = rem / 8, e = rem % 8;
    // swizzle: chunk index ^= (row index within the repeating pattern) masked to the swizzle width
    const int rowphase = (p * row_bytes / 128) & 7;     // address bits [7:9] of the row start
    const int cs = c ^ (rowphase & (chunks - 1));
    P[(size_t)p * (row_bytes / 2) + cs * 8 + e] = __float2half((float)((p * 64 + c * 8 + e) % 2003));
  }
  __half* B = reinterpret_cast<__half*>(smem + 64 * 1024);   // no-swizzle [chunk 0..1][n 0..15][8]
  for (int i = tid; i < 2 * 16 * 8; i += 128) {
    const int c = i / 128, n = (i / 8) % 16, e = i % 8;
    B[i] = __float2half((c * 8 + e) == n ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 32); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(0, 16);
    const uint64_t da = desc_full(smem_u32(smem) + a.a_off, a.a_lbo, a.a_sbo, a.a_layout, a.base_off);
    const uint64_t db = desc_full(smem_u32(B), 16 * 16, 128, 0, 0);
    umma_f16(tm, da, db, idesc, 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v[8];
  for (int g = 0; g < 2; ++g) {
    tmem_ld8(tm + ((uint32_t)(warp * 32) << 16) + g * 8, v);
    tmem_ld_wait();
    for (int e = 0; e < 8; ++e) a.out[(warp * 32 + lane) * 16 + g * 8 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 32); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  CK(cudaFuncSetAttribute(throughput_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  CK(cudaFuncSetAttribute(semantics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  long long* d_out;
  CK(cudaMalloc(&d_out, 2 * 148 * sizeof(long long)));
  const int reps = 2048;
  printf("== cycles per tcgen05.mma (M=128, K=16, f16), %d back-to-back, 1 CTA and 148 CTAs ==\n", reps);
  struct L { const char* name; uint32_t lbo_a, sbo, layout, kstep_a; int ksteps; bool lbo_is_n; };
  // lbo for "none/SBO128": A: 128 rows*16 = 2048 ; B: N*16.  K step (2 chunks) advances 2*LBO.
  for (int layout = 0; layout < 5; layout += 4) {
    for (int n : {16, 32, 64, 128, 256}) {
      TArgs a;
      memset(&a, 0, sizeof(a));
      a.n = n; a.reps = reps; a.out = d_out;
      const char* name = "";
      if (layout == 0) {        // no swizzle, rows at 16 B pitch (SBO 128), chunks LBO apart
        name = "none  SBO=128 (rows dense)";
        a.a_lbo = 2048; a.a_sbo = 128; a.a_layout = 0; a.a_kstep = 4096;
        a.b_lbo = n * 16; a.b_sbo = 128; a.b_layout = 0; a.b_kstep = 2 * n * 16; a.ksteps = 4;
      } else if (layout == 1) { // no swizzle, the two K chunks adjacent (LBO 128), 8-row groups 256 B apart
        name = "none  LBO=128 SBO=256";
        a.a_lbo = 128; a.a_sbo = 256; a.a_layout = 0; a.a_kstep = 4096;
        a.b_lbo = 128; a.b_sbo = 256; a.b_layout = 0; a.b_kstep = (n / 8) * 256; a.ksteps = 4;
      } else if (layout == 2) { // SW32: rows of 32 B, 8-row groups 256 B
        name = "SWIZZLE_32B  (32 B rows)";
        a.a_lbo = 16; a.a_sbo = 256; a.a_layout = 6; a.a_kstep = 4096;
        a.b_lbo = 16; a.b_sbo = 256; a.b_layout = 6; a.b_kstep = (n / 8) * 256; a.ksteps = 4;
      } else if (layout == 3) { // SW64: rows of 64 B, 2 K steps per row
        name = "SWIZZLE_64B  (64 B rows)";
        a.a_lbo = 16; a.a_sbo = 512; a.a_layout = 4; a.a_kstep = 32;
        a.b_lbo = 16; a.b_sbo = 512; a.b_layout = 4; a.b_kstep = 32; a.ksteps = 2;
      } else {                  // SW128: rows of 128 B, 4 K steps per row
        name = "SWIZZLE_128B (128 B rows)";
        a.a_lbo = 16; a.a_sbo = 1024; a.a_layout = 2; a.a_kstep = 32;
        a.b_lbo = 16; a.b_sbo = 1024; a.b_layout = 2; a.b_kstep = 32; a.ksteps = 4;
      }
      for (int nacc : {1, 2}) {
        if (nacc * n > 256) continue;
        a.nacc = nacc;
        long long h[2 * 148];
        double res[2], iss[2];
        for (int cfg = 0; cfg < 2; ++cfg) {
          const int grid = cfg ? 148 : 1;
          throughput_kernel<<<grid, 128, 128 * 1024>>>(a);
          CK(cudaDeviceSynchronize());
          throughput_kernel<<<grid, 128, 128 * 1024>>>(a);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h, d_out, 2 * 148 * sizeof(long long), cudaMemcpyDeviceToHost));
          long long mx = 0, mi = 0;
          for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mi = h[148 + i] > mi ? h[148 + i] : mi; }
          res[cfg] = (double)mx / reps;
          iss[cfg] = (double)mi / reps;
        }
        printf("%-28s N=%3d acc=%d  %7.1f cyc/mma (1 CTA; issue %5.1f)  %7.1f (148 CTAs)   ideal %5.1f\n", name, n, nacc,
               res[0], iss[0], res[1], 128.0 * n / 256.0);
      }
    }
  }


  printf("== A start-address alignment / LBO (N=128 and N=32, acc=2, 148 CTAs) ==\n");
  {
    struct C { const char* name; uint32_t layout, lbo, sbo, kstep, off; int ksteps; };
    const C cs[] = {
      {"none dense  off=0    LBO=2048", 0, 2048, 128, 4096, 0, 4},
      {"none dense  off=16   LBO=2048", 0, 2048, 128, 4096, 16, 4},
      {"none dense  off=48   LBO=2048", 0, 2048, 128, 4096, 48, 4},
      {"none dense  off=64   LBO=2048", 0, 2048, 128, 4096, 64, 4},
      {"none dense  off=0    LBO=9568", 0, 9568, 128, 19136, 0, 2},
      {"none dense  off=1072 LBO=9568", 0, 9568, 128, 19136, 1072, 2},
      {"none dense  off=0    LBO=16 (tap pair)", 0, 16, 128, 32, 0, 4},
      {"none dense  off=16   LBO=1056 (tap pair)", 0, 1056, 128, 32, 16, 4},
      {"SW128 off=0", 2, 16, 1024, 32, 0, 4},
      {"SW128 off=128 (1 row)", 2, 16, 1024, 32, 128, 4},
      {"SW128 off=384 (3 rows)", 2, 16, 1024, 32, 384, 4},
      {"SW128 off=8448+640 (66+5 rows)", 2, 16, 1024, 32, 9088, 4},
      {"SW64  off=0", 4, 16, 512, 32, 0, 2},
      {"SW64  off=192 (3 rows)", 4, 16, 512, 32, 192, 2},
      {"SW32  off=0", 6, 16, 256, 4096, 0, 4},
      {"SW32  off=96 (3 rows)", 6, 16, 256, 4096, 96, 4},
    };
    for (const C& c : cs) {
      for (int n : {32, 128}) {
        TArgs a;
        memset(&a, 0, sizeof(a));
        a.n = n; a.reps = reps; a.out = d_out; a.nacc = 2;
        a.a_lbo = c.lbo; a.a_sbo = c.sbo; a.a_layout = c.layout; a.a_kstep = c.kstep; a.a_off = c.off; a.ksteps = c.ksteps;
        a.b_lbo = n * 16; a.b_sbo = 128; a.b_layout = 0; a.b_kstep = 2 * n * 16;
        long long h[2 * 148];
        throughput_kernel<<<148, 128, 128 * 1024>>>(a);
        CK(cudaDeviceSynchronize());
        throughput_kernel<<<148, 128, 128 * 1024>>>(a);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_out, 2 * 148 * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-44s N=%3d  %7.1f cyc/mma\n", c.name, n, (double)mx / reps);
      }
    }
  }
  printf("== shifted start address on a swizzled K-major A operand ==\n");
  float* d_res;
  CK(cudaMalloc(&d_res, 128 * 16 * sizeof(float)));
  std::vector<float> hres(128 * 16);
  struct SW { const char* name; int bits; uint32_t layout, sbo; };
  const SW sws[3] = {{"SW128", 3, 2, 1024}, {"SW64", 2, 4, 512}, {"SW32", 1, 6, 256}};
  for (const SW& sw : sws) {
    const int row_bytes = 16 << sw.bits;
    for (int shift : {0, 1, 2, 3, 5, 8, 11}) {
      for (int kadv = 0; kadv < (sw.bits == 1 ? 1 : 2); ++kadv) {
        for (int variant = 0; variant < 2; ++variant) {
          SArgs s;
          s.a_off = shift * row_bytes + kadv * 32;
          const uint32_t phase = ((shift * row_bytes) >> 7) & 7;
          s.base_off = variant ? phase : 0;
          s.a_layout = sw.layout; s.a_lbo = 16; s.a_sbo = sw.sbo; s.out = d_res;
          semantics_kernel<<<1, 128, 128 * 1024>>>(s, sw.bits);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s shift %d: CUDA error %s\n", sw.name, shift, cudaGetErrorString(e)); return 1; }
          CK(cudaMemcpy(hres.data(), d_res, hres.size() * sizeof(float), cudaMemcpyDeviceToHost));
          int bad = 0;
          float first_got = 0, first_exp = 0;
          int first_m = -1, first_n = -1;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
              const int chn = kadv * 16 + n;
              const float exp_v = (float)(((m + shift) * 64 + chn) % 2003);
              const float got = hres[m * 16 + n];
              if (fabsf(got - exp_v) > 0.02f) {
                if (!bad) { first_got = got; first_exp = exp_v; first_m = m; first_n = n; }
                ++bad;
              }
            }
          printf("%-5s shift %2d rows, k-adv %d, base_offset=%u : %s (%d/2048 wrong", sw.name, shift, kadv, s.base_off,
                 bad ? "MISMATCH" : "ok", bad);
          if (bad) printf("; first D[%d][%d] = %.3f expected %.3f", first_m, first_n, first_got, first_exp);
          printf(")\n");
          if (phase == 0) break;   // both variants identical
        }
      }
    }
  }
  return 0;
}
