// Lightcone stitching on the device: Gaussian-edge weight map and weighted accumulation of painted
// tiles into a plane (reference baryon_painter/process_SLICS.py:85-99 make_weight_map and :211-220
// `painted_plane[slice] += w*tile; weight_plane[slice] += w; plane = painted/weight`), float64.
#include <math.h>

#include "bp_common.h"

namespace bp {

// 1-D profile of make_weight_map: rows i < fp are scaled by exp(-0.5 (fp-i)^2 / (fp*sigma)^2), rows
// counted from the far edge likewise; both factors apply when the ramps overlap.
__device__ __forceinline__ double edge_profile(int i, int T, int fp, double inv_s2) {
  double g = 1.0;
  if (i < fp) { const double d = (double)(fp - i); g *= exp(-0.5 * d * d * inv_s2); }
  const int k = T - 1 - i;
  if (k < fp) { const double d = (double)(fp - k); g *= exp(-0.5 * d * d * inv_s2); }
  return g;
}

__global__ void __launch_bounds__(256) stitch_kernel(double* __restrict__ num, double* __restrict__ den, int np,
                                                     const float* __restrict__ tiles,
                                                     const int32_t* __restrict__ origins, int T, int fp,
                                                     double inv_s2) {
  const int t = blockIdx.z;
  const int y0 = origins[2 * t], x0 = origins[2 * t + 1];
  const float* tile = tiles + (size_t)t * T * T;
  const int i = blockIdx.y;
  const double gi = edge_profile(i, T, fp, inv_s2);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < T; j += gridDim.x * blockDim.x) {
    const double w = gi * edge_profile(j, T, fp, inv_s2);
    const size_t p = (size_t)(y0 + i) * np + (x0 + j);
    atomicAdd(num + p, w * (double)tile[(size_t)i * T + j]);
    atomicAdd(den + p, w);
  }
}

__global__ void stitch_finalize_kernel(const double* __restrict__ num, const double* __restrict__ den,
                                       double* __restrict__ plane, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    plane[i] = num[i] / den[i];
}

}  // namespace bp

using namespace bp;

extern "C" int bp_stitch_accumulate(double* plane_num, double* plane_den, int n_pixel_plane, const float* tiles,
                                    const int32_t* origins, int n, int tile_size, float falloff, float sigma,
                                    void* stream) {
  BP_REQUIRE(plane_num && plane_den && tiles && origins, BP_E_INVALID, "null pointer");
  BP_REQUIRE(n >= 0 && tile_size > 0 && n_pixel_plane >= tile_size, BP_E_INVALID, "bad stitch geometry");
  if (n == 0) return BP_OK;
  // int(tile_shape[0]*falloff) in double like the reference (falloff arrives as the nearest float)
  const int fp = (int)((double)tile_size * (double)falloff + 1e-6);
  const double s = (double)fp * (double)sigma;
  const double inv_s2 = fp > 0 ? 1.0 / (s * s) : 0.0;
  dim3 grid((tile_size + 255) / 256, tile_size, n);
  stitch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(plane_num, plane_den, n_pixel_plane, tiles, origins,
                                                        tile_size, fp, inv_s2);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

extern "C" int bp_stitch_finalize(const double* plane_num, const double* plane_den, double* plane, size_t n_pixels,
                                  void* stream) {
  BP_REQUIRE(plane_num && plane_den && plane, BP_E_INVALID, "null pointer");
  if (n_pixels == 0) return BP_OK;
  const int blocks = (int)((n_pixels + 255) / 256 < 148 * 8 ? (n_pixels + 255) / 256 : 148 * 8);
  stitch_finalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(plane_num, plane_den, plane, n_pixels);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}
