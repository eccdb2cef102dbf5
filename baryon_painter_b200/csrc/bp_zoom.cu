// Tile extraction for the lightcone loop on the device:
//   periodic crop of a mass / delta plane  (reference process_SLICS.py:68-83, get_tile)
//   + cubic-spline resampling to the painter's tile size  (reference :200, :213: scipy.ndimage.zoom(tile,
//     zoom=n_pixel_tile / side, mode="reflect" | "mirror"), i.e. order 3, prefilter on, grid_mode off).
// The reference runs both on one host core per tile; here a batch of tiles is cropped, prefiltered and resampled
// by three small kernels and stays on the device for the painter.
// The same kernels, with the quintic spline (two poles, six taps), project the painted planes onto the Compton-y
// map (row f2): y_map += scale * zoom(plane, order=5, mode="mirror")  (reference process_SLICS.py:12-66).
//
// scipy's algorithm, restated (scipy/ndimage/src/ni_splines.c, ni_interpolation.c; oracle/zoom_oracle.py is the
// numpy restatement the tests pin against scipy itself):
//   1. B-spline coefficients: along each axis, c *= 6; causal recursion c[i] += z c[i-1] and anticausal recursion
//      c[i] = z (c[i+1] - c[i]) with the pole z = sqrt(3) - 2, started from the exact sums of the chosen boundary
//      extension (mirror: d c b | a b c d | c b a,  reflect: d c b a | a b c d | d c b a); float64 throughout.
//   2. output sample o reads the input coordinate o (n_in - 1)/(n_out - 1); the four cubic B-spline weights of its
//      fractional part multiply coefficients floor(x) - 1 .. floor(x) + 2, out-of-range indices folded by the same
//      boundary extension; float64 accumulation, result rounded to float32.
#include <cuda_runtime.h>

#include <cmath>
#include <mutex>
#include <vector>

#include "bp_common.h"

namespace bp {

constexpr double kPole = -0.26794919243112270647;      // sqrt(3) - 2

// ---- crop with wrap-around -> float64 ---------------------------------------------------------------
template <typename T>
__global__ void zoom_crop_kernel(const T* __restrict__ plane, int ph, int pw, const int* __restrict__ origins, int side,
                                 double* __restrict__ work) {
  const int n = blockIdx.z;
  const int r0 = origins ? origins[2 * n] : 0, c0 = origins ? origins[2 * n + 1] : 0;
  const int y = blockIdx.y;
  int yy = (r0 + y) % ph;
  if (yy < 0) yy += ph;
  // four elements per thread, a block's width apart: the four loads are in flight together (one element per thread
  // left the kernel latency bound at a few hundred GB/s)
  const int xb = blockIdx.x * (4 * blockDim.x) + threadIdx.x;
  double v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = xb + k * blockDim.x;
    int xx = (c0 + x) % pw;
    if (xx < 0) xx += pw;
    v[k] = x < side ? (double)plane[(size_t)yy * pw + xx] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = xb + k * blockDim.x;
    // float64 planes are the painted planes on their way into the y-map: NaN -> 0 as create_y_map does (:54)
    if (x < side) work[((size_t)n * side + y) * side + x] = (sizeof(T) == 8 && v[k] != v[k]) ? 0.0 : v[k];
  }
}

// ---- spline prefilter along one axis: one thread per line ------------------------------------------------
// element i of line l of tile n: work[n*side*side + l*line_stride + i*elem_stride]
__global__ void zoom_filter_kernel(double* __restrict__ work, int side, int ntiles, long long line_stride, long long elem_stride,
                                   int mirror, int npoles, double pole0, double pole1) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)ntiles * side) return;
  const int n = (int)(t / side), l = (int)(t % side);
  double* c = work + (size_t)n * side * side + (size_t)l * line_stride;
  const int len = side;
  if (len < 2) return;
#define C(i) c[(size_t)(i) * elem_stride]
  double gain = 1.0;
  for (int p = 0; p < npoles; ++p) {
    const double z = p == 0 ? pole0 : pole1;
    gain *= (1.0 - z) * (1.0 - 1.0 / z);
  }
  for (int i = 0; i < len; ++i) C(i) *= gain;
  for (int p = 0; p < npoles; ++p) {
    const double z = p == 0 ? pole0 : pole1;
    // causal initialisation
    if (mirror) {
      const double z_n_1 = pow(z, (double)(len - 1));
      double z_i = z, s = C(0) + z_n_1 * C(len - 1);
      for (int i = 1; i < len - 1; ++i) {
        s += z_i * (C(i) + z_n_1 * C(len - 1 - i));
        z_i *= z;
      }
      C(0) = s / (1.0 - z_n_1 * z_n_1);
    } else {
      const double z_n = pow(z, (double)len);
      const double c0 = C(0);
      double z_i = z, s = C(0) + z_n * C(len - 1);
      for (int i = 1; i < len; ++i) {
        s += z_i * (C(i) + z_n * C(len - 1 - i));
        z_i *= z;
      }
      C(0) = s * (z / (1.0 - z_n * z_n)) + c0;
    }
    for (int i = 1; i < len; ++i) C(i) += z * C(i - 1);
    // anticausal initialisation
    if (mirror) C(len - 1) = (z * C(len - 2) + C(len - 1)) * z / (z * z - 1.0);
    else C(len - 1) *= z / (z - 1.0);
    for (int i = len - 2; i >= 0; --i) C(i) = z * (C(i + 1) - C(i));
  }
#undef C
}

// ---- the same recursions for LONG lines, parallel along the line -------------------------------------------
// Both recursions forget their state geometrically: |z|^i drops below 1e-18 after `horizon` = ceil(ln 1e-18 / ln|z|)
// samples (32 for the cubic pole, 50 / 14 for the quintic ones).  So (1) the boundary sums of scipy's initialisation
// stop after `horizon` terms and z^(len-1) vanishes from them -- same float64 result to the last bits; (2) a line is
// cut into segments of kSeg samples, each warmed up from a zero state over the `horizon` samples in front of it
// (behind it for the anticausal sweep): (32 lines, segment) blocks are independent, which fills the machine even for
// the four 7000^2 crops of a near lightcone plane.  Sweeps are out of place (in -> out), because a segment's warm-up
// reads what its neighbour is about to overwrite.
constexpr int kSeg = 256;
__device__ __forceinline__ int zoom_horizon(double z) { return (int)ceil(-41.4465 / log(fabs(z))); }

// ---- both recursions of a direction in ONE pass over shared memory ------------------------------------------
// Block = 32 lines x one segment [i0, i1) plus a halo of Hh = sum of the poles' horizons on either side (clipped at
// the line ends).  The block loads the window, warp 0 runs, per pole, the causal and then the anticausal recursion of
// its 32 lines over the whole window in shared memory (exact scipy start / end conditions where the window touches a
// line end, a zero state elsewhere: what the halo is for) and the block writes the segment.  One read of 1.25-1.5x
// the data and one write per direction instead of two reads and two writes (the crops of a line of sight are ~20 GB per
// float64 pass).  Four blocks per SM: some load / store while the others recurse (ncu: 2.4-2.9 TB/s of DRAM traffic).
constexpr int kSweepThreads = 256;
constexpr int kSweepSeg = 128;      // 32 lines x (128 + 2 x 32) float64 = 48 KB: four blocks per SM, four recursing warps
__device__ __forceinline__ void zoom_cp8(double* dst_smem, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
// CROP (row sweep only): the window is read straight from the float32 plane with the periodic wrap of get_tile
// (`origins[n]` = top-left pixel of crop n) instead of from a float64 crop buffer: no separate crop pass
struct ZoomCrop {
  const float* plane;
  const int* origins;
  int ph, pw;
};
template <bool ROWS, bool CROP>
__global__ void __launch_bounds__(kSweepThreads) zoom_sweep_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                                   int side, int ntiles, int mirror, int npoles, double z0,
                                                                   double z1, double gain, int Hh, ZoomCrop cr) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = kSweepThreads / 32;
  const int nseg = (side + kSweepSeg - 1) / kSweepSeg, lblocks = (side + 31) / 32;
  const long long bid = blockIdx.x;
  const int j = (int)(bid % nseg), lb = (int)((bid / nseg) % lblocks), n = (int)(bid / ((long long)nseg * lblocks));
  const int l0 = lb * 32, nl = min(32, side - l0);
  const int i0 = j * kSweepSeg, i1 = min(side, i0 + kSweepSeg);
  const int a = max(0, i0 - Hh), b = min(side, i1 + Hh), len = b - a;
  const int pitch = len | 1;                       // ROWS: sm[line][i], odd pitch; columns: sm[i][line]
  const size_t base = (size_t)n * side * side;
  // ---- load: asynchronous 8-byte copies, the whole window of the block in flight at once (register-staged loads
  // left 8 KB in flight per block and the load phase latency bound)
  if (ROWS && CROP) {
    const int r0 = cr.origins[2 * n], c0 = cr.origins[2 * n + 1];
    int x0 = (c0 + a + lane) % cr.pw;
    if (x0 < 0) x0 += cr.pw;
    for (int line = warp; line < nl; line += NW) {
      int yy = (r0 + l0 + line) % cr.ph;
      if (yy < 0) yy += cr.ph;
      const float* src = cr.plane + (size_t)yy * cr.pw;
      double* dst = sm + line * pitch;
      int xx = x0;
      for (int i = lane; i < len; i += 256) {          // eight loads in flight, then their conversions
        float x[8];
        int xk = xx;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          x[k] = i + 32 * k < len ? src[xk] : 0.f;
          xk += 32;
          if (xk >= cr.pw) xk -= cr.pw;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (i + 32 * k < len) dst[i + 32 * k] = (double)x[k];
        xx = xk;
      }
    }
  } else if (ROWS) {
    for (int line = warp; line < nl; line += NW) {
      const double* src = in + base + (size_t)(l0 + line) * side + a;
      double* dst = sm + line * pitch;
      for (int i = lane; i < len; i += 32) zoom_cp8(dst + i, src + i);
    }
  } else {
    const double* src = in + base + (size_t)a * side + l0 + lane;
    if (lane < nl)
      for (int i = warp; i < len; i += NW) zoom_cp8(sm + i * 32 + lane, src + (size_t)i * side);
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- recursions: thread = line
  if (warp == 0 && lane < nl) {
    double* c = ROWS ? sm + lane * pitch : sm + lane;
    const int st = ROWS ? 1 : 32;
#define S(i) c[(i) * st]
    for (int p = 0; p < npoles; ++p) {
      const double z = p == 0 ? z0 : z1, g = p == 0 ? gain : 1.0;
      const int H = zoom_horizon(z);
      double prev;
      int i;
      if (a == 0) {
        double z_i = 1.0, sum = 0.0;
        for (int k = 0; k < H && k < len; ++k) {
          sum += z_i * (g * S(k));
          z_i *= z;
        }
        prev = mirror ? sum : sum * z + g * S(0);
        S(0) = prev;
        i = 1;
      } else {
        prev = 0.0;
        i = 0;
      }
      for (; i + 8 <= len; i += 8) {
        double x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = S(i + k);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          prev = g * x[k] + z * prev;
          S(i + k) = prev;
        }
      }
      for (; i < len; ++i) {
        prev = g * S(i) + z * prev;
        S(i) = prev;
      }
      double next;
      if (b == side) {
        const double ca = S(len - 2), cb = S(len - 1);
        next = mirror ? (z * ca + cb) * z / (z * z - 1.0) : cb * (z / (z - 1.0));
        S(len - 1) = next;
        i = len - 2;
      } else {
        next = 0.0;
        i = len - 1;
      }
      for (; i - 7 >= 0; i -= 8) {
        // z (next - x) as fma(z, next, -z x): the product z x is off the dependent chain, which is then one DFMA
        // per sample like the causal sweep (a DADD + DMUL chain made this sweep the longer of the two)
        double x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = -z * S(i - k);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          next = fma(z, next, x[k]);
          S(i - k) = next;
        }
      }
      for (; i >= 0; --i) {
        next = z * (next - S(i));
        S(i) = next;
      }
    }
#undef S
  }
  __syncthreads();
  // ---- store the segment
  const int o0 = i0 - a, no = i1 - i0;
  if (ROWS) {
    for (int line = warp; line < nl; line += NW) {
      double* dst = out + base + (size_t)(l0 + line) * side + i0;
      const double* src = sm + line * pitch + o0;
      for (int i = lane; i < no; i += 32) dst[i] = src[i];
    }
  } else if (lane < nl) {
    double* dst = out + base + (size_t)i0 * side + l0 + lane;
    for (int i = warp; i < no; i += NW) dst[(size_t)i * side] = sm[(o0 + i) * 32 + lane];
  }
}

__device__ __forceinline__ int zoom_fold(int idx, int len, int mirror) {
  if (idx >= 0 && idx < len) return idx;
  if (len <= 1) return 0;
  if (mirror) {
    const int s2 = 2 * len - 2;
    if (idx < 0) {
      idx = s2 * (-idx / s2) + idx;
      return idx <= 1 - len ? idx + s2 : -idx;
    }
    idx -= s2 * (idx / s2);
    return idx >= len ? s2 - idx : idx;
  }
  const int s2 = 2 * len;
  if (idx < 0) {
    if (idx < -s2) idx += s2 * (-idx / s2);
    return idx < -len ? idx + s2 : -idx - 1;
  }
  idx -= s2 * (idx / s2);
  return idx >= len ? s2 - idx - 1 : idx;
}

// B-spline basis weights of the ORDER + 1 coefficients around x (ORDER = 3 or 5), first coefficient index in *start
template <int ORDER>
__device__ __forceinline__ void zoom_weights(double x, int* start, double* w) {
  const double f = floor(x);
  const double t = x - f;
  *start = (int)f - ORDER / 2;
  if (ORDER == 3) {
    const double t1 = 1.0 - t;
    w[0] = t1 * t1 * t1 / 6.0;
    w[1] = (t * t * (t - 2.0) * 3.0 + 4.0) / 6.0;
    w[2] = (t1 * t1 * (t1 - 2.0) * 3.0 + 4.0) / 6.0;
    w[3] = t * t * t / 6.0;
  } else {
    // quintic B-spline beta5(y) at the distances y = |t + 2 - k|
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double y = fabs(t + 2.0 - (double)k);
      double v;
      if (y < 1.0) {
        const double y2 = y * y;
        v = y2 * (y2 * (0.25 - y / 12.0) - 0.5) + 0.55;
      } else if (y < 2.0) {
        v = y * (y * (y * (y * (y / 24.0 - 0.375) + 1.25) - 1.75) + 0.625) + 0.425;
      } else if (y < 3.0) {
        const double r = 3.0 - y, r2 = r * r;
        v = r * r2 * r2 / 120.0;
      } else {
        v = 0.0;
      }
      w[k] = v;
    }
  }
}

// ---- evaluation: one thread per output pixel -----------------------------------------------------------
// ACCUM = false: float32 tiles out[n][oy][ox] = value;  ACCUM = true: float64 map out[oy][ox] += scale * value
template <int ORDER, bool ACCUM>
__global__ void zoom_eval_kernel(const double* __restrict__ work, int side, int out_side, int mirror, void* __restrict__ out_,
                                 double scale_out) {
  constexpr int T = ORDER + 1;
  const int n = blockIdx.z;
  const int oy = blockIdx.y;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (ox >= out_side) return;
  const double scale = out_side > 1 ? (double)(side - 1) / (double)(out_side - 1) : 0.0;
  int sy, sx;
  double wy[T], wx[T];
  zoom_weights<ORDER>((double)oy * scale, &sy, wy);
  zoom_weights<ORDER>((double)ox * scale, &sx, wx);
  const double* c = work + (size_t)n * side * side;
  int ix[T];
#pragma unroll
  for (int b = 0; b < T; ++b) ix[b] = zoom_fold(sx + b, side, mirror);
  double acc = 0.0;
#pragma unroll
  for (int a = 0; a < T; ++a) {
    const double* row = c + (size_t)zoom_fold(sy + a, side, mirror) * side;
    double r = 0.0;
#pragma unroll
    for (int b = 0; b < T; ++b) r += wx[b] * row[ix[b]];
    acc += wy[a] * r;
  }
  const size_t o = ((size_t)n * out_side + oy) * out_side + ox;
  if (ACCUM) static_cast<double*>(out_)[o] += scale_out * acc;
  else static_cast<float*>(out_)[o] = (float)acc;
}

// ---- plane preprocessing: transpose + affine of the raw file contents (row f3) -----------------------------
// out[c][r] = (raw[r][c] + add) * mul, two separately rounded fp32 operations exactly as numpy's `+=` / `*=`
// (reference process_SLICS.py:157-159 mass planes, :187-189 delta planes); 32 x 32 tiles through shared memory
__global__ void plane_prepare_kernel(const float* __restrict__ raw, int rows, int cols, float add, float mul,
                                     float* __restrict__ out) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) t[j][threadIdx.x] = __fmul_rn(__fadd_rn(raw[(size_t)r * cols + c], add), mul);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = t[threadIdx.x][j];
  }
}

// grow-only float64 workspace per device
static double* zoom_workspace(int device, size_t elems) {
  static std::mutex mu;
  static std::vector<std::pair<double*, size_t>> ws(64, {nullptr, 0});
  std::lock_guard<std::mutex> lock(mu);
  auto& w = ws[device & 63];
  if (w.second < elems) {
    if (w.first) cudaFree(w.first);
    w.first = nullptr; w.second = 0;
    if (cudaMalloc(&w.first, elems * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    w.second = elems;
  }
  return w.first;
}

}  // namespace bp

using namespace bp;

namespace {
constexpr double kPole5a = -0.43057534709997381;       // sqrt(67.5 - sqrt(4436.25)) + sqrt(26.25) - 6.5
constexpr double kPole5b = -0.043096288203264652;      // sqrt(67.5 + sqrt(4436.25)) - sqrt(26.25) - 6.5

// prefilter (rows, then columns) of n cropped tiles already in `work`; `tmp` is a second buffer of the same size
// `crop`: read the tiles straight from a float32 plane (long lines only; the caller skips its crop pass)
bool zoom_fuses_crop(int side) { return side > 2 * kSeg; }
int zoom_prefilter(double* work, double* tmp, int side, int n, int order, int mirror, cudaStream_t s,
                   const ZoomCrop* crop = nullptr) {
  const long long lines = (long long)n * side;
  const int fb = 64;
  const int npoles = order == 3 ? 1 : 2;
  const double p0 = order == 3 ? kPole : kPole5a, p1 = kPole5b;
  // scipy filters axis by axis; the recursions along different axes commute
  if (side > 2 * kSeg) {     // long lines: segment-parallel out-of-place sweeps (every horizon is <= 50 < kSeg)
    double gain = 1.0;
    for (int p = 0; p < npoles; ++p) {
      const double z = p == 0 ? p0 : p1;
      gain *= (1.0 - z) * (1.0 - 1.0 / z);
    }
    // sum of the poles' horizons (host copy of zoom_horizon)
    int Hh = 0;
    for (int p = 0; p < npoles; ++p) Hh += (int)ceil(-41.4465 / log(fabs(p == 0 ? p0 : p1)));
    const int nseg = (side + kSweepSeg - 1) / kSweepSeg;
    const long long blocks = (long long)n * ((side + 31) / 32) * nseg;
    const size_t smem = sizeof(double) * 32 * (size_t)((kSweepSeg + 2 * Hh) | 1);
    static bool attr = false;
    if (!attr) {
      BP_CUDA_TRY(cudaFuncSetAttribute(zoom_sweep_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      BP_CUDA_TRY(cudaFuncSetAttribute(zoom_sweep_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      BP_CUDA_TRY(cudaFuncSetAttribute(zoom_sweep_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      attr = true;
    }
    BP_REQUIRE(smem <= 112 * 1024 && blocks < (1ll << 31), BP_E_UNSUPPORTED, "zoom prefilter: window of %zu bytes", smem);
    const ZoomCrop none{nullptr, nullptr, 0, 0};
    if (crop)
      zoom_sweep_kernel<true, true><<<(unsigned)blocks, kSweepThreads, smem, s>>>(nullptr, tmp, side, n, mirror, npoles, p0, p1, gain, Hh, *crop);
    else
      zoom_sweep_kernel<true, false><<<(unsigned)blocks, kSweepThreads, smem, s>>>(work, tmp, side, n, mirror, npoles, p0, p1, gain, Hh, none);
    zoom_sweep_kernel<false, false><<<(unsigned)blocks, kSweepThreads, smem, s>>>(tmp, work, side, n, mirror, npoles, p0, p1, gain, Hh, none);
    launch_counter() += 2;
  } else {
    zoom_filter_kernel<<<(unsigned)((lines + fb - 1) / fb), fb, 0, s>>>(work, side, n, (long long)side, 1LL, mirror, npoles, p0, p1);
    zoom_filter_kernel<<<(unsigned)((lines + fb - 1) / fb), fb, 0, s>>>(work, side, n, 1LL, (long long)side, mirror, npoles, p0, p1);
    launch_counter() += 2;
  }
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}
}  // namespace

extern "C" int bp_zoom_tiles(int device, const float* plane, int plane_h, int plane_w, const int* origins, int side, int n,
                             int out_side, int mode, float* out, void* stream) {
  BP_REQUIRE(plane && origins && out, BP_E_INVALID, "zoom_tiles: null pointer");
  BP_REQUIRE(plane_h > 0 && plane_w > 0 && side >= 2 && out_side >= 1 && n >= 0, BP_E_INVALID,
             "zoom_tiles: bad geometry (plane %dx%d, side %d, out %d, n %d)", plane_h, plane_w, side, out_side, n);
  BP_REQUIRE(mode == BP_ZOOM_REFLECT || mode == BP_ZOOM_MIRROR, BP_E_UNSUPPORTED, "zoom_tiles: boundary mode %d", mode);
  if (n == 0) return BP_OK;
  BP_CUDA_TRY(cudaSetDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* work = zoom_workspace(device, 2 * (size_t)n * side * side);
  BP_REQUIRE(work, BP_E_NOMEM, "zoom_tiles: %zu bytes of workspace", 2 * (size_t)n * side * side * sizeof(double));
  const int mirror = mode == BP_ZOOM_MIRROR ? 1 : 0;
  const ZoomCrop crop{plane, origins, plane_h, plane_w};
  const bool fused = zoom_fuses_crop(side);
  if (!fused) zoom_crop_kernel<float><<<dim3((side + 1023) / 1024, side, n), 256, 0, s>>>(plane, plane_h, plane_w, origins, side, work);
  int rc = zoom_prefilter(work, work + (size_t)n * side * side, side, n, 3, mirror, s, fused ? &crop : nullptr);
  if (rc != BP_OK) return rc;
  zoom_eval_kernel<3, false><<<dim3((out_side + 127) / 128, out_side, n), 128, 0, s>>>(work, side, out_side, mirror, out, 1.0);
  launch_counter() += 2;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

extern "C" int bp_zoom_accumulate(int device, const double* plane, int side, int out_side, int order, int mode, double scale,
                                  double* map, void* stream) {
  BP_REQUIRE(plane && map, BP_E_INVALID, "zoom_accumulate: null pointer");
  BP_REQUIRE(side >= 2 && out_side >= 1, BP_E_INVALID, "zoom_accumulate: bad geometry (side %d, out %d)", side, out_side);
  BP_REQUIRE(order == 3 || order == 5, BP_E_UNSUPPORTED, "zoom_accumulate: spline order %d (3 and 5 are implemented)", order);
  BP_REQUIRE(mode == BP_ZOOM_REFLECT || mode == BP_ZOOM_MIRROR, BP_E_UNSUPPORTED, "zoom_accumulate: boundary mode %d", mode);
  BP_CUDA_TRY(cudaSetDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* work = zoom_workspace(device, 2 * (size_t)side * side);
  BP_REQUIRE(work, BP_E_NOMEM, "zoom_accumulate: %zu bytes of workspace", 2 * (size_t)side * side * sizeof(double));
  const int mirror = mode == BP_ZOOM_MIRROR ? 1 : 0;
  zoom_crop_kernel<double><<<dim3((side + 1023) / 1024, side, 1), 256, 0, s>>>(plane, side, side, nullptr, side, work);
  int rc = zoom_prefilter(work, work + (size_t)side * side, side, 1, order, mirror, s);
  if (rc != BP_OK) return rc;
  const dim3 grid((out_side + 127) / 128, out_side, 1);
  if (order == 3) zoom_eval_kernel<3, true><<<grid, 128, 0, s>>>(work, side, out_side, mirror, map, scale);
  else zoom_eval_kernel<5, true><<<grid, 128, 0, s>>>(work, side, out_side, mirror, map, scale);
  launch_counter() += 2;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

extern "C" int bp_plane_prepare(int device, const float* raw, int rows, int cols, float add, float mul, float* out,
                                void* stream) {
  BP_REQUIRE(raw && out, BP_E_INVALID, "plane_prepare: null pointer");
  BP_REQUIRE(rows > 0 && cols > 0, BP_E_INVALID, "plane_prepare: bad shape %d x %d", rows, cols);
  BP_CUDA_TRY(cudaSetDevice(device));
  const dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  BP_REQUIRE(grid.y <= 65535, BP_E_UNSUPPORTED, "plane_prepare: %d rows", rows);
  plane_prepare_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(raw, rows, cols, add, mul, out);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}
