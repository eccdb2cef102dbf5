/*
 * baryon_painter_b200 -- C ABI of the B200 paint path.
 *
 * This is the boundary a maintainer of tilmantroester/baryon_painter binds to replace
 * the torch modules underneath `CVAEPainter.paint()` (reference
 * baryon_painter/painter.py:371-392) and the external `GAN_Painter.paint()`
 * (reference scripts/create_lightcone.py:41-54).  Plain pointers and sizes only; no
 * torch / Python types.  Every entry point returns BP_OK (0) or a negative BP_E_* code;
 * bp_last_error() gives the message for the calling thread.  There is no CPU fallback:
 * without a CUDA device of compute capability 10.x every create call fails with
 * BP_E_NO_DEVICE.
 *
 * Layout conventions
 *   tiles      float32 [n][H][W]        one dark-matter / pressure tile per sample
 *   latents    float32 [n][h][w]        h = H/32, w = W/32 for the fiducial CVAE
 *   weights    PyTorch layout, float32: conv (Cout,Cin,k,k); transposed conv (Cin,Cout,k,k)
 * "device" entry points take device pointers valid on the net's device and enqueue on
 * `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) without
 * synchronising.  "_host" entry points take host pointers, stage through pinned memory,
 * and return after the result is in `out`.
 */
#ifndef BARYON_PAINTER_B200_H
#define BARYON_PAINTER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_VERSION 100

/* status codes (mapped to Python exceptions by baryon_painter_b200/_lib.py) */
enum {
  BP_OK = 0,
  BP_E_INVALID = -1,      /* bad argument / shape mismatch        -> ValueError          */
  BP_E_UNSUPPORTED = -2,  /* layer or option without a kernel     -> NotImplementedError */
  BP_E_CUDA = -3,         /* CUDA runtime error                    -> RuntimeError        */
  BP_E_NO_DEVICE = -4,    /* no sm_100 device / wrong arch         -> RuntimeError        */
  BP_E_NOMEM = -5         /* host or device allocation failed      -> MemoryError         */
};

/* layer kinds: torch.nn.Conv2d / torch.nn.ConvTranspose2d as instantiated by
 * build_sequential (reference baryon_painter/models/utils.py:128-131) */
enum { BP_CONV = 0, BP_CONVT = 1 };

/* activations fused into the epilogue (reference models/utils.py:134-147) */
enum {
  BP_ACT_NONE = 0, BP_ACT_RELU = 1, BP_ACT_LEAKY = 2, BP_ACT_PRELU = 3,
  BP_ACT_SOFTPLUS = 4,   /* beta = 1, threshold = 20 (torch.nn.Softplus defaults) */
  BP_ACT_TANH = 5, BP_ACT_SIGMOID = 6
};

/* ResidualBlock membership (reference models/utils.py:22-38): the input of the OPEN
 * layer is added to the scaled/shifted output of the CLOSE layer before its activation */
enum { BP_RES_NONE = 0, BP_RES_OPEN = 1, BP_RES_CLOSE = 2 };

/* arithmetic of the convolution stacks */
enum {
  BP_PREC_F32 = 0,   /* fp32-accurate on the tensor cores: every activation and weight travels as two fp16
                        numbers (hi = fp16(v), lo = fp16(v - hi): 22 significant bits), each product is
                        x_hi*w_hi + x_lo*w_hi + x_hi*w_lo accumulated in fp32 in TMEM; <= 1e-4 rel-L2 per tile
                        vs the reference's fp32 torch.nn.Conv2d (reference models/utils.py:128-131)        */
  BP_PREC_BF16 = 1,  /* bf16 operands, fp32 accumulation in TMEM (tcgen05.mma .kind::f16)                   */
  BP_PREC_F16 = 2,   /* fp16 operands, same kernel and rate; 8x finer rounding than bf16 -- the
                        16-bit path that holds <= 1e-2 rel-L2 through the exp(4x) inverse transform        */
  BP_PREC_F32_FFMA = 3 /* scalar fp32 FFMA kernels (no tensor cores): the on-device cross-check of the others */
};

/* how the CVAE latent is obtained (reference cvae.py:63-66, 97-100, 149-155) */
enum {
  BP_LATENT_GIVEN = 0,  /* `latent` IS the latent z; the prior network is skipped (sample_P(z=...)) */
  BP_LATENT_EPS = 1,    /* `latent` is eps; z = z_mu + eps*(exp(z_log_var/2) + min_z_var)            */
  BP_LATENT_SEED = 2    /* eps drawn on the device from a counter-based normal generator (`seed`)    */
};

/* paint flags */
enum {
  BP_FLAG_TRANSFORM = 1,  /* input is raw density: apply ln(x/sigma_in + 1)/k_in - shift_in   */
  BP_FLAG_INVERSE = 2     /* output is pressure:  (exp((x + shift_out)*k_out) - 1)*sigma_out  */
};

/* one convolution with its fused epilogue:
 *   out = act( conv(in, weight) * scale + shift [+ skip] )                                  */
typedef struct bp_layer_desc {
  int32_t kind;            /* BP_CONV | BP_CONVT */
  int32_t cin, cout;
  int32_t kernel, stride, pad, out_pad;
  int32_t act;             /* BP_ACT_* */
  float act_param;         /* leaky / prelu slope */
  int32_t res;             /* BP_RES_* */
  const float* weight;     /* host, PyTorch layout */
  const float* scale;      /* host [cout] or NULL (ones): folded batch-norm gamma/sqrt(var+eps) */
  const float* shift;      /* host [cout] or NULL (zeros): folded batch-norm/bias shift         */
} bp_layer_desc;

/* the four sub-networks paint() runs (reference cvae.py:82-120): prior_network,
 * p_z_in, p_y_z_in, p_mu_out.  p_y_in is the identity in every shipped architecture. */
typedef struct bp_cvae_desc {
  int32_t tile_h, tile_w;        /* dim_y[1:], 512 x 512 for the fiducial model       */
  int32_t latent_h, latent_w;    /* dim_z[1:], 16 x 16                                */
  float min_z_var;               /* 1e-7, added to the latent std (reference cvae.py:65) */
  int32_t n_prior, n_p_z_in, n_p_y_z_in, n_p_mu_out;
  const bp_layer_desc* prior;    /* in: [y, z-plane] (2 ch) -> (z_mu, z_log_var)      */
  const bp_layer_desc* p_z_in;   /* in: latent (1 ch) -> 1 ch at tile resolution      */
  const bp_layer_desc* p_y_z_in; /* in: [p_z_in(latent), y, z-plane] (3 ch)           */
  const bp_layer_desc* p_mu_out; /* -> 1 ch x_mu                                      */
  /* recognition network (reference cvae.py:24-26, 68-80); optional (n_* = 0): needed by bp_cvae_elbo_host only */
  int32_t n_q_x_in, n_q_y_in, n_q_out;
  const bp_layer_desc* q_x_in;   /* in: x (1 ch)                                      */
  const bp_layer_desc* q_y_in;   /* in: [y, z-plane] (2 ch)                           */
  const bp_layer_desc* q_out;    /* in: cat(q_x_in, q_y_in) -> (z_mu, z_log_var)      */
  float likelihood_scaling;      /* reference cvae.py:56, default 1                   */
} bp_cvae_desc;

/* per-call elementwise transform parameters (host arrays of length n; sigma computed on
 * the host in float64 like reference data_transforms.py:52-64 and rounded to float32) */
typedef struct bp_transform_params {
  const float* sigma_in;   /* [n] sqrt(var_dm(z_i))       */
  const float* sigma_out;  /* [n] sqrt(var_pressure(z_i)) */
  const float* aux;        /* [n] value of the constant redshift plane (CVAE: z, CGAN: z-1) */
  float k_in, shift_in;    /* 4, 0 (CVAE shift-log);  4, 1 (CGAN shift-log-cam) */
  float k_out, shift_out;
} bp_transform_params;

typedef struct bp_net bp_net;   /* one network resident on one device */

/* ---- lifetime ------------------------------------------------------------------------ */
int bp_device_count(void);
/* replaces CVAE.__init__ + load_state_dict (reference painter.py:431-432) */
int bp_cvae_create(const bp_cvae_desc* desc, int precision, int max_batch, int device, bp_net** out);
/* replaces GAN_Painter's generator construction (reference create_lightcone.py:52-54) */
int bp_cgan_create(const bp_layer_desc* layers, int n_layers, int tile_h, int tile_w,
                   int precision, int max_batch, int device, bp_net** out);
void bp_net_destroy(bp_net* net);

/* ---- painting ------------------------------------------------------------------------- */
/* replaces CVAE.sample_P + the numpy transforms around it (reference painter.py:375-390),
 * batched over n tiles.  `latent` may be NULL for BP_LATENT_SEED. */
int bp_cvae_paint(bp_net* net, const float* tiles, const float* latent, int latent_mode,
                  uint64_t seed, const bp_transform_params* tp, int flags, float* out, int n,
                  void* stream);
int bp_cvae_paint_host(bp_net* net, const float* tiles, const float* latent, int latent_mode,
                       uint64_t seed, const bp_transform_params* tp, int flags, float* out, int n);
/* The same for a STREAM of batches (reference: a loop of paint() calls over the tiles of a plane, process_SLICS.py:201-218):
 * returns once the batch is enqueued on `slot` (0, 1 or 2); bp_net_wait(net, slot) returns when `out` holds the result.
 * Using the slots in turn and keeping two batches outstanding overlaps the upload of batch k+1 and the download of batch
 * k-1 with the kernels of batch k.
 * tiles / latent / out must stay valid until the wait and should be page-locked. */
int bp_cvae_paint_host_async(bp_net* net, const float* tiles, const float* latent, int latent_mode,
                             uint64_t seed, const bp_transform_params* tp, int flags, float* out, int n, int slot);
int bp_net_wait(bp_net* net, int slot);
/* (z_mu, z_log_var) of the last bp_cvae_paint* call in BP_LATENT_EPS/SEED mode: host [n][h][w] each */
int bp_cvae_read_prior(bp_net* net, float* z_mu, float* z_log_var, int n);
/* replaces GAN_Painter.paint (generator forward + transforms) */
int bp_cgan_paint(bp_net* net, const float* tiles, const bp_transform_params* tp, int flags,
                  float* out, int n, void* stream);
int bp_cgan_paint_host(bp_net* net, const float* tiles, const bp_transform_params* tp, int flags,
                       float* out, int n);
/* n_draws latent draws per tile -> per-pixel mean and (population) variance of the painted
 * pressure; host pointers (BASELINE config 4; reference: repeated paint() calls) */
int bp_cvae_paint_variance_host(bp_net* net, const float* tiles, const bp_transform_params* tp,
                                int n_draws, uint64_t seed, float* mean_out, float* var_out, int n);

/* replaces CVAE.forward with CVAE.Q (reference cvae.py:68-80, 122-147): the evidence lower bound of a batch of n
 * (x = pressure, y = dark matter) tile pairs, as CVAEPainter.validate evaluates it (reference painter.py:295-367).
 *   z ~ Q(x, y): (z_mu, z_log_var) = q_out(cat(q_x_in(x'), q_y_in([y', z]))), z = z_mu + eps*(exp(z_log_var/2) + min_z_var)
 *   KL_term = 0.5/n * sum((pm - z_mu)^2/pv + exp(z_log_var)/pv + plv - z_log_var - 1) over the prior network's (pm, plv)
 *   log_likelihood = -0.5 ln(2 pi) - 0.5/n * sum((x' - x_mu(z, y))^2)          (fixed variance, L = 1)
 *   ELBO = -KL_term + likelihood_scaling * log_likelihood
 * With BP_FLAG_TRANSFORM x' = ln(x/sigma_out + 1)/k_out - shift_out and y' = ln(y/sigma_in + 1)/k_in - shift_in (the
 * dataset transforms of reference scripts/CVAE_single_scale.py:34-65); eps as in bp_cvae_paint (BP_LATENT_EPS / _SEED).
 * Host pointers; stats = {ELBO, KL_term, log_likelihood}; z_mu / z_log_var ([n][h][w], optional) receive Q's output. */
int bp_cvae_elbo_host(bp_net* net, const float* x_tiles, const float* y_tiles, const float* eps, int latent_mode,
                      uint64_t seed, const bp_transform_params* tp, int flags, int n, double* stats, float* z_mu,
                      float* z_log_var);

/* ---- lightcone stitching (reference process_SLICS.py:85-99, 211-220) -------------------- */
/* plane_num[y0+i][x0+j] += w(i,j)*tile[t][i][j]; plane_den[...] += w(i,j) for every tile t;
 * device pointers, float64 planes of n_pixel_plane^2, origins int32 [n][2] = (row0, col0) */
int bp_stitch_accumulate(double* plane_num, double* plane_den, int n_pixel_plane,
                         const float* tiles, const int32_t* origins, int n, int tile_size,
                         float falloff, float sigma, void* stream);
/* plane[i] = num[i]/den[i] */
int bp_stitch_finalize(const double* plane_num, const double* plane_den, double* plane,
                       size_t n_pixels, void* stream);

/* ---- lightcone tile extraction (reference process_SLICS.py:68-83 get_tile + :200/:213 scipy.ndimage.zoom) ---- */
enum { BP_ZOOM_REFLECT = 0, BP_ZOOM_MIRROR = 1 };   /* scipy boundary modes "reflect" / "mirror" */
/* n tiles: periodic side x side crop of the float32 plane (plane_h x plane_w) at origins[t] = (row0, col0), then
 * cubic-spline resampling (order 3, prefilter, grid_mode off -- scipy.ndimage.zoom defaults) to out_side x out_side.
 * plane, origins (int32 [n][2]) and out (float32 [n][out_side][out_side]) are device pointers on `device`. */
int bp_zoom_tiles(int device, const float* plane, int plane_h, int plane_w, const int32_t* origins, int side, int n,
                  int out_side, int mode, float* out, void* stream);
/* Compton-y projection step (reference process_SLICS.py:12-66, create_y_map loop body):
 * map[out_side][out_side] += scale * scipy.ndimage.zoom(where(isnan(plane), 0, plane), out_side/side, order, mode)
 * for one float64 side x side painted plane; order 3 or 5; plane and map are device pointers. */
int bp_zoom_accumulate(int device, const double* plane, int side, int out_side, int order, int mode, double scale,
                       double* map, void* stream);
/* Plane preprocessing (reference process_SLICS.py:157-159, :187-189): the raw file is rows x cols float32, the
 * reference takes `.T`, then `+= add`, then `*= mul` in float32:  out[c][r] = (raw[r][c] + add) * mul
 * (two roundings, no fused multiply-add).  raw (after the caller skipped any header word) and out are device pointers. */
int bp_plane_prepare(int device, const float* raw, int rows, int cols, float add, float mul, float* out, void* stream);

/* ---- formulation table (determinism) ---------------------------------------------------------------
 * Each convolution has several tensor-core formulations (Toeplitz packings, tilings); which one runs is looked
 * up in a table (text, one layer per line: "key G Jy N mode Wt T_r gl nbst [ms]") so that every process paints
 * bit-identical tiles; a layer without an entry gets the cost model's first choice (deterministic).  The package
 * installs baryon_painter_b200/tuning_table.txt at load.  bp_tuning_mode(1, log) makes subsequently created nets
 * time the candidates on the device instead and record the winners (python -m baryon_painter_b200.tune);
 * bp_tuning_get returns the table (needed length excluding NUL).  No reference counterpart: torch picks cuDNN
 * algorithms per process (torch.backends.cudnn.benchmark) and is not bit-reproducible across algorithms either. */
int bp_tuning_set(const char* table_text);
int bp_tuning_get(char* buf, size_t cap);
int bp_tuning_mode(int on, int log);
/* n standard-normal draws of the BP_LATENT_SEED generator (counter-based: splitmix64(seed ^ splitmix64(offset + i))
 * -> Box-Muller) to a host array -- the eps that replace torch.randn of reference cvae.py:64; lets tests check the
 * distribution and restate the draws for the variance-map oracle */
int bp_rng_normal_host(int device, uint64_t seed, uint64_t offset, float* out, size_t n);

/* ---- introspection ---------------------------------------------------------------------- */
/* copy the activation after layer `layer` of sub-network `stack` (0 prior, 1 p_z_in, 2 p_y_z_in,
 * 3 p_mu_out; CGAN: 0) of the last paint call to host as float32 [n][C][H][W]; needs
 * bp_net_set_debug(net, 1) before the paint call.  Used by the layer-boundary parity tests. */
int bp_net_set_debug(bp_net* net, int keep_activations);
int bp_net_read_activation(bp_net* net, int stack, int layer, float* out, size_t out_floats);
/* per-layer timing: while on, every layer launch is bracketed by CUDA events on the launching
 * stream (adds launch gaps -- use for attribution, never for throughput numbers) */
int bp_net_set_profile(bp_net* net, int on);
int bp_net_read_profile(bp_net* net, int stack, int layer, double* total_ms, int* launches);
/* flops = 2*MACs per tile; geom[10] = {kind, cin, cout, kernel, stride, H, W, OH, OW, on_tensor_cores} */
int bp_net_layer_info(const bp_net* net, int stack, int layer, double* flops, int* geom);
/* kernels launched by this library on this thread since the last reset */
int64_t bp_launch_count(int reset);
/* tiles processed per internal launch group (layer outputs of one chunk are kept L2-/HBM-resident together) */
int bp_net_chunk(const bp_net* net);
/* algorithmic FLOPs (2*MACs) per tile of the network's convolutions */
double bp_net_flops_per_tile(const bp_net* net);
const char* bp_last_error(void);
int bp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BARYON_PAINTER_B200_H */
