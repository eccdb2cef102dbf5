// Host side of the CUDA library: weight packing, the per-network execution plan and the C ABI
// declared in include/baryon_painter_b200.h.
//
// Replaces, per reference call site:
//   CVAE.__init__ / load_state_dict       baryon_painter/models/cvae.py:9-61, painter.py:431-432
//   CVAE.prior / sample_prior / P / sample_P   baryon_painter/models/cvae.py:82-120, 149-162
//   CVAEPainter.paint (device part)       baryon_painter/painter.py:375-390
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <new>
#include <sstream>
#include <string>

#include "bp_common.h"
#include "bp_wconv.h"
#include "bp_front.h"

namespace bp {

static thread_local char g_err[1024] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int64_t& launch_counter() { return g_launches; }

// ------------------------------------------------------------------------------------------
// packing: PyTorch weight layout -> per-phase (channel, tap) tables + [k][n] weight matrices
// ------------------------------------------------------------------------------------------
static int npad_for(int cout) {
  if (cout <= 2) return cout;
  if (cout <= 8) return 8;
  if (cout <= 16) return 16;
  if (cout <= 32) return 32;
  if (cout <= 64) return 64;
  return ((cout + 127) / 128) * 128;
}

int pack_layer(const bp_layer_desc& d, int H, int W, Layer* out) {
  Layer& l = *out;
  l.d = d;
  l.H = H; l.W = W;
  BP_REQUIRE(d.kind == BP_CONV || d.kind == BP_CONVT, BP_E_UNSUPPORTED, "unknown layer kind %d", d.kind);
  BP_REQUIRE(d.cin > 0 && d.cout > 0 && d.kernel > 0 && d.stride > 0 && d.pad >= 0 && d.out_pad >= 0,
             BP_E_INVALID, "bad convolution geometry");
  BP_REQUIRE(d.weight != nullptr, BP_E_INVALID, "layer without weights");
  l.host_weight.assign(d.weight, d.weight + (size_t)d.cin * d.cout * d.kernel * d.kernel);
  l.d.weight = nullptr; l.d.scale = nullptr; l.d.shift = nullptr;
  const int k = d.kernel, s = d.stride, p = d.pad;
  std::vector<int4> ktab;
  std::vector<float> wmat;
  l.npad = npad_for(d.cout);
  int rows = 0;
  if (d.kind == BP_CONV) {
    BP_REQUIRE(d.out_pad == 0, BP_E_INVALID, "output_padding on a forward convolution");
    l.OHF = (H + 2 * p - k) / s + 1;
    l.OWF = (W + 2 * p - k) / s + 1;
    BP_REQUIRE(l.OHF > 0 && l.OWF > 0, BP_E_INVALID, "convolution output is empty");
    l.OH = l.OHF; l.OW = l.OWF; l.istride = s; l.os = 1; l.nphase = 1;
    l.phase[0] = PhaseDev{0, d.cin * k * k, 0, 0};
    rows = d.cin * k * k;
    ktab.resize(rows);
    wmat.assign((size_t)rows * l.npad, 0.f);
    for (int c = 0; c < d.cin; ++c)
      for (int r = 0; r < k; ++r)
        for (int q = 0; q < k; ++q) {
          const int row = (c * k + r) * k + q;
          ktab[row] = make_int4(c * H * W + (r - p) * W + (q - p), r - p, q - p, c);
          for (int n = 0; n < d.cout; ++n)
            wmat[(size_t)row * l.npad + n] = d.weight[(((size_t)n * d.cin + c) * k + r) * k + q];
        }
    l.flops = 2.0 * l.OHF * l.OWF * d.cin * d.cout * k * k;
  } else {
    BP_REQUIRE(s <= 4, BP_E_UNSUPPORTED, "transposed convolution stride %d > 4", s);
    l.OHF = (H - 1) * s - 2 * p + k + d.out_pad;
    l.OWF = (W - 1) * s - 2 * p + k + d.out_pad;
    BP_REQUIRE(l.OHF > 0 && l.OHF % s == 0 && l.OWF % s == 0, BP_E_UNSUPPORTED,
               "transposed convolution output %dx%d not a multiple of the stride", l.OHF, l.OWF);
    l.OH = l.OHF / s; l.OW = l.OWF / s; l.istride = 1; l.os = s; l.nphase = s * s;
    for (int ph = 0; ph < s; ++ph)
      for (int pw = 0; pw < s; ++pw) {
        PhaseDev& P = l.phase[ph * s + pw];
        P.k_begin = rows; P.ph = ph; P.pw = pw; P.K = 0;
        const int r0 = (ph + p) % s, qh = (ph + p) / s;
        const int c0 = (pw + p) % s, qw = (pw + p) / s;
        for (int c = 0; c < d.cin; ++c)
          for (int a = 0; r0 + s * a < k; ++a)
            for (int b = 0; c0 + s * b < k; ++b) {
              const int r = r0 + s * a, q = c0 + s * b;
              const int dr = qh - a, ds = qw - b;
              ktab.push_back(make_int4(c * H * W + dr * W + ds, dr, ds, c));
              wmat.resize(wmat.size() + l.npad, 0.f);
              float* wr = wmat.data() + (size_t)rows * l.npad;
              for (int n = 0; n < d.cout; ++n) wr[n] = d.weight[(((size_t)c * d.cout + n) * k + r) * k + q];
              ++rows; ++P.K;
            }
      }
    l.flops = 2.0 * H * W * d.cin * d.cout * k * k;
  }
  l.Kmax = 0;
  for (int i = 0; i < l.nphase; ++i) l.Kmax = std::max(l.Kmax, l.phase[i].K);
  if (ktab.empty()) {  // degenerate phase set (cannot happen for k >= s) -- keep allocations valid
    ktab.push_back(make_int4(0, 1 << 20, 1 << 20, 0));
    wmat.assign(l.npad, 0.f);
  }
  std::vector<float> scale(d.cout, 1.f), shift(d.cout, 0.f);
  if (d.scale) memcpy(scale.data(), d.scale, sizeof(float) * d.cout);
  if (d.shift) memcpy(shift.data(), d.shift, sizeof(float) * d.cout);
  l.host_scale = scale; l.host_shift = shift;
  BP_CUDA_TRY(cudaMalloc(&l.ktab, ktab.size() * sizeof(int4)));
  BP_CUDA_TRY(cudaMalloc(&l.wmat, wmat.size() * sizeof(float)));
  BP_CUDA_TRY(cudaMalloc(&l.scale, d.cout * sizeof(float)));
  BP_CUDA_TRY(cudaMalloc(&l.shift, d.cout * sizeof(float)));
  BP_CUDA_TRY(cudaMemcpy(l.ktab, ktab.data(), ktab.size() * sizeof(int4), cudaMemcpyHostToDevice));
  BP_CUDA_TRY(cudaMemcpy(l.wmat, wmat.data(), wmat.size() * sizeof(float), cudaMemcpyHostToDevice));
  BP_CUDA_TRY(cudaMemcpy(l.scale, scale.data(), d.cout * sizeof(float), cudaMemcpyHostToDevice));
  BP_CUDA_TRY(cudaMemcpy(l.shift, shift.data(), d.cout * sizeof(float), cudaMemcpyHostToDevice));
  return BP_OK;
}

void free_layer(Layer* l) {
  cudaFree(l->ktab); cudaFree(l->wmat); cudaFree(l->scale); cudaFree(l->shift);
  l->ktab = nullptr; l->wmat = nullptr; l->scale = nullptr; l->shift = nullptr;
}

}  // namespace bp

using namespace bp;

// ------------------------------------------------------------------------------------------
// the network object
// ------------------------------------------------------------------------------------------
struct Stack {
  std::vector<Layer> layers;
  int in_c = 0, H = 0, W = 0;        // input
  int out_c = 0, OH = 0, OW = 0;     // output
  size_t max_floats = 0;             // largest intermediate activation per sample
};

enum { NET_CVAE = 0, NET_CGAN = 1 };
enum { ST_PRIOR = 0, ST_PZ = 1, ST_PYZ = 2, ST_MU = 3, ST_QX = 4, ST_QY = 5, ST_QOUT = 6, ST_GEN = 0 };
constexpr int kMaxStacks = 7;

// ---- execution plan of the 16-bit window-GEMM engine (bp_wconv.cu) ------------------------------
enum { V2_WCONV = 0, V2_F32CONV = 1, V2_TO_NHWC16 = 2, V2_TO_F32 = 3, V2_TAIL = 4, V2_COPY_F32 = 5 };
struct V2Op {
  int kind = V2_WCONV;
  int stack = -1, index = -1;     // layer this op executes (-1: layout conversion)
  Layer* l = nullptr;
  WLayer* w = nullptr;
  int in = -1, out = -1, skip = -1;
  bool final = false;             // writes the caller's output buffer (with the fused inverse transform)
};
struct V2Plan {
  bool built = false;
  std::vector<ActDesc> acts;
  std::vector<V2Op> ops;          // decoder: p_y_z_in + p_mu_out
  std::vector<V2Op> prior_ops;    // prior_network (when it lowers to window GEMMs)
  std::vector<void*> owned;
  bool front_on = false;          // fused transform + p_z_in + NHWC pack feeds the decoder
  bool prior_on = false;
  int dec_in = -1, prior_in = -1, prior_out = -1;
  PzParams pz;
  TailParams tail;
  bool prior_conv0 = false;       // prior_network.0 runs inside the front kernel
  bool prior_from_cat = false;    // split precision: the prior network reads [y, z] of in_cat through a layout conversion
  FrontConvParams fc;
  // recognition network Q (ELBO evaluation): q_x_in and q_y_in write the two channel halves of q_cat, q_out reads it
  bool q_on = false;
  std::vector<V2Op> qx_ops, qy_ops, qout_ops;
  int q_cat = -1, q_out = -1;
};

constexpr int kIoSlots = 3;

struct bp_net {
  int device = 0, kind = NET_CVAE, prec = BP_PREC_F32, max_batch = 0, chunk = 0;
  bool split = false;           // fp32-accurate tensor-core path: split-precision fp16 operands (BP_PREC_F32)
  int H = 0, W = 0, lh = 0, lw = 0, in_c = 0;
  float min_z_var = 1e-7f;
  Stack st[kMaxStacks];
  int nstacks = 0;
  float* in_cat = nullptr;
  float* pool[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t pool_floats = 0;
  float* latent = nullptr;
  float* prior_all = nullptr;
  float* prior_keep = nullptr;  // [chunk][2][lh*lw], variance mode
  float* params = nullptr;  // [3][max_batch] sigma_in, sigma_out, aux
  float *d_in = nullptr, *d_out = nullptr, *d_lat = nullptr;
  float *h_in = nullptr, *h_out = nullptr, *h_lat = nullptr;
  double *var_mean = nullptr, *var_m2 = nullptr;                      // variance maps: running moments (float64)
  float* var_rep = nullptr;                                           // ... and replicated tiles
  float *xq = nullptr, *d_x = nullptr, *h_x = nullptr, *q_cat32 = nullptr, *q_out32 = nullptr;   // ELBO: transformed x, staging, fp32 Q tensors
  double* d_sums = nullptr;                                           // ELBO: {sum of KL terms, sum of squared residuals}
  float likelihood_scaling = 1.f;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr, out_stream = nullptr;
  // I/O slots of the asynchronous host entry point: batch k + 1 uploads and batch k - 1 downloads while batch k
  // computes; the caller keeps two batches outstanding, so a third slot is needed for the one being enqueued
  // (allocated on first use)
  struct IoSlot {
    float *d_in = nullptr, *d_out = nullptr, *d_lat = nullptr, *h_lat = nullptr, *h_par = nullptr;
    cudaEvent_t h2d = nullptr, done = nullptr, out = nullptr;
    bool used = false;
  } io[kIoSlots];
  std::vector<cudaEvent_t> ev_ready, ev_chunk_done, ev_out;   // per chunk of the pipelined host path
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  int prior_valid = 0;
  bool debug = false;
  std::vector<float*> dbg[kMaxStacks];
  int dbg_n = 0;
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;  // one pair per recorded layer launch
  std::vector<int> prof_ids;                                     // stack*1000 + layer
  double flops_per_tile = 0;
  V2Plan v2;
};

static int build_stack(const bp_layer_desc* descs, int n, int in_c, int H, int W, Stack* st, const char* name) {
  st->in_c = in_c; st->H = H; st->W = W;
  st->layers.resize(n);
  int c = in_c, h = H, w = W;
  int open = 0;
  int skip_c = 0, skip_h = 0, skip_w = 0;
  for (int i = 0; i < n; ++i) {
    BP_REQUIRE(descs[i].cin == c, BP_E_INVALID, "%s layer %d expects %d input channels, got %d", name, i,
               descs[i].cin, c);
    int rc = pack_layer(descs[i], h, w, &st->layers[i]);
    if (rc != BP_OK) return rc;
    if (descs[i].res == BP_RES_OPEN) {
      BP_REQUIRE(!open, BP_E_UNSUPPORTED, "%s: nested residual blocks", name);
      open = 1; skip_c = c; skip_h = h; skip_w = w;
    }
    c = descs[i].cout; h = st->layers[i].OHF; w = st->layers[i].OWF;
    if (descs[i].res == BP_RES_CLOSE) {
      BP_REQUIRE(open && skip_c == c && skip_h == h && skip_w == w, BP_E_INVALID,
                 "%s layer %d: residual block does not preserve the tensor shape", name, i);
      open = 0;
    }
    st->max_floats = std::max(st->max_floats, (size_t)c * h * w);
  }
  BP_REQUIRE(!open, BP_E_INVALID, "%s: residual block never closed", name);
  st->out_c = c; st->OH = h; st->OW = w;
  return BP_OK;
}

static void destroy_net(bp_net* net) {
  if (!net) return;
  cudaSetDevice(net->device);
  for (int s = 0; s < kMaxStacks; ++s) {
    for (auto& l : net->st[s].layers) free_layer(&l);
    for (float* p : net->dbg[s]) cudaFree(p);
  }
  for (auto& op : net->v2.ops) wconv_free(op.w);
  for (auto& op : net->v2.prior_ops) wconv_free(op.w);
  for (void* p : net->v2.owned) cudaFree(p);
  cudaFree(net->in_cat);
  for (int i = 0; i < 4; ++i) cudaFree(net->pool[i]);
  cudaFree(net->latent); cudaFree(net->prior_all); cudaFree(net->prior_keep); cudaFree(net->params);
  cudaFree(net->d_in); cudaFree(net->d_out); cudaFree(net->d_lat);
  for (auto& sl : net->io) {
    cudaFree(sl.d_in); cudaFree(sl.d_out); cudaFree(sl.d_lat);
    if (sl.h_lat) cudaFreeHost(sl.h_lat);
    if (sl.h_par) cudaFreeHost(sl.h_par);
    for (cudaEvent_t e : {sl.h2d, sl.done, sl.out})
      if (e) cudaEventDestroy(e);
  }
  cudaFree(net->var_mean); cudaFree(net->var_m2); cudaFree(net->var_rep);
  cudaFree(net->xq); cudaFree(net->d_x); cudaFree(net->q_cat32); cudaFree(net->q_out32); cudaFree(net->d_sums);
  if (net->h_x) cudaFreeHost(net->h_x);
  for (auto* v : {&net->v2.qx_ops, &net->v2.qy_ops, &net->v2.qout_ops})
    for (auto& op : *v) wconv_free(op.w);
  if (net->h_in) cudaFreeHost(net->h_in);
  if (net->h_out) cudaFreeHost(net->h_out);
  if (net->h_lat) cudaFreeHost(net->h_lat);
  for (int i = 0; i < 2; ++i) {
    if (net->ev_in[i]) cudaEventDestroy(net->ev_in[i]);
    if (net->ev_done[i]) cudaEventDestroy(net->ev_done[i]);
  }
  if (net->stream) cudaStreamDestroy(net->stream);
  if (net->copy_stream) cudaStreamDestroy(net->copy_stream);
  if (net->out_stream) cudaStreamDestroy(net->out_stream);
  for (auto* v : {&net->ev_ready, &net->ev_chunk_done, &net->ev_out})
    for (cudaEvent_t e : *v) cudaEventDestroy(e);
  delete net;
}

static int check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  BP_REQUIRE(e == cudaSuccess && count > 0, BP_E_NO_DEVICE,
             "no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(e));
  BP_REQUIRE(device >= 0 && device < count, BP_E_INVALID, "device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  BP_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  BP_REQUIRE(prop.major == 10, BP_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
             device, prop.major, prop.minor);
  BP_CUDA_TRY(cudaSetDevice(device));
  return BP_OK;
}

extern "C" const char* bp_last_error(void);
static int v2_build(bp_net* net);
static int v2_build_cgan(bp_net* net);

static int finish_create(bp_net* net) {
  const size_t HW = (size_t)net->H * net->W;
  size_t mx = 0;
  net->flops_per_tile = 0;
  // the rotating fp32 pool serves run_stack: every stack of the FFMA path, only p_z_in (debug taps, and the split
  // path's decoder input) on the tensor-core paths
  for (int s = 0; s < net->nstacks; ++s) {
    if (net->prec == BP_PREC_F32_FFMA || (net->kind == NET_CVAE && s == ST_PZ)) mx = std::max(mx, net->st[s].max_floats);
    if (net->kind == NET_CVAE && s >= ST_QX) continue;           // (the recognition network is not part of a paint)
    for (auto& l : net->st[s].layers) net->flops_per_tile += l.flops;
  }
  mx = std::max<size_t>(mx, 16);
  // chunk: keep the three rotating activation buffers of one chunk around L2 size (126 MB) so that
  // a layer's output is still cache-resident when the next layer reads it
  int chunk = net->prec == BP_PREC_F32_FFMA ? 4 : 256;
  if (const char* e = getenv("BP_CHUNK")) chunk = std::max(1, atoi(e));
  net->chunk = std::min(chunk, net->max_batch);
  net->pool_floats = mx * net->chunk;
  BP_CUDA_TRY(cudaMalloc(&net->in_cat, sizeof(float) * net->in_c * HW * net->chunk));
  for (int i = 0; i < 4; ++i) BP_CUDA_TRY(cudaMalloc(&net->pool[i], sizeof(float) * net->pool_floats));
  const size_t lhw = (size_t)net->lh * net->lw;
  if (net->kind == NET_CVAE) {
    BP_CUDA_TRY(cudaMalloc(&net->latent, sizeof(float) * lhw * net->chunk));
    BP_CUDA_TRY(cudaMalloc(&net->prior_all, sizeof(float) * 2 * lhw * net->max_batch));
    BP_CUDA_TRY(cudaMalloc(&net->prior_keep, sizeof(float) * 2 * lhw * net->chunk));
    BP_CUDA_TRY(cudaMalloc(&net->d_lat, sizeof(float) * lhw * net->max_batch));
    BP_CUDA_TRY(cudaMallocHost(&net->h_lat, sizeof(float) * lhw * net->max_batch));
  }
  BP_CUDA_TRY(cudaMalloc(&net->params, sizeof(float) * 3 * net->max_batch));
  BP_CUDA_TRY(cudaMalloc(&net->d_in, sizeof(float) * HW * net->max_batch));
  BP_CUDA_TRY(cudaMalloc(&net->d_out, sizeof(float) * HW * net->max_batch));
  BP_CUDA_TRY(cudaMallocHost(&net->h_in, sizeof(float) * HW * net->max_batch));
  BP_CUDA_TRY(cudaMallocHost(&net->h_out, sizeof(float) * HW * net->max_batch));
  BP_CUDA_TRY(cudaStreamCreateWithFlags(&net->stream, cudaStreamNonBlocking));
  BP_CUDA_TRY(cudaStreamCreateWithFlags(&net->copy_stream, cudaStreamNonBlocking));
  BP_CUDA_TRY(cudaStreamCreateWithFlags(&net->out_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    BP_CUDA_TRY(cudaEventCreateWithFlags(&net->ev_in[i], cudaEventDisableTiming));
    BP_CUDA_TRY(cudaEventCreateWithFlags(&net->ev_done[i], cudaEventDisableTiming));
  }
  if (net->kind == NET_CVAE && !net->st[ST_QOUT].layers.empty()) {
    const Stack& qx = net->st[ST_QX];
    const Stack& qy = net->st[ST_QY];
    BP_CUDA_TRY(cudaMalloc(&net->xq, sizeof(float) * HW * net->chunk));
    BP_CUDA_TRY(cudaMalloc(&net->q_cat32, sizeof(float) * (size_t)(qx.out_c + qy.out_c) * qx.OH * qx.OW * net->chunk));
    BP_CUDA_TRY(cudaMalloc(&net->q_out32, sizeof(float) * 2 * lhw * net->chunk));
    BP_CUDA_TRY(cudaMalloc(&net->d_sums, sizeof(double) * 2));
  }
  if (net->prec != BP_PREC_F32_FFMA) {
    // the tensor-core paths (16-bit, and split-precision fp32) are the window-GEMM engine (bp_wconv.cu) or nothing: a network it cannot lower fails here
    int rc = net->kind == NET_CVAE ? v2_build(net) : v2_build_cgan(net);
    if (rc != BP_OK) return rc;
    BP_REQUIRE(net->v2.built, BP_E_UNSUPPORTED,
               "the tensor-core engine cannot lower this network's first layer; use precision fp32-ffma");
  }
  return BP_OK;
}


// ------------------------------------------------------------------------------------------
// 16-bit engine, version 2: decoder (p_y_z_in + p_mu_out) as a chain of window GEMMs over NHWC
// activations; every layer output has its own zero-initialised buffer (space-to-depth borders are
// the zero padding of the consumer and are never written)
// ------------------------------------------------------------------------------------------
static int v2_new_act(bp_net* net, ActDesc d, int* idx) {
  void* p = nullptr;
  const size_t bytes = d.bytes_per_sample() * (size_t)net->chunk + 4096;   // slack: 16-byte vector tails
  BP_CUDA_TRY(cudaMalloc(&p, bytes));
  BP_CUDA_TRY(cudaMemset(p, 0, bytes));
  d.ptr = p;
  net->v2.owned.push_back(p);
  net->v2.acts.push_back(d);
  *idx = (int)net->v2.acts.size() - 1;
  return BP_OK;
}

struct V2Ref { int stack, index; Layer* l; bool w; int need_b; };

// split-precision activations are stored times 2^7: values down to ~2e-3 keep a normal fp16 lo half, the largest
// representable activation is 511 (the paint networks' post-BN activations stay below ~50)
constexpr int kSplitSexp = 7;

// ---- tuning table ----------------------------------------------------------------------------------
// Which window-GEMM formulation (pixel packing G x Jy, GEMM N) and tiling (strip width, M-tiles per region, tap
// lines per weight stage, ring depth) a layer runs with is DATA, not a per-process measurement: formulations sum K
// in different orders, so a process that timed candidates at load could paint bit-different tiles from the same
// inputs as its neighbour.  Net creation looks the layer up in this table (shipped with the package as
// baryon_painter_b200/tuning_table.txt, installed through bp_tuning_set); a layer that is not listed gets the
// cost model's first formulation that fits -- deterministic too.  Only `bp_tuning_mode(1)` (python -m
// baryon_painter_b200.tune, on a B200) times candidates, and records what it picked for bp_tuning_get.
struct TuneEntry { int G, Jy, N, mode; WTiling t; float ms; };
static std::mutex g_tune_mutex;
static std::map<std::string, TuneEntry> g_tune_table;
static bool g_tune_mode = false, g_tune_log = false;
static int g_tune_launches = 0;

static std::string tune_key(const Layer& l, int fmt, int split) {
  char b[128];
  snprintf(b, sizeof(b), "%s:%d:%d:k%d:s%d:%dx%d:%s%s", l.d.kind == BP_CONV ? "conv" : "convT", l.d.cin, l.d.cout, l.d.kernel,
           l.d.stride, l.H, l.W, fmt == TC_FMT_BF16 ? "bf16" : "f16", split ? ":split" : "");
  return b;
}
static bool tune_lookup(const std::string& key, TuneEntry* e) {
  std::lock_guard<std::mutex> g(g_tune_mutex);
  auto it = g_tune_table.find(key);
  if (it == g_tune_table.end()) return false;
  *e = it->second;
  return true;
}
static void tune_record(const std::string& key, const TuneEntry& e) {
  std::lock_guard<std::mutex> g(g_tune_mutex);
  g_tune_table[key] = e;
}
static int v2_time_layer(const WLayer* w, const ActDesc& out, const void* skip, int nb, float* ms) {
  cudaEvent_t e[4];
  g_tune_launches += 4;
  for (auto& x : e) BP_CUDA_TRY(cudaEventCreate(&x));
  int rc = wconv_launch(w, out, skip, nb, 0);
  for (int i = 0; i < 4 && rc == BP_OK; ++i) {
    if (i) rc = wconv_launch(w, out, skip, nb, 0);
    if (rc == BP_OK && cudaEventRecord(e[i], 0) != cudaSuccess) rc = BP_E_CUDA;
  }
  if (rc == BP_OK && cudaStreamSynchronize(0) != cudaSuccess) rc = BP_E_CUDA;
  float t[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < 3 && rc == BP_OK; ++i)
    if (cudaEventElapsedTime(&t[i], e[i], e[i + 1]) != cudaSuccess) rc = BP_E_CUDA;
  for (auto& x : e) cudaEventDestroy(x);
  if (rc != BP_OK) { cudaGetLastError(); return rc; }
  std::sort(t, t + 3);
  *ms = t[1];
  return BP_OK;
}

// lowers one layer sequence starting from act `cur`; the last layer of a `caller_out` sequence writes the
// caller's fp32 tiles (tail stencil or fp32 kernel, inverse transform fused)
static int v2_build_seq(bp_net* net, std::vector<V2Ref>& seq, int cur, std::vector<V2Op>& ops, bool caller_out,
                        int* last_act, int final_f32_act = -1) {
  V2Plan& P = net->v2;
  const int fmt = net->prec == BP_PREC_BF16 ? TC_FMT_BF16 : TC_FMT_F16;
  // the caller's fp32 tiles come from the tail stencil (1 -> 1 convolution) or, for a wider last layer, from its
  // window GEMM (fp32 plane) followed by an identity tail that applies the inverse transform
  const bool wide_tail = caller_out && seq.back().w && seq.back().l->d.cout == 1 && seq.back().l->d.cin > 1 &&
                         !dev_env("BP_V2_NOTAIL");
  if (caller_out && !wide_tail) seq.back().w = false;
  for (size_t i = 0; i < seq.size(); ++i)
    if (seq[i].l->d.res == BP_RES_OPEN) {
      size_t j = i;
      while (seq[j].l->d.res != BP_RES_CLOSE) ++j;
      bool all = true;
      for (size_t q = i; q <= j; ++q) all = all && seq[q].w;
      for (size_t q = i; q <= j; ++q) seq[q].w = all;
    }
  int skip = -1;
  for (int i = 0; i < (int)seq.size(); ++i) {
    V2Ref& r = seq[i];
    const bp_layer_desc& d = r.l->d;
    const bool last = (i + 1 == (int)seq.size());
    V2Op op;
    op.stack = r.stack; op.index = r.index; op.l = r.l;
    if (r.w) {
      const int Cp = v2_padc(d.cin);
      if (P.acts[cur].f32) {
        ActDesc a; a.C = d.cin; a.Cp = Cp; a.H = r.l->H; a.W = r.l->W; a.b = r.need_b; a.split = net->split; a.sexp = net->split ? kSplitSexp : 0;
        V2Op cv; cv.kind = V2_TO_NHWC16; cv.in = cur;
        int rc = v2_new_act(net, a, &cv.out);
        if (rc != BP_OK) return rc;
        ops.push_back(cv);
        cur = cv.out;
      }
      BP_REQUIRE(P.acts[cur].b == r.need_b && P.acts[cur].Cp == Cp, BP_E_INVALID,
                 "internal: layer %d.%d reads layout b=%d Cp=%d, producer wrote b=%d Cp=%d", r.stack, r.index,
                 r.need_b, Cp, P.acts[cur].b, P.acts[cur].Cp);
      if (d.res == BP_RES_OPEN) skip = cur;
      ActDesc o; o.C = d.cout; o.H = r.l->OHF; o.W = r.l->OWF;
      const bool next_w = !last && seq[i + 1].w;
      if (d.cout == 1) { o.f32 = true; o.Cp = 1; }
      else { o.Cp = std::max(8, v2_padc(d.cout)); o.b = next_w ? seq[i + 1].need_b : 1; o.split = net->split; o.sexp = net->split ? kSplitSexp : 0; }   // 16-byte stores
      std::vector<WSpec> cands;
      int rc = v2_candidates(*r.l, fmt, Cp, net->split, &cands);
      if (rc != BP_OK) return rc;
      op.kind = V2_WCONV; op.in = cur;
      r.l->v2 = true;
      if (d.res == BP_RES_CLOSE) { op.skip = skip; skip = -1; }
      rc = v2_new_act(net, o, &op.out);
      if (rc != BP_OK) return rc;
      // several formulations of one layer (pixel packing of the narrow stride-1 convolutions) and several tilings
      // of each: the tuning table names the one to build; without an entry the cost model's first that fits is used
      const std::string key = tune_key(*r.l, fmt, net->split ? 1 : 0);
      const void* skip_ptr = op.skip >= 0 ? P.acts[op.skip].ptr : nullptr;
      // no formulation fits (e.g. a 9x9 convolution over 128-byte split-precision pixels: more k-step slots than the
      // offset table holds): this one layer runs on the fp32 FFMA kernel instead, bridged by layout conversions
      auto fallback = [&]() {
        cudaFree(P.acts[op.out].ptr);
        net->v2.owned.pop_back();
        P.acts.pop_back();
        r.l->v2 = false;
        r.w = false;
      };
      TuneEntry te;
      if (tune_lookup(key, &te) && !g_tune_mode) {
        rc = BP_E_UNSUPPORTED;
        for (const WSpec& sp : cands)
          if (sp.G == te.G && sp.Jy == te.Jy && sp.N == te.N && sp.mode == te.mode) {
            rc = wconv_build(sp, P.acts[cur], net->chunk, &op.w, 0, &te.t);
            break;
          }
        if (rc == BP_E_UNSUPPORTED) { op.w = nullptr; }          // stale entry (other chunk size / code version): default
        else if (rc != BP_OK) return rc;
      }
      if (!op.w && !g_tune_mode) {
        rc = BP_E_UNSUPPORTED;
        std::string why;
        for (const WSpec& sp : cands) {
          rc = wconv_build(sp, P.acts[cur], net->chunk, &op.w, 0, nullptr);
          if (rc != BP_E_UNSUPPORTED) break;
          char b[64];
          snprintf(b, sizeof(b), " [G=%d Jy=%d N=%d mode=%d: ", sp.G, sp.Jy, sp.N, sp.mode);
          why += std::string(b) + bp_last_error() + "]";
        }
        if (rc == BP_E_UNSUPPORTED)
          set_error("no window-GEMM formulation of %s fits:%s", key.c_str(), why.substr(0, 800).c_str());
        if (rc == BP_E_UNSUPPORTED && d.res == BP_RES_NONE) { fallback(); --i; continue; }
        if (rc != BP_OK) { op.w = nullptr; return rc; }
      }
      if (!op.w) {
        // tuning mode: every formulation that fits is timed on a full chunk with the model's best tiling, then the
        // model's next three tilings of the winner; the fastest is kept and recorded
        const int max_tune = 24, max_rank = 4;
        float best_ms = 0.f;
        int built = 0, ci = -1, best_ci = -1;
        rc = BP_E_UNSUPPORTED;
        for (const WSpec& sp : cands) {
          ++ci;
          if (built >= max_tune) break;
          WLayer* w = nullptr;
          int rb = wconv_build(sp, P.acts[cur], net->chunk, &w, 0, nullptr);
          if (rb == BP_E_UNSUPPORTED) continue;
          if (rb != BP_OK) { rc = rb; break; }
          ++built;
          rc = BP_OK;
          float ms = 0.f;
          rb = v2_time_layer(w, P.acts[op.out], skip_ptr, net->chunk, &ms);
          if (rb != BP_OK) { wconv_free(w); rc = rb; break; }
          if (g_tune_log) fprintf(stderr, "[tune]   %s N=%d G=%d Jy=%d mode=%d: %.3f ms\n", key.c_str(), sp.N, sp.G, sp.Jy, sp.mode, ms);
          if (!op.w || ms < best_ms) { if (op.w) wconv_free(op.w); op.w = w; best_ms = ms; best_ci = ci; }
          else wconv_free(w);
        }
        if (rc == BP_E_UNSUPPORTED && !op.w && d.res == BP_RES_NONE) { fallback(); --i; continue; }
        if (rc != BP_OK) { if (op.w) wconv_free(op.w); op.w = nullptr; return rc; }
        for (int rank = 1; rank < max_rank; ++rank) {
          WLayer* w = nullptr;
          int rb = wconv_build(cands[best_ci], P.acts[cur], net->chunk, &w, rank, nullptr);
          if (rb == BP_E_UNSUPPORTED) break;
          if (rb != BP_OK) { wconv_free(op.w); op.w = nullptr; return rb; }
          float ms = 0.f;
          rb = v2_time_layer(w, P.acts[op.out], skip_ptr, net->chunk, &ms);
          if (rb != BP_OK) { wconv_free(w); wconv_free(op.w); op.w = nullptr; return rb; }
          if (g_tune_log) fprintf(stderr, "[tune]   %s tiling %d: %.3f ms (best so far %.3f)\n", key.c_str(), rank, ms, best_ms);
          if (ms < best_ms) { wconv_free(op.w); op.w = w; best_ms = ms; }
          else wconv_free(w);
        }
        TuneEntry ne;
        ne.G = cands[best_ci].G; ne.Jy = cands[best_ci].Jy; ne.N = cands[best_ci].N; ne.mode = cands[best_ci].mode;
        wconv_tiling(op.w, &ne.t);
        ne.ms = best_ms;
        tune_record(key, ne);
        if (g_tune_log)
          fprintf(stderr, "[tune] %s: %d formulations timed, best %.3f ms (G=%d Jy=%d N=%d; tune launches so far %d)\n", key.c_str(),
                  built, best_ms, ne.G, ne.Jy, ne.N, g_tune_launches);
      }
      ops.push_back(op);
      cur = op.out;
      if (last && wide_tail) {
        V2Op t;
        t.kind = V2_TAIL; t.stack = -1; t.index = -1; t.in = cur; t.final = true;
        TailParams& tp = P.tail;
        memset(&tp, 0, sizeof(tp));
        tp.k = 1; tp.w[0] = 1.f; tp.scale = 1.f; tp.shift = 0.f; tp.act = BP_ACT_NONE;
        tp.precise = net->split ? 1 : 0;
        ops.push_back(t);
      }
    } else {
      if (!P.acts[cur].f32) {
        ActDesc a; a.C = d.cin; a.Cp = d.cin; a.H = r.l->H; a.W = r.l->W; a.f32 = true;
        V2Op cv; cv.kind = V2_TO_F32; cv.in = cur;
        int rc = v2_new_act(net, a, &cv.out);
        if (rc != BP_OK) return rc;
        ops.push_back(cv);
        cur = cv.out;
      }
      if (d.res == BP_RES_OPEN) skip = cur;
      op.kind = V2_F32CONV; op.in = cur; op.final = last && caller_out;
      if (op.final && d.kind == BP_CONV && d.cin == 1 && d.cout == 1 && d.stride == 1 && d.kernel == 2 * d.pad + 1 &&
          d.kernel <= 7 && d.res == BP_RES_NONE && !dev_env("BP_V2_NOTAIL")) {
        op.kind = V2_TAIL;
        TailParams& t = P.tail;
        memset(&t, 0, sizeof(t));
        t.k = d.kernel;
        for (int q = 0; q < d.kernel * d.kernel; ++q) t.w[q] = r.l->host_weight[q];
        t.scale = r.l->host_scale[0]; t.shift = r.l->host_shift[0];
        t.act = d.act; t.act_param = d.act_param;
        t.precise = net->split ? 1 : 0;
      }
      if (d.res == BP_RES_CLOSE) { op.skip = skip; skip = -1; }
      if (!op.final) {
        ActDesc o; o.C = d.cout; o.Cp = d.cout; o.H = r.l->OHF; o.W = r.l->OWF; o.f32 = true;
        int rc = v2_new_act(net, o, &op.out);
        if (rc != BP_OK) return rc;
        cur = op.out;
      }
      ops.push_back(op);
    }
  }
  if (!caller_out && !P.acts[cur].f32) {
    // the sequence's consumer (latent sampling; the concatenation in front of q_out) reads fp32 NCHW
    const ActDesc& c = P.acts[cur];
    V2Op cv; cv.kind = V2_TO_F32; cv.in = cur;
    if (final_f32_act >= 0) {
      cv.out = final_f32_act;              // a caller-owned view (channel slice of a wider fp32 tensor)
    } else {
      ActDesc a; a.C = c.C; a.Cp = c.C; a.H = c.H; a.W = c.W; a.f32 = true;
      int rc = v2_new_act(net, a, &cv.out);
      if (rc != BP_OK) return rc;
    }
    ops.push_back(cv);
    cur = cv.out;
  } else if (!caller_out && final_f32_act >= 0) {
    V2Op cp; cp.kind = V2_COPY_F32; cp.in = cur; cp.out = final_f32_act;
    ops.push_back(cp);
    cur = cp.out;
  }
  if (last_act) *last_act = cur;
  return BP_OK;
}

static int v2_build_cgan(bp_net* net) {
  V2Plan& P = net->v2;
  std::vector<V2Ref> seq;
  for (size_t i = 0; i < net->st[ST_GEN].layers.size(); ++i) {
    V2Ref r{ST_GEN, (int)i, &net->st[ST_GEN].layers[i], false, 1};
    r.w = v2_eligible(*r.l, &r.need_b, net->split);
    seq.push_back(r);
  }
  if (!seq[0].w || net->in_c != 2) return BP_OK;           // first layer not lowerable (the caller reports it)
  int rc;
  if (net->split) {
    // split precision: [x', z - 1] as fp32 planes (launch_prepare), brought into the split layout by a conversion
    ActDesc in0;
    in0.ptr = net->in_cat; in0.C = 2; in0.Cp = 2; in0.H = net->H; in0.W = net->W; in0.f32 = true;
    P.acts.push_back(in0);
    P.dec_in = (int)P.acts.size() - 1;
  } else {
    // generator input [x', z - 1] written straight into the first layer's NHWC layout by the front kernel
    ActDesc a; a.C = 2; a.Cp = 4; a.H = net->H; a.W = net->W; a.b = seq[0].need_b;
    rc = v2_new_act(net, a, &P.dec_in);
    if (rc != BP_OK) return rc;
  }
  rc = v2_build_seq(net, seq, P.dec_in, P.ops, true, nullptr);
  if (rc != BP_OK) return rc;
  P.built = true;
  return BP_OK;
}

static int v2_build(bp_net* net) {
  V2Plan& P = net->v2;
  // ---- decoder
  std::vector<V2Ref> seq;
  for (int sidx : {ST_PYZ, ST_MU})
    for (size_t i = 0; i < net->st[sidx].layers.size(); ++i) {
      V2Ref r{sidx, (int)i, &net->st[sidx].layers[i], false, 1};
      r.w = v2_eligible(*r.l, &r.need_b, net->split);
      seq.push_back(r);
    }
  // fused front: p_z_in is a pyramid of single-channel k = 2s transposed convolutions
  const std::vector<Layer>& pzl = net->st[ST_PZ].layers;
  // (the fused front kernels use the fast log and write plain 16-bit pixels: 16-bit path only)
  bool front = seq[0].w && seq[0].need_b == 1 && net->in_c == 3 && pzl.size() <= 4 && !net->split && !dev_env("BP_V2_NOFRONT");
  for (const Layer& l : pzl)
    front = front && l.d.kind == BP_CONVT && l.d.cin == 1 && l.d.cout == 1 && l.d.kernel == 2 * l.d.stride &&
            2 * l.d.pad == l.d.stride && l.d.kernel <= 8 && l.d.out_pad == 0 && l.d.res == BP_RES_NONE;
  int start;
  if (front) {
    memset(&P.pz, 0, sizeof(P.pz));
    P.pz.nl = (int)pzl.size();
    for (int i = 0; i < P.pz.nl; ++i) {
      const Layer& l = pzl[i];
      P.pz.k[i] = l.d.kernel; P.pz.s[i] = l.d.stride; P.pz.p[i] = l.d.pad; P.pz.act[i] = l.d.act;
      P.pz.act_param[i] = l.d.act_param; P.pz.scale[i] = l.host_scale[0]; P.pz.shift[i] = l.host_shift[0];
      for (int q = 0; q < l.d.kernel * l.d.kernel; ++q) P.pz.w[i][q] = l.host_weight[q];
    }
    ActDesc a; a.C = 3; a.Cp = 4; a.H = net->H; a.W = net->W; a.b = 1;
    int rc = v2_new_act(net, a, &start);
    if (rc != BP_OK) return rc;
    P.front_on = true;
  } else {
    // the assembled decoder input [latent plane, y, z plane], fp32 NCHW (written by prepare + p_z_in)
    ActDesc in0;
    in0.ptr = net->in_cat; in0.C = net->in_c; in0.Cp = net->in_c; in0.H = net->H; in0.W = net->W; in0.f32 = true;
    P.acts.push_back(in0);
    start = (int)P.acts.size() - 1;
  }
  P.dec_in = start;
  int rc = v2_build_seq(net, seq, start, P.ops, true, nullptr);
  if (rc != BP_OK) return rc;
  // ---- prior network
  std::vector<Layer>& prl = net->st[ST_PRIOR].layers;
  if (!prl.empty() && !dev_env("BP_V2_NOPRIOR")) {
    std::vector<V2Ref> ps;
    for (size_t i = 0; i < prl.size(); ++i) {
      V2Ref r{ST_PRIOR, (int)i, &prl[i], false, 1};
      r.w = v2_eligible(*r.l, &r.need_b, net->split);
      ps.push_back(r);
    }
    if (ps[0].w && prl[0].d.cin == 2) {
      const bp_layer_desc& d0 = prl[0].d;
      // k4 s2 p1, 2 -> <= 8 channels followed by another window GEMM: folded into the front pass (FFMA stencil)
      const bool fuse0 = d0.kind == BP_CONV && d0.kernel == 4 && d0.stride == 2 && d0.pad == 1 && d0.cout <= 8 &&
                         d0.res == BP_RES_NONE && ps.size() > 1 && ps[1].w && (net->H % 2) == 0 && (net->W % 2) == 0 &&
                         !net->split && !dev_env("BP_V2_NOFUSE0");
      if (net->split) {
        // [y, z] = channels 1, 2 of in_cat (fp32 planes, per-sample stride 3 HW), converted to the split layout
        ActDesc v;
        v.ptr = net->in_cat + (size_t)net->H * net->W; v.C = 2; v.Cp = 2; v.H = net->H; v.W = net->W; v.f32 = true;
        v.f32_bs = 3ll * net->H * net->W;
        P.acts.push_back(v);
        P.prior_in = (int)P.acts.size() - 1;
        P.prior_from_cat = true;
      } else if (fuse0) {
        memset(&P.fc, 0, sizeof(P.fc));
        P.fc.cout = d0.cout; P.fc.act = d0.act; P.fc.act_param = d0.act_param;
        for (int co = 0; co < d0.cout; ++co) {
          P.fc.shift[co] = prl[0].host_shift[co];
          for (int ci = 0; ci < 2; ++ci)
            for (int t = 0; t < 16; ++t) P.fc.w[co][ci][t] = prl[0].host_weight[((size_t)co * 2 + ci) * 16 + t] * prl[0].host_scale[co];
          // z-plane tap sums by border class (first / interior / last row and column): the padded taps drop out
          for (int rc_ = 0; rc_ < 3; ++rc_)
            for (int cc = 0; cc < 3; ++cc) {
              float sz = 0.f;
              for (int r = (rc_ == 0 ? 1 : 0); r < (rc_ == 2 ? 3 : 4); ++r)
                for (int q = (cc == 0 ? 1 : 0); q < (cc == 2 ? 3 : 4); ++q) sz += P.fc.w[co][1][r * 4 + q];
              P.fc.zsum[rc_][cc][co] = sz;
            }
        }
        ActDesc a; a.C = d0.cout; a.Cp = 8; a.H = net->H / 2; a.W = net->W / 2; a.b = ps[1].need_b;
        rc = v2_new_act(net, a, &P.prior_in);
        if (rc != BP_OK) return rc;
        ps.erase(ps.begin());
        P.prior_conv0 = true;
        prl[0].v2 = true;
      } else {
        ActDesc a; a.C = 2; a.Cp = 4; a.H = net->H; a.W = net->W; a.b = ps[0].need_b;
        rc = v2_new_act(net, a, &P.prior_in);
        if (rc != BP_OK) return rc;
      }
      rc = v2_build_seq(net, ps, P.prior_in, P.prior_ops, false, &P.prior_out);
      if (rc != BP_OK) return rc;
      BP_REQUIRE(P.acts[P.prior_out].f32, BP_E_INVALID, "internal: prior head is not fp32");
      P.prior_on = true;
    }
  }
  // ---- recognition network (ELBO evaluation only): inputs are fp32 planes (x' in net->xq, [y', z] in in_cat),
  // brought into the engine's layout by conversions; outputs meet in the fp32 tensor q_cat32
  if (!net->st[ST_QOUT].layers.empty()) {
    const Stack& qx = net->st[ST_QX];
    const Stack& qy = net->st[ST_QY];
    const size_t hw = (size_t)net->H * net->W, qhw = (size_t)qx.OH * qx.OW;
    const int cc = qx.out_c + qy.out_c;
    auto view = [&](float* ptr, int C, int H, int W, long long bs) {
      ActDesc v; v.ptr = ptr; v.C = C; v.Cp = C; v.H = H; v.W = W; v.f32 = true; v.f32_bs = bs;
      P.acts.push_back(v);
      return (int)P.acts.size() - 1;
    };
    const int x_in = view(net->xq, 1, net->H, net->W, (long long)hw);
    const int y_in = view(net->in_cat + hw, 2, net->H, net->W, 3ll * hw);
    const int cat_x = view(net->q_cat32, qx.out_c, qx.OH, qx.OW, (long long)cc * qhw);
    const int cat_y = view(net->q_cat32 + (size_t)qx.out_c * qhw, qy.out_c, qy.OH, qy.OW, (long long)cc * qhw);
    P.q_cat = view(net->q_cat32, cc, qx.OH, qx.OW, (long long)cc * qhw);
    struct QSeq { int stack, in, fin; std::vector<V2Op>* ops; };
    const QSeq seqs[3] = {{ST_QX, x_in, cat_x, &P.qx_ops}, {ST_QY, y_in, cat_y, &P.qy_ops}, {ST_QOUT, P.q_cat, -1, &P.qout_ops}};
    for (const QSeq& q : seqs) {
      std::vector<V2Ref> qs;
      for (size_t i = 0; i < net->st[q.stack].layers.size(); ++i) {
        V2Ref r{q.stack, (int)i, &net->st[q.stack].layers[i], false, 1};
        r.w = v2_eligible(*r.l, &r.need_b, net->split);
        qs.push_back(r);
      }
      int last = -1;
      rc = v2_build_seq(net, qs, q.in, *q.ops, false, &last, q.fin);
      if (rc != BP_OK) return rc;
      if (q.stack == ST_QOUT) P.q_out = last;
    }
    BP_REQUIRE(P.acts[P.q_out].f32, BP_E_INVALID, "internal: q_out head is not fp32");
    P.q_on = true;
  }
  P.built = true;
  return BP_OK;
}

// ------------------------------------------------------------------------------------------
// execution
// ------------------------------------------------------------------------------------------
struct PostOp {
  int post = POST_NONE;
  const float* sigma = nullptr;
  float k = 0.f, shift = 0.f;
};

static float* pick_buffer(bp_net* net, const void* a, const void* b, const void* c = nullptr) {
  for (int i = 0; i < 4; ++i)
    if (net->pool[i] != a && net->pool[i] != b && net->pool[i] != c) return net->pool[i];
  return nullptr;
}

// an fp32 NCHW activation tensor with per-sample stride `bs`
struct ActRef {
  const void* ptr = nullptr;
  long long bs = 0;
};

static int record_debug(bp_net* net, int sidx, int i, int nl, const Layer& l, const void* out, long long out_bs, int nb,
                        cudaStream_t s) {
  const size_t per = (size_t)l.d.cout * l.OHF * l.OWF;
  if (net->dbg[sidx].size() < (size_t)nl) net->dbg[sidx].resize(nl, nullptr);
  if (!net->dbg[sidx][i]) BP_CUDA_TRY(cudaMalloc(&net->dbg[sidx][i], sizeof(float) * per * net->chunk));
  BP_CUDA_TRY(cudaMemcpy2DAsync(net->dbg[sidx][i], per * sizeof(float), out, out_bs * sizeof(float),
                                per * sizeof(float), nb, cudaMemcpyDeviceToDevice, s));
  return BP_OK;
}

// run one sub-network on nb samples with the fp32 kernels (precision fp32; p_z_in taps of the debug mode).  The last
// layer writes to final_out (stride final_bs) when given; otherwise to a pool buffer returned in *result.
static int run_stack(bp_net* net, int sidx, ActRef in, float* final_out, long long final_bs, const PostOp& post,
                     int nb, cudaStream_t s, ActRef* result) {
  Stack& st = net->st[sidx];
  ActRef cur = in;
  ActRef skip;
  const int nl = (int)st.layers.size();
  for (int i = 0; i < nl; ++i) {
    Layer& l = st.layers[i];
    const bool last = (i == nl - 1);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (net->profile) {
      BP_CUDA_TRY(cudaEventCreate(&e0)); BP_CUDA_TRY(cudaEventCreate(&e1));
      BP_CUDA_TRY(cudaEventRecord(e0, s));
    }
    if (l.d.res == BP_RES_OPEN) skip = cur;
    const long long dense_bs = (long long)l.d.cout * l.OHF * l.OWF;
    ActRef out;
    if (last && final_out) {
      out.ptr = final_out; out.bs = final_bs;
    } else {
      out.ptr = pick_buffer(net, cur.ptr, skip.ptr, in.ptr);
      BP_REQUIRE(out.ptr, BP_E_INVALID, "internal: no free activation buffer");
      out.bs = dense_bs;
    }
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.in = static_cast<const float*>(cur.ptr); a.in_bs = cur.bs;
    a.out = static_cast<float*>(const_cast<void*>(out.ptr)); a.out_bs = out.bs; a.nb = nb;
    if (l.d.res == BP_RES_CLOSE) { a.skip = static_cast<const float*>(skip.ptr); a.skip_bs = skip.bs; }
    if (last && post.post != POST_NONE) {
      a.post = post.post; a.post_sigma = post.sigma; a.post_k = post.k; a.post_shift = post.shift;
    }
    int rc = launch_conv_f32(l, a, s);
    if (rc != BP_OK) return rc;
    if (net->profile) {
      BP_CUDA_TRY(cudaEventRecord(e1, s));
      net->prof_events.emplace_back(e0, e1);
      net->prof_ids.push_back(sidx * 1000 + i);
    }
    if (l.d.res == BP_RES_CLOSE) skip = ActRef();
    if (net->debug) {
      rc = record_debug(net, sidx, i, nl, l, out.ptr, out.bs, nb, s);
      if (rc != BP_OK) return rc;
    }
    cur = out;
  }
  if (result) *result = cur;
  return BP_OK;
}


// decoder of one chunk on the window-GEMM engine: in_cat (fp32) -> painted tiles
static int v2_run(bp_net* net, std::vector<V2Op>& ops, float* final_out, long long final_bs, const PostOp& post, int nb,
                  cudaStream_t s) {
  V2Plan& P = net->v2;
  const int fmt = net->prec == BP_PREC_BF16 ? TC_FMT_BF16 : TC_FMT_F16;
  int n_wconv = 0;
  for (V2Op& op : ops) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool layer_op = op.index >= 0;
    if (net->profile && layer_op) {
      BP_CUDA_TRY(cudaEventCreate(&e0)); BP_CUDA_TRY(cudaEventCreate(&e1));
      BP_CUDA_TRY(cudaEventRecord(e0, s));
    }
    int rc = BP_OK;
    const ActDesc& in = P.acts[op.in];
    switch (op.kind) {
      case V2_TO_NHWC16:
        rc = launch_nchw32_to_nhwc16(static_cast<const float*>(in.ptr), (long long)in.elems_per_sample(), P.acts[op.out],
                                     nb, fmt, s);
        break;
      case V2_TO_F32: {
        const ActDesc& o = P.acts[op.out];
        rc = launch_nhwc16_to_nchw32(in, static_cast<float*>(o.ptr), (long long)o.elems_per_sample(), nb, fmt, s);
        break;
      }
      case V2_WCONV:
        // alternate directions along the chain (the first layer runs in reverse: its producer, a front kernel or a
        // layout conversion, walks the tiles forward)
        rc = wconv_launch(op.w, P.acts[op.out], op.skip >= 0 ? P.acts[op.skip].ptr : nullptr, nb, s, (n_wconv++ & 1) == 0);
        break;
      case V2_COPY_F32: {
        const ActDesc& o = P.acts[op.out];
        const size_t row = (size_t)in.C * in.H * in.W * sizeof(float);
        BP_CUDA_TRY(cudaMemcpy2DAsync(o.ptr, o.elems_per_sample() * sizeof(float), in.ptr, in.elems_per_sample() * sizeof(float),
                                      row, nb, cudaMemcpyDeviceToDevice, s));
        break;
      }
      case V2_TAIL: {
        TailParams t = P.tail;
        if (post.post != POST_NONE) { t.post = 1; t.post_k = post.k; t.post_shift = post.shift; }
        rc = launch_tail_stencil(static_cast<const float*>(in.ptr), final_out, final_bs, t, post.sigma, in.H, in.W, nb, s);
        break;
      }
      case V2_F32CONV: {
        ConvArgs a;
        memset(&a, 0, sizeof(a));
        a.in = static_cast<const float*>(in.ptr); a.in_bs = (long long)in.elems_per_sample();
        if (op.final) {
          a.out = final_out; a.out_bs = final_bs;
          if (post.post != POST_NONE) { a.post = post.post; a.post_sigma = post.sigma; a.post_k = post.k; a.post_shift = post.shift; }
        } else {
          a.out = static_cast<float*>(P.acts[op.out].ptr); a.out_bs = (long long)P.acts[op.out].elems_per_sample();
        }
        a.nb = nb;
        if (op.skip >= 0) { a.skip = static_cast<const float*>(P.acts[op.skip].ptr); a.skip_bs = (long long)P.acts[op.skip].elems_per_sample(); }
        rc = launch_conv_f32(*op.l, a, s);
        break;
      }
    }
    if (rc != BP_OK) return rc;
    if (net->profile && layer_op) {
      BP_CUDA_TRY(cudaEventRecord(e1, s));
      net->prof_events.emplace_back(e0, e1);
      net->prof_ids.push_back(op.stack * 1000 + op.index);
    }
    if (net->debug && layer_op) {
      const Layer& l = *op.l;
      const size_t per = (size_t)l.d.cout * l.OHF * l.OWF;
      const int nl = (int)net->st[op.stack].layers.size();
      if (net->dbg[op.stack].size() < (size_t)nl) net->dbg[op.stack].resize(nl, nullptr);
      float*& dst = net->dbg[op.stack][op.index];
      if (!dst) BP_CUDA_TRY(cudaMalloc(&dst, sizeof(float) * per * net->chunk));
      if (op.final) {
        BP_CUDA_TRY(cudaMemcpy2DAsync(dst, per * sizeof(float), final_out, final_bs * sizeof(float), per * sizeof(float), nb,
                                      cudaMemcpyDeviceToDevice, s));
      } else if (P.acts[op.out].f32) {
        BP_CUDA_TRY(cudaMemcpyAsync(dst, P.acts[op.out].ptr, per * sizeof(float) * nb, cudaMemcpyDeviceToDevice, s));
      } else {
        rc = launch_nhwc16_to_nchw32(P.acts[op.out], dst, (long long)per, nb, fmt, s);
        if (rc != BP_OK) return rc;
      }
    }
  }
  return BP_OK;
}

// `staging`: page-locked [3][max_batch] floats; the per-tile parameters go through it so that the copies are truly
// asynchronous (a cudaMemcpyAsync from pageable memory first waits for the stream to drain)
static int upload_params(bp_net* net, const bp_transform_params* tp, int flags, int n, cudaStream_t s, float* staging = nullptr) {
  BP_REQUIRE(tp != nullptr && tp->aux != nullptr, BP_E_INVALID, "transform params / aux plane values missing");
  BP_REQUIRE(!(flags & BP_FLAG_TRANSFORM) || tp->sigma_in, BP_E_INVALID, "sigma_in missing");
  BP_REQUIRE(!(flags & BP_FLAG_INVERSE) || tp->sigma_out, BP_E_INVALID, "sigma_out missing");
  const int mb = net->max_batch;
  if (staging) {
    if (tp->sigma_in) {
      memcpy(staging, tp->sigma_in, sizeof(float) * n);
      BP_CUDA_TRY(cudaMemcpyAsync(net->params, staging, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    }
    if (tp->sigma_out) {
      memcpy(staging + mb, tp->sigma_out, sizeof(float) * n);
      BP_CUDA_TRY(cudaMemcpyAsync(net->params + mb, staging + mb, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    }
    memcpy(staging + 2 * mb, tp->aux, sizeof(float) * n);
    BP_CUDA_TRY(cudaMemcpyAsync(net->params + 2 * mb, staging + 2 * mb, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    return BP_OK;
  }
  if (tp->sigma_in)
    BP_CUDA_TRY(cudaMemcpyAsync(net->params, tp->sigma_in, sizeof(float) * n, cudaMemcpyHostToDevice, s));
  if (tp->sigma_out)
    BP_CUDA_TRY(cudaMemcpyAsync(net->params + mb, tp->sigma_out, sizeof(float) * n, cudaMemcpyHostToDevice, s));
  BP_CUDA_TRY(cudaMemcpyAsync(net->params + 2 * mb, tp->aux, sizeof(float) * n, cudaMemcpyHostToDevice, s));
  return BP_OK;
}

// stage A of one CVAE chunk: forward transform + aux plane (+ prior network)
static int cvae_chunk_front(bp_net* net, const float* tiles, const bp_transform_params* tp, int flags, int c0,
                            int nb, bool need_prior, cudaStream_t s, float** prior_out) {
  ActRef in, res;
  const size_t HW = (size_t)net->H * net->W;
  const int mb = net->max_batch;
  const int do_t = (flags & BP_FLAG_TRANSFORM) ? 1 : 0;
  const int fmt = net->prec == BP_PREC_BF16 ? TC_FMT_BF16 : TC_FMT_F16;
  V2Plan& P = net->v2;
  const bool v2_prior = P.built && P.prior_on;
  const bool need_in_cat = !(P.built && P.front_on) || (need_prior && !v2_prior) || net->debug;
  int rc;
  if (need_in_cat) {
    rc = launch_prepare(tiles + (size_t)c0 * HW, net->in_cat, 3 * (long long)HW, 1, 2, net->params + c0,
                        net->params + 2 * mb + c0, tp->k_in, tp->shift_in, do_t, nb, (int)HW, s);
    if (rc != BP_OK) return rc;
  }
  if (!need_prior) return BP_OK;
  PostOp none;
  if (v2_prior) {
    if (P.prior_from_cat)
      rc = BP_OK;                                  // launch_prepare above wrote [y, z]; prior_ops start with the conversion
    else if (P.prior_conv0)
      rc = launch_front_prior_conv(tiles + (size_t)c0 * HW, P.acts[P.prior_in], net->params + c0, net->params + 2 * mb + c0,
                                   P.fc, tp->k_in, tp->shift_in, do_t, net->H, net->W, nb, fmt, s);
    else
      rc = launch_front_prior(tiles + (size_t)c0 * HW, P.acts[P.prior_in], net->params + c0, net->params + 2 * mb + c0,
                              tp->k_in, tp->shift_in, do_t, nb, fmt, s);
    if (rc != BP_OK) return rc;
    if (net->debug && P.prior_conv0) {
      // record the fused kernel's own output as the activation of prior_network layer 0
      const Layer& l0 = net->st[ST_PRIOR].layers[0];
      const size_t per = (size_t)l0.d.cout * l0.OHF * l0.OWF;
      std::vector<float*>& dv = net->dbg[ST_PRIOR];
      if (dv.size() < net->st[ST_PRIOR].layers.size()) dv.resize(net->st[ST_PRIOR].layers.size(), nullptr);
      if (!dv[0]) BP_CUDA_TRY(cudaMalloc(&dv[0], sizeof(float) * per * net->chunk));
      rc = launch_nhwc16_to_nchw32(P.acts[P.prior_in], dv[0], (long long)per, nb, fmt, s);
      if (rc != BP_OK) return rc;
    }
    rc = v2_run(net, P.prior_ops, nullptr, 0, none, nb, s);
    if (rc != BP_OK) return rc;
    *prior_out = static_cast<float*>(P.acts[P.prior_out].ptr);
    return BP_OK;
  }
  in.ptr = net->in_cat + HW; in.bs = 3 * (long long)HW;
  rc = run_stack(net, ST_PRIOR, in, nullptr, 0, none, nb, s, &res);
  if (rc != BP_OK) return rc;
  *prior_out = static_cast<float*>(const_cast<void*>(res.ptr));
  return BP_OK;
}

// stage B: latent -> painted tile
static int cvae_chunk_back(bp_net* net, const float* tiles, const float* latent, const bp_transform_params* tp,
                           int flags, int c0, int nb, float* out, cudaStream_t s) {
  const size_t HW = (size_t)net->H * net->W;
  const size_t lhw = (size_t)net->lh * net->lw;
  const int mb = net->max_batch;
  const int fmt = net->prec == BP_PREC_BF16 ? TC_FMT_BF16 : TC_FMT_F16;
  PostOp none, post;
  if (flags & BP_FLAG_INVERSE) {
    post.post = POST_INV_SHIFT_LOG; post.sigma = net->params + mb + c0;
    post.k = tp->k_out; post.shift = tp->shift_out;
  }
  ActRef in, h;
  in.ptr = latent; in.bs = (long long)lhw;
  V2Plan& P = net->v2;
  int rc;
  if (!(P.built && P.front_on) || net->debug) {
    // fp32 p_z_in into channel 0 of in_cat (the fused front kernel has no per-layer outputs to record)
    rc = run_stack(net, ST_PZ, in, net->in_cat, 3 * (long long)HW, none, nb, s, nullptr);
    if (rc != BP_OK) return rc;
  }
  if (P.built) {
    if (P.front_on) {
      rc = launch_front_latent(tiles + (size_t)c0 * HW, latent, P.acts[P.dec_in], net->params + c0,
                               net->params + 2 * mb + c0, P.pz, tp->k_in, tp->shift_in,
                               (flags & BP_FLAG_TRANSFORM) ? 1 : 0, net->lh, net->lw, nb, fmt, s);
      if (rc != BP_OK) return rc;
      if (net->debug) {
        // record what the decoder really reads as the last p_z_in tap: channel 0 of the fused kernel's output
        const int li = (int)net->st[ST_PZ].layers.size() - 1;
        ActDesc v = P.acts[P.dec_in];
        v.C = 1;
        rc = launch_nhwc16_to_nchw32(v, net->dbg[ST_PZ][li], (long long)HW, nb, fmt, s);
        if (rc != BP_OK) return rc;
      }
    }
    return v2_run(net, P.ops, out, (long long)HW, post, nb, s);
  }
  in.ptr = net->in_cat; in.bs = 3 * (long long)HW;
  rc = run_stack(net, ST_PYZ, in, nullptr, 0, none, nb, s, &h);
  if (rc != BP_OK) return rc;
  return run_stack(net, ST_MU, h, out, (long long)HW, post, nb, s, nullptr);
}

// per-chunk hooks of the pipelined host path: the compute stream waits for `ready[c]` (inputs of chunk c are on the
// device) before the chunk's first kernel and records `done[c]` after its last one
struct ChunkHooks {
  std::vector<cudaEvent_t>* ready = nullptr;
  std::vector<cudaEvent_t>* done = nullptr;
  const std::vector<int>* bounds = nullptr;   // pipeline chunk c = tiles [bounds[c], bounds[c + 1]); null = the plan's chunks
};

static int cvae_paint_device(bp_net* net, const float* tiles, const float* latent, int mode, uint64_t seed,
                             const bp_transform_params* tp, int flags, float* out, int n, cudaStream_t s,
                             const ChunkHooks* hooks = nullptr, float* param_staging = nullptr) {
  BP_REQUIRE(net && net->kind == NET_CVAE, BP_E_INVALID, "not a CVAE network");
  BP_REQUIRE(n >= 0 && n <= net->max_batch, BP_E_INVALID, "batch %d exceeds max_batch %d", n, net->max_batch);
  BP_REQUIRE(mode == BP_LATENT_GIVEN || mode == BP_LATENT_EPS || mode == BP_LATENT_SEED, BP_E_INVALID,
             "bad latent mode %d", mode);
  BP_REQUIRE(mode == BP_LATENT_SEED || latent != nullptr, BP_E_INVALID, "latent/eps array missing");
  if (n == 0) return BP_OK;
  BP_REQUIRE(tiles && out, BP_E_INVALID, "null tile pointer");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  int rc = upload_params(net, tp, flags, n, s, param_staging);
  if (rc != BP_OK) return rc;
  const size_t HW = (size_t)net->H * net->W;
  const size_t lhw = (size_t)net->lh * net->lw;
  net->prior_valid = 0;
  if (net->debug) net->dbg_n = std::min(n, net->chunk);
  std::vector<int> plain;
  if (!(hooks && hooks->bounds)) {
    for (int c0 = 0; c0 < n; c0 += net->chunk) plain.push_back(c0);
    plain.push_back(n);
  }
  const std::vector<int>& bounds = (hooks && hooks->bounds) ? *hooks->bounds : plain;
  for (size_t ci = 0; ci + 1 < bounds.size(); ++ci) {
    const int c0 = bounds[ci], nb = bounds[ci + 1] - c0;
    BP_REQUIRE(nb > 0 && nb <= net->chunk, BP_E_INVALID, "internal: pipeline chunk of %d tiles", nb);
    float* prior_out = nullptr;
    if (hooks && hooks->ready) BP_CUDA_TRY(cudaStreamWaitEvent(s, (*hooks->ready)[ci], 0));
    rc = cvae_chunk_front(net, tiles, tp, flags, c0, nb, mode != BP_LATENT_GIVEN, s, &prior_out);
    if (rc != BP_OK) return rc;
    const float* lat;
    if (mode == BP_LATENT_GIVEN) {
      lat = latent + (size_t)c0 * lhw;
    } else {
      rc = launch_sample_z(prior_out, mode == BP_LATENT_EPS ? latent + (size_t)c0 * lhw : nullptr, net->latent,
                           net->prior_all + (size_t)c0 * lhw, net->prior_all + ((size_t)net->max_batch + c0) * lhw,
                           net->min_z_var, nb, (int)lhw, mode, seed, (uint64_t)c0 * lhw, s);
      if (rc != BP_OK) return rc;
      lat = net->latent;
    }
    rc = cvae_chunk_back(net, tiles, lat, tp, flags, c0, nb, out + (size_t)c0 * HW, s);
    if (rc != BP_OK) return rc;
    if (hooks && hooks->done) BP_CUDA_TRY(cudaEventRecord((*hooks->done)[ci], s));
    if (net->debug) break;  // debug buffers hold one chunk
  }
  if (mode != BP_LATENT_GIVEN) net->prior_valid = n;
  return BP_OK;
}

static int cgan_paint_device(bp_net* net, const float* tiles, const bp_transform_params* tp, int flags, float* out,
                             int n, cudaStream_t s, const ChunkHooks* hooks = nullptr) {
  BP_REQUIRE(net && net->kind == NET_CGAN, BP_E_INVALID, "not a CGAN network");
  BP_REQUIRE(n >= 0 && n <= net->max_batch, BP_E_INVALID, "batch %d exceeds max_batch %d", n, net->max_batch);
  if (n == 0) return BP_OK;
  BP_REQUIRE(tiles && out, BP_E_INVALID, "null tile pointer");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  int rc = upload_params(net, tp, flags, n, s);
  if (rc != BP_OK) return rc;
  const size_t HW = (size_t)net->H * net->W;
  const int mb = net->max_batch;
  if (net->debug) net->dbg_n = std::min(n, net->chunk);
  std::vector<int> plain;
  if (!(hooks && hooks->bounds)) {
    for (int c0 = 0; c0 < n; c0 += net->chunk) plain.push_back(c0);
    plain.push_back(n);
  }
  const std::vector<int>& bounds = (hooks && hooks->bounds) ? *hooks->bounds : plain;
  struct DoneGuard {                       // records done[ci] on every way out of a loop iteration
    const ChunkHooks* h; size_t ci; cudaStream_t s;
    ~DoneGuard() { if (h && h->done) cudaEventRecord((*h->done)[ci], s); }
  };
  for (size_t ci = 0; ci + 1 < bounds.size(); ++ci) {
    const int c0 = bounds[ci], nb = bounds[ci + 1] - c0;
    BP_REQUIRE(nb > 0 && nb <= net->chunk, BP_E_INVALID, "internal: pipeline chunk of %d tiles", nb);
    if (hooks && hooks->ready) BP_CUDA_TRY(cudaStreamWaitEvent(s, (*hooks->ready)[ci], 0));
    DoneGuard guard{hooks, ci, s};
    if (net->v2.built) {
      PostOp post2;
      if (flags & BP_FLAG_INVERSE) {
        post2.post = POST_INV_SHIFT_LOG; post2.sigma = net->params + mb + c0;
        post2.k = tp->k_out; post2.shift = tp->shift_out;
      }
      if (net->split)
        rc = launch_prepare(tiles + (size_t)c0 * HW, net->in_cat, 2 * (long long)HW, 0, 1, net->params + c0,
                            net->params + 2 * mb + c0, tp->k_in, tp->shift_in, (flags & BP_FLAG_TRANSFORM) ? 1 : 0, nb,
                            (int)HW, s);
      else
        rc = launch_front_prior(tiles + (size_t)c0 * HW, net->v2.acts[net->v2.dec_in], net->params + c0,
                                net->params + 2 * mb + c0, tp->k_in, tp->shift_in, (flags & BP_FLAG_TRANSFORM) ? 1 : 0, nb,
                                net->prec == BP_PREC_BF16 ? TC_FMT_BF16 : TC_FMT_F16, s);
      if (rc != BP_OK) return rc;
      rc = v2_run(net, net->v2.ops, out + (size_t)c0 * HW, (long long)HW, post2, nb, s);
      if (rc != BP_OK) return rc;
      if (net->debug) break;
      continue;
    }
    rc = launch_prepare(tiles + (size_t)c0 * HW, net->in_cat, 2 * (long long)HW, 0, 1, net->params + c0,
                        net->params + 2 * mb + c0, tp->k_in, tp->shift_in, (flags & BP_FLAG_TRANSFORM) ? 1 : 0, nb,
                        (int)HW, s);
    if (rc != BP_OK) return rc;
    PostOp post;
    if (flags & BP_FLAG_INVERSE) {
      post.post = POST_INV_SHIFT_LOG; post.sigma = net->params + mb + c0;
      post.k = tp->k_out; post.shift = tp->shift_out;
    }
    ActRef in;
    in.ptr = net->in_cat; in.bs = 2 * (long long)HW;
    rc = run_stack(net, ST_GEN, in, out + (size_t)c0 * HW, (long long)HW, post, nb, s, nullptr);
    if (rc != BP_OK) return rc;
    if (net->debug) break;
  }
  return BP_OK;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
// page-locked (or otherwise CUDA-registered) host memory can be the source / target of an async copy directly
static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// Host-buffer painting, shared by both painters: copy-in / compute / copy-out of neighbouring pipeline chunks
// overlap on three streams.  `run(hooks)` enqueues the device work, waiting for hooks.ready[c] before chunk c and
// recording hooks.done[c] after it.
template <class Run>
static int paint_host_pipelined(bp_net* net, const float* tiles, float* out, int n, Run run) {
  const size_t HW = (size_t)net->H * net->W;
  cudaStream_t s = net->stream, sin = net->copy_stream, sout = net->out_stream;
  // pipeline chunks: copy-in / compute / copy-out of neighbouring chunks overlap.  The first chunk's copy-in and
  // the last chunk's copy-out cannot hide behind anything, so the schedule ramps 16, 48, 64, ..., 64, 48, 16
  // (BP_HOST_STEP / BP_HOST_EDGE override): the exposed copies shrink to 16 tiles each while the bulk of the tiles
  // still run in launches that fill the machine
  static const int env_step = getenv("BP_HOST_STEP") ? atoi(getenv("BP_HOST_STEP")) : 64;
  static const int env_edge = getenv("BP_HOST_EDGE") ? atoi(getenv("BP_HOST_EDGE")) : 16;
  int step = std::max(1, std::min(net->chunk, env_step));
  if (n <= step && n >= 2 * env_edge && env_edge > 0) step = std::max(env_edge, (n + 3) / 4);   // small batch: still 2-4 chunks
  std::vector<int> bounds;
  {
    std::vector<int> head, tail;
    int left = n;
    const int edge = std::max(1, std::min(env_edge, step));
    for (int sz = edge; sz < step && left >= 2 * sz + step; sz *= 3) {   // ramp while a full middle chunk remains
      head.push_back(sz); tail.push_back(sz);
      left -= 2 * sz;
    }
    int c0 = 0;
    bounds.push_back(0);
    for (int sz : head) bounds.push_back(c0 += sz);
    for (; left > 0; left -= std::min(step, left)) bounds.push_back(c0 += std::min(step, left));
    for (size_t i = tail.size(); i-- > 0;) bounds.push_back(c0 += tail[i]);
  }
  const int nchunks = (int)bounds.size() - 1;
  static const bool trace = dev_env("BP_HOST_TRACE") != nullptr;      // per-chunk device timeline on stderr
  const unsigned evflags = trace ? cudaEventDefault : cudaEventDisableTiming;
  const auto cpu_now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_cpu0 = cpu_now();
  cudaEvent_t ev_t0 = nullptr;
  if (trace) {
    BP_CUDA_TRY(cudaEventCreate(&ev_t0));
    BP_CUDA_TRY(cudaEventRecord(ev_t0, net->copy_stream));
  }
  while ((int)net->ev_ready.size() < nchunks) {
    cudaEvent_t a, b, c;
    BP_CUDA_TRY(cudaEventCreateWithFlags(&a, evflags));
    BP_CUDA_TRY(cudaEventCreateWithFlags(&b, evflags));
    BP_CUDA_TRY(cudaEventCreateWithFlags(&c, evflags));
    net->ev_ready.push_back(a); net->ev_chunk_done.push_back(b); net->ev_out.push_back(c);
  }
  const bool in_pinned = is_pinned(tiles), out_pinned = is_pinned(out);
  // inputs: chunk by chunk on the copy-in stream (staging a pageable chunk overlaps the device's work on earlier ones)
  for (int c = 0; c < nchunks; ++c) {
    const size_t c0 = (size_t)bounds[c], nb = (size_t)(bounds[c + 1] - bounds[c]);
    const float* src = tiles + c0 * HW;
    if (!in_pinned) {
      memcpy(net->h_in + c0 * HW, src, sizeof(float) * HW * nb);
      src = net->h_in + c0 * HW;
    }
    BP_CUDA_TRY(cudaMemcpyAsync(net->d_in + c0 * HW, src, sizeof(float) * HW * nb, cudaMemcpyHostToDevice, sin));
    BP_CUDA_TRY(cudaEventRecord(net->ev_ready[c], sin));
  }
  ChunkHooks hooks;
  hooks.ready = &net->ev_ready; hooks.done = &net->ev_chunk_done; hooks.bounds = &bounds;
  int rc = run(hooks);
  if (rc != BP_OK) return rc;
  const int done_chunks = net->debug ? 1 : nchunks;
  for (int c = 0; c < done_chunks; ++c) {
    const size_t c0 = (size_t)bounds[c], nb = (size_t)(bounds[c + 1] - bounds[c]);
    float* dst = out_pinned ? out + c0 * HW : net->h_out + c0 * HW;
    BP_CUDA_TRY(cudaStreamWaitEvent(sout, net->ev_chunk_done[c], 0));
    BP_CUDA_TRY(cudaMemcpyAsync(dst, net->d_out + c0 * HW, sizeof(float) * HW * nb, cudaMemcpyDeviceToHost, sout));
    BP_CUDA_TRY(cudaEventRecord(net->ev_out[c], sout));
  }
  if (!out_pinned) {
    for (int c = 0; c < done_chunks; ++c) {
      const size_t c0 = (size_t)bounds[c], nb = (size_t)(bounds[c + 1] - bounds[c]);
      BP_CUDA_TRY(cudaEventSynchronize(net->ev_out[c]));
      memcpy(out + c0 * HW, net->h_out + c0 * HW, sizeof(float) * HW * nb);
    }
  }
  const double t_cpu1 = cpu_now();
  BP_CUDA_TRY(cudaStreamSynchronize(sout));
  BP_CUDA_TRY(cudaStreamSynchronize(s));
  BP_CUDA_TRY(cudaStreamSynchronize(sin));
  if (trace) {
    fprintf(stderr, "[host] n=%d chunks=%d: enqueue %.2f ms, total %.2f ms (cpu)\n", n, nchunks, t_cpu1 - t_cpu0, cpu_now() - t_cpu0);
    for (int c = 0; c < done_chunks; ++c) {
      float a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&a, ev_t0, net->ev_ready[c]);
      cudaEventElapsedTime(&b, ev_t0, net->ev_chunk_done[c]);
      cudaEventElapsedTime(&d, ev_t0, net->ev_out[c]);
      fprintf(stderr, "[host]   chunk %d tiles [%d, %d): in %.2f  computed %.2f  out %.2f ms\n", c, bounds[c], bounds[c + 1], a, b, d);
    }
    cudaEventDestroy(ev_t0);
  }
  return BP_OK;
}

extern "C" {

int bp_version(void) { return BP_VERSION; }

// ---- tuning table (see the comment above tune_key) ----
// text: one layer per line, "key G Jy N mode Wt T_r gl nbst [ms]"; '#' starts a comment.  Replaces the table.
int bp_tuning_set(const char* text) {
  BP_REQUIRE(text, BP_E_INVALID, "null tuning table");
  std::map<std::string, TuneEntry> t;
  std::istringstream in(text);
  std::string line;
  int lineno = 0;
  while (std::getline(in, line)) {
    ++lineno;
    const size_t h = line.find('#');
    if (h != std::string::npos) line.resize(h);
    std::istringstream ls(line);
    std::string key;
    if (!(ls >> key)) continue;
    TuneEntry e;
    e.ms = 0.f;
    BP_REQUIRE(bool(ls >> e.G >> e.Jy >> e.N >> e.mode >> e.t.Wt >> e.t.T_r >> e.t.gl >> e.t.nbst), BP_E_INVALID,
               "tuning table line %d: expected 'key G Jy N mode Wt T_r gl nbst'", lineno);
    ls >> e.ms;
    t[key] = e;
  }
  std::lock_guard<std::mutex> g(g_tune_mutex);
  g_tune_table.swap(t);
  return BP_OK;
}
// writes the current table (as bp_tuning_set reads it) into buf; returns the length needed (excluding the NUL)
int bp_tuning_get(char* buf, size_t cap) {
  std::ostringstream o;
  {
    std::lock_guard<std::mutex> g(g_tune_mutex);
    for (const auto& kv : g_tune_table) {
      const TuneEntry& e = kv.second;
      char ms[32];
      snprintf(ms, sizeof(ms), "%.4f", e.ms);
      o << kv.first << ' ' << e.G << ' ' << e.Jy << ' ' << e.N << ' ' << e.mode << ' ' << e.t.Wt << ' ' << e.t.T_r << ' ' << e.t.gl
        << ' ' << e.t.nbst << ' ' << ms << '\n';
    }
  }
  const std::string str = o.str();
  if (buf && cap > 0) {
    const size_t n = std::min(cap - 1, str.size());
    memcpy(buf, str.data(), n);
    buf[n] = 0;
  }
  return (int)str.size();
}
// on: nets created from now on TIME the candidate formulations of every layer on the device and record the winners
// (non-deterministic across runs by nature -- a tool for producing the table, never the default); log: print timings
int bp_tuning_mode(int on, int log) {
  std::lock_guard<std::mutex> g(g_tune_mutex);
  g_tune_mode = on != 0;
  g_tune_log = log != 0;
  return BP_OK;
}
const char* bp_last_error(void) { return g_err; }
int64_t bp_launch_count(int reset) {
  const int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}
int bp_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
  return count;
}

int bp_cvae_create(const bp_cvae_desc* d, int precision, int max_batch, int device, bp_net** out) {
  BP_REQUIRE(d && out, BP_E_INVALID, "null argument");
  *out = nullptr;
  BP_REQUIRE(precision == BP_PREC_F32 || precision == BP_PREC_BF16 || precision == BP_PREC_F16 || precision == BP_PREC_F32_FFMA,
             BP_E_INVALID, "bad precision %d", precision);
  BP_REQUIRE(max_batch > 0, BP_E_INVALID, "max_batch must be positive");
  BP_REQUIRE(d->tile_h > 0 && d->tile_w > 0 && d->latent_h > 0 && d->latent_w > 0, BP_E_INVALID, "bad tile shape");
  BP_REQUIRE((d->tile_h * d->tile_w) % 4 == 0, BP_E_INVALID, "tile area must be a multiple of 4");
  BP_REQUIRE(d->n_p_z_in > 0 && d->n_p_y_z_in > 0 && d->n_p_mu_out > 0, BP_E_INVALID, "empty decoder stack");
  int rc = check_device(device);
  if (rc != BP_OK) return rc;
  bp_net* net = new (std::nothrow) bp_net();
  BP_REQUIRE(net, BP_E_NOMEM, "out of host memory");
  net->device = device; net->kind = NET_CVAE; net->prec = precision; net->max_batch = max_batch;
  net->split = precision == BP_PREC_F32;
  net->H = d->tile_h; net->W = d->tile_w; net->lh = d->latent_h; net->lw = d->latent_w;
  net->min_z_var = d->min_z_var; net->in_c = 3; net->nstacks = 4;
  do {
    if (d->n_prior > 0) {
      rc = build_stack(d->prior, d->n_prior, 2, net->H, net->W, &net->st[ST_PRIOR], "prior_network");
      if (rc != BP_OK) break;
      const Stack& p = net->st[ST_PRIOR];
      if (!(p.out_c == 2 && p.OH == net->lh && p.OW == net->lw)) {
        set_error("Dimension of z_mu does not match dim_z: (%d, %d, %d) vs (1, %d, %d).", p.out_c / 2, p.OH, p.OW,
                  net->lh, net->lw);
        rc = BP_E_INVALID; break;
      }
    }
    rc = build_stack(d->p_z_in, d->n_p_z_in, 1, net->lh, net->lw, &net->st[ST_PZ], "p_z_in");
    if (rc != BP_OK) break;
    const Stack& z = net->st[ST_PZ];
    if (!(z.out_c == 1 && z.OH == net->H && z.OW == net->W)) {
      set_error("p_z_in output (%d, %d, %d) does not match the tile (1, %d, %d)", z.out_c, z.OH, z.OW, net->H,
                net->W);
      rc = BP_E_INVALID; break;
    }
    rc = build_stack(d->p_y_z_in, d->n_p_y_z_in, 3, net->H, net->W, &net->st[ST_PYZ], "p_y_z_in");
    if (rc != BP_OK) break;
    const Stack& y = net->st[ST_PYZ];
    rc = build_stack(d->p_mu_out, d->n_p_mu_out, y.out_c, y.OH, y.OW, &net->st[ST_MU], "p_mu_out");
    if (rc != BP_OK) break;
    const Stack& m = net->st[ST_MU];
    if (!(m.out_c == 1 && m.OH == net->H && m.OW == net->W)) {
      set_error("Dimension of x_mu does not match dim_x: (%d, %d, %d) vs (1, %d, %d).", m.out_c, m.OH, m.OW,
                net->H, net->W);
      rc = BP_E_INVALID; break;
    }
    if (d->n_q_out > 0) {
      BP_REQUIRE(d->n_q_x_in > 0 && d->n_q_y_in > 0, BP_E_INVALID, "q_out without q_x_in / q_y_in");
      rc = build_stack(d->q_x_in, d->n_q_x_in, 1, net->H, net->W, &net->st[ST_QX], "q_x_in");
      if (rc != BP_OK) break;
      rc = build_stack(d->q_y_in, d->n_q_y_in, 2, net->H, net->W, &net->st[ST_QY], "q_y_in");
      if (rc != BP_OK) break;
      const Stack& qx = net->st[ST_QX];
      const Stack& qy = net->st[ST_QY];
      if (!(qx.OH == qy.OH && qx.OW == qy.OW)) {
        set_error("q_x_in and q_y_in outputs differ in size: (%d, %d) vs (%d, %d)", qx.OH, qx.OW, qy.OH, qy.OW);
        rc = BP_E_INVALID; break;
      }
      rc = build_stack(d->q_out, d->n_q_out, qx.out_c + qy.out_c, qx.OH, qx.OW, &net->st[ST_QOUT], "q_out");
      if (rc != BP_OK) break;
      const Stack& qo = net->st[ST_QOUT];
      if (!(qo.out_c == 2 && qo.OH == net->lh && qo.OW == net->lw)) {
        set_error("Dimension of z_mu does not match dim_z: (%d, %d, %d) vs (1, %d, %d).", qo.out_c / 2, qo.OH, qo.OW, net->lh, net->lw);
        rc = BP_E_INVALID; break;
      }
      net->nstacks = kMaxStacks;
      net->likelihood_scaling = d->likelihood_scaling != 0.f ? d->likelihood_scaling : 1.f;
    }
    rc = finish_create(net);
  } while (0);
  if (rc != BP_OK) { destroy_net(net); return rc; }
  *out = net;
  return BP_OK;
}

int bp_cgan_create(const bp_layer_desc* layers, int n_layers, int tile_h, int tile_w, int precision, int max_batch,
                   int device, bp_net** out) {
  BP_REQUIRE(layers && out && n_layers > 0, BP_E_INVALID, "null argument");
  *out = nullptr;
  BP_REQUIRE(precision == BP_PREC_F32 || precision == BP_PREC_BF16 || precision == BP_PREC_F16 || precision == BP_PREC_F32_FFMA,
             BP_E_INVALID, "bad precision %d", precision);
  BP_REQUIRE(max_batch > 0 && tile_h > 0 && tile_w > 0 && (tile_h * tile_w) % 4 == 0, BP_E_INVALID, "bad shape");
  int rc = check_device(device);
  if (rc != BP_OK) return rc;
  bp_net* net = new (std::nothrow) bp_net();
  BP_REQUIRE(net, BP_E_NOMEM, "out of host memory");
  net->device = device; net->kind = NET_CGAN; net->prec = precision; net->max_batch = max_batch;
  net->split = precision == BP_PREC_F32;
  net->H = tile_h; net->W = tile_w; net->in_c = 2; net->nstacks = 1;
  rc = build_stack(layers, n_layers, 2, tile_h, tile_w, &net->st[ST_GEN], "generator");
  if (rc == BP_OK) {
    const Stack& g = net->st[ST_GEN];
    if (!(g.out_c == 1 && g.OH == tile_h && g.OW == tile_w)) {
      set_error("generator output (%d, %d, %d) does not match the tile (1, %d, %d)", g.out_c, g.OH, g.OW, tile_h,
                tile_w);
      rc = BP_E_INVALID;
    }
  }
  if (rc == BP_OK) rc = finish_create(net);
  if (rc != BP_OK) { destroy_net(net); return rc; }
  *out = net;
  return BP_OK;
}

void bp_net_destroy(bp_net* net) { destroy_net(net); }

int bp_cvae_paint(bp_net* net, const float* tiles, const float* latent, int latent_mode, uint64_t seed,
                  const bp_transform_params* tp, int flags, float* out, int n, void* stream) {
  return cvae_paint_device(net, tiles, latent, latent_mode, seed, tp, flags, out, n, (cudaStream_t)stream);
}

int bp_cgan_paint(bp_net* net, const float* tiles, const bp_transform_params* tp, int flags, float* out, int n,
                  void* stream) {
  return cgan_paint_device(net, tiles, tp, flags, out, n, (cudaStream_t)stream);
}


// Host buffers in, host buffers out.  The batch is processed in the network's chunks on three streams:
// copy-in (H2D of chunk c+1), compute (chunk c), copy-out (D2H of chunk c-1), chained by events, so PCIe traffic in
// both directions overlaps the kernels.  Buffers that are already page-locked (cudaHostAlloc / cudaHostRegister /
// torch pin_memory) are used in place; pageable ones are staged through the network's pinned buffers, chunk by chunk,
// while the device works on the previous chunk.
int bp_cvae_paint_host(bp_net* net, const float* tiles, const float* latent, int latent_mode, uint64_t seed,
                       const bp_transform_params* tp, int flags, float* out, int n) {
  BP_REQUIRE(net && net->kind == NET_CVAE, BP_E_INVALID, "not a CVAE network");
  BP_REQUIRE(n >= 0 && n <= net->max_batch, BP_E_INVALID, "batch %d exceeds max_batch %d", n, net->max_batch);
  if (n == 0) return BP_OK;
  BP_REQUIRE(tiles && out, BP_E_INVALID, "null tile pointer");
  BP_REQUIRE(latent_mode == BP_LATENT_SEED || latent != nullptr, BP_E_INVALID, "latent/eps array missing");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  const size_t lhw = (size_t)net->lh * net->lw;
  cudaStream_t s = net->stream, sin = net->copy_stream;
  if (latent_mode != BP_LATENT_SEED) {
    memcpy(net->h_lat, latent, sizeof(float) * lhw * n);
    BP_CUDA_TRY(cudaMemcpyAsync(net->d_lat, net->h_lat, sizeof(float) * lhw * n, cudaMemcpyHostToDevice, sin));
  }
  return paint_host_pipelined(net, tiles, out, n, [&](const ChunkHooks& hooks) {
    return cvae_paint_device(net, net->d_in, net->d_lat, latent_mode, seed, tp, flags, net->d_out, n, s, &hooks);
  });
}

int bp_cgan_paint_host(bp_net* net, const float* tiles, const bp_transform_params* tp, int flags, float* out,
                       int n) {
  BP_REQUIRE(net && net->kind == NET_CGAN, BP_E_INVALID, "not a CGAN network");
  BP_REQUIRE(n >= 0 && n <= net->max_batch, BP_E_INVALID, "batch %d exceeds max_batch %d", n, net->max_batch);
  if (n == 0) return BP_OK;
  BP_REQUIRE(tiles && out, BP_E_INVALID, "null tile pointer");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  cudaStream_t s = net->stream;
  return paint_host_pipelined(net, tiles, out, n, [&](const ChunkHooks& hooks) {
    return cgan_paint_device(net, net->d_in, tp, flags, net->d_out, n, s, &hooks);
  });
}

// Asynchronous host entry point for a STREAM of batches: returns once the batch is enqueued; bp_net_wait(net, slot) returns
// when its result is in `out`.  With the three slots used in turn and two batches kept outstanding, the upload of batch
// k + 1 and the download of batch k - 1 overlap the kernels of batch k, which run as whole plan chunks (the synchronous entry point has to cut one batch
// into small pipeline chunks to overlap anything, and pays their launch overheads).  `tiles`, `out` (and `latent`) must stay
// valid and untouched until the wait; they should be page-locked (pageable buffers make the copies synchronous).
int bp_cvae_paint_host_async(bp_net* net, const float* tiles, const float* latent, int latent_mode, uint64_t seed,
                             const bp_transform_params* tp, int flags, float* out, int n, int slot) {
  BP_REQUIRE(net && net->kind == NET_CVAE, BP_E_INVALID, "not a CVAE network");
  BP_REQUIRE(slot >= 0 && slot < kIoSlots, BP_E_INVALID, "slot must be 0 .. %d", kIoSlots - 1);
  BP_REQUIRE(n > 0 && n <= net->max_batch && tiles && out, BP_E_INVALID, "bad batch / null tile pointer");
  BP_REQUIRE(latent_mode == BP_LATENT_SEED || latent != nullptr, BP_E_INVALID, "latent/eps array missing");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  const size_t HW = (size_t)net->H * net->W, lhw = (size_t)net->lh * net->lw;
  bp_net::IoSlot& io = net->io[slot];
  if (!io.d_in) {
    BP_CUDA_TRY(cudaMalloc(&io.d_in, sizeof(float) * HW * net->max_batch));
    BP_CUDA_TRY(cudaMalloc(&io.d_out, sizeof(float) * HW * net->max_batch));
    BP_CUDA_TRY(cudaMalloc(&io.d_lat, sizeof(float) * lhw * net->max_batch));
    BP_CUDA_TRY(cudaMallocHost(&io.h_lat, sizeof(float) * lhw * net->max_batch));
    BP_CUDA_TRY(cudaMallocHost(&io.h_par, sizeof(float) * 3 * net->max_batch));
    for (cudaEvent_t* e : {&io.h2d, &io.done, &io.out}) BP_CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  cudaStream_t s = net->stream, sin = net->copy_stream, sout = net->out_stream;
  if (io.used) {
    BP_CUDA_TRY(cudaEventSynchronize(io.out));             // the slot's previous result has left the device
    BP_CUDA_TRY(cudaStreamWaitEvent(sin, io.done, 0));     // ... and its inputs are no longer read
  }
  if (latent_mode != BP_LATENT_SEED) {
    memcpy(io.h_lat, latent, sizeof(float) * lhw * n);
    BP_CUDA_TRY(cudaMemcpyAsync(io.d_lat, io.h_lat, sizeof(float) * lhw * n, cudaMemcpyHostToDevice, sin));
  }
  BP_CUDA_TRY(cudaMemcpyAsync(io.d_in, tiles, sizeof(float) * HW * n, cudaMemcpyHostToDevice, sin));
  BP_CUDA_TRY(cudaEventRecord(io.h2d, sin));
  BP_CUDA_TRY(cudaStreamWaitEvent(s, io.h2d, 0));
  int rc = cvae_paint_device(net, io.d_in, io.d_lat, latent_mode, seed, tp, flags, io.d_out, n, s, nullptr, io.h_par);
  if (rc != BP_OK) return rc;
  BP_CUDA_TRY(cudaEventRecord(io.done, s));
  BP_CUDA_TRY(cudaStreamWaitEvent(sout, io.done, 0));
  BP_CUDA_TRY(cudaMemcpyAsync(out, io.d_out, sizeof(float) * HW * n, cudaMemcpyDeviceToHost, sout));
  BP_CUDA_TRY(cudaEventRecord(io.out, sout));
  io.used = true;
  return BP_OK;
}

int bp_net_wait(bp_net* net, int slot) {
  BP_REQUIRE(net && slot >= 0 && slot < kIoSlots, BP_E_INVALID, "bad net / slot");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  if (net->io[slot].used) BP_CUDA_TRY(cudaEventSynchronize(net->io[slot].out));
  return BP_OK;
}

int bp_cvae_read_prior(bp_net* net, float* z_mu, float* z_log_var, int n) {
  BP_REQUIRE(net && net->kind == NET_CVAE && z_mu && z_log_var, BP_E_INVALID, "bad argument");
  BP_REQUIRE(n > 0 && n <= net->prior_valid, BP_E_INVALID,
             "prior of %d tiles requested but the last paint computed %d", n, net->prior_valid);
  BP_CUDA_TRY(cudaSetDevice(net->device));
  const size_t lhw = (size_t)net->lh * net->lw;
  BP_CUDA_TRY(cudaDeviceSynchronize());
  BP_CUDA_TRY(cudaMemcpy(z_mu, net->prior_all, sizeof(float) * lhw * n, cudaMemcpyDeviceToHost));
  BP_CUDA_TRY(cudaMemcpy(z_log_var, net->prior_all + (size_t)net->max_batch * lhw, sizeof(float) * lhw * n,
                         cudaMemcpyDeviceToHost));
  return BP_OK;
}

int bp_cvae_paint_variance_host(bp_net* net, const float* tiles, const bp_transform_params* tp, int n_draws,
                                uint64_t seed, float* mean_out, float* var_out, int n) {
  BP_REQUIRE(net && net->kind == NET_CVAE, BP_E_INVALID, "not a CVAE network");
  BP_REQUIRE(n > 0 && n <= net->max_batch && n_draws > 0, BP_E_INVALID, "bad batch / draw count");
  BP_REQUIRE(tiles && mean_out && var_out, BP_E_INVALID, "null pointer");
  BP_REQUIRE(net->st[ST_PRIOR].layers.size() > 0, BP_E_UNSUPPORTED, "network has no prior network");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  const size_t HW = (size_t)net->H * net->W, lhw = (size_t)net->lh * net->lw;
  cudaStream_t s = net->stream;
  if (!net->var_mean) {
    BP_CUDA_TRY(cudaMalloc(&net->var_mean, sizeof(double) * HW * net->max_batch));
    BP_CUDA_TRY(cudaMalloc(&net->var_m2, sizeof(double) * HW * net->max_batch));
  }
  const int flags = BP_FLAG_TRANSFORM | BP_FLAG_INVERSE;
  memcpy(net->h_in, tiles, sizeof(float) * HW * n);
  BP_CUDA_TRY(cudaMemcpyAsync(net->d_in, net->h_in, sizeof(float) * HW * n, cudaMemcpyHostToDevice, s));
  int rc = upload_params(net, tp, flags, n, s);
  if (rc != BP_OK) return rc;
  // Draws are batched with tiles: a pass paints R draws of every tile of the group at once (the tiles and their
  // transform parameters replicated R times), so the launches stay plan-chunk sized whatever the tile count --
  // one 16-tile launch per draw ran the kernels at 60 % of their full-chunk rate.  (16-bit engine only: the fp32
  // path keeps (y, z) of the front pass in place per batch entry.)
  // (the split-precision path keeps (y, z) of the front pass in in_cat per batch entry: one draw per pass there)
  const bool can_rep = net->v2.built && net->v2.front_on && !net->debug && !dev_env("BP_VAR_NOREP");   // env: test aid
  static const bool vtrace = dev_env("BP_HOST_TRACE") != nullptr;
  const auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double vt0 = now_ms();
  for (int c0 = 0; c0 < n; c0 += net->chunk) {
    const int nb = std::min(net->chunk, n - c0);
    const int R = can_rep ? std::max(1, std::min(n_draws, net->chunk / nb)) : 1;
    rc = upload_params(net, tp, flags, n, s);              // (a previous group may have left replicated ones)
    if (rc != BP_OK) return rc;
    float* prior_out = nullptr;
    rc = cvae_chunk_front(net, net->d_in, tp, flags, c0, nb, true, s, &prior_out);
    if (rc != BP_OK) return rc;
    for (int r = 0; r < R; ++r)
      BP_CUDA_TRY(cudaMemcpyAsync(net->prior_keep + (size_t)r * nb * 2 * lhw, prior_out, sizeof(float) * 2 * lhw * nb,
                                  cudaMemcpyDeviceToDevice, s));
    bp_transform_params tpr = *tp;
    std::vector<float> r_in, r_out, r_aux;
    const float* rep_tiles = net->d_in;
    int rep_c0 = c0;
    if (R > 1) {
      if (!net->var_rep) BP_CUDA_TRY(cudaMalloc(&net->var_rep, sizeof(float) * HW * net->chunk));
      for (int r = 0; r < R; ++r) {
        BP_CUDA_TRY(cudaMemcpyAsync(net->var_rep + (size_t)r * nb * HW, net->d_in + (size_t)c0 * HW, sizeof(float) * HW * nb,
                                    cudaMemcpyDeviceToDevice, s));
        for (int t = 0; t < nb; ++t) {
          r_in.push_back(tp->sigma_in ? tp->sigma_in[c0 + t] : 0.f);
          r_out.push_back(tp->sigma_out ? tp->sigma_out[c0 + t] : 0.f);
          r_aux.push_back(tp->aux[c0 + t]);
        }
      }
      tpr.sigma_in = tp->sigma_in ? r_in.data() : nullptr;
      tpr.sigma_out = tp->sigma_out ? r_out.data() : nullptr;
      tpr.aux = r_aux.data();
      rc = upload_params(net, &tpr, flags, R * nb, s);
      if (rc != BP_OK) return rc;
      BP_CUDA_TRY(cudaStreamSynchronize(s));               // the replicated host vectors are read by the copies
      rep_tiles = net->var_rep;
      rep_c0 = 0;
    }
    for (int d0 = 0; d0 < n_draws; d0 += R) {
      const int Rd = std::min(R, n_draws - d0), ne = Rd * nb;
      // counter-RNG offsets: group c0 owns [c0 * n_draws, (c0 + nb) * n_draws) * lhw, pass d0 a slice of it
      rc = launch_sample_z(net->prior_keep, nullptr, net->latent, nullptr, nullptr, net->min_z_var, ne, (int)lhw,
                           BP_LATENT_SEED, seed, ((uint64_t)c0 * n_draws + (uint64_t)d0 * nb) * lhw, s);
      if (rc != BP_OK) return rc;
      rc = cvae_chunk_back(net, rep_tiles, net->latent, R > 1 ? &tpr : tp, flags, rep_c0, ne, net->d_out, s);
      if (rc != BP_OK) return rc;
      for (int r = 0; r < Rd; ++r) {
        rc = launch_welford(net->d_out + (size_t)r * nb * HW, net->var_mean + (size_t)c0 * HW, net->var_m2 + (size_t)c0 * HW,
                            d0 + r + 1, HW * nb, s);
        if (rc != BP_OK) return rc;
      }
    }
  }
  if (vtrace) {
    cudaStreamSynchronize(s);
    fprintf(stderr, "[host] variance maps: %d tiles x %d draws painted in %.2f ms\n", n, n_draws, now_ms() - vt0);
  }
  // float32 mean into d_out, variance into d_in (the tiles are no longer needed), then to the host
  rc = launch_var_finalize(net->var_mean, net->var_m2, net->d_out, net->d_in, n_draws, HW * n, s);
  if (rc != BP_OK) return rc;
  BP_CUDA_TRY(cudaMemcpyAsync(net->h_out, net->d_out, sizeof(float) * HW * n, cudaMemcpyDeviceToHost, s));
  BP_CUDA_TRY(cudaMemcpyAsync(net->h_in, net->d_in, sizeof(float) * HW * n, cudaMemcpyDeviceToHost, s));
  BP_CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(mean_out, net->h_out, sizeof(float) * HW * n);
  memcpy(var_out, net->h_in, sizeof(float) * HW * n);
  return BP_OK;
}

// Evidence lower bound of a batch (reference cvae.py:122-147 with Q :68-80); see the header for the formulas.
int bp_cvae_elbo_host(bp_net* net, const float* x_tiles, const float* y_tiles, const float* eps, int latent_mode,
                      uint64_t seed, const bp_transform_params* tp, int flags, int n, double* stats, float* z_mu,
                      float* z_log_var) {
  BP_REQUIRE(net && net->kind == NET_CVAE, BP_E_INVALID, "not a CVAE network");
  BP_REQUIRE(!net->st[ST_QOUT].layers.empty(), BP_E_UNSUPPORTED, "the network was created without its recognition network (q_x_in / q_y_in / q_out)");
  BP_REQUIRE(!net->st[ST_PRIOR].layers.empty(), BP_E_UNSUPPORTED, "network has no prior network");
  BP_REQUIRE(n > 0 && n <= net->max_batch && x_tiles && y_tiles && stats && tp, BP_E_INVALID, "bad batch / null pointer");
  BP_REQUIRE(latent_mode == BP_LATENT_EPS || latent_mode == BP_LATENT_SEED, BP_E_INVALID, "ELBO needs eps or a seed");
  BP_REQUIRE(latent_mode == BP_LATENT_SEED || eps, BP_E_INVALID, "eps array missing");
  BP_REQUIRE(!(flags & BP_FLAG_TRANSFORM) || (tp->sigma_in && tp->sigma_out), BP_E_INVALID, "sigma_in / sigma_out missing");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  const size_t HW = (size_t)net->H * net->W, lhw = (size_t)net->lh * net->lw;
  const int mb = net->max_batch;
  cudaStream_t s = net->stream;
  if (!net->d_x) {
    BP_CUDA_TRY(cudaMalloc(&net->d_x, sizeof(float) * HW * net->max_batch));
    BP_CUDA_TRY(cudaMallocHost(&net->h_x, sizeof(float) * HW * net->max_batch));
  }
  memcpy(net->h_in, y_tiles, sizeof(float) * HW * n);
  memcpy(net->h_x, x_tiles, sizeof(float) * HW * n);
  BP_CUDA_TRY(cudaMemcpyAsync(net->d_in, net->h_in, sizeof(float) * HW * n, cudaMemcpyHostToDevice, s));
  BP_CUDA_TRY(cudaMemcpyAsync(net->d_x, net->h_x, sizeof(float) * HW * n, cudaMemcpyHostToDevice, s));
  if (latent_mode == BP_LATENT_EPS) {
    memcpy(net->h_lat, eps, sizeof(float) * lhw * n);
    BP_CUDA_TRY(cudaMemcpyAsync(net->d_lat, net->h_lat, sizeof(float) * lhw * n, cudaMemcpyHostToDevice, s));
  }
  // sigma_in -> params[0..], sigma_out -> params[mb..] (here: the FORWARD transform of x), aux -> params[2 mb..]
  int rc = upload_params(net, tp, (flags & BP_FLAG_TRANSFORM) ? (BP_FLAG_TRANSFORM | BP_FLAG_INVERSE) : 0, n, s);
  if (rc != BP_OK) return rc;
  BP_CUDA_TRY(cudaMemsetAsync(net->d_sums, 0, sizeof(double) * 2, s));
  const int do_t = (flags & BP_FLAG_TRANSFORM) ? 1 : 0;
  const Stack& qx = net->st[ST_QX];
  const Stack& qy = net->st[ST_QY];
  const size_t qhw = (size_t)qx.OH * qx.OW;
  const int cc = qx.out_c + qy.out_c;
  V2Plan& P = net->v2;
  PostOp none;
  for (int c0 = 0; c0 < n; c0 += net->chunk) {
    const int nb = std::min(net->chunk, n - c0);
    // x' and [y', z] as fp32 planes
    rc = launch_prepare(net->d_x + (size_t)c0 * HW, net->xq, (long long)HW, 0, -1, net->params + mb + c0, net->params + 2 * mb + c0,
                        tp->k_out, tp->shift_out, do_t, nb, (int)HW, s);
    if (rc != BP_OK) return rc;
    rc = launch_prepare(net->d_in + (size_t)c0 * HW, net->in_cat, 3 * (long long)HW, 1, 2, net->params + c0,
                        net->params + 2 * mb + c0, tp->k_in, tp->shift_in, do_t, nb, (int)HW, s);
    if (rc != BP_OK) return rc;
    // recognition network -> (z_mu, z_log_var), into buffers of its own (the prior network below and the decoder
    // reuse the rotating pool)
    const float* q_out = nullptr;
    if (P.built && P.q_on) {
      rc = v2_run(net, P.qx_ops, nullptr, 0, none, nb, s);
      if (rc == BP_OK) rc = v2_run(net, P.qy_ops, nullptr, 0, none, nb, s);
      if (rc == BP_OK) rc = v2_run(net, P.qout_ops, nullptr, 0, none, nb, s);
      if (rc != BP_OK) return rc;
      q_out = static_cast<const float*>(P.acts[P.q_out].ptr);
    } else {
      ActRef in;
      in.ptr = net->xq; in.bs = (long long)HW;
      rc = run_stack(net, ST_QX, in, net->q_cat32, (long long)cc * qhw, none, nb, s, nullptr);
      if (rc != BP_OK) return rc;
      in.ptr = net->in_cat + HW; in.bs = 3 * (long long)HW;
      rc = run_stack(net, ST_QY, in, net->q_cat32 + (size_t)qx.out_c * qhw, (long long)cc * qhw, none, nb, s, nullptr);
      if (rc != BP_OK) return rc;
      in.ptr = net->q_cat32; in.bs = (long long)cc * qhw;
      rc = run_stack(net, ST_QOUT, in, net->q_out32, 2 * (long long)lhw, none, nb, s, nullptr);
      if (rc != BP_OK) return rc;
      q_out = net->q_out32;
    }
    // prior network
    float* prior_out = nullptr;
    rc = cvae_chunk_front(net, net->d_in, tp, flags, c0, nb, true, s, &prior_out);
    if (rc != BP_OK) return rc;
    // z = z_mu + eps * (exp(z_log_var / 2) + min_z_var); Q's (z_mu, z_log_var) are kept in prior_all for the caller
    rc = launch_sample_z(q_out, latent_mode == BP_LATENT_EPS ? net->d_lat + (size_t)c0 * lhw : nullptr, net->latent,
                         net->prior_all + (size_t)c0 * lhw, net->prior_all + ((size_t)mb + c0) * lhw, net->min_z_var, nb,
                         (int)lhw, latent_mode, seed, (uint64_t)c0 * lhw, s);
    if (rc != BP_OK) return rc;
    rc = launch_kl_sum(q_out, prior_out, nb, (int)lhw, net->d_sums, s);
    if (rc != BP_OK) return rc;
    // x_mu = P(z, y) (no inverse transform), then the squared residual against x'
    rc = cvae_chunk_back(net, net->d_in, net->latent, tp, flags & ~BP_FLAG_INVERSE, c0, nb, net->d_out, s);
    if (rc != BP_OK) return rc;
    rc = launch_sqdiff_sum(net->xq, net->d_out, HW * nb, net->d_sums + 1, s);
    if (rc != BP_OK) return rc;
  }
  double sums[2] = {0, 0};
  BP_CUDA_TRY(cudaMemcpyAsync(sums, net->d_sums, sizeof(sums), cudaMemcpyDeviceToHost, s));
  BP_CUDA_TRY(cudaStreamSynchronize(s));
  const double kl = 0.5 / n * sums[0];
  const double ll = -0.5 * log(2.0 * 3.14159265358979323846) + (-0.5 * sums[1]) / n;
  stats[0] = -kl + (double)net->likelihood_scaling * ll;
  stats[1] = kl;
  stats[2] = ll;
  net->prior_valid = n;
  if (z_mu && z_log_var) {
    BP_CUDA_TRY(cudaMemcpy(z_mu, net->prior_all, sizeof(float) * lhw * n, cudaMemcpyDeviceToHost));
    BP_CUDA_TRY(cudaMemcpy(z_log_var, net->prior_all + (size_t)mb * lhw, sizeof(float) * lhw * n, cudaMemcpyDeviceToHost));
  }
  return BP_OK;
}

int bp_rng_normal_host(int device, uint64_t seed, uint64_t offset, float* out, size_t n) {
  BP_REQUIRE(out || n == 0, BP_E_INVALID, "null output");
  if (n == 0) return BP_OK;
  int rc = check_device(device);
  if (rc != BP_OK) return rc;
  float* d = nullptr;
  BP_CUDA_TRY(cudaMalloc(&d, sizeof(float) * n));
  rc = launch_rng_normal(d, seed, offset, n, 0);
  cudaError_t e = rc == BP_OK ? cudaMemcpy(out, d, sizeof(float) * n, cudaMemcpyDeviceToHost) : cudaSuccess;
  cudaFree(d);
  if (rc != BP_OK) return rc;
  BP_CUDA_TRY(e);
  return BP_OK;
}

int bp_net_set_debug(bp_net* net, int keep) {
  BP_REQUIRE(net, BP_E_INVALID, "null net");
  net->debug = keep != 0;
  return BP_OK;
}

int bp_net_read_activation(bp_net* net, int stack, int layer, float* out, size_t out_floats) {
  BP_REQUIRE(net && out, BP_E_INVALID, "null argument");
  BP_REQUIRE(stack >= 0 && stack < net->nstacks, BP_E_INVALID, "bad stack index %d", stack);
  BP_REQUIRE(layer >= 0 && (size_t)layer < net->dbg[stack].size() && net->dbg[stack][layer], BP_E_INVALID,
             "activation (%d, %d) was not recorded; call bp_net_set_debug(net, 1) before painting", stack, layer);
  const Layer& l = net->st[stack].layers[layer];
  const size_t want = (size_t)net->dbg_n * l.d.cout * l.OHF * l.OWF;
  BP_REQUIRE(out_floats == want, BP_E_INVALID, "activation (%d, %d) holds %zu floats, caller asked for %zu", stack,
             layer, want, out_floats);
  BP_CUDA_TRY(cudaSetDevice(net->device));
  BP_CUDA_TRY(cudaDeviceSynchronize());
  BP_CUDA_TRY(cudaMemcpy(out, net->dbg[stack][layer], sizeof(float) * want, cudaMemcpyDeviceToHost));
  return BP_OK;
}

double bp_net_flops_per_tile(const bp_net* net) { return net ? net->flops_per_tile : 0.0; }
int bp_net_chunk(const bp_net* net) { return net ? net->chunk : 0; }

int bp_net_set_profile(bp_net* net, int on) {
  BP_REQUIRE(net, BP_E_INVALID, "null net");
  for (auto& p : net->prof_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  net->prof_events.clear();
  net->prof_ids.clear();
  net->profile = on != 0;
  return BP_OK;
}

int bp_net_layer_info(const bp_net* net, int stack, int layer, double* flops, int* geom) {
  BP_REQUIRE(net && stack >= 0 && stack < net->nstacks, BP_E_INVALID, "bad stack index");
  BP_REQUIRE(layer >= 0 && (size_t)layer < net->st[stack].layers.size(), BP_E_INVALID, "bad layer index");
  const Layer& l = net->st[stack].layers[layer];
  if (flops) *flops = l.flops;
  if (geom) {
    geom[0] = l.d.kind; geom[1] = l.d.cin; geom[2] = l.d.cout; geom[3] = l.d.kernel; geom[4] = l.d.stride;
    geom[5] = l.H; geom[6] = l.W; geom[7] = l.OHF; geom[8] = l.OWF; geom[9] = l.v2 ? 3 : 0;
  }
  return BP_OK;
}

int bp_net_read_profile(bp_net* net, int stack, int layer, double* total_ms, int* launches) {
  BP_REQUIRE(net && total_ms && launches, BP_E_INVALID, "null argument");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  BP_CUDA_TRY(cudaDeviceSynchronize());
  double ms = 0;
  int cnt = 0;
  for (size_t i = 0; i < net->prof_ids.size(); ++i) {
    if (net->prof_ids[i] != stack * 1000 + layer) continue;
    float t = 0.f;
    BP_CUDA_TRY(cudaEventElapsedTime(&t, net->prof_events[i].first, net->prof_events[i].second));
    ms += t; ++cnt;
  }
  *total_ms = ms; *launches = cnt;
  return BP_OK;
}

}  // extern "C"
