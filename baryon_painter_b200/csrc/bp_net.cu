// Host side of the CUDA library: weight packing, the per-network execution plan and the C ABI
// declared in include/baryon_painter_b200.h.
//
// Replaces, per reference call site:
//   CVAE.__init__ / load_state_dict       baryon_painter/models/cvae.py:9-61, painter.py:431-432
//   CVAE.prior / sample_prior / P / sample_P   baryon_painter/models/cvae.py:82-120, 149-162
//   CVAEPainter.paint (device part)       baryon_painter/painter.py:375-390
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "bp_common.h"
#include "bp_tc.h"

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
  tc_free_layer(l);
  win_free_layer(l);
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
enum { ST_PRIOR = 0, ST_PZ = 1, ST_PYZ = 2, ST_MU = 3, ST_GEN = 0 };

struct bp_net {
  int device = 0, kind = NET_CVAE, prec = BP_PREC_F32, max_batch = 0, chunk = 0;
  int H = 0, W = 0, lh = 0, lw = 0, in_c = 0;
  float min_z_var = 1e-7f;
  Stack st[4];
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
  float *var_mean = nullptr, *var_m2 = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  int prior_valid = 0;
  bool debug = false;
  std::vector<float*> dbg[4];
  int dbg_n = 0;
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;  // one pair per recorded layer launch
  std::vector<int> prof_ids;                                     // stack*1000 + layer
  double flops_per_tile = 0;
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
  for (int s = 0; s < 4; ++s) {
    for (auto& l : net->st[s].layers) free_layer(&l);
    for (float* p : net->dbg[s]) cudaFree(p);
  }
  cudaFree(net->in_cat);
  for (int i = 0; i < 4; ++i) cudaFree(net->pool[i]);
  cudaFree(net->latent); cudaFree(net->prior_all); cudaFree(net->prior_keep); cudaFree(net->params);
  cudaFree(net->d_in); cudaFree(net->d_out); cudaFree(net->d_lat);
  cudaFree(net->var_mean); cudaFree(net->var_m2);
  if (net->h_in) cudaFreeHost(net->h_in);
  if (net->h_out) cudaFreeHost(net->h_out);
  if (net->h_lat) cudaFreeHost(net->h_lat);
  for (int i = 0; i < 2; ++i) {
    if (net->ev_in[i]) cudaEventDestroy(net->ev_in[i]);
    if (net->ev_done[i]) cudaEventDestroy(net->ev_done[i]);
  }
  if (net->stream) cudaStreamDestroy(net->stream);
  if (net->copy_stream) cudaStreamDestroy(net->copy_stream);
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

static int finish_create(bp_net* net) {
  const size_t HW = (size_t)net->H * net->W;
  size_t mx = 0;
  net->flops_per_tile = 0;
  for (int s = 0; s < net->nstacks; ++s) {
    mx = std::max(mx, net->st[s].max_floats);
    for (auto& l : net->st[s].layers) net->flops_per_tile += l.flops;
  }
  // chunk: keep the three rotating activation buffers of one chunk around L2 size (126 MB) so that
  // a layer's output is still cache-resident when the next layer reads it
  int chunk = net->prec == BP_PREC_F32 ? 4 : 8;
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
  for (int i = 0; i < 2; ++i) {
    BP_CUDA_TRY(cudaEventCreateWithFlags(&net->ev_in[i], cudaEventDisableTiming));
    BP_CUDA_TRY(cudaEventCreateWithFlags(&net->ev_done[i], cudaEventDisableTiming));
  }
  if (net->prec != BP_PREC_F32) {
    const int fmt = net->prec == BP_PREC_BF16 ? TC_FMT_BF16 : TC_FMT_F16;
    // sequences that stay in the 16-bit c8 layout: the prior network; p_z_in; p_y_z_in + p_mu_out
    // (CGAN: the generator).  A layer runs on the tensor cores when tc_layer_eligible() says so.
    for (int s = 0; s < net->nstacks; ++s) {
      const bool continues = net->kind == NET_CVAE && s == ST_MU;   // p_mu_out continues p_y_z_in
      std::vector<Layer>& L = net->st[s].layers;
      for (size_t i = 0; i < L.size(); ++i) {
        const bool first = (i == 0) && !continues;
        if (!tc_layer_eligible(L[i].d, first)) continue;
        if (!first) {
          // the producer of this layer's input must be a tensor-core layer too, unless we repack
          const Layer* prev = i > 0 ? &L[i - 1] : &net->st[ST_PYZ].layers.back();
          if (!prev->tc && !prev->win && (L[i].d.cin % 8) != 0) continue;
        }
        int rc = BP_E_UNSUPPORTED;
        if (win_layer_eligible(L[i].d, first)) rc = win_pack_layer(&L[i], fmt);
        if (rc == BP_E_UNSUPPORTED) rc = tc_pack_layer(&L[i], fmt);
        if (rc != BP_OK) return rc;
      }
      // a residual block runs on one path only
      for (size_t i = 0; i < L.size(); ++i)
        if (L[i].d.res == BP_RES_OPEN) {
          size_t j = i;
          while (L[j].d.res != BP_RES_CLOSE) ++j;
          bool all = true;
          for (size_t q = i; q <= j; ++q) all = all && (L[q].tc || L[q].win);
          if (!all) for (size_t q = i; q <= j; ++q) { tc_free_layer(&L[q]); win_free_layer(&L[q]); }
        }
    }
  }
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

// an activation tensor: fp32 NCHW with per-sample stride `bs`, or the 16-bit c8 layout of the
// tensor-core path ([nb][C/8][H][W][8], dense)
struct ActRef {
  const void* ptr = nullptr;
  long long bs = 0;
  bool c8 = false;
};

static int record_debug(bp_net* net, int sidx, int i, int nl, const Layer& l, const void* out, long long out_bs,
                        bool c8, int nb, cudaStream_t s) {
  const size_t per = (size_t)l.d.cout * l.OHF * l.OWF;
  if (net->dbg[sidx].size() < (size_t)nl) net->dbg[sidx].resize(nl, nullptr);
  if (!net->dbg[sidx][i]) BP_CUDA_TRY(cudaMalloc(&net->dbg[sidx][i], sizeof(float) * per * net->chunk));
  if (c8)
    return launch_unpack_c8(out, l.d.cout, l.OHF * l.OWF, net->dbg[sidx][i], (long long)per, nb,
                            net->prec == BP_PREC_BF16 ? TC_FMT_BF16 : TC_FMT_F16, s);
  BP_CUDA_TRY(cudaMemcpy2DAsync(net->dbg[sidx][i], per * sizeof(float), out, out_bs * sizeof(float),
                                per * sizeof(float), nb, cudaMemcpyDeviceToDevice, s));
  return BP_OK;
}

// run one sub-network on nb samples.  The last layer writes fp32 NCHW to final_out (stride final_bs)
// when given; otherwise to a pool buffer returned in *result (c8 if `c8_result_ok` and the last layer
// runs on the tensor cores).
static int run_stack(bp_net* net, int sidx, ActRef in, float* final_out, long long final_bs, const PostOp& post,
                     int nb, cudaStream_t s, ActRef* result, bool c8_result_ok = false) {
  Stack& st = net->st[sidx];
  ActRef cur = in;
  ActRef skip;
  const int nl = (int)st.layers.size();
  const int fmt = net->prec == BP_PREC_BF16 ? TC_FMT_BF16 : TC_FMT_F16;
  for (int i = 0; i < nl; ++i) {
    Layer& l = st.layers[i];
    const bool last = (i == nl - 1);
    const bool use_tc = l.tc != nullptr || l.win != nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (net->profile) {
      BP_CUDA_TRY(cudaEventCreate(&e0)); BP_CUDA_TRY(cudaEventCreate(&e1));
      BP_CUDA_TRY(cudaEventRecord(e0, s));
    }
    // ---- bring the input into the representation this layer's kernel reads
    if (use_tc && !cur.c8) {
      float* pk = pick_buffer(net, cur.ptr, skip.ptr);
      BP_REQUIRE(pk, BP_E_INVALID, "internal: no free activation buffer");
      int rc = launch_pack_c8(static_cast<const float*>(cur.ptr), cur.bs, l.d.cin, l.H * l.W, pk, nb, fmt, s);
      if (rc != BP_OK) return rc;
      cur.ptr = pk; cur.bs = 0; cur.c8 = true;
    } else if (!use_tc && cur.c8) {
      float* up = pick_buffer(net, cur.ptr, skip.ptr);
      BP_REQUIRE(up, BP_E_INVALID, "internal: no free activation buffer");
      const long long bs = (long long)l.d.cin * l.H * l.W;
      int rc = launch_unpack_c8(cur.ptr, l.d.cin, l.H * l.W, up, bs, nb, fmt, s);
      if (rc != BP_OK) return rc;
      cur.ptr = up; cur.bs = bs; cur.c8 = false;
    }
    if (l.d.res == BP_RES_OPEN) skip = cur;
    // ---- where the output goes
    const long long dense_bs = (long long)l.d.cout * l.OHF * l.OWF;
    ActRef out;
    if (last && final_out) {
      out.ptr = final_out; out.bs = final_bs; out.c8 = false;
    } else {
      const bool next_tc = last ? c8_result_ok : (st.layers[i + 1].tc != nullptr || st.layers[i + 1].win != nullptr);
      out.ptr = pick_buffer(net, cur.ptr, skip.ptr, in.ptr);
      BP_REQUIRE(out.ptr, BP_E_INVALID, "internal: no free activation buffer");
      out.c8 = use_tc && next_tc;
      out.bs = out.c8 ? 0 : dense_bs;
    }
    int rc;
    if (use_tc) {
      BP_REQUIRE(!(last && post.post != POST_NONE), BP_E_UNSUPPORTED,
                 "inverse transform fused into a tensor-core layer is not implemented");
      const void* sk = nullptr;
      if (l.d.res == BP_RES_CLOSE) {
        BP_REQUIRE(skip.c8, BP_E_INVALID, "internal: residual skip is not in the c8 layout");
        sk = skip.ptr;
      }
      void* o16 = out.c8 ? const_cast<void*>(out.ptr) : nullptr;
      float* o32 = out.c8 ? nullptr : static_cast<float*>(const_cast<void*>(out.ptr));
      rc = l.win ? launch_conv_win(l, cur.ptr, o16, o32, out.bs, sk, nb, s)
                 : launch_conv_tc(l, cur.ptr, o16, o32, out.bs, sk, nb, s);
    } else {
      ConvArgs a;
      memset(&a, 0, sizeof(a));
      a.in = static_cast<const float*>(cur.ptr); a.in_bs = cur.bs;
      a.out = static_cast<float*>(const_cast<void*>(out.ptr)); a.out_bs = out.bs; a.nb = nb;
      if (l.d.res == BP_RES_CLOSE) {
        BP_REQUIRE(!skip.c8, BP_E_INVALID, "internal: residual skip is not fp32");
        a.skip = static_cast<const float*>(skip.ptr); a.skip_bs = skip.bs;
      }
      if (last && post.post != POST_NONE) {
        a.post = post.post; a.post_sigma = post.sigma; a.post_k = post.k; a.post_shift = post.shift;
      }
      rc = launch_conv_f32(l, a, s);
    }
    if (rc != BP_OK) return rc;
    if (net->profile) {
      BP_CUDA_TRY(cudaEventRecord(e1, s));
      net->prof_events.emplace_back(e0, e1);
      net->prof_ids.push_back(sidx * 1000 + i);
    }
    if (l.d.res == BP_RES_CLOSE) skip = ActRef();
    if (net->debug) {
      rc = record_debug(net, sidx, i, nl, l, out.ptr, out.bs, out.c8, nb, s);
      if (rc != BP_OK) return rc;
    }
    cur = out;
  }
  if (result) *result = cur;
  return BP_OK;
}

static int upload_params(bp_net* net, const bp_transform_params* tp, int flags, int n, cudaStream_t s) {
  BP_REQUIRE(tp != nullptr && tp->aux != nullptr, BP_E_INVALID, "transform params / aux plane values missing");
  BP_REQUIRE(!(flags & BP_FLAG_TRANSFORM) || tp->sigma_in, BP_E_INVALID, "sigma_in missing");
  BP_REQUIRE(!(flags & BP_FLAG_INVERSE) || tp->sigma_out, BP_E_INVALID, "sigma_out missing");
  const int mb = net->max_batch;
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
  int rc = launch_prepare(tiles + (size_t)c0 * HW, net->in_cat, 3 * (long long)HW, 1, 2, net->params + c0,
                          net->params + 2 * mb + c0, tp->k_in, tp->shift_in, (flags & BP_FLAG_TRANSFORM) ? 1 : 0,
                          nb, (int)HW, s);
  if (rc != BP_OK) return rc;
  if (need_prior) {
    PostOp none;
    in.ptr = net->in_cat + HW; in.bs = 3 * (long long)HW;
    rc = run_stack(net, ST_PRIOR, in, nullptr, 0, none, nb, s, &res);
    if (rc != BP_OK) return rc;
    *prior_out = static_cast<float*>(const_cast<void*>(res.ptr));
  }
  return BP_OK;
}

// stage B: latent -> painted tile
static int cvae_chunk_back(bp_net* net, const float* latent, const bp_transform_params* tp, int flags, int c0,
                           int nb, float* out, cudaStream_t s) {
  const size_t HW = (size_t)net->H * net->W;
  const size_t lhw = (size_t)net->lh * net->lw;
  PostOp none;
  ActRef in, h;
  in.ptr = latent; in.bs = (long long)lhw;
  int rc = run_stack(net, ST_PZ, in, net->in_cat, 3 * (long long)HW, none, nb, s, nullptr);
  if (rc != BP_OK) return rc;
  in.ptr = net->in_cat; in.bs = 3 * (long long)HW;
  rc = run_stack(net, ST_PYZ, in, nullptr, 0, none, nb, s, &h, net->st[ST_MU].layers[0].tc != nullptr || net->st[ST_MU].layers[0].win != nullptr);
  if (rc != BP_OK) return rc;
  PostOp post;
  if (flags & BP_FLAG_INVERSE) {
    post.post = POST_INV_SHIFT_LOG; post.sigma = net->params + net->max_batch + c0;
    post.k = tp->k_out; post.shift = tp->shift_out;
  }
  return run_stack(net, ST_MU, h, out, (long long)HW, post, nb, s, nullptr);
}

static int cvae_paint_device(bp_net* net, const float* tiles, const float* latent, int mode, uint64_t seed,
                             const bp_transform_params* tp, int flags, float* out, int n, cudaStream_t s) {
  BP_REQUIRE(net && net->kind == NET_CVAE, BP_E_INVALID, "not a CVAE network");
  BP_REQUIRE(n >= 0 && n <= net->max_batch, BP_E_INVALID, "batch %d exceeds max_batch %d", n, net->max_batch);
  BP_REQUIRE(mode == BP_LATENT_GIVEN || mode == BP_LATENT_EPS || mode == BP_LATENT_SEED, BP_E_INVALID,
             "bad latent mode %d", mode);
  BP_REQUIRE(mode == BP_LATENT_SEED || latent != nullptr, BP_E_INVALID, "latent/eps array missing");
  if (n == 0) return BP_OK;
  BP_REQUIRE(tiles && out, BP_E_INVALID, "null tile pointer");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  int rc = upload_params(net, tp, flags, n, s);
  if (rc != BP_OK) return rc;
  const size_t HW = (size_t)net->H * net->W;
  const size_t lhw = (size_t)net->lh * net->lw;
  net->prior_valid = 0;
  if (net->debug) net->dbg_n = std::min(n, net->chunk);
  for (int c0 = 0; c0 < n; c0 += net->chunk) {
    const int nb = std::min(net->chunk, n - c0);
    float* prior_out = nullptr;
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
    rc = cvae_chunk_back(net, lat, tp, flags, c0, nb, out + (size_t)c0 * HW, s);
    if (rc != BP_OK) return rc;
    if (net->debug) break;  // debug buffers hold one chunk
  }
  if (mode != BP_LATENT_GIVEN) net->prior_valid = n;
  return BP_OK;
}

static int cgan_paint_device(bp_net* net, const float* tiles, const bp_transform_params* tp, int flags, float* out,
                             int n, cudaStream_t s) {
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
  for (int c0 = 0; c0 < n; c0 += net->chunk) {
    const int nb = std::min(net->chunk, n - c0);
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
extern "C" {

int bp_version(void) { return BP_VERSION; }
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
  BP_REQUIRE(precision == BP_PREC_F32 || precision == BP_PREC_BF16 || precision == BP_PREC_F16, BP_E_INVALID,
             "bad precision %d", precision);
  BP_REQUIRE(max_batch > 0, BP_E_INVALID, "max_batch must be positive");
  BP_REQUIRE(d->tile_h > 0 && d->tile_w > 0 && d->latent_h > 0 && d->latent_w > 0, BP_E_INVALID, "bad tile shape");
  BP_REQUIRE((d->tile_h * d->tile_w) % 4 == 0, BP_E_INVALID, "tile area must be a multiple of 4");
  BP_REQUIRE(d->n_p_z_in > 0 && d->n_p_y_z_in > 0 && d->n_p_mu_out > 0, BP_E_INVALID, "empty decoder stack");
  int rc = check_device(device);
  if (rc != BP_OK) return rc;
  bp_net* net = new (std::nothrow) bp_net();
  BP_REQUIRE(net, BP_E_NOMEM, "out of host memory");
  net->device = device; net->kind = NET_CVAE; net->prec = precision; net->max_batch = max_batch;
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
  BP_REQUIRE(precision == BP_PREC_F32 || precision == BP_PREC_BF16 || precision == BP_PREC_F16, BP_E_INVALID,
             "bad precision %d", precision);
  BP_REQUIRE(max_batch > 0 && tile_h > 0 && tile_w > 0 && (tile_h * tile_w) % 4 == 0, BP_E_INVALID, "bad shape");
  int rc = check_device(device);
  if (rc != BP_OK) return rc;
  bp_net* net = new (std::nothrow) bp_net();
  BP_REQUIRE(net, BP_E_NOMEM, "out of host memory");
  net->device = device; net->kind = NET_CGAN; net->prec = precision; net->max_batch = max_batch;
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

int bp_cvae_paint_host(bp_net* net, const float* tiles, const float* latent, int latent_mode, uint64_t seed,
                       const bp_transform_params* tp, int flags, float* out, int n) {
  BP_REQUIRE(net && net->kind == NET_CVAE, BP_E_INVALID, "not a CVAE network");
  BP_REQUIRE(n >= 0 && n <= net->max_batch, BP_E_INVALID, "batch %d exceeds max_batch %d", n, net->max_batch);
  if (n == 0) return BP_OK;
  BP_REQUIRE(tiles && out, BP_E_INVALID, "null tile pointer");
  BP_REQUIRE(latent_mode == BP_LATENT_SEED || latent != nullptr, BP_E_INVALID, "latent/eps array missing");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  const size_t HW = (size_t)net->H * net->W, lhw = (size_t)net->lh * net->lw;
  cudaStream_t s = net->stream;
  // stage through pinned memory so the copies are truly asynchronous DMA transfers
  memcpy(net->h_in, tiles, sizeof(float) * HW * n);
  BP_CUDA_TRY(cudaMemcpyAsync(net->d_in, net->h_in, sizeof(float) * HW * n, cudaMemcpyHostToDevice, s));
  if (latent_mode != BP_LATENT_SEED) {
    memcpy(net->h_lat, latent, sizeof(float) * lhw * n);
    BP_CUDA_TRY(cudaMemcpyAsync(net->d_lat, net->h_lat, sizeof(float) * lhw * n, cudaMemcpyHostToDevice, s));
  }
  int rc = cvae_paint_device(net, net->d_in, net->d_lat, latent_mode, seed, tp, flags, net->d_out, n, s);
  if (rc != BP_OK) return rc;
  BP_CUDA_TRY(cudaMemcpyAsync(net->h_out, net->d_out, sizeof(float) * HW * n, cudaMemcpyDeviceToHost, s));
  BP_CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(out, net->h_out, sizeof(float) * HW * n);
  return BP_OK;
}

int bp_cgan_paint_host(bp_net* net, const float* tiles, const bp_transform_params* tp, int flags, float* out,
                       int n) {
  BP_REQUIRE(net && net->kind == NET_CGAN, BP_E_INVALID, "not a CGAN network");
  BP_REQUIRE(n >= 0 && n <= net->max_batch, BP_E_INVALID, "batch %d exceeds max_batch %d", n, net->max_batch);
  if (n == 0) return BP_OK;
  BP_REQUIRE(tiles && out, BP_E_INVALID, "null tile pointer");
  BP_CUDA_TRY(cudaSetDevice(net->device));
  const size_t HW = (size_t)net->H * net->W;
  cudaStream_t s = net->stream;
  memcpy(net->h_in, tiles, sizeof(float) * HW * n);
  BP_CUDA_TRY(cudaMemcpyAsync(net->d_in, net->h_in, sizeof(float) * HW * n, cudaMemcpyHostToDevice, s));
  int rc = cgan_paint_device(net, net->d_in, tp, flags, net->d_out, n, s);
  if (rc != BP_OK) return rc;
  BP_CUDA_TRY(cudaMemcpyAsync(net->h_out, net->d_out, sizeof(float) * HW * n, cudaMemcpyDeviceToHost, s));
  BP_CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(out, net->h_out, sizeof(float) * HW * n);
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
    BP_CUDA_TRY(cudaMalloc(&net->var_mean, sizeof(float) * HW * net->max_batch));
    BP_CUDA_TRY(cudaMalloc(&net->var_m2, sizeof(float) * HW * net->max_batch));
  }
  const int flags = BP_FLAG_TRANSFORM | BP_FLAG_INVERSE;
  memcpy(net->h_in, tiles, sizeof(float) * HW * n);
  BP_CUDA_TRY(cudaMemcpyAsync(net->d_in, net->h_in, sizeof(float) * HW * n, cudaMemcpyHostToDevice, s));
  int rc = upload_params(net, tp, flags, n, s);
  if (rc != BP_OK) return rc;
  for (int c0 = 0; c0 < n; c0 += net->chunk) {
    const int nb = std::min(net->chunk, n - c0);
    for (int d = 0; d < n_draws; ++d) {
      // the prior depends only on the tile, but the rotating activation buffers are reused by the
      // decoder, so (z_mu, z_log_var) are kept in prior_all after the first draw
      float* prior_out = nullptr;
      if (d == 0) {
        rc = cvae_chunk_front(net, net->d_in, tp, flags, c0, nb, true, s, &prior_out);
        if (rc != BP_OK) return rc;
        BP_CUDA_TRY(cudaMemcpyAsync(net->prior_keep, prior_out, sizeof(float) * 2 * lhw * nb, cudaMemcpyDeviceToDevice, s));
      }  // later draws: in_cat channel 0 is overwritten by p_z_in; channels 1-2 (y, z) stay valid
      rc = launch_sample_z(net->prior_keep, nullptr, net->latent, nullptr, nullptr, net->min_z_var, nb, (int)lhw,
                           BP_LATENT_SEED, seed, ((uint64_t)d * n + c0) * lhw, s);
      if (rc != BP_OK) return rc;
      rc = cvae_chunk_back(net, net->latent, tp, flags, c0, nb, net->d_out + (size_t)c0 * HW, s);
      if (rc != BP_OK) return rc;
      rc = launch_welford(net->d_out + (size_t)c0 * HW, net->var_mean + (size_t)c0 * HW,
                          net->var_m2 + (size_t)c0 * HW, d + 1, HW * nb, s);
      if (rc != BP_OK) return rc;
    }
  }
  rc = launch_var_finalize(net->var_m2, n_draws, HW * n, s);
  if (rc != BP_OK) return rc;
  BP_CUDA_TRY(cudaMemcpyAsync(net->h_out, net->var_mean, sizeof(float) * HW * n, cudaMemcpyDeviceToHost, s));
  BP_CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(mean_out, net->h_out, sizeof(float) * HW * n);
  BP_CUDA_TRY(cudaMemcpyAsync(net->h_out, net->var_m2, sizeof(float) * HW * n, cudaMemcpyDeviceToHost, s));
  BP_CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(var_out, net->h_out, sizeof(float) * HW * n);
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
    geom[5] = l.H; geom[6] = l.W; geom[7] = l.OHF; geom[8] = l.OWF; geom[9] = l.win ? 2 : (l.tc ? 1 : 0);
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
