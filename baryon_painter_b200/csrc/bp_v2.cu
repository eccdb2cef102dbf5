// Lowering of the network's convolution layers to window GEMMs (bp_wconv.h) -- host only.
//
//   stride-1 Conv2d (k, p)                   taps (r-p, q-p); G pixels x Jy rows of outputs per M row when
//                                            the layer is narrow (Toeplitz-expanded weights)
//   Conv2d with k = 2s, p = s/2 (k4s2p1,     a 2x2-tap stride-1 GEMM over the shifted space-to-depth layout its
//   k8s4p2)                                  producer writes (block = s x s pixels, s*s*Cin channels)
//   ConvTranspose2d with k = 2s, p = s/2     s*s sub-pixel phases, each a 2x2-tap GEMM on the input grid; phases
//                                            are merged into the GEMM N dimension (9 taps, zero-padded weights)
//                                            while s*s*Cout <= 128
// Reference modules: baryon_painter/models/utils.py:40-77 (conv_block scale 1/2/4), :128-131.
#include <cuda_fp16.h>
#include <string.h>

#include <algorithm>
#include <functional>

#include "bp_wconv.h"

namespace bp {

// ---- split precision (the fp32-accurate tensor-core path) ------------------------------------------------------
// A value v travels as hi = fp16(v), lo = fp16(v - hi); a product x*w is evaluated as x_hi*w_hi + x_lo*w_hi (pass 0)
// + x_hi*w_lo (pass 1) in fp32 accumulators -- the dropped x_lo*w_lo term is ~2^-22 of the product.  A pixel stores
// [Cp hi | Cp lo] channels, so the lowerings below see 2*Cp "channels": their weight functions return the full fp32
// weight and report which part (0 = hi, 1 = lo) the element belongs to; the builder (wconv_build) splits per pass.
// stored channels per pixel of a 16-bit NHWC tensor with `c` logical channels
int v2_padc(int c) { return c <= 4 ? 4 : (c + 7) / 8 * 8; }

static int round_pow2(int v, int lo) {
  int r = lo;
  while (r < v) r <<= 1;
  return r;
}

// can `l` run as a window GEMM?  *need_b = space-to-depth block of the input layout it reads
bool v2_eligible(const Layer& l, int* need_b, bool split) {
  const bp_layer_desc& d = l.d;
  const int sm = split ? 2 : 1;
  *need_b = 1;
  if (dev_env("BP_V2_OFF")) return false;
  if (d.cout > 128 || (d.cout & (d.cout - 1)) != 0) return false;
  if (d.cout < 8 && !(d.kind == BP_CONV && d.stride == 1 && (d.cout == 1 || d.cout == 2 || d.cout == 4))) return false;
  if (d.kind == BP_CONV) {
    if (d.stride == 1) return d.kernel == 2 * d.pad + 1;   // unit / packing chosen by v2_candidates
    // k <= 2s with p = s/2 embeds into the k = 2s family (zero-padded kernel): k4s2p1, k8s4p2, k3s2p1
    if (d.kernel > 2 * d.stride || d.pad * 2 != d.stride) return false;
    if ((d.stride & (d.stride - 1)) != 0) return false;
    if (l.H % d.stride || l.W % d.stride) return false;
    if (l.OHF != l.H / d.stride || l.OWF != l.W / d.stride) return false;
    const int cs = v2_padc(d.cin) * d.stride * d.stride * sm;
    if (cs % 16 != 0 || !(cs * 2 <= 128 || (cs * 2) % 128 == 0)) return false;
    *need_b = d.stride;
    return true;
  }
  // transposed
  // k <= 2s, p = s/2, output_padding = 2s - k (output = s * input): k4s2p1, k3s2p1 with output_padding 1
  if (d.kernel > 2 * d.stride || d.pad * 2 != d.stride || d.out_pad != 2 * d.stride - d.kernel) return false;
  return d.cin % 16 == 0 && (d.cin * sm * 2 <= 128 || (d.cin * sm * 2) % 128 == 0);
}

int v2_make_spec(const Layer& l, int fmt, bool split, WSpec* sp) {
  const bp_layer_desc& d = l.d;
  const int k = d.kernel, s = d.stride, p = d.pad;
  const std::vector<float>& w = l.host_weight;
  const std::vector<float>& sc = l.host_scale;
  const std::vector<float>& sh = l.host_shift;
  sp->fmt = fmt;
  sp->act = d.act;
  sp->act_param = d.act_param;
  sp->G = 1; sp->Jy = 1; sp->mode = W_FLAT;
  sp->split = split;
  const int cout = d.cout, cin = d.cin;
  const int coutp = round_pow2(cout, 8);
  const int Cp = v2_padc(cin);                 // stored channels per pixel and part
  if (d.kind == BP_CONV && s == 1) {
    sp->nphase = 1;
    sp->OHl = l.OHF; sp->OWl = l.OWF;
    for (int r = 0; r < k; ++r)
      for (int q = 0; q < k; ++q) sp->taps[0].push_back({r - p, q - p});
    sp->N = std::max(16, coutp);
    sp->seg_len = sp->N; sp->seg_valid = coutp;
    sp->segs[0].push_back({0, 0});
    sp->ry = sp->rx = 1;
    sp->shift.assign(sp->N, 0.f);
    for (int n = 0; n < cout; ++n) sp->shift[n] = sh[n];
    sp->weight = [=, &w, &sc](int, int t, int elem, int n, int* part_out) -> float {
      const int part = split ? elem / Cp : 0, c = split ? elem % Cp : elem;
      if (n >= cout || c >= cin || part > 1) return 0.f;
      const int r = t / k, q = t % k;
      { *part_out = part; return w[(((size_t)n * cin + c) * k + r) * k + q] * sc[n]; }
    };
    return BP_OK;
  }
  if (d.kind == BP_CONV) {
    // k = 2s on the shifted space-to-depth input: element = (sy*s + sx)*cin + c
    sp->nphase = 1;
    sp->OHl = l.OHF; sp->OWl = l.OWF;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) sp->taps[0].push_back({a, b});
    sp->N = std::max(16, coutp);
    sp->seg_len = sp->N; sp->seg_valid = coutp;
    sp->segs[0].push_back({0, 0});
    sp->ry = sp->rx = 1;
    sp->shift.assign(sp->N, 0.f);
    for (int n = 0; n < cout; ++n) sp->shift[n] = sh[n];
    sp->weight = [=, &w, &sc](int, int t, int elem, int n, int* part_out) -> float {
      if (n >= cout) return 0.f;
      const int a = t / 2, b = t % 2;
      const int Cpx = split ? 2 * Cp : Cp;
      const int sub = elem / Cpx, cc = elem % Cpx;
      const int part = cc / Cp, c = cc % Cp;
      if (sub >= s * s || c >= cin) return 0.f;
      const int sy = sub / s, sx = sub % s;
      const int r = a * s + sy, q = b * s + sx;
      if (r >= k || q >= k) return 0.f;                       // kernel zero-padded to 2s x 2s
      { *part_out = part; return w[(((size_t)n * cin + c) * k + r) * k + q] * sc[n]; }
    };
    return BP_OK;
  }
  // transposed convolution: phase (ph, pw) of output pixel (s*y + ph, s*x + pw)
  // every phase gets the full 2 x 2 tap grid of the k = 2s family; taps beyond a shorter kernel carry zero weights
  struct PTap { int dy, dx, r, q; };
  std::vector<std::vector<PTap>> ptaps(l.nphase);
  for (int pi = 0; pi < l.nphase; ++pi) {
    const int ph = l.phase[pi].ph, pw = l.phase[pi].pw;
    const int r0 = (ph + p) % s, qh = (ph + p) / s, c0 = (pw + p) % s, qw = (pw + p) / s;
    for (int aa = 0; aa < 2; ++aa)
      for (int bb = 0; bb < 2; ++bb) ptaps[pi].push_back({qh - aa, qw - bb, r0 + s * aa, c0 + s * bb});
  }
  sp->OHl = l.H; sp->OWl = l.W;
  sp->ry = sp->rx = s;
  const bool merge = l.nphase * coutp <= 128 && !dev_env("BP_V2_NOMERGE");
  if (merge) {
    sp->nphase = 1;
    std::vector<WTap> all;
    for (int pi = 0; pi < l.nphase; ++pi)
      for (const PTap& t : ptaps[pi]) {
        bool seen = false;
        for (const WTap& u : all) seen = seen || (u.dl == t.dy && u.du == t.dx);
        if (!seen) all.push_back({t.dy, t.dx});
      }
    std::sort(all.begin(), all.end(), [](const WTap& a, const WTap& b) { return a.dl != b.dl ? a.dl < b.dl : a.du < b.du; });
    sp->taps[0] = all;
    sp->N = std::max(16, l.nphase * coutp);
    sp->seg_len = coutp; sp->seg_valid = coutp;
    for (int pi = 0; pi < l.nphase; ++pi) sp->segs[0].push_back({l.phase[pi].ph, l.phase[pi].pw});
    sp->shift.assign(sp->N, 0.f);
    for (int pi = 0; pi < l.nphase; ++pi)
      for (int n = 0; n < cout; ++n) sp->shift[pi * coutp + n] = sh[n];
    const int nph = l.nphase;
    sp->weight = [=, &w, &sc](int, int t, int elem, int n, int* part_out) -> float {
      const int pi = n / coutp, co = n % coutp;
      const int part = elem / cin, c = elem % cin;             // transposed layers have cin % 16 == 0: Cp = cin
      if (pi >= nph || co >= cout || part > (split ? 1 : 0)) return 0.f;
      for (const PTap& pt : ptaps[pi])
        if (pt.dy == all[t].dl && pt.dx == all[t].du)
          { *part_out = part; return (pt.r < k && pt.q < k) ? w[(((size_t)c * cout + co) * k + pt.r) * k + pt.q] * sc[co] : 0.f; }
      return 0.f;
    };
    return BP_OK;
  }
  sp->nphase = l.nphase;
  sp->N = std::max(16, coutp);
  sp->seg_len = sp->N; sp->seg_valid = coutp;
  sp->shift.assign(sp->N, 0.f);
  for (int n = 0; n < cout; ++n) sp->shift[n] = sh[n];
  for (int pi = 0; pi < l.nphase; ++pi) {
    for (const PTap& t : ptaps[pi]) sp->taps[pi].push_back({t.dy, t.dx});
    sp->segs[pi].push_back({l.phase[pi].ph, l.phase[pi].pw});
  }
  sp->weight = [=, &w, &sc](int pi, int t, int elem, int n, int* part_out) -> float {
    const int part = elem / cin, c = elem % cin;
    if (n >= cout || part > (split ? 1 : 0)) return 0.f;
    const PTap& pt = ptaps[pi][t];
    if (pt.r >= k || pt.q >= k) return 0.f;
    { *part_out = part; return w[(((size_t)c * cout + n) * k + pt.r) * k + pt.q] * sc[n]; }
  };
  return BP_OK;
}


static int floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// stride-1 convolution producing a Jy x G pixel block per M row from units of G pixels (Cp stored channels each)
static int make_spec_packed(const Layer& l, int fmt, int Cp, bool split, int G, int Jy, WSpec* sp) {
  const bp_layer_desc& d = l.d;
  const int k = d.kernel, p = d.pad, cin = d.cin, cout = d.cout;
  const std::vector<float>& w = l.host_weight;
  const std::vector<float>& sc = l.host_scale;
  const std::vector<float>& sh = l.host_shift;
  sp->fmt = fmt; sp->act = d.act; sp->act_param = d.act_param;
  sp->G = G; sp->Jy = Jy;
  sp->mode = (G == 1 && Jy == 1) ? W_FLAT : W_LINE;
  sp->nphase = 1;
  sp->split = split;
  sp->OHl = (l.OHF + Jy - 1) / Jy;
  sp->OWl = l.OWF / G;
  for (int dl = -p; dl <= Jy - 1 + p; ++dl)
    for (int du = floordiv(-p, G); du <= floordiv(G - 1 + p, G); ++du) sp->taps[0].push_back({dl, du});
  sp->ry = Jy; sp->rx = G;
  const int coutp = cout == 1 ? 1 : round_pow2(cout, 8);
  if (cout == 1) {
    sp->N = std::max(16, (Jy * G + 15) / 16 * 16);
    sp->seg_len = G; sp->seg_valid = G;
    for (int jy = 0; jy < Jy; ++jy) sp->segs[0].push_back({jy, 0});
  } else {
    sp->N = std::max(16, (Jy * G * coutp + 15) / 16 * 16);
    sp->seg_len = coutp; sp->seg_valid = coutp;
    for (int jy = 0; jy < Jy; ++jy)
      for (int jx = 0; jx < G; ++jx) sp->segs[0].push_back({jy, jx});
  }
  sp->shift.assign(sp->N, 0.f);
  for (int n = 0; n < Jy * G * coutp; ++n)
    if (n % coutp < cout) sp->shift[n] = sh[n % coutp];
  std::vector<WTap> taps = sp->taps[0];
  sp->weight = [=, &w, &sc](int, int t, int elem, int n, int* part_out) -> float {
    if (n >= Jy * G * coutp) return 0.f;
    const int co = n % coutp, jx = (n / coutp) % G, jy = n / (coutp * G);
    const int Cpx = split ? 2 * Cp : Cp;
    const int pp = elem / Cpx, cc = elem % Cpx;
    const int part = cc / Cp, c = cc % Cp;
    if (co >= cout || c >= cin || pp >= G) return 0.f;
    const int r = taps[t].dl - jy + p, q = taps[t].du * G + pp - jx + p;
    if (r < 0 || r >= k || q < 0 || q >= k) return 0.f;
    { *part_out = part; return w[(((size_t)co * cin + c) * k + r) * k + q] * sc[co]; }
  };
  return BP_OK;
}

// J consecutive M lines of a single-phase formulation folded into one (Toeplitz expansion along y): the tap grid
// grows by (J - 1) * Jy lines, N by the factor J.  Narrow strided / transposed layers trade a few zero weights
// for J times the work per A-operand read.
static bool pack_lines(const WSpec& base, int J, WSpec* out) {
  if (base.nphase != 1 || base.N * J > 128 || (int)base.segs[0].size() * base.seg_len != base.N) return false;
  *out = base;
  out->mode = W_LINE;
  out->Jy = base.Jy * J;
  out->ry = base.ry * J;
  out->OHl = (base.OHl + J - 1) / J;
  out->N = base.N * J;
  std::vector<WTap> taps;
  for (int j = 0; j < J; ++j)
    for (const WTap& t : base.taps[0]) {
      const WTap u{t.dl + j * base.Jy, t.du};
      bool seen = false;
      for (const WTap& v : taps) seen = seen || (v.dl == u.dl && v.du == u.du);
      if (!seen) taps.push_back(u);
    }
  std::sort(taps.begin(), taps.end(), [](const WTap& a, const WTap& b) { return a.dl != b.dl ? a.dl < b.dl : a.du < b.du; });
  out->taps[0] = taps;
  out->segs[0].clear();
  for (int j = 0; j < J; ++j)
    for (const WSegOff& sg : base.segs[0]) out->segs[0].push_back({sg.oy + j * base.ry, sg.ox});
  out->shift.resize(out->N);
  for (int n = 0; n < out->N; ++n) out->shift[n] = base.shift[n % base.N];
  const std::vector<WTap> btaps = base.taps[0];
  const std::function<float(int, int, int, int, int*)> bw = base.weight;
  const int bN = base.N, bJy = base.Jy;
  out->weight = [=](int, int t, int elem, int n, int* part_out) -> float {
    const int j = n / bN, dl = taps[t].dl - j * bJy;
    for (size_t i = 0; i < btaps.size(); ++i)
      if (btaps[i].dl == dl && btaps[i].du == taps[t].du) return bw(0, (int)i, elem, n % bN, part_out);
    return 0.f;
  };
  return true;
}

static double mma_cycles(int N) { return std::max(45.5, std::max((4096.0 + 32.0 * N) / 128.0, N / 2.0)); }

// window-GEMM formulations of `l` reading a 16-bit NHWC input with `Cp` stored channels per pixel,
// cheapest (tensor-pipe issue cycles per output pixel) first; the caller takes the first that fits
int v2_candidates(const Layer& l, int fmt, int Cp, bool split, std::vector<WSpec>* out) {
  const bp_layer_desc& d = l.d;
  out->clear();
  const int Cpx = split ? 2 * Cp : Cp;         // stored 16-bit channels per pixel
  if (!(d.kind == BP_CONV && d.stride == 1)) {
    WSpec sp;
    int rc = v2_make_spec(l, fmt, split, &sp);
    if (rc != BP_OK) return rc;
    out->push_back(sp);
    // wide single-phase layers: also offer two / four output lines per M row (the build times them)
    if (sp.nphase == 1 && sp.OWl >= 128 && !dev_env("BP_V2_NOLINEPACK"))
      for (int J : {2, 4}) {
        WSpec pk;
        if (pack_lines(sp, J, &pk)) out->push_back(pk);
      }
    return BP_OK;
  }
  struct Cand { double cost; int G, Jy; };
  std::vector<Cand> cands;
  const int coutp = d.cout == 1 ? 1 : round_pow2(d.cout, 8);
  for (int G : {1, 2, 4, 8}) {
    const int ub = G * Cpx * 2;
    if (!(ub == 32 || ub == 64 || ub == 128 || (G == 1 && ub % 128 == 0))) continue;
    if (l.OWF % G) continue;
    for (int Jy : {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16}) {
      const int N = std::max(16, (Jy * G * coutp + 15) / 16 * 16);     // UMMA N: multiple of 16 (dead columns: zero weights)
      if (N > 128) continue;
      if ((Jy & (Jy - 1)) != 0 && N - Jy * G * coutp >= 8 && N > 16) continue;   // odd packings only where they fill N
      const bool flat = (G == 1 && Jy == 1);
      // k-steps per tap line: the 32-byte slices of the tap units that hold a pixel of [-p, G - 1 + p] (the builder
      // drops the all-zero slices at both ends)
      const int du0 = floordiv(-d.pad, G), du1 = floordiv(G - 1 + d.pad, G);
      int kline = 0;
      for (int b0 = du0 * ub; b0 < (du1 + 1) * ub; b0 += 32) {
        const int px0 = floordiv(b0, Cpx * 2), px1 = floordiv(b0 + 31, Cpx * 2);    // pixels (relative to unit 0) in this slice
        if (px1 >= -d.pad && px0 <= G - 1 + d.pad) ++kline;
      }
      const int rows = flat ? 128 : std::min(l.OWF / G, 128);
      cands.push_back({(Jy + d.kernel - 1) * kline * mma_cycles(N) / ((double)rows * G * Jy), G, Jy});
    }
  }
  std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.cost < b.cost; });
  if (const char* e = dev_env("BP_V2_PACK")) {       // "cin:cout:k:G:Jy" forces one formulation (tuning aid)
    int ci, co, kk, g, jy;
    if (sscanf(e, "%d:%d:%d:%d:%d", &ci, &co, &kk, &g, &jy) == 5 && ci == d.cin && co == d.cout && kk == d.kernel) {
      cands.clear();
      cands.push_back({0.0, g, jy});
    }
  }
  for (const Cand& c : cands) {
    WSpec sp;
    make_spec_packed(l, fmt, Cp, split, c.G, c.Jy, &sp);
    out->push_back(sp);
    // the same packing on 8-unit x 16-block-line M-tiles: a small patch where full-width lines do not fit (wide-halo
    // layers with multi-pixel units); the builder / table decide
    if (sp.OWl % 8 == 0 && sp.OHl >= 16 && (c.G >= 2 || d.kernel >= 7 || split) && !dev_env("BP_V2_NOBLOCK")) {
      sp.mode = W_BLOCK;
      out->push_back(sp);
    }
  }
  BP_REQUIRE(!out->empty(), BP_E_UNSUPPORTED, "no window-GEMM formulation for conv %d->%d k%d (Cp=%d)", d.cin, d.cout,
             d.kernel, Cp);
  return BP_OK;
}

}  // namespace bp
