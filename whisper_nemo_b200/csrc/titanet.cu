// TitaNet-L behind two composite C-ABI entry points:
//
//   b200d_titanet_pack_weights   one-time, HOST: upstream NeMo's state_dict (names + fp32 arrays) -> one position-independent
//                                blob of device-ready operands + a plain descriptor of where each lies.  BatchNorm folded
//                                into the fp16 pointwise weights, the k = 1 depthwise of the last block folded into its
//                                pointwise, the TDNN context term [mean | std] split off the per-frame GEMM, the embedding
//                                BatchNorm folded into the 6144 -> 192 projection, the 192 -> 16 681 logits layer dropped,
//                                the mel filterbank in sparse (start, offset, weights) form.
//   b200d_titanet_forward        waveform + window descriptors -> float32 [n_seg][192] embeddings: featurizer, 14 sub-blocks
//                                of depthwise conv + tcgen05 pointwise GEMM, squeeze-excite, attentive statistics pooling and
//                                the embedding projection, ~60 launches per group of windows, all enqueued on the caller's
//                                stream from C++ (no Python between kernels).
//
// Upstream: nemo/collections/asr/models/label_models.py (EncDecSpeakerLabelModel.forward), modules/conv_asr.py
// (ConvASREncoder, SpeakerDecoder), parts/submodules/jasper.py (JasperBlock), parts/submodules/tdnn_attention.py.
#include "common.cuh"
#include "composite.cuh"

#include <atomic>
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace b200d {

// ------------------------------------------------------------------------------------ profiling spans (composite.cuh)
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<Span> g_spans;
std::atomic<long long> g_composite_launches{0};  // kernels launched from inside composite entry points (b200d_launch_count)

ProfScope::ProfScope(const char* name, double work, cudaStream_t stream, int kernels) : on(false), st(stream) {
  g_composite_launches.fetch_add(kernels, std::memory_order_relaxed);
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    on = g_prof_on;
  }
  if (!on) return;
  snprintf(sp.name, sizeof(sp.name), "%s", name);
  sp.work = work;
  cudaEventCreate(&sp.e0);
  cudaEventCreate(&sp.e1);
  cudaEventRecord(sp.e0, st);
}
ProfScope::~ProfScope() {
  if (!on) return;
  cudaEventRecord(sp.e1, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_spans.push_back(sp);
}



// ------------------------------------------------------------------------------------ packing
struct Tensor {
  const float* p;
  int64_t n;
};
using Dict = std::map<std::string, Tensor>;

static const Tensor* find(const Dict& d, const std::string& k) {
  auto it = d.find(k);
  return it == d.end() ? nullptr : &it->second;
}

static inline uint16_t f2h(double v) {  // torch's double -> Half goes through float; mirror it
  const __half h = __float2half_rn(static_cast<float>(v));
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

struct Writer {  // layout pass (base == nullptr) and fill pass share the code
  uint8_t* base;
  int64_t pos = 0;
  int64_t take(int64_t bytes) {
    const int64_t at = pos;
    pos = (pos + bytes + 255) & ~static_cast<int64_t>(255);
    return at;
  }
  template <typename T>
  T* at(int64_t off) const { return base ? reinterpret_cast<T*>(base + off) : nullptr; }
};

static void bn_fold(const Dict& d, const std::string& pre, double eps, std::vector<double>& s, std::vector<double>& t) {
  const Tensor *g = find(d, pre + ".weight"), *b = find(d, pre + ".bias"), *m = find(d, pre + ".running_mean"), *v = find(d, pre + ".running_var");
  const int64_t n = g->n;
  s.resize(n);
  t.resize(n);
  for (int64_t i = 0; i < n; ++i) {
    s[i] = static_cast<double>(g->p[i]) / std::sqrt(static_cast<double>(v->p[i]) + eps);
    t[i] = static_cast<double>(b->p[i]) - static_cast<double>(m->p[i]) * s[i];
  }
}
static bool has_bn(const Dict& d, const std::string& pre) {
  return find(d, pre + ".weight") && find(d, pre + ".bias") && find(d, pre + ".running_mean") && find(d, pre + ".running_var");
}

// Slaney-scale, slaney-normalised triangular filterbank (librosa.filters.mel defaults, what FilterbankFeatures uses),
// float32 [80][257]; only used when the state_dict does not carry `preprocessor.featurizer.fb`.
static void slaney_fb(std::vector<float>& fb, int sr, int n_fft, int n_mels) {
  const double f_sp = 200.0 / 3.0, brk = 1000.0, brk_mel = brk / f_sp, step = std::log(6.4) / 27.0;
  auto to_mel = [&](double f) { return f >= brk ? brk_mel + std::log(std::max(f, 1e-12) / brk) / step : f / f_sp; };
  auto to_hz = [&](double m) { return m >= brk_mel ? brk * std::exp(step * (m - brk_mel)) : f_sp * m; };
  const int bins = n_fft / 2 + 1;
  std::vector<double> edges(n_mels + 2);
  const double m0 = to_mel(0.0), m1 = to_mel(sr / 2.0);
  for (int i = 0; i < n_mels + 2; ++i) edges[i] = to_hz(m0 + (m1 - m0) * i / (n_mels + 1));
  fb.assign(static_cast<size_t>(n_mels) * bins, 0.f);
  for (int m = 0; m < n_mels; ++m)
    for (int k = 0; k < bins; ++k) {
      const double f = (sr / 2.0) * k / (bins - 1);
      const double up = (f - edges[m]) / (edges[m + 1] - edges[m]), down = (edges[m + 2] - f) / (edges[m + 2] - edges[m + 1]);
      const double v = std::max(0.0, std::min(up, down)) * (2.0 / (edges[m + 2] - edges[m]));
      fb[static_cast<size_t>(m) * bins + k] = static_cast<float>(v);
    }
}

static int pack(const Dict& d, b200d_titanet_desc* desc, Writer& w) {
  memset(desc, 0, sizeof(*desc));
  // ---- discover the encoder blocks
  int n_blocks = 0;
  for (const auto& kv : d) {
    int i = -1;
    if (sscanf(kv.first.c_str(), "encoder.encoder.%d.", &i) == 1 && i + 1 > n_blocks) n_blocks = i + 1;
  }
  if (n_blocks < 1 || n_blocks > B200D_TITANET_MAX_BLOCKS) return set_error(B200D_EINVAL, "%s: no (or too many) encoder.encoder.<i> blocks in the state_dict%s", "b200d_titanet_pack_weights");
  desc->n_blocks = n_blocks;
  int enc_out = 0;
  for (int i = 0; i < n_blocks; ++i) {
    const std::string pre = "encoder.encoder." + std::to_string(i) + ".mconv.";
    std::vector<int> convs, bns, ses;
    for (int j = 0; j < 64; ++j) {
      const std::string q = pre + std::to_string(j);
      if (find(d, q + ".conv.weight")) convs.push_back(j);
      if (find(d, q + ".running_mean")) bns.push_back(j);
      if (find(d, q + ".fc.0.weight")) ses.push_back(j);
    }
    if (convs.size() != 2 * bns.size() || ses.size() != 1 || bns.empty() || bns.size() > 3)
      return set_error(B200D_EINVAL, "%s: unexpected module layout in %s", "b200d_titanet_pack_weights", pre.c_str());
    auto& blk = desc->block[i];
    blk.repeat = static_cast<int32_t>(bns.size());
    int block_in = 0;
    for (size_t r = 0; r < bns.size(); ++r) {
      const Tensor* dw = find(d, pre + std::to_string(convs[2 * r]) + ".conv.weight");      // [Cin][1][k]
      const Tensor* pw = find(d, pre + std::to_string(convs[2 * r + 1]) + ".conv.weight");  // [Cout][Cin][1]
      const std::string bn = pre + std::to_string(bns[r]);
      if (!has_bn(d, bn)) return set_error(B200D_EINVAL, "%s: incomplete BatchNorm %s", "b200d_titanet_pack_weights", bn.c_str());
      const int64_t cout = find(d, bn + ".weight")->n;
      const int64_t cin = pw->n / cout;
      const int64_t k = dw->n / cin;
      if (cin * cout != pw->n || cin * k != dw->n || (k != 1 && k != 3 && k != 7 && k != 11 && k != 15) || cout % 128 != 0)
        return set_error(B200D_EINVAL, "%s: unsupported conv shapes in %s", "b200d_titanet_pack_weights", pre.c_str());
      const int64_t cin_pad = (cin % 128 == 0) ? cin : (cin + 127) / 128 * 128;
      if (r == 0) { block_in = static_cast<int>(cin); blk.cin = static_cast<int32_t>(cin); blk.cin_pad = static_cast<int32_t>(cin_pad); blk.ksize = static_cast<int32_t>(k); }
      else if (k != blk.ksize) return set_error(B200D_EINVAL, "%s: kernel size changes inside %s", "b200d_titanet_pack_weights", pre.c_str());
      blk.cout = static_cast<int32_t>(cout);
      std::vector<double> s, t;
      bn_fold(d, bn, 1e-3, s, t);
      blk.dw[r] = (k == 1) ? -1 : w.take(k * cin_pad * 4);
      blk.w[r] = w.take(cout * cin_pad * 2);
      blk.bias[r] = w.take(cout * 4);
      if (w.base) {
        if (k != 1) {
          float* o = w.at<float>(blk.dw[r]);  // tap-major [k][cin_pad]
          memset(o, 0, k * cin_pad * 4);
          for (int64_t c = 0; c < cin; ++c)
            for (int64_t j = 0; j < k; ++j) o[j * cin_pad + c] = dw->p[c * k + j];
        }
        uint16_t* ow = w.at<uint16_t>(blk.w[r]);
        memset(ow, 0, cout * cin_pad * 2);
        for (int64_t n = 0; n < cout; ++n)
          for (int64_t c = 0; c < cin; ++c) {
            double v = static_cast<double>(pw->p[n * cin + c]) * s[n];
            if (k == 1) v *= static_cast<double>(dw->p[c]);
            ow[n * cin_pad + c] = f2h(v);
          }
        float* ob = w.at<float>(blk.bias[r]);
        for (int64_t n = 0; n < cout; ++n) ob[n] = static_cast<float>(t[n]);
      }
    }
    const std::string se = pre + std::to_string(ses[0]) + ".fc.";
    const Tensor *w1 = find(d, se + "0.weight"), *w2 = find(d, se + "2.weight");
    if (!w2) return set_error(B200D_EINVAL, "%s: squeeze-excite without fc.2 in %s", "b200d_titanet_pack_weights", pre.c_str());
    const int64_t C = blk.cout, hid = w1->n / C;
    if (hid * C != w1->n || w2->n != w1->n || hid % 128 != 0) return set_error(B200D_EINVAL, "%s: unsupported squeeze-excite shape in %s", "b200d_titanet_pack_weights", pre.c_str());
    blk.se_hidden = static_cast<int32_t>(hid);
    blk.se_w1 = w.take(w1->n * 2);
    blk.se_w2 = w.take(w2->n * 2);
    if (w.base) {
      uint16_t *o1 = w.at<uint16_t>(blk.se_w1), *o2 = w.at<uint16_t>(blk.se_w2);
      for (int64_t q = 0; q < w1->n; ++q) { o1[q] = f2h(w1->p[q]); o2[q] = f2h(w2->p[q]); }
    }
    const std::string rp = "encoder.encoder." + std::to_string(i) + ".res.0.";
    const Tensor* rw = find(d, rp + "0.conv.weight");
    blk.residual = rw ? 1 : 0;
    blk.res_w = blk.res_bias = -1;
    if (rw) {
      if (!has_bn(d, rp + "1") || rw->n != static_cast<int64_t>(block_in) * C || block_in % 64 != 0)
        return set_error(B200D_EINVAL, "%s: unsupported residual branch in %s", "b200d_titanet_pack_weights", rp.c_str());
      std::vector<double> s, t;
      bn_fold(d, rp + "1", 1e-3, s, t);
      blk.res_w = w.take(rw->n * 2);
      blk.res_bias = w.take(C * 4);
      if (w.base) {
        uint16_t* o = w.at<uint16_t>(blk.res_w);
        float* ob = w.at<float>(blk.res_bias);
        for (int64_t n = 0; n < C; ++n) {
          for (int64_t c = 0; c < block_in; ++c) o[n * block_in + c] = f2h(static_cast<double>(rw->p[n * block_in + c]) * s[n]);
          ob[n] = static_cast<float>(t[n]);
        }
      }
    }
    enc_out = static_cast<int>(C);
  }
  // the orchestration below is TitaNet's: block 0 and the last block without residual, 1x1 last block
  if (desc->block[0].residual || desc->block[n_blocks - 1].residual || desc->block[n_blocks - 1].ksize != 1 || desc->block[0].repeat != 1 ||
      desc->block[n_blocks - 1].repeat != 1 || desc->block[0].ksize == 1)
    return set_error(B200D_EINVAL, "%s: encoder is not TitaNet-shaped (prologue block, residual blocks, 1x1 epilogue block)%s", "b200d_titanet_pack_weights");
  for (int i = 1; i < n_blocks; ++i) {
    const bool middle = i + 1 < n_blocks;
    if (desc->block[i].cin != desc->block[0].cout || (middle && (desc->block[i].cout != desc->block[0].cout || !desc->block[i].residual || desc->block[i].ksize == 1)))
      return set_error(B200D_EINVAL, "%s: middle blocks must be residual, k > 1, and as wide as block 0%s", "b200d_titanet_pack_weights");
  }
  desc->feat_in = desc->block[0].cin;
  desc->feat_pad = desc->block[0].cin_pad;
  desc->enc_out = enc_out;
  // ---- decoder
  const std::string ap = "decoder._pooling.attention_layer.";
  const Tensor *tw = find(d, ap + "0.conv_layer.weight"), *tb = find(d, ap + "0.conv_layer.bias");
  const Tensor *a2 = find(d, ap + "2.weight"), *a2b = find(d, ap + "2.bias");
  const Tensor *ew = find(d, "decoder.emb_layers.0.1.weight"), *eb = find(d, "decoder.emb_layers.0.1.bias");
  if (!tw || !tb || !a2 || !a2b || !ew || !eb || !has_bn(d, ap + "0.bn") || !has_bn(d, "decoder.emb_layers.0.0"))
    return set_error(B200D_EINVAL, "%s: decoder parameters missing (attention_layer / emb_layers)%s", "b200d_titanet_pack_weights");
  const int64_t C = enc_out, attn = tb->n, emb = eb->n;
  if (tw->n != attn * 3 * C || a2->n != C * attn || a2b->n != C || ew->n != emb * 2 * C || attn % 128 != 0 || emb > 256)
    return set_error(B200D_EINVAL, "%s: unsupported decoder shapes%s", "b200d_titanet_pack_weights");
  desc->attn = static_cast<int32_t>(attn);
  desc->emb = static_cast<int32_t>(emb);
  desc->emb_pad = 256;
  desc->tdnn_wx = w.take(attn * C * 2);
  desc->tdnn_wctx = w.take(attn * 2 * C * 2);
  desc->tdnn_b = w.take(attn * 4);
  desc->tdnn_scale = w.take(attn * 4);
  desc->tdnn_shift = w.take(attn * 4);
  desc->attn_w2 = w.take(C * attn * 2);
  desc->attn_b2 = w.take(C * 4);
  desc->emb_w = w.take(256 * 2 * C * 2);
  desc->emb_b = w.take(256 * 4);
  desc->zeros = w.take(4096 * 4);
  if (w.base) {
    uint16_t *ox = w.at<uint16_t>(desc->tdnn_wx), *oc = w.at<uint16_t>(desc->tdnn_wctx);
    for (int64_t n = 0; n < attn; ++n) {
      for (int64_t c = 0; c < C; ++c) ox[n * C + c] = f2h(tw->p[n * 3 * C + c]);
      for (int64_t c = 0; c < 2 * C; ++c) oc[n * 2 * C + c] = f2h(tw->p[n * 3 * C + C + c]);
    }
    std::vector<double> s, t;
    bn_fold(d, ap + "0.bn", 1e-5, s, t);
    float *ob = w.at<float>(desc->tdnn_b), *os = w.at<float>(desc->tdnn_scale), *ot = w.at<float>(desc->tdnn_shift);
    for (int64_t n = 0; n < attn; ++n) { ob[n] = tb->p[n]; os[n] = static_cast<float>(s[n]); ot[n] = static_cast<float>(t[n]); }
    uint16_t* o2 = w.at<uint16_t>(desc->attn_w2);
    for (int64_t q = 0; q < a2->n; ++q) o2[q] = f2h(a2->p[q]);
    memcpy(w.at<float>(desc->attn_b2), a2b->p, C * 4);
    bn_fold(d, "decoder.emb_layers.0.0", 1e-5, s, t);
    uint16_t* oe = w.at<uint16_t>(desc->emb_w);
    float* oeb = w.at<float>(desc->emb_b);
    memset(oe, 0, 256 * 2 * C * 2);
    memset(oeb, 0, 256 * 4);
    for (int64_t n = 0; n < emb; ++n) {
      double acc = static_cast<double>(eb->p[n]);
      for (int64_t c = 0; c < 2 * C; ++c) {
        const double wv = static_cast<double>(ew->p[n * 2 * C + c]);
        acc += wv * t[c];
        oe[n * 2 * C + c] = f2h(wv * s[c]);
      }
      oeb[n] = static_cast<float>(acc);
    }
    memset(w.at<float>(desc->zeros), 0, 4096 * 4);
  }
  // ---- featurizer tables
  const int bins = 257, mels = desc->feat_in;
  std::vector<float> fb_dense;
  const Tensor* fbt = find(d, "preprocessor.featurizer.fb");
  if (fbt && fbt->n == static_cast<int64_t>(mels) * bins) fb_dense.assign(fbt->p, fbt->p + fbt->n);
  else slaney_fb(fb_dense, 16000, 512, mels);
  std::vector<int32_t> fs(mels), fo(mels + 1, 0);
  std::vector<float> fw;
  for (int m = 0; m < mels; ++m) {
    int a = -1, b = -1;
    for (int k = 0; k < bins; ++k)
      if (fb_dense[static_cast<size_t>(m) * bins + k] != 0.f) { if (a < 0) a = k; b = k + 1; }
    fs[m] = a < 0 ? 0 : a;
    if (a >= 0) fw.insert(fw.end(), fb_dense.begin() + static_cast<size_t>(m) * bins + a, fb_dense.begin() + static_cast<size_t>(m) * bins + b);
    fo[m + 1] = static_cast<int32_t>(fw.size());
  }
  if (mels != 80 || fw.empty() || fw.size() > 1024) return set_error(B200D_EINVAL, "%s: the featurizer kernel is built for 80 mel filters with <= 1024 non-zero weights%s", "b200d_titanet_pack_weights");
  desc->fb_nnz = static_cast<int32_t>(fw.size());
  desc->fb_start = w.take(mels * 4);
  desc->fb_off = w.take((mels + 1) * 4);
  desc->fb_w = w.take(1024 * 4);
  desc->window = w.take(400 * 4);
  if (w.base) {
    memcpy(w.at<int32_t>(desc->fb_start), fs.data(), mels * 4);
    memcpy(w.at<int32_t>(desc->fb_off), fo.data(), (mels + 1) * 4);
    memset(w.at<float>(desc->fb_w), 0, 1024 * 4);
    memcpy(w.at<float>(desc->fb_w), fw.data(), fw.size() * 4);
    float* win = w.at<float>(desc->window);
    const Tensor* wt = find(d, "preprocessor.featurizer.window");
    if (wt && wt->n == 400) memcpy(win, wt->p, 400 * 4);
    else
      for (int i = 0; i < 400; ++i) win[i] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * M_PI * i / 399.0));  // hann(400, periodic=False)
  }
  desc->packed_bytes = w.pos;
  return B200D_OK;
}

// ------------------------------------------------------------------------------------ forward
struct Workspace {
  uint8_t* base;
  size_t pos = 0;
  template <typename T>
  T* take(size_t count) {
    T* p = reinterpret_cast<T*>(base + pos);
    pos = (pos + count * sizeof(T) + 255) & ~static_cast<size_t>(255);
    return p;
  }
};

struct Buffers {
  __half *x0, *a, *b, *y, *d, *x, *e, *hid, *mean16, *sehid, *stats16, *pool16;
  float *gate, *segbias, *emb, *colsum, *feat_scratch;
};

static size_t carve(const b200d_titanet_desc& ds, uint8_t* base, size_t frames, size_t segs, Buffers* out) {
  Workspace w{base};
  const size_t C = ds.block[0].cout, E = ds.enc_out;
  size_t H = 128;
  for (int i = 0; i < ds.n_blocks; ++i) H = H > static_cast<size_t>(ds.block[i].se_hidden) ? H : ds.block[i].se_hidden;
  Buffers b;
  b.x0 = w.take<__half>(frames * ds.feat_pad);
  b.a = w.take<__half>(frames * C);
  b.b = w.take<__half>(frames * C);
  b.y = w.take<__half>(frames * C);
  b.d = w.take<__half>(frames * C);
  b.x = w.take<__half>(frames * E);
  b.e = w.take<__half>(frames * E);
  b.hid = w.take<__half>(frames * ds.attn);
  b.mean16 = w.take<__half>(segs * E);
  b.sehid = w.take<__half>(segs * H);
  b.gate = w.take<float>(segs * E);
  b.stats16 = w.take<__half>(segs * 2 * E);
  b.segbias = w.take<float>(segs * ds.attn);
  b.pool16 = w.take<__half>(segs * 2 * E);
  b.emb = w.take<float>(segs * ds.emb_pad);
  b.feat_scratch = w.take<float>(frames * 80);  // un-normalised log-mel of the frames the featurizer computes per window
  b.colsum = w.take<float>((frames / 32 + 1) * 2 * E);  // per-32-row column sums left by the GEMM feeding each squeeze-excite
  if (out) *out = b;
  return w.pos;
}

static int gemm(const char* what, const void* A, int lda, const void* W, int ldw, int M, int N, int K, void* out, int ldo, b200d_gemm_epilogue epi,
                int flags, cudaStream_t st) {
  static const char* kMode[] = {"bias", "bias_relu", "se_res", "tdnn", "bias_f32", "cheb", "sigmoid_f32"};
  epi.flags = flags;
  char name[48];
  snprintf(name, sizeof(name), "gemm[%s%s]", kMode[epi.mode], gemm_uses_pair_kernel(M, N, epi.mode, flags) ? "|2cta" : "");
  (void)what;
  ProfScope ps(name, 2.0 * M * N * K, st);
  return b200d_gemm_f16(A, lda, W, ldw, M, N, K, out, ldo, &epi, st);
}

// SqueezeExcite gate of block `blk_i` for the activation y; `from_colsum`: the GEMM that produced y left per-32-row column
// sums in b.colsum (b200d_gemm_epilogue.colsum), so the time mean needs no second pass over y.
static int se_gate(const b200d_titanet_desc& ds, const uint8_t* pk, const Buffers& b, int blk_i, const __half* y, int n_seg, int T, int flags,
                   bool from_colsum, cudaStream_t st) {
  const auto& blk = ds.block[blk_i];
  const int C = blk.cout, H = blk.se_hidden;
  if (from_colsum) { ProfScope ps("se_mean_from_colsum", 0, st); RC(b200d_se_mean_from_colsum(b.colsum, n_seg, T, C, b.mean16, st)); }
  else { ProfScope ps("time_stats", 0, st); RC(b200d_time_stats(y, n_seg, T, C, 0, b.mean16, st)); }
  b200d_gemm_epilogue e1{};
  e1.mode = B200D_EPI_BIAS_RELU;
  e1.bias = reinterpret_cast<const float*>(pk + ds.zeros);
  RC(gemm("se_fc1", b.mean16, C, pk + blk.se_w1, C, n_seg, H, C, b.sehid, H, e1, flags, st));
  b200d_gemm_epilogue e2{};
  e2.mode = B200D_EPI_SIGMOID_F32;
  RC(gemm("se_fc2", b.sehid, H, pk + blk.se_w2, H, n_seg, C, H, b.gate, C, e2, flags, st));
  return B200D_OK;
}

static int forward_group(const b200d_titanet_desc& ds, const uint8_t* pk, const Buffers& b, int n_seg, int T, int flags, cudaStream_t st) {
  const int M = n_seg * T;
  const int nb = ds.n_blocks;
  auto dwc = [&](const __half* x, __half* y, int64_t w_off, int C, int k) -> int {
    ProfScope ps("depthwise", 0, st);
    return b200d_depthwise_conv(x, y, reinterpret_cast<const float*>(pk + w_off), n_seg, T, C, k, st);
  };
  auto bias_epi = [&](int mode, int64_t bias_off) {
    b200d_gemm_epilogue e{};
    e.mode = mode;
    e.bias = reinterpret_cast<const float*>(pk + bias_off);
    return e;
  };
  // the GEMM in front of a squeeze-excite also leaves the column sums its time mean needs (pair kernel, windows >= 32 frames)
  auto se_input_epi = [&](int64_t bias_off, int N, bool* from_colsum) {
    b200d_gemm_epilogue e = bias_epi(B200D_EPI_BIAS, bias_off);
    *from_colsum = T >= 32 && gemm_uses_pair_kernel(M, N, B200D_EPI_BIAS, flags);
    if (*from_colsum) {
      e.colsum = b.colsum;
      e.rows_per_seg = T;
    }
    return e;
  };
  bool cs = false;
  // ---- block 0: dw -> pw feat -> C -> BN -> SE -> ReLU
  const auto& b0 = ds.block[0];
  const int C = b0.cout;
  RC(dwc(b.x0, b.d, b0.dw[0], ds.feat_pad, b0.ksize));
  RC(gemm("b0", b.d, ds.feat_pad, pk + b0.w[0], ds.feat_pad, M, C, ds.feat_pad, b.y, C, se_input_epi(b0.bias[0], C, &cs), flags, st));
  RC(se_gate(ds, pk, b, 0, b.y, n_seg, T, flags, cs, st));
  { ProfScope ps("se_apply_relu", 0, st); RC(b200d_se_apply_relu(b.y, b.gate, b.a, n_seg, T, C, st)); }
  __half *cur = b.a, *nxt = b.b;
  // ---- residual blocks: repeat x (dw -> pw -> BN [-> ReLU]) -> SE ; + BN(conv1x1(in)) ; ReLU
  for (int bi = 1; bi + 1 < nb; ++bi) {
    const auto& blk = ds.block[bi];
    const __half* src = cur;
    for (int r = 0; r < blk.repeat; ++r) {
      RC(dwc(src, b.d, blk.dw[r], C, blk.ksize));
      const bool last = r == blk.repeat - 1;
      RC(gemm("pw", b.d, C, pk + blk.w[r], C, M, C, C, b.y, C, last ? se_input_epi(blk.bias[r], C, &cs) : bias_epi(B200D_EPI_BIAS_RELU, blk.bias[r]), flags, st));
      src = b.y;  // the next depthwise reads y and writes d; its GEMM then overwrites y
    }
    RC(se_gate(ds, pk, b, bi, b.y, n_seg, T, flags, cs, st));
    b200d_gemm_epilogue e = bias_epi(B200D_EPI_SE_RES, blk.res_bias);
    e.rowvec = b.gate;
    e.aux16 = b.y;
    e.rows_per_seg = T;
    RC(gemm("res", cur, C, pk + blk.res_w, C, M, C, C, nxt, C, e, flags, st));
    __half* t = cur; cur = nxt; nxt = t;
  }
  // ---- last block: (dw k = 1 folded) pw C -> E -> BN -> SE -> ReLU, with the [mean | std] of the result in the same pass
  const auto& bl = ds.block[nb - 1];
  const int E = ds.enc_out;
  RC(gemm("b_last", cur, C, pk + bl.w[0], C, M, E, C, b.e, E, se_input_epi(bl.bias[0], E, &cs), flags, st));
  RC(se_gate(ds, pk, b, nb - 1, b.e, n_seg, T, flags, cs, st));
  { ProfScope ps("se_apply_relu_stats", 0, st); RC(b200d_se_apply_relu_stats(b.e, b.gate, b.x, n_seg, T, E, b.stats16, st)); }
  // ---- decoder: attentive statistics pooling + embedding projection
  RC(gemm("tdnn_ctx", b.stats16, 2 * E, pk + ds.tdnn_wctx, 2 * E, n_seg, ds.attn, 2 * E, b.segbias, ds.attn, bias_epi(B200D_EPI_BIAS_F32, ds.tdnn_b), flags, st));
  b200d_gemm_epilogue et{};
  et.mode = B200D_EPI_TDNN;
  et.scale = reinterpret_cast<const float*>(pk + ds.tdnn_scale);
  et.shift = reinterpret_cast<const float*>(pk + ds.tdnn_shift);
  et.rowvec = b.segbias;
  et.rows_per_seg = T;
  RC(gemm("tdnn_x", b.x, E, pk + ds.tdnn_wx, E, M, ds.attn, E, b.hid, ds.attn, et, flags, st));
  RC(gemm("attn", b.hid, ds.attn, pk + ds.attn_w2, ds.attn, M, E, ds.attn, b.e, E, bias_epi(B200D_EPI_BIAS, ds.attn_b2), flags, st));
  { ProfScope ps("attn_pool", 0, st); RC(b200d_attn_pool(b.x, b.e, n_seg, T, E, b.pool16, st)); }
  RC(gemm("emb", b.pool16, 2 * E, pk + ds.emb_w, 2 * E, n_seg, ds.emb_pad, 2 * E, b.emb, ds.emb_pad, bias_epi(B200D_EPI_BIAS_F32, ds.emb_b), flags, st));
  return B200D_OK;
}

}  // namespace b200d

using namespace b200d;

extern "C" int b200d_titanet_pack_weights(int32_t n_tensors, const char* const* names, const float* const* data_host, const int64_t* numel,
                                          b200d_titanet_desc* desc, void* packed_host, size_t packed_bytes) {
  B200D_CHECK_ARG(n_tensors > 0 && names && data_host && numel && desc);
  Dict d;
  for (int i = 0; i < n_tensors; ++i) {
    B200D_CHECK_ARG(names[i] && data_host[i] && numel[i] > 0);
    d[names[i]] = Tensor{data_host[i], numel[i]};
  }
  Writer layout{nullptr};
  b200d_titanet_desc tmp;
  if (int rc = pack(d, &tmp, layout)) return rc;
  if (packed_host == nullptr) {  // size query: desc->packed_bytes
    *desc = tmp;
    return B200D_OK;
  }
  if (packed_bytes < static_cast<size_t>(tmp.packed_bytes)) return set_error(B200D_EWORKSPACE, "%s: packed_bytes too small%s", "b200d_titanet_pack_weights");
  Writer fill{reinterpret_cast<uint8_t*>(packed_host)};
  return pack(d, desc, fill);
}

extern "C" size_t b200d_titanet_workspace_bytes(const b200d_titanet_desc* desc, int32_t max_frames, int32_t max_segs) {
  if (!desc || max_frames <= 0 || max_segs <= 0) return 0;
  return carve(*desc, nullptr, static_cast<size_t>(max_frames), static_cast<size_t>(max_segs), nullptr);
}

// windows per launch group: the largest count (<= n_seg) whose activations fit ws_bytes
static long long group_windows(const b200d_titanet_desc& desc, size_t ws_bytes, int T, long long n_seg) {
  long long lo = 0, hi = n_seg;
  while (lo < hi) {
    const long long mid = (lo + hi + 1) / 2;
    if (carve(desc, nullptr, static_cast<size_t>(mid) * T, static_cast<size_t>(mid), nullptr) <= ws_bytes) lo = mid; else hi = mid - 1;
  }
  return lo;
}

extern "C" int32_t b200d_titanet_group_windows(const b200d_titanet_desc* desc, size_t ws_bytes, int32_t fixed_len, int32_t variant) {
  if (!desc || fixed_len <= 0) return 0;
  const int T = fixed_len / 160 + ((variant & B200D_FEAT_NO_PLUS_ONE) ? 0 : 1);
  return static_cast<int32_t>(group_windows(*desc, ws_bytes, T, 1 << 24));
}

extern "C" int b200d_titanet_forward(const b200d_titanet_desc* desc, const void* packed_dev, const float* wav, int64_t n_wav, const float* logmel,
                                     const int32_t* seg_start, const int32_t* seg_len, const int32_t* seg_row0, int32_t n_on_stream, int32_t n_seg,
                                     int32_t fixed_len, int32_t variant, int32_t flags, float* emb_out, int32_t ld_emb, void* ws, size_t ws_bytes,
                                     void* stream) {
  B200D_CHECK_ARG(desc && packed_dev && wav && seg_start && seg_len && emb_out && ws);
  B200D_CHECK_ARG(n_seg > 0 && n_on_stream >= 0 && n_on_stream <= n_seg && fixed_len > 0 && ld_emb >= desc->emb && desc->n_blocks >= 3);
  B200D_CHECK_ARG((reinterpret_cast<uintptr_t>(packed_dev) & 255) == 0 && (reinterpret_cast<uintptr_t>(ws) & 255) == 0);
  const int T = fixed_len / 160 + ((variant & B200D_FEAT_NO_PLUS_ONE) ? 0 : 1);
  B200D_CHECK_ARG(T >= 2);
  const long long lo = group_windows(*desc, ws_bytes, T, n_seg);
  if (lo < 1) return set_error(B200D_EWORKSPACE, "%s: workspace too small for one window (b200d_titanet_workspace_bytes)%s", "b200d_titanet_forward");
  const int group = static_cast<int>(lo);
  Buffers b;
  carve(*desc, reinterpret_cast<uint8_t*>(ws), static_cast<size_t>(group) * T, static_cast<size_t>(group), &b);
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed_dev);
  cudaStream_t st = as_stream(stream);
  for (int c0 = 0; c0 < n_seg; c0 += group) {
    const int n = n_seg - c0 < group ? n_seg - c0 : group;
    {
      ProfScope ps("featurize_windows", 0, st, 3);
      const int fast = n_on_stream - c0 < 0 ? 0 : (n_on_stream - c0 > n ? n : n_on_stream - c0);
      RC(b200d_featurize_windows(wav, n_wav, logmel, seg_start + c0, seg_len + c0, seg_row0 ? seg_row0 + c0 : nullptr, fast, n, fixed_len,
                                 reinterpret_cast<const int32_t*>(pk + desc->fb_start), reinterpret_cast<const int32_t*>(pk + desc->fb_off),
                                 reinterpret_cast<const float*>(pk + desc->fb_w), desc->fb_nnz, reinterpret_cast<const float*>(pk + desc->window),
                                 variant, b.feat_scratch, b.x0, desc->feat_pad, nullptr, st));
    }
    RC(forward_group(*desc, pk, b, n, T, flags, st));
    B200D_CHECK_CUDA(cudaMemcpy2DAsync(emb_out + static_cast<size_t>(c0) * ld_emb, static_cast<size_t>(ld_emb) * 4, b.emb,
                                       static_cast<size_t>(desc->emb_pad) * 4, static_cast<size_t>(desc->emb) * 4, n, cudaMemcpyDeviceToDevice, st));
  }
  return B200D_OK;
}

extern "C" int b200d_titanet_mel_stream(const b200d_titanet_desc* desc, const void* packed_dev, const float* wav, int64_t n_wav,
                                        const int64_t* stream_start, const int32_t* stream_off, int32_t n_streams, int32_t total_rows,
                                        float* logmel, void* stream) {
  B200D_CHECK_ARG(desc && packed_dev);
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed_dev);
  ProfScope ps("mel_stream", 0, as_stream(stream));
  return b200d_mel_stream(wav, n_wav, stream_start, stream_off, n_streams, total_rows, reinterpret_cast<const int32_t*>(pk + desc->fb_start),
                          reinterpret_cast<const int32_t*>(pk + desc->fb_off), reinterpret_cast<const float*>(pk + desc->fb_w), desc->fb_nnz,
                          reinterpret_cast<const float*>(pk + desc->window), logmel, stream);
}

extern "C" int64_t b200d_launch_count(void) { return g_composite_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------ profiling C ABI
extern "C" int b200d_profile_start(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& s : g_spans) { cudaEventDestroy(s.e0); cudaEventDestroy(s.e1); }
  g_spans.clear();
  g_prof_on = true;
  return B200D_OK;
}

extern "C" int b200d_profile_stop(b200d_profile_span* spans, int32_t capacity, int32_t* n_spans) {
  B200D_CHECK_ARG(n_spans && (spans || capacity == 0));
  B200D_CHECK_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  int n = 0;
  for (auto& s : g_spans) {
    if (n < capacity) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, s.e0, s.e1);
      memcpy(spans[n].name, s.name, sizeof(spans[n].name));
      spans[n].ms = ms;
      spans[n].work = s.work;
      ++n;
    }
    cudaEventDestroy(s.e0);
    cudaEventDestroy(s.e1);
  }
  *n_spans = static_cast<int32_t>(g_spans.size());
  g_spans.clear();
  return B200D_OK;
}
