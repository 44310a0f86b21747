// C ABI of the B200-native attention-lvcsr hot path (see include/lvsr_b200.h).
//
// Host-side orchestration only: which kernel runs when, on which buffers.  The
// compiled-function seam it replaces is SURVEY.md section 8b tier b3
// (libs/blocks/blocks/search.py:97-142; lvsr/bricks/recognizer.py:375-390,490-494).
#include "lvsr_b200.h"

#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "model.h"

namespace lvsr {

thread_local std::string g_last_error;
long long g_launch_count = 0;

int set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return 1;
}

// ---- per-kernel-class event timing ---------------------------------------------------
struct ProfEntry { std::string cls; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfEntry> g_prof;

ProfScope::ProfScope(const char* kernel_class, cudaStream_t stream) : slot(-1), st(stream) {
  if (!g_prof_on) return;
  ProfEntry e;
  e.cls = kernel_class;
  if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
  cudaEventRecord(e.a, st);
  g_prof.push_back(e);
  slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
}

}  // namespace lvsr

using namespace lvsr;

namespace lvsr {

void add_param(lvsr_model* m, const std::string& name, int64_t d0, int64_t d1 = -1) {
  Param p;
  p.name = name;
  p.shape[0] = d0;
  p.shape[1] = d1 > 0 ? d1 : 1;
  p.ndim = d1 > 0 ? 2 : 1;
  p.count = d0 * (d1 > 0 ? d1 : 1);
  p.dev = nullptr;
  m->index[name] = (int)m->params.size();
  m->params.push_back(p);
}


// Blocks initialisation order (oracle/lvsr_oracle.py: param_shapes)
void build_param_table(lvsr_model* m) {
  const lvsr_config& c = m->cfg;
  int din = c.num_features;
  for (int l = 0; l < c.num_layers; ++l) {
    const int D = c.dims_bidir[l];
    for (int dir = 0; dir < 2; ++dir) {
      const std::string b = enc_base(l, dir);
      add_param(m, b + "/gatedrecurrent.state_to_state", D, D);
      add_param(m, b + "/gatedrecurrent.state_to_gates", D, 2 * D);
      add_param(m, b + "/gatedrecurrent.initial_state", D);
      add_param(m, b + "/fork/fork_inputs.b", D);
      add_param(m, b + "/fork/fork_inputs.W", din, D);
      add_param(m, b + "/fork/fork_gate_inputs.b", 2 * D);
      add_param(m, b + "/fork/fork_gate_inputs.W", din, 2 * D);
    }
    din = 2 * D;
  }
  const int E = m->E, C = c.dim_dec, M = c.dim_matcher, K = c.conv_num_filters, w = 2 * c.conv_n + 1;
  const int V = c.num_phonemes, Cfb = c.dim_feedback, Cpm = c.post_merge_dim;
  const std::string g = GEN, t = TR, a = ATT;
  if (!c.one_of_n_feedback) add_param(m, g + "/readout/lookupfeedback/lookuptable.W", V + 1, Cfb);
  if (c.use_states_for_readout) add_param(m, g + "/readout/merge/transform_states.W", C, Cpm);
  add_param(m, g + "/readout/merge/transform_weighted_averages.W", E, Cpm);
  add_param(m, g + "/readout/post_merge/bias.b", Cpm);
  add_param(m, g + "/readout/post_merge/mlp/linear_0.b", V);
  add_param(m, g + "/readout/post_merge/mlp/linear_0.W", Cpm / c.maxout_pieces, V);
  add_param(m, g + "/fork/fork_inputs.b", C);
  add_param(m, g + "/fork/fork_inputs.W", Cfb, C);
  add_param(m, g + "/fork/fork_gate_inputs.b", 2 * C);
  add_param(m, g + "/fork/fork_gate_inputs.W", Cfb, 2 * C);
  add_param(m, t + "/transition.state_to_state", C, C);
  add_param(m, t + "/transition.state_to_gates", C, 2 * C);
  add_param(m, t + "/transition.initial_state", C);
  add_param(m, a + "/state_trans/transform_states.W", C, M);
  add_param(m, a + "/preprocess.b", M);
  add_param(m, a + "/preprocess.W", E, M);
  if (c.energy_normalizer != LVSR_NORM_SOFTMAX) add_param(m, a + "/energy_comp/linear.b", 1);
  add_param(m, a + "/energy_comp/linear.W", M, 1);
  add_param(m, a + "/handler.W", K, M);
  add_param(m, a + "/conv1d.filters", K, w);
  add_param(m, t + "/distribute/fork_inputs.W", E, C);
  add_param(m, t + "/distribute/fork_gate_inputs.W", E, 2 * C);
}

int copy_cols(float* dst, int ld_dst, int col0, const float* src, int rows, int cols, cudaStream_t s) {
  LVSR_CUDA_OK(cudaMemcpy2DAsync(dst + col0, (size_t)ld_dst * sizeof(float), src, (size_t)cols * sizeof(float),
                                 (size_t)cols * sizeof(float), rows, cudaMemcpyDeviceToDevice, s));
  return 0;
}



// take_glimpses for R rows: q = s.W_state, window, attention step.
struct Segments {           // batched beam search: hypotheses of one utterance = one segment (the reference's batch)
  const int* seg_start = nullptr; int nseg = 0; const int* seg_len = nullptr; const int* row_seg = nullptr;
};
int glimpses(lvsr_model* m, const float* H, const float* P, const float* maskH, int Tp, int U,
             const int* row_utt, int R, const float* states, const float* w_prev, const long long* step,
             long long step_offset, float* w_out, float* e_out, float* ctx, cudaStream_t st, Segments sg = Segments()) {
  const lvsr_config& c = m->cfg;
  Arena& ws = m->ws;
  float* q = ws.f32((size_t)R * c.dim_matcher);
  int* win = ws.i32((size_t)2 * std::max(1, sg.nseg));
  float* lohi = ws.f32((size_t)2 * R);
  LVSR_CHECK(q && win && lohi, "out of device memory (workspace)");
  DenseArgs d = {};
  d.X1 = states; d.K1 = c.dim_dec; d.W1 = m->P(std::string(ATT) + "/state_trans/transform_states.W");
  d.R = R; d.N = c.dim_matcher; d.mode = DENSE_PLAIN; d.out = q;
  if (int rc = dense_step(d, st)) return rc;
  WindowArgs wa = {};
  wa.weights = w_prev; wa.step = step; wa.step_offset = step_offset; wa.R = R; wa.Tp = Tp;
  wa.prior = prior_of(c); wa.win = win; wa.lohi = lohi;
  wa.seg_start = sg.seg_start; wa.nseg = sg.nseg; wa.seg_len = sg.seg_len;
  if (int rc = attention_window(wa, st)) return rc;
  AttStepArgs a = {};
  a.P = P; a.H = H; a.maskH = maskH; a.row_utt = row_utt; a.q = q; a.w_prev = w_prev; a.win = win; a.lohi = lohi;
  a.row_seg = sg.seg_start ? sg.row_seg : nullptr;
  a.filt = m->P(std::string(ATT) + "/conv1d.filters");
  a.Wh = m->P(std::string(ATT) + "/handler.W");
  a.v = m->P(std::string(ATT) + "/energy_comp/linear.W");
  a.v_bias = m->v_bias;   // energy bias exists only when the normaliser is not softmax
  a.w_out = w_out; a.e_out = e_out; a.ctx = ctx;
  a.R = R; a.U = U; a.Tp = Tp; a.M = c.dim_matcher; a.E = m->E; a.K = c.conv_num_filters; a.n = c.conv_n;
  a.normalizer = c.energy_normalizer;
  return attention_step(a, st);
}

// compute_states for R rows: distribute + fork(feedback) + GRU step.
int transition(lvsr_model* m, int R, const float* states, const float* ctx, const long long* outputs,
               const float* rmask, float* next_states, cudaStream_t st) {
  const lvsr_config& c = m->cfg;
  Arena& ws = m->ws;
  const int C = c.dim_dec;
  float* z = ws.f32((size_t)R * C);
  float* hr = ws.f32((size_t)R * C);
  float* ai = ws.f32((size_t)R * C);
  LVSR_CHECK(z && hr && ai, "out of device memory (workspace)");
  DenseArgs g = {};
  g.X1 = ctx; g.K1 = m->E; g.W1 = m->Wd_cat;
  g.X2 = states; g.K2 = C; g.W2 = m->P(std::string(TR) + "/transition.state_to_gates"); g.N2 = 2 * C;
  g.add = m->FF; g.arow = outputs; g.add_rows = c.num_phonemes + 1; g.R = R; g.N = 3 * C; g.mode = DENSE_GATES;
  g.s = states; g.z = z; g.hr = hr; g.ai = ai; g.C = C;
  if (int rc = dense_step(g, st)) return rc;
  DenseArgs k = {};
  k.X1 = hr; k.K1 = C; k.W1 = m->P(std::string(TR) + "/transition.state_to_state");
  k.add = ai; k.arow = nullptr; k.R = R; k.N = C; k.mode = DENSE_CAND;
  k.s = states; k.z = z; k.rmask = rmask; k.out = next_states; k.C = C;
  return dense_step(k, st);
}

int readout_merged(lvsr_model* m, int R, const float* states, const float* ctx, float* merged, cudaStream_t st) {
  const lvsr_config& c = m->cfg;
  DenseArgs d = {};
  d.X1 = ctx; d.K1 = m->E; d.W1 = m->P(std::string(GEN) + "/readout/merge/transform_weighted_averages.W");
  if (c.use_states_for_readout) {
    d.X2 = states; d.K2 = c.dim_dec; d.W2 = m->P(std::string(GEN) + "/readout/merge/transform_states.W");
    d.N2 = c.post_merge_dim;
  }
  d.R = R; d.N = c.post_merge_dim; d.mode = DENSE_PLAIN; d.out = merged;
  return dense_step(d, st);
}

ReadoutArgs readout_args(lvsr_model* m, int R, const float* merged) {
  const lvsr_config& c = m->cfg;
  ReadoutArgs r = {};
  r.merged = merged;
  r.b_pm = m->P(std::string(GEN) + "/readout/post_merge/bias.b");
  r.Wo = m->P(std::string(GEN) + "/readout/post_merge/mlp/linear_0.W");
  r.bo = m->P(std::string(GEN) + "/readout/post_merge/mlp/linear_0.b");
  r.R = R; r.Cpm = c.post_merge_dim; r.pieces = c.maxout_pieces; r.V = c.num_phonemes; r.act = c.post_merge_activation;
  return r;
}

size_t encoder_ws_bytes(const lvsr_model* m, int T, int B) {
  size_t total = 0;
  int Tl = T;
  for (int l = 0; l < m->cfg.num_layers; ++l) {
    const int D = m->cfg.dims_bidir[l], k = m->cfg.subsample[l];
    const int Tout = ceil_div(Tl, k);
    total += ((size_t)Tl * B * 6 * D + (size_t)Tout * B * 2 * D) * sizeof(float) + 1024;
    total += (size_t)2 * Tl * B * gemm_tc_kpad(l == 0 ? m->cfg.num_features : 2 * m->cfg.dims_bidir[l - 1]) * sizeof(float) + 1024;
    Tl = Tout;
  }
  return total + (1 << 16);
}
size_t cost_ws_bytes(const lvsr_model* m, int Tp, int B, int L) {
  const lvsr_config& c = m->cfg;
  size_t f = (size_t)Tp * B * c.dim_matcher + (size_t)2 * Tp * B * m->E + (size_t)(L + 1) * B * c.dim_dec + (size_t)L * B * m->E +
             (size_t)4 * B * Tp + (size_t)L * B * c.post_merge_dim + (size_t)B * c.dim_matcher +
             (size_t)L * B * (Tp + c.dim_matcher + c.dim_dec + 1) +
             (size_t)3 * B * c.dim_dec + 4 * B + 64;
  return f * sizeof(float) + (1 << 16);
}



}  // namespace lvsr

extern "C" {

const char* lvsr_last_error(void) { return g_last_error.c_str(); }
int lvsr_version(void) { return 100; }
int64_t lvsr_launch_count(int reset) {
  const int64_t v = g_launch_count;
  if (reset) g_launch_count = 0;
  return v;
}

int lvsr_profile_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}
int lvsr_profile_read(const char* kernel_class, double* total_ms, int64_t* count) {
  LVSR_CHECK(kernel_class && total_ms && count, "null argument");
  LVSR_CUDA_OK(cudaDeviceSynchronize());
  double tot = 0.0;
  int64_t n = 0;
  std::vector<ProfEntry> keep;
  for (auto& e : g_prof) {
    if (e.cls == kernel_class) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) { tot += ms; n++; }
      cudaEventDestroy(e.a);
      cudaEventDestroy(e.b);
    } else {
      keep.push_back(e);
    }
  }
  g_prof.swap(keep);
  *total_ms = tot;
  *count = n;
  return 0;
}

int lvsr_model_create(const lvsr_config* cfg, lvsr_model** out) {
  LVSR_CHECK(cfg && out, "null argument");
  LVSR_CHECK(cfg->num_layers >= 1 && cfg->num_layers <= LVSR_MAX_LAYERS, "num_layers %d out of range", cfg->num_layers);
  for (int l = 0; l < cfg->num_layers; ++l) {
    LVSR_CHECK(bigru_supported(cfg->dims_bidir[l]), "encoder dim %d unsupported (128 or 256)", cfg->dims_bidir[l]);
    LVSR_CHECK(cfg->subsample[l] >= 1, "subsample must be >= 1");
  }
  LVSR_CHECK(cfg->dim_dec % 8 == 0 && cfg->post_merge_dim % 8 == 0, "dim_dec and post_merge_dim must be multiples of 8");
  LVSR_CHECK(cfg->dim_matcher == 128 || cfg->dim_matcher == 256 || cfg->dim_matcher == 512,
             "dim_matcher %d unsupported by the attention kernel (128, 256 or 512)", cfg->dim_matcher);
  LVSR_CHECK(cfg->one_of_n_feedback ? cfg->dim_feedback == cfg->num_phonemes + 1 : cfg->dim_feedback % 4 == 0,
             "dim_feedback must be a multiple of 4 (LookupFeedback) or num_phonemes + 1 (OneOfNFeedback)");
  LVSR_CHECK(cfg->maxout_pieces >= 1 && cfg->post_merge_dim % cfg->maxout_pieces == 0, "bad maxout_pieces");
  LVSR_CHECK(cfg->post_merge_activation >= LVSR_ACT_MAXOUT && cfg->post_merge_activation <= LVSR_ACT_IDENTITY,
             "bad post_merge_activation");
  LVSR_CHECK(cfg->post_merge_activation == LVSR_ACT_MAXOUT || cfg->maxout_pieces == 1,
             "maxout_pieces must be 1 unless the activation is Maxout");
  LVSR_CHECK(cfg->conv_num_filters >= 1 && cfg->conv_num_filters <= 16, "conv_num_filters %d not in [1,16]", cfg->conv_num_filters);
  // the centre crop [:, :, n:-n] of the reference is EMPTY for n = 0 (lvsr/bricks/attention.py:109-110)
  LVSR_CHECK(cfg->conv_n >= 1, "conv_n must be >= 1 (got %d)", cfg->conv_n);
  LVSR_CHECK(cfg->num_phonemes >= 1 && cfg->num_phonemes <= 128, "num_phonemes out of range");
  int dev_count = 0;
  LVSR_CUDA_OK(cudaGetDeviceCount(&dev_count));
  LVSR_CHECK(dev_count > 0, "no CUDA device: the B200 path has no CPU fallback");
  lvsr_model* m = new lvsr_model();
  m->cfg = *cfg;
  LVSR_CUDA_OK(cudaGetDevice(&m->device));
  m->E = 2 * cfg->dims_bidir[cfg->num_layers - 1];
  build_param_table(m);
  int64_t total = 0;
  for (auto& p : m->params) {
    p.offset = total;
    total += (p.count + 63) & ~(int64_t)63;
  }
  m->flat_count = total;
  {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&m->flat), (size_t)total * sizeof(float));
    if (e != cudaSuccess) {
      lvsr_model_destroy(m);
      return set_error("cudaMalloc(parameters, %lld floats) failed: %s", (long long)total, cudaGetErrorString(e));
    }
    cudaMemset(m->flat, 0, (size_t)total * sizeof(float));
  }
  for (auto& p : m->params) p.dev = m->flat + p.offset;
  if (cudaMalloc(reinterpret_cast<void**>(&m->status), 64) != cudaSuccess) {
    lvsr_model_destroy(m);
    return set_error("cudaMalloc(status) failed");
  }
  cudaMemset(m->status, 0, 64);
  *out = m;
  return 0;
}

int lvsr_model_destroy(lvsr_model* m) {
  if (!m) return 0;
  DeviceGuard device_guard(m);
  cudaDeviceSynchronize();
  if (m->flat) cudaFree(m->flat);
  for (float* p : m->Wcat) if (p) cudaFree(p);
  for (float* p : m->bcat) if (p) cudaFree(p);
  for (float* p : m->Wcat_hi) if (p) cudaFree(p);
  for (float* p : m->Wcat_lo) if (p) cudaFree(p);
  for (void* p : m->Wcat_h16_head) if (p) cudaFree(p);
  for (void* p : m->Wcat_h16_tail) if (p) cudaFree(p);
  for (float* p : m->Wcat_h16_scale) if (p) cudaFree(p);
  if (m->Wp_h16_head) cudaFree(m->Wp_h16_head);
  if (m->Wp_h16_tail) cudaFree(m->Wp_h16_tail);
  if (m->Wp_h16_scale) cudaFree(m->Wp_h16_scale);
  if (m->Wp_hi) cudaFree(m->Wp_hi);
  if (m->Wp_lo) cudaFree(m->Wp_lo);
  if (m->Wd_cat) cudaFree(m->Wd_cat);
  if (m->Wb1) cudaFree(m->Wb1);
  if (m->Wff_cat) cudaFree(m->Wff_cat);
  if (m->bff_cat) cudaFree(m->bff_cat);
  if (m->FF) cudaFree(m->FF);
  if (m->status) cudaFree(m->status);
  if (m->opt_velocity) cudaFree(m->opt_velocity);
  if (m->opt_ms_step) cudaFree(m->opt_ms_step);
  if (m->opt_ms_dx) cudaFree(m->opt_ms_dx);
  if (m->opt_scratch) cudaFree(m->opt_scratch);
  if (m->opt_desc) cudaFree(m->opt_desc);
  m->tws.destroy();
  m->ws.destroy();
  delete m;
  return 0;
}

int lvsr_model_status(lvsr_model* m, int32_t* launch_status, int64_t* stepwise_fallbacks) {
  DeviceGuard device_guard(m);
  LVSR_CHECK(m && launch_status, "null argument");
  unsigned hst = 0;
  LVSR_CUDA_OK(cudaMemcpy(&hst, m->status, sizeof(hst), cudaMemcpyDeviceToHost));   // synchronises with the device
  *launch_status = (int32_t)hst;
  if (stepwise_fallbacks) *stepwise_fallbacks = m->dec_fallbacks;
  return 0;
}

int lvsr_model_num_params(const lvsr_model* m) { return m ? (int)m->params.size() : 0; }
const char* lvsr_model_param_name(const lvsr_model* m, int i) {
  if (!m || i < 0 || i >= (int)m->params.size()) return nullptr;
  return m->params[i].name.c_str();
}
int lvsr_model_param_shape(const lvsr_model* m, int i, int64_t shape[2], int32_t* ndim) {
  LVSR_CHECK(m && i >= 0 && i < (int)m->params.size(), "bad parameter index %d", i);
  shape[0] = m->params[i].shape[0];
  shape[1] = m->params[i].shape[1];
  *ndim = m->params[i].ndim;
  return 0;
}
int64_t lvsr_model_flat_size(const lvsr_model* m) { return m ? m->flat_count : 0; }
int lvsr_model_param_offset(const lvsr_model* m, int i, int64_t* offset, int64_t* count) {
  LVSR_CHECK(m && offset && i >= 0 && i < (int)m->params.size(), "bad parameter index %d", i);
  *offset = m->params[i].offset;
  if (count) *count = m->params[i].count;
  return 0;
}
float* lvsr_model_flat_params(lvsr_model* m) { return m ? m->flat : nullptr; }
int lvsr_model_set_param(lvsr_model* m, const char* name, const float* host, int64_t count) {
  DeviceGuard device_guard(m);
  LVSR_CHECK(m && name && host, "null argument");
  auto it = m->index.find(name);
  LVSR_CHECK(it != m->index.end(), "unknown parameter '%s'", name);
  Param& p = m->params[it->second];
  LVSR_CHECK(count == p.count, "parameter '%s' expects %lld values, got %lld", name, (long long)p.count, (long long)count);
  LVSR_CUDA_OK(cudaMemcpy(p.dev, host, (size_t)count * sizeof(float), cudaMemcpyHostToDevice));
  m->finalized = false;
  return 0;
}
int lvsr_model_get_param(const lvsr_model* m, const char* name, float* host, int64_t count) {
  DeviceGuard device_guard(m);
  LVSR_CHECK(m && name && host, "null argument");
  auto it = m->index.find(name);
  LVSR_CHECK(it != m->index.end(), "unknown parameter '%s'", name);
  const Param& p = m->params[it->second];
  LVSR_CHECK(count == p.count, "parameter '%s' holds %lld values, asked for %lld", name, (long long)p.count, (long long)count);
  LVSR_CUDA_OK(cudaMemcpy(host, p.dev, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int lvsr_model_finalize(lvsr_model* m) {
  DeviceGuard device_guard(m);
  LVSR_CHECK(m, "null model");
  return finalize_on_stream(m, 0, true);
}

}  // extern "C"

namespace lvsr {
// fp16 head/tail operands for the projections that read BiGRU outputs (|h| <= 1): layers >= 1 and preprocess.
// Allocated and split at their first use after a parameter change, i.e. AFTER the caller has reserved the workspace arena:
// the arena keeps the device pages it gets without these buffers (the persistent decoder's step time moves by up to 10 %
// with the physical placement of its buffers, profiles/r2k_summary.md).
int ensure_h16(lvsr_model* m, cudaStream_t st) {
  if (!m->use_tc || !m->use_h16 || !m->h16_stale) return 0;
  const lvsr_config& c = m->cfg;
  if (m->Wcat_h16_head.empty()) {
    int dk2 = c.num_features;
    for (int l = 0; l < c.num_layers; ++l) {
      const int D = c.dims_bidir[l];
      void *h = nullptr, *t = nullptr;
      float* sc = nullptr;
      if (l >= 1 && gemm_tc_h16_supported(128, 6 * D, dk2)) {
        const size_t bytes = (size_t)gemm_tc_kpad_h16(dk2) * 6 * D * 2;
        LVSR_CUDA_OK(cudaMalloc(&h, bytes));
        LVSR_CUDA_OK(cudaMalloc(&t, bytes));
        LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&sc), 2 * sizeof(float)));
      }
      m->Wcat_h16_head.push_back(h);
      m->Wcat_h16_tail.push_back(t);
      m->Wcat_h16_scale.push_back(sc);
      dk2 = 2 * D;
    }
    if (gemm_tc_h16_supported(128, c.dim_matcher, m->E)) {
      const size_t bytes = (size_t)gemm_tc_kpad_h16(m->E) * c.dim_matcher * 2;
      LVSR_CUDA_OK(cudaMalloc(&m->Wp_h16_head, bytes));
      LVSR_CUDA_OK(cudaMalloc(&m->Wp_h16_tail, bytes));
      LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&m->Wp_h16_scale), 2 * sizeof(float)));
    }
  }
  int dk2 = c.num_features;
  for (int l = 0; l < c.num_layers; ++l) {
    const int D = c.dims_bidir[l];
    if (m->Wcat_h16_head[l])
      if (int rc = split_weight_h16(m->Wcat[l], dk2, 6 * D, m->Wcat_h16_head[l], m->Wcat_h16_tail[l], m->Wcat_h16_scale[l], st))
        return rc;
    dk2 = 2 * D;
  }
  if (m->Wp_h16_head)
    if (int rc = split_weight_h16(m->P(std::string(ATT) + "/preprocess.W"), m->E, c.dim_matcher, m->Wp_h16_head, m->Wp_h16_tail,
                                  m->Wp_h16_scale, st))
      return rc;
  m->h16_stale = false;
  return 0;
}

int finalize_on_stream(lvsr_model* m, cudaStream_t st, bool synchronise) {
  const lvsr_config& c = m->cfg;
  if (m->Wcat.empty()) {
    int din = c.num_features;
    for (int l = 0; l < c.num_layers; ++l) {
      const int D = c.dims_bidir[l];
      float *W = nullptr, *b = nullptr;
      LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&W), (size_t)din * 6 * D * sizeof(float)));
      LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b), (size_t)6 * D * sizeof(float)));
      m->Wcat.push_back(W);
      m->bcat.push_back(b);
      din = 2 * D;
    }
    const int C = c.dim_dec;
    LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&m->Wd_cat), (size_t)m->E * 3 * C * sizeof(float)));
    LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&m->Wb1), (size_t)(m->E + C) * 3 * C * sizeof(float)));
    LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&m->Wff_cat), (size_t)c.dim_feedback * 3 * C * sizeof(float)));
    LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&m->bff_cat), (size_t)3 * C * sizeof(float)));
    LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&m->FF), (size_t)(c.num_phonemes + 1) * 3 * C * sizeof(float)));
  }
  int din = c.num_features;
  for (int l = 0; l < c.num_layers; ++l) {
    const int D = c.dims_bidir[l];
    for (int dir = 0; dir < 2; ++dir) {
      const std::string b = enc_base(l, dir);
      const int c0 = dir * 3 * D;   // per direction: [inputs D | gate_inputs 2D (update | reset)]
      if (int rc = copy_cols(m->Wcat[l], 6 * D, c0, m->P(b + "/fork/fork_inputs.W"), din, D, st)) return rc;
      if (int rc = copy_cols(m->Wcat[l], 6 * D, c0 + D, m->P(b + "/fork/fork_gate_inputs.W"), din, 2 * D, st)) return rc;
      if (int rc = copy_cols(m->bcat[l], 6 * D, c0, m->P(b + "/fork/fork_inputs.b"), 1, D, st)) return rc;
      if (int rc = copy_cols(m->bcat[l], 6 * D, c0 + D, m->P(b + "/fork/fork_gate_inputs.b"), 1, 2 * D, st)) return rc;
    }
    din = 2 * D;
  }
  const int C = c.dim_dec, Cfb = c.dim_feedback, V = c.num_phonemes;
  const std::string g = GEN, t = TR;
  // decoder-side packing: gate columns first (update | reset), then the candidate inputs
  if (int rc = copy_cols(m->Wd_cat, 3 * C, 0, m->P(t + "/distribute/fork_gate_inputs.W"), m->E, 2 * C, st)) return rc;
  if (int rc = copy_cols(m->Wd_cat, 3 * C, 2 * C, m->P(t + "/distribute/fork_inputs.W"), m->E, C, st)) return rc;
  LVSR_CUDA_OK(cudaMemsetAsync(m->Wb1, 0, (size_t)(m->E + C) * 3 * C * sizeof(float), st));
  if (int rc = copy_cols(m->Wb1, 3 * C, 0, m->Wd_cat, m->E, 3 * C, st)) return rc;
  if (int rc = copy_cols(m->Wb1 + (size_t)m->E * 3 * C, 3 * C, 0, m->P(t + "/transition.state_to_gates"), C, 2 * C, st)) return rc;
  if (int rc = copy_cols(m->Wff_cat, 3 * C, 0, m->P(g + "/fork/fork_gate_inputs.W"), Cfb, 2 * C, st)) return rc;
  if (int rc = copy_cols(m->Wff_cat, 3 * C, 2 * C, m->P(g + "/fork/fork_inputs.W"), Cfb, C, st)) return rc;
  if (int rc = copy_cols(m->bff_cat, 3 * C, 0, m->P(g + "/fork/fork_gate_inputs.b"), 1, 2 * C, st)) return rc;
  if (int rc = copy_cols(m->bff_cat, 3 * C, 2 * C, m->P(g + "/fork/fork_inputs.b"), 1, C, st)) return rc;
  // tensor-core operands: K-major tf32 hi/lo pairs of the fork and preprocess weights
  m->use_tc = getenv("LVSR_NO_TC_GEMM") == nullptr;
  if (m->use_tc) {
    if (m->Wcat_hi.empty()) {
      int dk = c.num_features;
      for (int l = 0; l < c.num_layers; ++l) {
        const int D = c.dims_bidir[l];
        float *h = nullptr, *lo = nullptr;
        if (gemm_tc_supported(128, 6 * D, dk)) {
          LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h), (size_t)gemm_tc_kpad(dk) * 6 * D * sizeof(float)));
          LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&lo), (size_t)gemm_tc_kpad(dk) * 6 * D * sizeof(float)));
        }
        m->Wcat_hi.push_back(h);
        m->Wcat_lo.push_back(lo);
        dk = 2 * D;
      }
      if (gemm_tc_supported(128, c.dim_matcher, m->E)) {
        LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&m->Wp_hi), (size_t)gemm_tc_kpad(m->E) * c.dim_matcher * sizeof(float)));
        LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&m->Wp_lo), (size_t)gemm_tc_kpad(m->E) * c.dim_matcher * sizeof(float)));
      }
    }
    int dk = c.num_features;
    for (int l = 0; l < c.num_layers; ++l) {
      const int D = c.dims_bidir[l];
      if (m->Wcat_hi[l])
        if (int rc = split_weight_tf32(m->Wcat[l], dk, 6 * D, m->Wcat_hi[l], m->Wcat_lo[l], st)) return rc;
      dk = 2 * D;
    }
    if (m->Wp_hi)
      if (int rc = split_weight_tf32(m->P(std::string(ATT) + "/preprocess.W"), m->E, c.dim_matcher, m->Wp_hi, m->Wp_lo, st))
        return rc;
    // fp16 head/tail operands of the same weights: re-split lazily at their next use (ensure_h16).  Opt-in
    // (LVSR_F16_GEMM=1): the GEMM class drops from 1.89 to 1.63 ms at the metric batch, but the extra device allocations move
    // the workspace to other physical pages, and the persistent decoder happened to lose more than that on the benchmarked
    // allocation sequence (51.2 -> 55.3 us per step; placement sweep in profiles/r2k_summary.md).
    m->use_h16 = getenv("LVSR_F16_GEMM") != nullptr && atoi(getenv("LVSR_F16_GEMM")) != 0;
    m->h16_stale = true;
  }
  // fork(feedback(y)) for every symbol y, once: [(V+1), 3C]
  if (c.one_of_n_feedback) {
    // one-hot feedback: fork(feedback(y)) is row y of the fork weights plus the bias
    if (int rc = add_bias_rows(m->FF, m->Wff_cat, m->bff_cat, V + 1, 3 * C, st)) return rc;
  } else {
    GemmArgs ff = make_gemm(m->P(g + "/readout/lookupfeedback/lookuptable.W"), V + 1, Cfb, m->Wff_cat, 3 * C,
                            m->bff_cat, m->FF);
    if (int rc = gemm_bias(ff, st)) return rc;
  }
  m->v_bias = 0.f;
  if (c.energy_normalizer != LVSR_NORM_SOFTMAX)
    LVSR_CUDA_OK(cudaMemcpy(&m->v_bias, m->P(std::string(ATT) + "/energy_comp/linear.b"), sizeof(float),
                            cudaMemcpyDeviceToHost));
  if (synchronise) LVSR_CUDA_OK(cudaStreamSynchronize(st));
  m->finalized = true;
  return 0;
}
}  // namespace lvsr

extern "C" {

int lvsr_encoded_length(const lvsr_model* m, int32_t T) {
  if (!m) return 0;
  int t = T;
  for (int l = 0; l < m->cfg.num_layers; ++l) t = ceil_div(t, m->cfg.subsample[l]);
  return t;
}
int lvsr_encoded_dim(const lvsr_model* m) { return m ? m->E : 0; }

int lvsr_encoder_forward(lvsr_model* m, const float* x, const float* mask, int32_t T, int32_t B,
                         float* attended, float* attended_mask, void* stream) {
  DeviceGuard device_guard(m);
  if (int rc = check_ready(m)) return rc;
  LVSR_CHECK(x && attended && attended_mask && T > 0 && B > 0, "encoder_forward: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  m->ws.reserve(encoder_ws_bytes(m, T, B), st);
  ArenaScope scope(m, st);
  const lvsr_config& c = m->cfg;
  const float* cur = x;
  int Tl = T, din = c.num_features;
  long long mstride = B;
  int kcum = 1;
  for (int l = 0; l < c.num_layers; ++l) {
    const int D = c.dims_bidir[l], k = c.subsample[l];
    const int rows = Tl * B;
    float* pre = m->ws.f32((size_t)rows * 6 * D);
    LVSR_CHECK(pre, "out of device memory (encoder pre-activations)");
    static const int h16_mask_e = getenv("LVSR_F16_GEMM_MASK") ? atoi(getenv("LVSR_F16_GEMM_MASK")) : 3;
    if (l == 1 && (h16_mask_e & 1))
      if (int rc = ensure_h16(m, st)) return rc;
    if ((h16_mask_e & 1) && m->use_tc && m->use_h16 && l < (int)m->Wcat_h16_head.size() && m->Wcat_h16_head[l] && gemm_tc_h16_supported(rows, 6 * D, din)) {
      // input = the previous layer's BiGRU output: fp16 head/tail operands (gemm_tc.cu)
      const size_t mark = m->ws.off;
      const size_t halfs = (size_t)rows * gemm_tc_kpad_h16(din);
      float* a_h = m->ws.f32((halfs + 1) / 2);
      float* a_t = m->ws.f32((halfs + 1) / 2);
      LVSR_CHECK(a_h && a_t, "out of device memory (fp16 split scratch)");
      if (int rc = gemm_tc_h16(cur, a_h, a_t, rows, din, m->Wcat_h16_head[l], m->Wcat_h16_tail[l], m->Wcat_h16_scale[l], 6 * D,
                               m->bcat[l], pre, 6 * D, st))
        return rc;
      if (m->ws.off <= m->ws.cap) m->ws.off = mark;
    } else if (m->use_tc && l < (int)m->Wcat_hi.size() && m->Wcat_hi[l] && gemm_tc_supported(rows, 6 * D, din)) {
      const size_t mark = m->ws.off;
      float* a_hi = m->ws.f32((size_t)rows * gemm_tc_kpad(din));
      float* a_lo = m->ws.f32((size_t)rows * gemm_tc_kpad(din));
      LVSR_CHECK(a_hi && a_lo, "out of device memory (tf32 split scratch)");
      if (int rc = gemm_tc(cur, a_hi, a_lo, rows, din, m->Wcat_hi[l], m->Wcat_lo[l], 6 * D, m->bcat[l], pre, 6 * D, st)) return rc;
      if (m->ws.off <= m->ws.cap) m->ws.off = mark;     // scratch is dead once the GEMM is enqueued (stream order)
    } else {
      GemmArgs g = make_gemm(cur, rows, din, m->Wcat[l], 6 * D, m->bcat[l], pre);
      if (int rc = gemm_bias(g, st)) return rc;
    }
    const int Tout = ceil_div(Tl, k);
    float* out = (l == c.num_layers - 1) ? attended : m->ws.f32((size_t)Tout * B * 2 * D);
    LVSR_CHECK(out, "out of device memory (encoder layer output)");
    BiGruArgs a = {};
    a.pre = pre; a.mask = mask; a.mask_tstride = mstride;
    const std::string bf = enc_base(l, 0) + "/gatedrecurrent", bb = enc_base(l, 1) + "/gatedrecurrent";
    a.Wg_f = m->P(bf + ".state_to_gates"); a.Ws_f = m->P(bf + ".state_to_state"); a.h0_f = m->P(bf + ".initial_state");
    a.Wg_b = m->P(bb + ".state_to_gates"); a.Ws_b = m->P(bb + ".state_to_state"); a.h0_b = m->P(bb + ".initial_state");
    a.out = out; a.T = Tl; a.B = B; a.D = D; a.subsample = k;
    if (int rc = bigru_layer(a, st)) return rc;
    cur = out; Tl = Tout; din = 2 * D; mstride *= k; kcum *= k;
  }
  if (mask) {
    if (int rc = gather_time_subsample(attended_mask, mask, Tl, kcum, B, st)) return rc;
  } else {
    if (int rc = fill_f32(attended_mask, (long long)Tl * B, 1.f, st)) return rc;   // lvsr/bricks/__init__.py:78
  }
  return 0;
}

int lvsr_preprocess(lvsr_model* m, const float* attended, int32_t Tp, int32_t U, float* out, void* stream) {
  DeviceGuard device_guard(m);
  if (int rc = check_ready(m)) return rc;
  LVSR_CHECK(attended && out && Tp > 0 && U > 0, "preprocess: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static const int h16_mask_p = getenv("LVSR_F16_GEMM_MASK") ? atoi(getenv("LVSR_F16_GEMM_MASK")) : 3;
  if (h16_mask_p & 2)
    if (int rc = ensure_h16(m, st)) return rc;
  if ((h16_mask_p & 2) && m->use_tc && m->use_h16 && m->Wp_h16_head && gemm_tc_h16_supported(Tp * U, m->cfg.dim_matcher, m->E)) {
    ArenaScope scope(m, st);
    const size_t halfs = (size_t)Tp * U * gemm_tc_kpad_h16(m->E);
    float* a_h = m->ws.f32((halfs + 1) / 2);
    float* a_t = m->ws.f32((halfs + 1) / 2);
    LVSR_CHECK(a_h && a_t, "out of device memory (fp16 split scratch)");
    return gemm_tc_h16(attended, a_h, a_t, Tp * U, m->E, m->Wp_h16_head, m->Wp_h16_tail, m->Wp_h16_scale, m->cfg.dim_matcher,
                       m->P(std::string(ATT) + "/preprocess.b"), out, m->cfg.dim_matcher, st);
  }
  if (m->use_tc && m->Wp_hi && gemm_tc_supported(Tp * U, m->cfg.dim_matcher, m->E)) {
    ArenaScope scope(m, st);
    float* a_hi = m->ws.f32((size_t)Tp * U * gemm_tc_kpad(m->E));
    float* a_lo = m->ws.f32((size_t)Tp * U * gemm_tc_kpad(m->E));
    LVSR_CHECK(a_hi && a_lo, "out of device memory (tf32 split scratch)");
    return gemm_tc(attended, a_hi, a_lo, Tp * U, m->E, m->Wp_hi, m->Wp_lo, m->cfg.dim_matcher,
                   m->P(std::string(ATT) + "/preprocess.b"), out, m->cfg.dim_matcher, st);
  }
  GemmArgs g = make_gemm(attended, Tp * U, m->E, m->P(std::string(ATT) + "/preprocess.W"), m->cfg.dim_matcher,
                         m->P(std::string(ATT) + "/preprocess.b"), out);
  return gemm_bias(g, st);
}

int lvsr_cost_matrix(lvsr_model* m, const float* attended, const float* attended_mask, int32_t Tp, int32_t B,
                     const int64_t* labels, const float* labels_mask, int32_t L, float* costs,
                     float* weights_out, float* energies_out, float* states_out, float* wavg_out, void* stream) {
  DeviceGuard device_guard(m);
  if (int rc = check_ready(m)) return rc;
  LVSR_CHECK(attended && attended_mask && labels && costs && Tp > 0 && B > 0 && L > 0, "cost_matrix: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  m->ws.reserve(cost_ws_bytes(m, Tp, B, L), st);
  ArenaScope scope(m, st);
  const lvsr_config& c = m->cfg;
  Arena& ws = m->ws;
  const int C = c.dim_dec, E = m->E, M = c.dim_matcher;
  const long long* lab = reinterpret_cast<const long long*>(labels);

  float* P = ws.f32((size_t)Tp * B * M);
  float* s_all = ws.f32((size_t)(L + 1) * B * C);
  float* ctx_all = wavg_out ? wavg_out : ws.f32((size_t)L * B * E);
  float* w0 = ws.f32((size_t)B * Tp);
  float* wpp[2] = {ws.f32((size_t)B * Tp), ws.f32((size_t)B * Tp)};   // step-wise fallback only
  float* e_scratch = energies_out ? nullptr : ws.f32((size_t)B * Tp);
  float* merged = ws.f32((size_t)L * B * c.post_merge_dim);
  LVSR_CHECK(P && s_all && ctx_all && w0 && merged && wpp[0] && wpp[1] && (energies_out || e_scratch),
             "out of device memory (decoder workspace)");

  if (int rc = lvsr_preprocess(m, attended, Tp, B, P, stream)) return rc;          // hoisted: B/bricks/attention.py:733-738
  if (int rc = broadcast_rows(s_all, m->P(std::string(TR) + "/transition.initial_state"), B, C, st)) return rc;
  if (int rc = onehot_rows(w0, B, Tp, st)) return rc;                               // lvsr/bricks/attention.py:215-222
  if (int rc = fill_f32(costs, (long long)L * B, 0.f, st)) return rc;

  bool scanned = false;
  LVSR_CUDA_OK(cudaMemsetAsync(m->status, 0, sizeof(unsigned), st));
  if (getenv("LVSR_NO_DEC_SCAN") == nullptr && !m->force_stepwise) {
    DecScanArgs d = {};
    d.P = P; d.H = attended; d.maskH = attended_mask;
    d.filt = m->P(std::string(ATT) + "/conv1d.filters");
    d.Wh = m->P(std::string(ATT) + "/handler.W");
    d.v = m->P(std::string(ATT) + "/energy_comp/linear.W");
    d.v_bias = m->v_bias;
    d.prior = prior_of(c);
    d.Wb1 = m->Wb1;
    d.Wstate = m->P(std::string(TR) + "/transition.state_to_state");
    d.Ws = m->P(std::string(ATT) + "/state_trans/transform_states.W");
    d.FF = m->FF;
    d.labels = lab; d.lmask = labels_mask;
    d.s_all = s_all; d.ctx_all = ctx_all; d.w0 = w0;
    d.e_seq = energies_out; d.e_scratch = e_scratch;
    d.status = m->status;
    d.Tp = Tp; d.B = B; d.L = L; d.M = M; d.E = E; d.C = C; d.K = c.conv_num_filters; d.n = c.conv_n;
    d.normalizer = c.energy_normalizer;
    d.V = c.num_phonemes;
    // per-step hand-over buffers of the data-flow decoder; everything another CTA polls starts
    // as the sentinel (0xFF bytes)
    d.w_all = weights_out ? weights_out : ws.f32((size_t)L * B * Tp);
    d.q_all = ws.f32((size_t)L * B * M);
    d.hr_all = ws.f32((size_t)L * B * C);
    d.rowpos_all = ws.f32((size_t)(L + 1) * B);
    LVSR_CHECK(d.w_all && d.q_all && d.hr_all && d.rowpos_all,
               "out of device memory (decoder scan workspace)");
    LVSR_CUDA_OK(cudaMemsetAsync(d.w_all, 0xFF, (size_t)L * B * Tp * sizeof(float), st));
    LVSR_CUDA_OK(cudaMemsetAsync(d.q_all, 0xFF, (size_t)L * B * M * sizeof(float), st));
    LVSR_CUDA_OK(cudaMemsetAsync(d.hr_all, 0xFF, (size_t)L * B * C * sizeof(float), st));
    LVSR_CUDA_OK(cudaMemsetAsync(d.rowpos_all + B, 0xFF, (size_t)L * B * sizeof(float), st));
    LVSR_CUDA_OK(cudaMemsetAsync(d.rowpos_all, 0, (size_t)B * sizeof(float), st));
    LVSR_CUDA_OK(cudaMemsetAsync(s_all + (size_t)B * C, 0xFF, (size_t)L * B * C * sizeof(float), st));
    LVSR_CUDA_OK(cudaMemsetAsync(ctx_all, 0xFF, (size_t)L * B * E * sizeof(float), st));
    const bool trace = getenv("LVSR_DEC_TRACE") != nullptr;
    if (trace) {
      d.trace = reinterpret_cast<unsigned long long*>(ws.i64((size_t)2 * L * 9 + (size_t)L * 12 + (size_t)L * B));
      LVSR_CUDA_OK(cudaMemsetAsync(d.trace, 0, ((size_t)2 * L * 9 + (size_t)L * 12 + (size_t)L * B) * 8, st));
    }
    int supported = 0;
    if (int rc = dec_scan_try(d, &supported, st)) return rc;
    scanned = supported != 0;
    if (scanned && getenv("LVSR_DEC_CHECK") != nullptr) {
      // debug post-condition: the launch reported success and every hand-over word was written
      LVSR_CUDA_OK(cudaStreamSynchronize(st));
      unsigned hst = 0;
      LVSR_CUDA_OK(cudaMemcpy(&hst, m->status, sizeof(hst), cudaMemcpyDeviceToHost));
      LVSR_CHECK(hst == 0, "LVSR_DEC_CHECK: persistent decoder launch status %u (2 = a value never arrived, "
                 "3 = launched without its cluster shape)", hst);
      struct { const char* name; const float* p; size_t n; } bufs[] = {
          {"weights", d.w_all, (size_t)L * B * Tp}, {"queries", d.q_all, (size_t)L * B * M},
          {"reset-gated states", d.hr_all, (size_t)L * B * C}, {"states", s_all, (size_t)(L + 1) * B * C},
          {"weighted averages", ctx_all, (size_t)L * B * E},
          {"row positions", d.rowpos_all, c.prior_type == LVSR_PRIOR_EXPANDING ? (size_t)B : (size_t)(L + 1) * B}};
      for (auto& b : bufs) {
        long long left = 0;
        if (int rc = count_sentinels(b.p, (long long)b.n, &left, st)) return rc;
        // the query of step L is never needed; everything else must have been produced
        LVSR_CHECK(left == 0, "LVSR_DEC_CHECK: %lld sentinel words left in the %s buffer", left, b.name);
      }
    }
    if (trace && scanned) {
      std::vector<unsigned long long> h((size_t)2 * L * 9 + (size_t)L * 12 + (size_t)L * B);
      LVSR_CUDA_OK(cudaMemcpyAsync(h.data(), d.trace, h.size() * 8, cudaMemcpyDeviceToHost, st));
      LVSR_CUDA_OK(cudaStreamSynchronize(st));
      const char* names[8] = {"A", "syncA", "B1", "sync1", "B2", "sync2", "B3", "sync3"};
      for (int slot = 0; slot < 2; ++slot) {
        double sum[8] = {0};
        int n = 0;
        for (int i = 1; i + 1 < L; ++i, ++n)
          for (int j = 0; j < 8; ++j) sum[j] += (double)(h[((size_t)slot * L + i) * 9 + j + 1] - h[((size_t)slot * L + i) * 9 + j]);
        fprintf(stderr, "[dec_scan trace] CTA %s:", slot == 0 ? "first" : "last");
        for (int j = 0; j < 8; ++j) fprintf(stderr, " %s=%.2fus", names[j], n ? sum[j] / n * 1e-3 : 0.0);
        fprintf(stderr, "\n");
      }
      {
        const char* an[7] = {"stage", "conv", "energy", "stats", "ctx", "exchange", "combine"};
        double sum[7] = {0};
        int n = 0;
        for (int i = 1; i + 1 < L; ++i, ++n)
          for (int j = 0; j < 7; ++j)
            sum[j] += (double)(h[(size_t)2 * L * 9 + (size_t)i * 8 + j + 1] - h[(size_t)2 * L * 9 + (size_t)i * 8 + j]);
        fprintf(stderr, "[dec_scan trace] attention row 0:");
        for (int j = 0; j < 7; ++j) fprintf(stderr, " %s=%.2fus", an[j], n ? sum[j] / n * 1e-3 : 0.0);
        fprintf(stderr, "\n");
      }
      {
        // gate tile of CTA 0: start of B1 -> x arrived -> products done -> cross-warp sums done -> end of B1
        double sum[4] = {0};
        int n = 0;
        for (int i = 1; i + 1 < L; ++i, ++n) {
          const unsigned long long* b = &h[(size_t)0 * L * 9 + (size_t)i * 9];
          const unsigned long long* t = &h[(size_t)2 * L * 9 + (size_t)L * 8 + (size_t)i * 4];
          sum[0] += (double)(t[0] - b[2]); sum[1] += (double)(t[1] - t[0]);
          sum[2] += (double)(t[2] - t[1]); sum[3] += (double)(b[3] - t[2]);
        }
        {
          // when each row's attention phase ended, relative to row 0 (mean over steps)
          fprintf(stderr, "[dec_scan trace] end of attention vs row 0 (us):");
          for (int r = 0; r < B; ++r) {
            double acc = 0;
            for (int i = 1; i + 1 < L; ++i) {
              const unsigned long long* e = &h[(size_t)2 * L * 9 + (size_t)L * 12 + (size_t)i * B];
              acc += (double)((long long)e[r] - (long long)e[0]);
            }
            fprintf(stderr, " %.1f", acc / (L - 2) * 1e-3);
          }
          fprintf(stderr, "\n");
        }
        fprintf(stderr, "[dec_scan trace] gate tile: wait_x=%.2fus products=%.2fus sums=%.2fus epilogue=%.2fus\n",
                n ? sum[0] / n * 1e-3 : 0.0, n ? sum[1] / n * 1e-3 : 0.0, n ? sum[2] / n * 1e-3 : 0.0,
                n ? sum[3] / n * 1e-3 : 0.0);
      }
    }
  }
  const float* w_prev = w0;
  for (int i = 0; i < L && !scanned; ++i) {
    float* w_i = weights_out ? weights_out + (size_t)i * B * Tp : wpp[i & 1];
    float* e_i = energies_out ? energies_out + (size_t)i * B * Tp : e_scratch;
    float* ctx_i = ctx_all + (size_t)i * B * E;
    const float* s_i = s_all + (size_t)i * B * C;
    const size_t mark = ws.off;
    if (int rc = glimpses(m, attended, P, attended_mask, Tp, B, nullptr, B, s_i, w_prev, nullptr, i, w_i, e_i, ctx_i, st)) return rc;
    if (int rc = transition(m, B, s_i, ctx_i, lab + (size_t)i * B, labels_mask ? labels_mask + (size_t)i * B : nullptr,
                            s_all + (size_t)(i + 1) * B * C, st)) return rc;
    if (ws.off <= ws.cap) ws.off = mark;   // per-step scratch is reusable (stream order)
    w_prev = w_i;
  }
  // readout(states[:-1], glimpses[1:]) for all steps at once, then the emitter cost
  {
    const int R = L * B;
    const std::string g = GEN;
    bool acc = false;
    if (c.use_states_for_readout) {
      GemmArgs a = make_gemm(s_all, R, C, m->P(g + "/readout/merge/transform_states.W"), c.post_merge_dim, nullptr, merged);
      if (int rc = gemm_bias(a, st)) return rc;
      acc = true;
    }
    GemmArgs b = make_gemm(ctx_all, R, E, m->P(g + "/readout/merge/transform_weighted_averages.W"), c.post_merge_dim,
                           nullptr, merged, acc);
    if (int rc = gemm_bias(b, st)) return rc;
    ReadoutArgs r = readout_args(m, R, merged);
    r.labels = lab; r.lmask = labels_mask; r.costs_picked = costs;
    r.poison = scanned ? m->status : nullptr;
    if (int rc = readout_costs(r, st)) return rc;
  }
  if (states_out)
    LVSR_CUDA_OK(cudaMemcpyAsync(states_out, s_all, (size_t)L * B * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int lvsr_initial_states(lvsr_model* m, int32_t Tp, int32_t R, float* states, int64_t* outputs, float* wavg,
                        float* weights, float* energies, int64_t* step, void* stream) {
  DeviceGuard device_guard(m);
  if (int rc = check_ready(m)) return rc;
  LVSR_CHECK(states && outputs && wavg && weights && energies && step && Tp > 0 && R > 0, "initial_states: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const lvsr_config& c = m->cfg;
  if (int rc = broadcast_rows(states, m->P(std::string(TR) + "/transition.initial_state"), R, c.dim_dec, st)) return rc;
  if (int rc = fill_i64(reinterpret_cast<long long*>(outputs), R, c.num_phonemes, st)) return rc;   // recognizer.py:286
  if (int rc = fill_f32(wavg, (long long)R * m->E, 0.f, st)) return rc;
  if (int rc = onehot_rows(weights, R, Tp, st)) return rc;
  if (int rc = onehot_rows(energies, R, Tp, st)) return rc;
  return fill_i64(reinterpret_cast<long long*>(step), R, 0, st);
}

int lvsr_logprobs(lvsr_model* m, const float* attended, const float* preprocessed, const float* attended_mask,
                  int32_t Tp, int32_t U, const int32_t* row_utt, int32_t R, const float* states,
                  const float* weights, const int64_t* step, float* out, void* stream) {
  DeviceGuard device_guard(m);
  if (int rc = check_ready(m)) return rc;
  LVSR_CHECK(attended && attended_mask && states && weights && step && out && Tp > 0 && U > 0 && R > 0,
             "logprobs: bad arguments");
  LVSR_CHECK(row_utt || U == R, "logprobs: without row_utt the contexts must be replicated (U == R)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ArenaScope scope(m, st);
  const lvsr_config& c = m->cfg;
  Arena& ws = m->ws;
  const float* P = preprocessed;
  if (!P) {
    float* Pb = ws.f32((size_t)Tp * U * c.dim_matcher);
    LVSR_CHECK(Pb, "out of device memory (preprocessed)");
    if (int rc = lvsr_preprocess(m, attended, Tp, U, Pb, stream)) return rc;
    P = Pb;
  }
  float* w_tmp = ws.f32((size_t)R * Tp);
  float* e_tmp = ws.f32((size_t)R * Tp);
  float* ctx = ws.f32((size_t)R * m->E);
  float* merged = ws.f32((size_t)R * c.post_merge_dim);
  LVSR_CHECK(w_tmp && e_tmp && ctx && merged, "out of device memory (logprobs workspace)");
  if (int rc = glimpses(m, attended, P, attended_mask, Tp, U, row_utt, R, states, weights,
                        reinterpret_cast<const long long*>(step), 0, w_tmp, e_tmp, ctx, st)) return rc;
  if (int rc = readout_merged(m, R, states, ctx, merged, st)) return rc;
  ReadoutArgs r = readout_args(m, R, merged);
  r.costs_all = out;
  return readout_costs(r, st);
}

int lvsr_next_states(lvsr_model* m, const float* attended, const float* preprocessed, const float* attended_mask,
                     int32_t Tp, int32_t U, const int32_t* row_utt, int32_t R, const float* states,
                     const float* weights, const int64_t* step, const int64_t* outputs, float* next_states,
                     float* next_wavg, float* next_weights, float* next_energies, int64_t* next_step, void* stream) {
  DeviceGuard device_guard(m);
  if (int rc = check_ready(m)) return rc;
  LVSR_CHECK(attended && attended_mask && states && weights && step && outputs && next_states && next_wavg &&
                 next_weights && next_energies && next_step && Tp > 0 && U > 0 && R > 0,
             "next_states: bad arguments");
  LVSR_CHECK(row_utt || U == R, "next_states: without row_utt the contexts must be replicated (U == R)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ArenaScope scope(m, st);
  const lvsr_config& c = m->cfg;
  Arena& ws = m->ws;
  const float* P = preprocessed;
  if (!P) {
    float* Pb = ws.f32((size_t)Tp * U * c.dim_matcher);
    LVSR_CHECK(Pb, "out of device memory (preprocessed)");
    if (int rc = lvsr_preprocess(m, attended, Tp, U, Pb, stream)) return rc;
    P = Pb;
  }
  if (int rc = glimpses(m, attended, P, attended_mask, Tp, U, row_utt, R, states, weights,
                        reinterpret_cast<const long long*>(step), 0, next_weights, next_energies, next_wavg, st)) return rc;
  if (int rc = transition(m, R, states, next_wavg, reinterpret_cast<const long long*>(outputs), nullptr, next_states, st)) return rc;
  return add_i64(reinterpret_cast<long long*>(next_step), reinterpret_cast<const long long*>(step), R, 1, st);
}

int lvsr_search_expand(lvsr_model* m, const float* attended, const float* preprocessed, const float* attended_mask,
                       int32_t Tp, int32_t U, const int32_t* utt_len, const int32_t* row_utt, const int32_t* row_seg,
                       const int32_t* seg_start, int32_t nseg, int32_t R, const float* states, const float* weights,
                       const int64_t* step, const float* cost_so_far, int32_t k, float* wavg, float* new_weights,
                       float* new_energies, int32_t* top_parent, int32_t* top_symbol, float* top_cost, int32_t* top_count,
                       void* stream) {
  DeviceGuard device_guard(m);
  if (int rc = check_ready(m)) return rc;
  LVSR_CHECK(attended && preprocessed && attended_mask && row_utt && row_seg && seg_start && states && weights && step &&
                 cost_so_far && wavg && new_weights && new_energies && top_parent && top_symbol && top_cost && top_count &&
                 Tp > 0 && U > 0 && nseg > 0 && R > 0 && k > 0,
             "search_expand: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ArenaScope scope(m, st);
  const lvsr_config& c = m->cfg;
  Arena& ws = m->ws;
  float* merged = ws.f32((size_t)R * c.post_merge_dim);
  float* neglogp = ws.f32((size_t)R * c.num_phonemes);
  LVSR_CHECK(merged && neglogp, "out of device memory (search workspace)");
  Segments sg;
  sg.seg_start = seg_start; sg.nseg = nseg; sg.seg_len = utt_len; sg.row_seg = row_seg;
  // take_glimpses ONCE per hypothesis: the same glimpse feeds the readout (logprobs_computer) and, for the
  // surviving parents, the state update (next_state_computer) -- B/search.py:109-142 computes it twice
  if (int rc = glimpses(m, attended, preprocessed, attended_mask, Tp, U, row_utt, R, states, weights,
                        reinterpret_cast<const long long*>(step), 0, new_weights, new_energies, wavg, st, sg)) return rc;
  if (int rc = readout_merged(m, R, states, wavg, merged, st)) return rc;
  ReadoutArgs r = readout_args(m, R, merged);
  r.costs_all = neglogp;
  if (int rc = readout_costs(r, st)) return rc;
  return segment_topk(neglogp, cost_so_far, seg_start, nseg, c.num_phonemes, k, top_parent, top_symbol, top_cost, top_count, st);
}

int lvsr_search_advance(lvsr_model* m, const float* attended, const float* preprocessed, const float* attended_mask,
                        int32_t Tp, int32_t U, const int32_t* utt_len, int32_t Rn, const int32_t* parent,
                        const int64_t* symbols, const int32_t* row_utt, const int32_t* row_seg, const int32_t* seg_start,
                        int32_t nseg, const float* states, const float* weights, const int64_t* step, const float* wavg,
                        const float* new_weights, const float* new_energies, int32_t reuse_glimpses, float* n_states,
                        float* n_wavg, float* n_weights, float* n_energies, int64_t* n_step, void* stream) {
  DeviceGuard device_guard(m);
  if (int rc = check_ready(m)) return rc;
  LVSR_CHECK(parent && symbols && states && step && n_states && n_wavg && n_weights && n_energies && n_step && Rn > 0 && Tp > 0,
             "search_advance: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ArenaScope scope(m, st);
  const lvsr_config& c = m->cfg;
  Arena& ws = m->ws;
  const int C = c.dim_dec, E = m->E;
  float* s_sel = ws.f32((size_t)Rn * C);
  LVSR_CHECK(s_sel, "out of device memory (search workspace)");
  if (int rc = gather_rows(s_sel, states, parent, Rn, C, st)) return rc;
  if (reuse_glimpses) {
    LVSR_CHECK(wavg && new_weights && new_energies, "search_advance: glimpses of the parents are missing");
    if (int rc = gather_rows(n_wavg, wavg, parent, Rn, E, st)) return rc;
    if (int rc = gather_rows(n_weights, new_weights, parent, Rn, Tp, st)) return rc;
    if (int rc = gather_rows(n_energies, new_energies, parent, Rn, Tp, st)) return rc;
  } else {
    // window priors: the reference recomputes the glimpses over the SELECTED parents, whose batch-global cut
    // (lvsr/bricks/attention.py:151-152) can differ from the cut over the whole beam
    LVSR_CHECK(attended && preprocessed && attended_mask && row_utt && row_seg && seg_start && weights && nseg > 0,
               "search_advance: contexts are required to recompute the glimpses");
    float* w_sel = ws.f32((size_t)Rn * Tp);
    long long* st_sel = ws.i64((size_t)Rn);
    LVSR_CHECK(w_sel && st_sel, "out of device memory (search workspace)");
    if (int rc = gather_rows(w_sel, weights, parent, Rn, Tp, st)) return rc;
    if (int rc = gather_i64(st_sel, reinterpret_cast<const long long*>(step), parent, Rn, 0, st)) return rc;
    Segments sg;
    sg.seg_start = seg_start; sg.nseg = nseg; sg.seg_len = utt_len; sg.row_seg = row_seg;
    if (int rc = glimpses(m, attended, preprocessed, attended_mask, Tp, U, row_utt, Rn, s_sel, w_sel, st_sel, 0, n_weights,
                          n_energies, n_wavg, st, sg)) return rc;
  }
  if (int rc = transition(m, Rn, s_sel, n_wavg, reinterpret_cast<const long long*>(symbols), nullptr, n_states, st)) return rc;
  return gather_i64(reinterpret_cast<long long*>(n_step), reinterpret_cast<const long long*>(step), parent, Rn, 1, st);
}

int lvsr_recognizer_cost_host(lvsr_model* m, const float* x_h, const float* mask_h, const int64_t* labels_h,
                              const float* lmask_h, int32_t T, int32_t B, int32_t L, float* costs_h, void* stream) {
  DeviceGuard device_guard(m);
  if (int rc = check_ready(m)) return rc;
  LVSR_CHECK(x_h && labels_h && costs_h && T > 0 && B > 0 && L > 0, "recognizer_cost_host: bad arguments");
  for (long long i = 0; i < (long long)L * B; ++i)       // host memory: the lookup's IndexError, up front
    LVSR_CHECK(labels_h[i] >= 0 && labels_h[i] < m->cfg.num_phonemes,
               "recognizer_cost_host: label %lld at [%lld, %lld] outside [0, %d)", (long long)labels_h[i], i / B, i % B,
               m->cfg.num_phonemes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = 0;
  {
    const int F = m->cfg.num_features, Tp = lvsr_encoded_length(m, T), E = m->E;
    m->ws.reserve(encoder_ws_bytes(m, T, B) + cost_ws_bytes(m, Tp, B, L) +
                      ((size_t)T * B * (F + 1) + (size_t)3 * L * B + (size_t)Tp * B * (E + 1)) * sizeof(float) + (1 << 16),
                  st);
    ArenaScope scope(m, st);
    Arena& ws = m->ws;
    float* x = ws.f32((size_t)T * B * F);
    float* mask = mask_h ? ws.f32((size_t)T * B) : nullptr;
    long long* lab = ws.i64((size_t)L * B);
    float* lmask = lmask_h ? ws.f32((size_t)L * B) : nullptr;
    float* att = ws.f32((size_t)Tp * B * E);
    float* attm = ws.f32((size_t)Tp * B);
    float* costs = ws.f32((size_t)L * B);
    LVSR_CHECK(x && lab && att && attm && costs && (!mask_h || mask) && (!lmask_h || lmask), "out of device memory (host call)");
    LVSR_CUDA_OK(cudaMemcpyAsync(x, x_h, (size_t)T * B * F * sizeof(float), cudaMemcpyHostToDevice, st));
    if (mask) LVSR_CUDA_OK(cudaMemcpyAsync(mask, mask_h, (size_t)T * B * sizeof(float), cudaMemcpyHostToDevice, st));
    LVSR_CUDA_OK(cudaMemcpyAsync(lab, labels_h, (size_t)L * B * sizeof(long long), cudaMemcpyHostToDevice, st));
    if (lmask) LVSR_CUDA_OK(cudaMemcpyAsync(lmask, lmask_h, (size_t)L * B * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = lvsr_encoder_forward(m, x, mask, T, B, att, attm, stream);
    if (!rc) rc = lvsr_cost_matrix(m, att, attm, Tp, B, reinterpret_cast<const int64_t*>(lab), lmask, L, costs,
                                   nullptr, nullptr, nullptr, nullptr, stream);
    if (!rc) {
      unsigned hst = 0;
      LVSR_CUDA_OK(cudaMemcpyAsync(costs_h, costs, (size_t)L * B * sizeof(float), cudaMemcpyDeviceToHost, st));
      LVSR_CUDA_OK(cudaMemcpyAsync(&hst, m->status, sizeof(hst), cudaMemcpyDeviceToHost, st));
      LVSR_CUDA_OK(cudaStreamSynchronize(st));
      if (hst != 0) {
        // the persistent decoder gave up (status in common.cuh): same math on the step-wise kernels
        if (m->dec_fallbacks++ == 0)
          fprintf(stderr, "[lvsr_b200] persistent decoder launch failed (status %u); re-running on the step-wise kernels\n", hst);
        m->force_stepwise = true;
        rc = lvsr_cost_matrix(m, att, attm, Tp, B, reinterpret_cast<const int64_t*>(lab), lmask, L, costs,
                              nullptr, nullptr, nullptr, nullptr, stream);
        m->force_stepwise = false;
        if (!rc) {
          LVSR_CUDA_OK(cudaMemcpyAsync(costs_h, costs, (size_t)L * B * sizeof(float), cudaMemcpyDeviceToHost, st));
          LVSR_CUDA_OK(cudaStreamSynchronize(st));
        }
      }
    }
  }
  return rc;
}

}  // extern "C"
