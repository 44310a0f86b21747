// Training step of the recognizer: cost, gradients of every parameter, and the step rules.
//
// Replaces GradientDescent._function (libs/blocks/blocks/algorithms/__init__.py:244-256,284-287) built by
// lvsr/main.py:340-345,480-519: cost = sum(cost_matrix) / batch size, gradients by back-propagation
// through the decoder scan (attention + GRU), the readout and the pyramidal BiGRU encoder, then
// StepClipping -> Momentum -> AdaDelta -> Restrict(VariableClipping(axis=0), WEIGHT) -> RemoveNotFinite(0.0)
// -> BurnIn and the in-place update.  The gradient buffer uses the flat parameter layout, so a data-parallel
// caller all-reduces ONE buffer between lvsr_train_cost_and_grads and lvsr_train_apply_updates.
//
// Structure of the backward pass (B200 view):
//   * everything that does not depend on the recurrence is a LARGE GEMM over all steps at once
//     (weight gradients X^T dY as split-R TN products, input gradients dY W^T, the decoder's gate values
//     re-computed for all L steps from the saved states and glimpses);
//   * the two recurrences are reverse-time scans: bigru_bwd.cu (persistent cluster kernel, same DSMEM
//     all-gathers as the forward) and the decoder loop below (per step: 4 skinny products, the attention
//     backward kernel -- 2 CTAs per utterance re-computing tanh(match) instead of storing [L,T',B,M] --
//     and two element-wise kernels).
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "model.h"
#include "train_kernels.cuh"

using namespace lvsr;
using namespace lvsr::train;

namespace {

inline int grid1d(long long n, int block = 256, int cap = 2048) {
  return (int)std::min<long long>(cap, std::max<long long>(1, (n + block - 1) / block));
}

// C[Mo,N] (ldc) (+)= A[:, m-cols]^T . B over R rows
int gemm_tn(Arena& ws, const float* A, int lda, const float* B, int ldb, int R, int Mo, int N, float* C, int ldc,
            bool accumulate, cudaStream_t st) {
  ProfScope prof("gemm_tn", st);
  if (R <= 0 || Mo <= 0 || N <= 0) return 0;
  const int tiles = ceil_div(Mo, TN_BM) * ceil_div(N, TN_BN);
  int splits = std::max(1, std::min(ceil_div(2 * device_sm_count(), tiles), ceil_div(R, 4 * TN_BK)));
  int rps = ceil_div(ceil_div(R, splits), TN_BK) * TN_BK;
  splits = ceil_div(R, rps);
  const size_t mark = ws.off;
  float* part = ws.f32((size_t)splits * Mo * N);
  LVSR_CHECK(part, "out of device memory (TN partials)");
  TnArgs g;
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.R = R; g.Mo = Mo; g.N = N; g.part = part; g.rows_per_split = rps;
  dim3 grid(ceil_div(N, TN_BN), ceil_div(Mo, TN_BM), splits);
  gemm_tn_kernel<<<grid, 256, 0, st>>>(g);
  LVSR_LAUNCH_CHECK();
  tn_reduce_kernel<<<grid1d((long long)Mo * N), 256, 0, st>>>(part, splits, Mo, N, C, ldc, accumulate ? 1 : 0);
  LVSR_LAUNCH_CHECK();
  if (ws.off <= ws.cap) ws.off = mark;      // partials are dead once the reduce is enqueued (stream order)
  return 0;
}

int gemm_nn(const float* A, int M, int K, int lda, const float* W, int N, int ldw, const float* bias, float* C, int ldc,
            bool accumulate, cudaStream_t st) {
  GemmArgs g = make_gemm(A, M, K, W, N, bias, C, accumulate);
  g.lda = lda; g.ldw = ldw; g.ldc = ldc;
  return gemm_bias(g, st);
}

int colsum(const float* X, int R, int N, int ldx, float* out, bool accumulate, cudaStream_t st) {
  if (N <= 0) return 0;
  colsum_kernel<<<ceil_div(N, 32), 256, 0, st>>>(X, R, N, ldx, out, accumulate ? 1 : 0);
  LVSR_LAUNCH_CHECK();
  return 0;
}

int transpose(const float* src, float* dst, int K, int N, cudaStream_t st) {
  dim3 grid(ceil_div(N, 32), ceil_div(K, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(src, dst, K, N);
  LVSR_LAUNCH_CHECK();
  return 0;
}

int skinny(const float* X0, int K0, int ldx0, const float* W0, const float* X1, int K1, int ldx1, const float* W1,
           const float* add0, int lda0, const float* add1, int lda1, float* out, int ldo, int R, int N, cudaStream_t st,
           int split = 0, float* out1 = nullptr, int ldo1 = 0) {
  LVSR_CHECK(N % 4 == 0 && K0 % 4 == 0 && (X1 == nullptr || K1 % 4 == 0) && ldx0 % 4 == 0, "skinny: dimensions must be multiples of 4");
  SkinnyArgs a = {};
  a.X[0] = X0; a.K[0] = K0; a.ldx[0] = ldx0; a.W[0] = W0;
  a.X[1] = X1; a.K[1] = K1; a.ldx[1] = ldx1; a.W[1] = W1;
  a.add[0] = add0; a.lda[0] = lda0; a.add[1] = add1; a.lda[1] = lda1;
  a.out = out; a.ldo = ldo; a.R = R; a.N = N;
  a.split = split; a.out1 = out1; a.ldo1 = ldo1;
  LVSR_CHECK(split % SK_N == 0, "skinny: the column split must be a multiple of %d", SK_N);
  dim3 grid(ceil_div(N, SK_N), ceil_div(R, SK_R));
  ProfScope prof("skinny", st);
  skinny_kernel<<<grid, SK_WARPS * 32, 0, st>>>(a);
  LVSR_LAUNCH_CHECK();
  return 0;
}

int copy2d(float* dst, int ld_dst, const float* src, int ld_src, int rows, int cols, cudaStream_t st) {
  LVSR_CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)ld_dst * sizeof(float), src, (size_t)ld_src * sizeof(float),
                                 (size_t)cols * sizeof(float), rows, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// K-major (contraction-major) tf32 hi/lo operand of the tcgen05 GEMM: [rows, Kpad] built from a [K, rows] matrix
struct TcOperand { float* hi = nullptr; float* lo = nullptr; int rows = 0, Kpad = 0; };

// src [R, cols] (leading dimension ld) -> transposed hi/lo pair [cols, kpad(R)]
int make_tc_operand(Arena& ws, const float* src, int R, int cols, int ld, TcOperand* out, cudaStream_t st) {
  out->rows = cols;
  out->Kpad = gemm_tc_kpad(R);
  out->hi = ws.f32((size_t)cols * out->Kpad);
  out->lo = ws.f32((size_t)cols * out->Kpad);
  LVSR_CHECK(out->hi && out->lo, "out of device memory (tensor-core operand)");
  return transpose_split_tf32(src, R, cols, ld, out->hi, out->lo, st);
}

// C[Mo, N] (ldc) (+)= A^T B on the tensor cores: A, B given as K-major operands (rows a0.. / b0..), split-K over the
// contraction (R) so that every SM gets a tile; partials are summed in a fixed order.
int gemm_tn_tc(Arena& ws, const TcOperand& A, int a0, int Mo, const TcOperand& B, int b0, int N, float* C, int ldc,
               bool accumulate, cudaStream_t st) {
  ProfScope prof("gemm_tn", st);
  const int tiles = ceil_div(Mo, 128) * (N / (N % 256 == 0 ? 256 : 128));
  const int want = std::max(1, std::min(32, ceil_div(device_sm_count(), tiles)));
  const int splits = gemm_tc_splits_launched(A.Kpad, want);
  const size_t mark = ws.off;
  float* part = ws.f32((size_t)splits * Mo * N);
  LVSR_CHECK(part, "out of device memory (TN partials)");
  if (int rc = gemm_tc_presplit(A.hi + (size_t)a0 * A.Kpad, A.lo + (size_t)a0 * A.Kpad, Mo, B.hi + (size_t)b0 * B.Kpad,
                                B.lo + (size_t)b0 * B.Kpad, N, A.Kpad, nullptr, part, N, want, (long long)Mo * N, st)) return rc;
  tn_reduce_kernel<<<grid1d((long long)Mo * N), 256, 0, st>>>(part, splits, Mo, N, C, ldc, accumulate ? 1 : 0);
  LVSR_LAUNCH_CHECK();
  if (ws.off <= ws.cap) ws.off = mark;
  return 0;
}

struct LayerTape {
  const float* X;      // input of the layer [T*B, Din]
  float* pre;          // [T*B, 6D] forward tape, then dPre
  float* hext;         // [(T+2), B, 2D]
  float* out;          // [Tout, B, 2D]
  int T, Tout, Din, D, k;
  long long mstride;
};

float* grad_of(lvsr_model* m, float* grads, const std::string& name) {
  auto it = m->index.find(name);
  return it == m->index.end() ? nullptr : grads + m->params[it->second].offset;
}

}  // namespace

extern "C" {

int lvsr_train_cost_and_grads(lvsr_model* m, const float* x, const float* mask, const int64_t* labels, const float* lmask,
                              int32_t T, int32_t B, int32_t L, float gscale, float* cost_out, float* grads, void* stream) {
  DeviceGuard device_guard(m);
  if (int rc = check_ready(m)) return rc;
  LVSR_CHECK(x && labels && cost_out && grads && T > 0 && B > 0 && L > 0, "train_cost_and_grads: bad arguments");
  const lvsr_config& c = m->cfg;
  LVSR_CHECK(c.energy_normalizer == LVSR_NORM_SOFTMAX,
             "training supports the softmax energy normaliser only (logistic / relu: inference only)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = c.dim_dec, E = m->E, M = c.dim_matcher, K = c.conv_num_filters, n = c.conv_n, w = 2 * n + 1;
  const int V = c.num_phonemes, Cfb = c.dim_feedback, Cpm = c.post_merge_dim, Hd = Cpm / c.maxout_pieces;
  const int Tp = lvsr_encoded_length(m, T);
  const int R = L * B;
  const long long* lab = reinterpret_cast<const long long*>(labels);
  Arena& ws = m->tws;
  // size the tape arena once per shape
  {
    size_t bytes = (size_t)64 << 20;
    int Tl = T, din = c.num_features;
    for (int l = 0; l < c.num_layers; ++l) {
      const int D = c.dims_bidir[l], Tout = ceil_div(Tl, c.subsample[l]);
      bytes += ((size_t)Tl * B * 6 * D * 2 + (size_t)(Tl + 2) * B * 2 * D * 2 + (size_t)Tout * B * 2 * D * 2 + (size_t)3 * Tl * B * gemm_tc_kpad(din)) * sizeof(float);
      bytes += (size_t)80 * std::max(din, 2 * D) * 6 * D * sizeof(float);      // TN partials
      bytes += ((size_t)2 * (6 * D + din + 3 * D + 32) * (Tl * (size_t)B + 32) + (size_t)2 * Tl * B * 6 * D) * sizeof(float);   // K-major tf32 operands
      Tl = Tout; din = 2 * D;
    }
    bytes += ((size_t)Tp * B * (2 * M + 2 * E) + (size_t)R * (Tp + 8 * C + 3 * E + 2 * M + 3 * Cpm + V + 16) + (size_t)4 * B * Tp +
              (size_t)2 * B * (M + (size_t)K * M + (size_t)K * w) + (size_t)4 * (E + C) * 3 * C + (size_t)80 * E * M) * sizeof(float);
    ws.reserve(bytes, st);
  }
  ArenaScope scope(ws, st);
  LVSR_CUDA_OK(cudaMemsetAsync(grads, 0, (size_t)m->flat_count * sizeof(float), st));

  // =========================== forward, keeping the tape ===========================
  std::vector<LayerTape> tape(c.num_layers);
  {
    const float* cur = x;
    int Tl = T, din = c.num_features;
    long long mstride = B;
    for (int l = 0; l < c.num_layers; ++l) {
      const int D = c.dims_bidir[l], k = c.subsample[l], rows = Tl * B, Tout = ceil_div(Tl, k);
      LayerTape& tp = tape[l];
      tp.X = cur; tp.T = Tl; tp.Tout = Tout; tp.Din = din; tp.D = D; tp.k = k; tp.mstride = mstride;
      tp.pre = ws.f32((size_t)rows * 6 * D);
      tp.hext = ws.f32((size_t)(Tl + 2) * B * 2 * D);
      tp.out = ws.f32((size_t)Tout * B * 2 * D);
      LVSR_CHECK(tp.pre && tp.hext && tp.out, "out of device memory (encoder tape)");
      if (m->use_tc && l < (int)m->Wcat_hi.size() && m->Wcat_hi[l] && gemm_tc_supported(rows, 6 * D, din)) {
        const size_t mark = ws.off;
        float* a_hi = ws.f32((size_t)rows * gemm_tc_kpad(din));
        float* a_lo = ws.f32((size_t)rows * gemm_tc_kpad(din));
        LVSR_CHECK(a_hi && a_lo, "out of device memory (tf32 split scratch)");
        if (int rc = gemm_tc(cur, a_hi, a_lo, rows, din, m->Wcat_hi[l], m->Wcat_lo[l], 6 * D, m->bcat[l], tp.pre, 6 * D, st)) return rc;
        if (ws.off <= ws.cap) ws.off = mark;
      } else {
        if (int rc = gemm_nn(cur, rows, din, din, m->Wcat[l], 6 * D, 6 * D, m->bcat[l], tp.pre, 6 * D, false, st)) return rc;
      }
      BiGruArgs a = {};
      a.pre = tp.pre; a.mask = mask; a.mask_tstride = mstride;
      const std::string bf = enc_base(l, 0) + "/gatedrecurrent", bb = enc_base(l, 1) + "/gatedrecurrent";
      a.Wg_f = m->P(bf + ".state_to_gates"); a.Ws_f = m->P(bf + ".state_to_state"); a.h0_f = m->P(bf + ".initial_state");
      a.Wg_b = m->P(bb + ".state_to_gates"); a.Ws_b = m->P(bb + ".state_to_state"); a.h0_b = m->P(bb + ".initial_state");
      a.out = tp.out; a.T = Tl; a.B = B; a.D = D; a.subsample = k;
      a.tape = tp.pre; a.hext = tp.hext;
      if (int rc = bigru_layer(a, st)) return rc;
      cur = tp.out; Tl = Tout; din = 2 * D; mstride *= k;
    }
  }
  const float* Hatt = tape.back().out;                       // attended [Tp, B, E]
  float* attm = ws.f32((size_t)Tp * B);
  LVSR_CHECK(attm, "out of device memory");
  if (mask) {
    int kcum = 1;
    for (int l = 0; l < c.num_layers; ++l) kcum *= c.subsample[l];
    if (int rc = gather_time_subsample(attm, mask, Tp, kcum, B, st)) return rc;
  } else {
    if (int rc = fill_f32(attm, (long long)Tp * B, 1.f, st)) return rc;
  }
  float* costs = ws.f32((size_t)R);
  float* W_all = ws.f32((size_t)R * Tp);        // alignments alpha_i
  float* S_prev = ws.f32((size_t)R * C);        // s_{i-1}
  float* CTX = ws.f32((size_t)R * E);           // weighted averages
  LVSR_CHECK(costs && W_all && S_prev && CTX, "out of device memory (decoder tape)");
  if (int rc = lvsr_cost_matrix(m, Hatt, attm, Tp, B, labels, lmask, L, costs, W_all, nullptr, S_prev, CTX, stream)) return rc;
  sum_all_kernel<<<1, 1024, 0, st>>>(costs, R, cost_out, gscale);
  LVSR_LAUNCH_CHECK();

  // =========================== backward ===========================
  const std::string g = GEN, t = TR, at = ATT;
  // ---- transposed weights used as right-hand sides of dY . W^T -------------------------------
  float* WoT_unused = nullptr; (void)WoT_unused;
  float* WmsT = c.use_states_for_readout ? ws.f32((size_t)Cpm * C) : nullptr;     // [Cpm, C]
  float* WmcT = ws.f32((size_t)Cpm * E);
  float* WstateT = ws.f32((size_t)C * C);
  float* WgT = ws.f32((size_t)2 * C * C);             // [2C, C]
  float* WdcatT = ws.f32((size_t)3 * C * E);          // [3C, E]
  float* WsT = ws.f32((size_t)M * C);                 // [M, C]
  float* WpT = ws.f32((size_t)M * E);                 // [M, E]
  LVSR_CHECK(WmcT && WstateT && WgT && WdcatT && WsT && WpT, "out of device memory (transposed weights)");
  if (WmsT) if (int rc = transpose(m->P(g + "/readout/merge/transform_states.W"), WmsT, C, Cpm, st)) return rc;
  if (int rc = transpose(m->P(g + "/readout/merge/transform_weighted_averages.W"), WmcT, E, Cpm, st)) return rc;
  if (int rc = transpose(m->P(t + "/transition.state_to_state"), WstateT, C, C, st)) return rc;
  if (int rc = transpose(m->P(t + "/transition.state_to_gates"), WgT, C, 2 * C, st)) return rc;
  if (int rc = transpose(m->Wd_cat, WdcatT, E, 3 * C, st)) return rc;
  // [dG (3C)] . WcombT [3C, C + E] = [ grad of s_{i-1} through the gates | grad of the glimpse ]: one product per step
  float* WcombT = ws.f32((size_t)3 * C * (C + E));
  LVSR_CHECK(WcombT, "out of device memory (transposed weights)");
  LVSR_CUDA_OK(cudaMemsetAsync(WcombT, 0, (size_t)3 * C * (C + E) * sizeof(float), st));
  if (int rc = copy2d(WcombT, C + E, WgT, C, 2 * C, C, st)) return rc;
  if (int rc = copy2d(WcombT + C, C + E, WdcatT, E, 3 * C, E, st)) return rc;
  if (int rc = transpose(m->P(at + "/state_trans/transform_states.W"), WsT, C, M, st)) return rc;
  if (int rc = transpose(m->P(at + "/preprocess.W"), WpT, E, M, st)) return rc;

  // ---- readout + emitter backward (all steps at once) ----------------------------------------
  float* merged = ws.f32((size_t)R * Cpm);
  float* hid = ws.f32((size_t)R * Hd);
  float* dlogits = ws.f32((size_t)R * V);
  float* dmerged = ws.f32((size_t)R * Cpm);
  float* dS_ro = ws.f32((size_t)R * C);
  float* dCtx_ro = ws.f32((size_t)R * E);
  LVSR_CHECK(merged && hid && dlogits && dmerged && dS_ro && dCtx_ro, "out of device memory (readout backward)");
  {
    bool acc = false;
    if (c.use_states_for_readout) {
      if (int rc = gemm_nn(S_prev, R, C, C, m->P(g + "/readout/merge/transform_states.W"), Cpm, Cpm, nullptr, merged, Cpm, false, st)) return rc;
      acc = true;
    }
    if (int rc = gemm_nn(CTX, R, E, E, m->P(g + "/readout/merge/transform_weighted_averages.W"), Cpm, Cpm, nullptr, merged, Cpm, acc, st)) return rc;
    ReadoutBwdArgs rb = {};
    rb.merged = merged; rb.b_pm = m->P(g + "/readout/post_merge/bias.b"); rb.Wo = m->P(g + "/readout/post_merge/mlp/linear_0.W");
    rb.bo = m->P(g + "/readout/post_merge/mlp/linear_0.b");
    rb.R = R; rb.Cpm = Cpm; rb.pieces = c.maxout_pieces; rb.V = V; rb.act = c.post_merge_activation;
    rb.labels = lab; rb.lmask = lmask; rb.gscale = gscale; rb.hid = hid; rb.dlogits = dlogits; rb.dmerged = dmerged;
    const size_t smem = (size_t)8 * (Hd + 128) * sizeof(float);
    LVSR_CHECK(smem <= 48 * 1024 && V <= 128, "readout backward: post_merge_dim / num_phonemes too large");
    readout_bwd_kernel<<<ceil_div(R, 8), 256, smem, st>>>(rb);
    LVSR_LAUNCH_CHECK();
    if (int rc = gemm_tn(ws, hid, Hd, dlogits, V, R, Hd, V, grad_of(m, grads, g + "/readout/post_merge/mlp/linear_0.W"), V, false, st)) return rc;
    if (int rc = colsum(dlogits, R, V, V, grad_of(m, grads, g + "/readout/post_merge/mlp/linear_0.b"), false, st)) return rc;
    if (int rc = colsum(dmerged, R, Cpm, Cpm, grad_of(m, grads, g + "/readout/post_merge/bias.b"), false, st)) return rc;
    if (int rc = gemm_tn(ws, CTX, E, dmerged, Cpm, R, E, Cpm, grad_of(m, grads, g + "/readout/merge/transform_weighted_averages.W"), Cpm, false, st)) return rc;
    if (int rc = gemm_nn(dmerged, R, Cpm, Cpm, WmcT, E, E, nullptr, dCtx_ro, E, false, st)) return rc;
    if (c.use_states_for_readout) {
      if (int rc = gemm_tn(ws, S_prev, C, dmerged, Cpm, R, C, Cpm, grad_of(m, grads, g + "/readout/merge/transform_states.W"), Cpm, false, st)) return rc;
      if (int rc = gemm_nn(dmerged, R, Cpm, Cpm, WmsT, C, C, nullptr, dS_ro, C, false, st)) return rc;
    } else {
      LVSR_CUDA_OK(cudaMemsetAsync(dS_ro, 0, (size_t)R * C * sizeof(float), st));
    }
  }

  // ---- decoder: gate values of all steps, then the reverse-time loop --------------------------
  float* G = ws.f32((size_t)R * 3 * C);          // pre-activations -> third block keeps the candidate input
  float* Z = ws.f32((size_t)R * C);
  float* Rg = ws.f32((size_t)R * C);
  float* HR = ws.f32((size_t)R * C);
  float* Cc = ws.f32((size_t)R * C);
  float* Q = ws.f32((size_t)R * M);
  float* P = ws.f32((size_t)Tp * B * M);
  float* dP = ws.f32((size_t)Tp * B * M);
  float* dG = ws.f32((size_t)R * 3 * C);         // [dGz | dGr | dA] of every step
  float* dCTX = ws.f32((size_t)R * E);
  float* dQp = ws.f32((size_t)2 * R * M);        // the two CTAs' partial dq of every step
  float* dsbuf[2] = {ws.f32((size_t)B * C), ws.f32((size_t)B * C)};
  float* keep = ws.f32((size_t)B * C);
  float* dHR = ws.f32((size_t)B * C);
  float* dspart = ws.f32((size_t)B * C);
  float* dAbuf[2] = {ws.f32((size_t)2 * B * Tp), ws.f32((size_t)2 * B * Tp)};
  float* w0 = ws.f32((size_t)B * Tp);
  const int nct = AB_CS * B;
  float* acc_v = ws.f32((size_t)nct * M);
  float* acc_Wh = ws.f32((size_t)nct * K * M);
  float* acc_filt = ws.f32((size_t)nct * K * w);
  int* win = ws.i32(2);
  float* lohi = ws.f32((size_t)2 * B);
  LVSR_CHECK(G && Z && Rg && HR && Cc && Q && P && dP && dG && dCTX && dQp && dsbuf[0] && dsbuf[1] && keep && dHR && dspart &&
                 dAbuf[0] && dAbuf[1] && w0 && acc_v && acc_Wh && acc_filt && win && lohi,
             "out of device memory (decoder backward)");
  if (int rc = lvsr_preprocess(m, Hatt, Tp, B, P, stream)) return rc;
  if (int rc = gemm_nn(CTX, R, E, E, m->Wd_cat, 3 * C, 3 * C, nullptr, G, 3 * C, false, st)) return rc;
  if (int rc = gemm_nn(S_prev, R, C, C, m->P(t + "/transition.state_to_gates"), 2 * C, 2 * C, nullptr, G, 3 * C, true, st)) return rc;
  dec_gates_kernel<<<grid1d((long long)R * 3 * C), 256, 0, st>>>(G, m->FF, lab, S_prev, R, C, Z, Rg, HR);
  LVSR_LAUNCH_CHECK();
  {
    const size_t mark = ws.off;
    float* Cpre = ws.f32((size_t)R * C);
    LVSR_CHECK(Cpre, "out of device memory");
    if (int rc = gemm_nn(HR, R, C, C, m->P(t + "/transition.state_to_state"), C, C, nullptr, Cpre, C, false, st)) return rc;
    dec_cand_kernel<<<grid1d((long long)R * C), 256, 0, st>>>(Cpre, G, R, C, Cc);
    LVSR_LAUNCH_CHECK();
    if (ws.off <= ws.cap) ws.off = mark;
  }
  if (int rc = gemm_nn(S_prev, R, C, C, m->P(at + "/state_trans/transform_states.W"), M, M, nullptr, Q, M, false, st)) return rc;
  LVSR_CUDA_OK(cudaMemsetAsync(dP, 0, (size_t)Tp * B * M * sizeof(float), st));
  LVSR_CUDA_OK(cudaMemsetAsync(acc_v, 0, (size_t)nct * M * sizeof(float), st));
  LVSR_CUDA_OK(cudaMemsetAsync(acc_Wh, 0, (size_t)nct * K * M * sizeof(float), st));
  LVSR_CUDA_OK(cudaMemsetAsync(acc_filt, 0, (size_t)nct * K * w * sizeof(float), st));
  LVSR_CUDA_OK(cudaMemsetAsync(dsbuf[0], 0, (size_t)B * C * sizeof(float), st));
  if (int rc = onehot_rows(w0, B, Tp, st)) return rc;
  {
    const int tc_cap = ceil_div(Tp, AB_CS);
    const size_t ab_smem = att_bwd_smem_floats(M, E, K, n, tc_cap) * sizeof(float);
    LVSR_CHECK(ab_smem <= 227 * 1024 && M <= AB_NT && M % 128 == 0 && K <= 16 && E % 4 == 0,
               "attention backward: shape unsupported (Tp=%d M=%d)", Tp, M);
    const bool kp12 = att_bwd_kp(K) == 12;
    LVSR_CUDA_OK(cudaFuncSetAttribute(att_bwd_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ab_smem));
    LVSR_CUDA_OK(cudaFuncSetAttribute(att_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ab_smem));
    const int ew = ceil_div(B * C, 256);
    for (int i = L - 1; i >= 0; --i) {
      ProfScope prof("dec_bwd_step", st);
      const float* ds = dsbuf[(L - 1 - i) & 1];
      float* ds_next = dsbuf[(L - i) & 1];
      const float* lm_i = lmask ? lmask + (size_t)i * B : nullptr;
      float* dG_i = dG + (size_t)i * B * 3 * C;
      const float* Sp_i = S_prev + (size_t)i * B * C;
      dec_bwd_a_kernel<<<ew, 256, 0, st>>>(ds, Z + (size_t)i * B * C, Cc + (size_t)i * B * C, Sp_i, lm_i, B, C, dG_i, keep);
      LVSR_LAUNCH_CHECK();
      if (int rc = skinny(dG_i + 2 * C, C, 3 * C, WstateT, nullptr, 0, 0, nullptr, nullptr, 0, nullptr, 0, dHR, C, B, C, st)) return rc;
      dec_bwd_b_kernel<<<ew, 256, 0, st>>>(dHR, Rg + (size_t)i * B * C, Sp_i, B, C, dG_i, keep);
      LVSR_LAUNCH_CHECK();
      // grad of s_{i-1} through the gates (+ the element-wise paths), grad of the glimpse
      float* dctx_i = dCTX + (size_t)i * B * E;
      if (int rc = skinny(dG_i, 3 * C, 3 * C, WcombT, nullptr, 0, 0, nullptr, keep, C, dCtx_ro + (size_t)i * B * E, E, dspart, C, B, C + E, st,
                          C, dctx_i, E)) return rc;
      // attention backward
      const float* w_prev = i == 0 ? w0 : W_all + (size_t)(i - 1) * B * Tp;
      WindowArgs wa = {};
      wa.weights = w_prev; wa.step = nullptr; wa.step_offset = i; wa.R = B; wa.Tp = Tp; wa.prior = prior_of(c); wa.win = win; wa.lohi = lohi;
      if (int rc = attention_window(wa, st)) return rc;
      AttBwdArgs ab = {};
      ab.P = P; ab.H = Hatt; ab.maskH = attm; ab.q = Q + (size_t)i * B * M; ab.w_prev = w_prev; ab.w_cur = W_all + (size_t)i * B * Tp;
      ab.ctx = CTX + (size_t)i * B * E; ab.dctx = dctx_i;
      ab.dA_in = (i == L - 1) ? nullptr : dAbuf[(L - 1 - i) & 1];
      ab.win = win;
      ab.filt = m->P(at + "/conv1d.filters"); ab.Wh = m->P(at + "/handler.W"); ab.v = m->P(at + "/energy_comp/linear.W");
      ab.dP = dP; ab.dq_part = dQp + (size_t)i * 2 * B * M; ab.dA_out = dAbuf[(L - i) & 1];
      ab.acc_v = acc_v; ab.acc_Wh = acc_Wh; ab.acc_filt = acc_filt;
      ab.B = B; ab.Tp = Tp; ab.M = M; ab.E = E; ab.K = K; ab.n = n;
      {
        ProfScope prof_ab("att_bwd", st);
        if (kp12) att_bwd_kernel<12><<<nct, AB_NT, ab_smem, st>>>(ab, tc_cap);
        else att_bwd_kernel<16><<<nct, AB_NT, ab_smem, st>>>(ab, tc_cap);
        LVSR_LAUNCH_CHECK();
      }
      // ds_{i-1} = gates/elementwise part + dq . W_s^T (two partials) + readout of step i (which saw s_{i-1})
      const float* q0 = dQp + (size_t)i * 2 * B * M;
      if (int rc = skinny(q0, M, M, WsT, q0 + (size_t)B * M, M, M, WsT, dspart, C, dS_ro + (size_t)i * B * C, C, ds_next, C, B, C, st)) return rc;
    }
  }
  const float* ds_init = dsbuf[L & 1];                    // gradient of the broadcast initial state, per row
  if (int rc = colsum(ds_init, B, C, C, grad_of(m, grads, t + "/transition.initial_state"), false, st)) return rc;

  // ---- decoder weight gradients: large GEMMs over all steps -----------------------------------
  {
    // state_to_state = HR^T dA ; state_to_gates = S_prev^T [dGz|dGr] ; distribute = CTX^T dG (gate columns first in Wd_cat)
    if (int rc = gemm_tn(ws, HR, C, dG + 2 * C, 3 * C, R, C, C, grad_of(m, grads, t + "/transition.state_to_state"), C, false, st)) return rc;
    if (int rc = gemm_tn(ws, S_prev, C, dG, 3 * C, R, C, 2 * C, grad_of(m, grads, t + "/transition.state_to_gates"), 2 * C, false, st)) return rc;
    if (int rc = gemm_tn(ws, CTX, E, dG, 3 * C, R, E, 2 * C, grad_of(m, grads, t + "/distribute/fork_gate_inputs.W"), 2 * C, false, st)) return rc;
    if (int rc = gemm_tn(ws, CTX, E, dG + 2 * C, 3 * C, R, E, C, grad_of(m, grads, t + "/distribute/fork_inputs.W"), C, false, st)) return rc;
    // state transformer: S_prev^T dQ (two partials)
    float* gWs = grad_of(m, grads, at + "/state_trans/transform_states.W");
    {
      // dQp is [L][2][B][M]: view partial p as rows of length M with stride 2*B*M per step -> gather into [R, M] first
      const size_t mark = ws.off;
      float* dQ = ws.f32((size_t)R * M);
      LVSR_CHECK(dQ, "out of device memory");
      for (int p = 0; p < 2; ++p) {
        // rows of step i live at dQp + (i*2 + p)*B*M: a 2-D copy with pitch 2*B*M
        LVSR_CUDA_OK(cudaMemcpy2DAsync(dQ, (size_t)B * M * sizeof(float), dQp + (size_t)p * B * M, (size_t)2 * B * M * sizeof(float),
                                       (size_t)B * M * sizeof(float), L, cudaMemcpyDeviceToDevice, st));
        if (int rc = gemm_tn(ws, S_prev, C, dQ, M, R, C, M, gWs, M, p == 1, st)) return rc;
      }
      if (ws.off <= ws.cap) ws.off = mark;
    }
    // fork(feedback(y)): dFF by label, then lookup / fork weights / biases
    const size_t mark = ws.off;
    float* dFF = ws.f32((size_t)(V + 1) * 3 * C);
    float* WffT = ws.f32((size_t)3 * C * Cfb);
    float* dlook = ws.f32((size_t)(V + 1) * Cfb);
    float* dWff = ws.f32((size_t)Cfb * 3 * C);
    float* dbff = ws.f32((size_t)3 * C);
    LVSR_CHECK(dFF && WffT && dlook && dWff && dbff, "out of device memory (feedback gradients)");
    scatter_rows_kernel<<<V + 1, 256, 0, st>>>(dG, lab, R, 3 * C, dFF);
    LVSR_LAUNCH_CHECK();
    if (c.one_of_n_feedback) {
      // FF[y] = W_fork[y, :] + b: the gradient of the fork weights IS dFF
      LVSR_CUDA_OK(cudaMemcpyAsync(dWff, dFF, (size_t)(V + 1) * 3 * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
      if (int rc = transpose(m->Wff_cat, WffT, Cfb, 3 * C, st)) return rc;
      const float* look = m->P(g + "/readout/lookupfeedback/lookuptable.W");
      if (int rc = gemm_nn(dFF, V + 1, 3 * C, 3 * C, WffT, Cfb, Cfb, nullptr, grad_of(m, grads, g + "/readout/lookupfeedback/lookuptable.W"), Cfb, false, st)) return rc;
      if (int rc = gemm_tn(ws, look, Cfb, dFF, 3 * C, V + 1, Cfb, 3 * C, dWff, 3 * C, false, st)) return rc;
    }
    if (int rc = colsum(dFF, V + 1, 3 * C, 3 * C, dbff, false, st)) return rc;
    // Wff_cat columns: [gate_inputs 2C | inputs C]
    if (int rc = copy2d(grad_of(m, grads, g + "/fork/fork_gate_inputs.W"), 2 * C, dWff, 3 * C, Cfb, 2 * C, st)) return rc;
    if (int rc = copy2d(grad_of(m, grads, g + "/fork/fork_inputs.W"), C, dWff + 2 * C, 3 * C, Cfb, C, st)) return rc;
    if (int rc = copy2d(grad_of(m, grads, g + "/fork/fork_gate_inputs.b"), 2 * C, dbff, 3 * C, 1, 2 * C, st)) return rc;
    if (int rc = copy2d(grad_of(m, grads, g + "/fork/fork_inputs.b"), C, dbff + 2 * C, 3 * C, 1, C, st)) return rc;
    (void)dlook;
    if (ws.off <= ws.cap) ws.off = mark;
    // attention constants: sums of the per-CTA partials
    reduce_partials_kernel<<<grid1d(M), 256, 0, st>>>(acc_v, nct, M, grad_of(m, grads, at + "/energy_comp/linear.W"));
    LVSR_LAUNCH_CHECK();
    reduce_partials_kernel<<<grid1d((long long)K * M), 256, 0, st>>>(acc_Wh, nct, (long long)K * M, grad_of(m, grads, at + "/handler.W"));
    LVSR_LAUNCH_CHECK();
    reduce_partials_kernel<<<grid1d((long long)K * w), 256, 0, st>>>(acc_filt, nct, (long long)K * w, grad_of(m, grads, at + "/conv1d.filters"));
    LVSR_LAUNCH_CHECK();
  }
  // ---- gradient of the attended sequence: glimpses + preprocess --------------------------------
  float* dH = ws.f32((size_t)Tp * B * E);
  LVSR_CHECK(dH, "out of device memory (dH)");
  {
    dim3 grid(ceil_div(Tp, 8), B);
    LVSR_CHECK(E <= 1024, "encoded dim %d > 1024 unsupported in training", E);
    dh_from_ctx_kernel<<<grid, 256, 0, st>>>(W_all, dCTX, L, B, Tp, E, dH, 0);
    LVSR_LAUNCH_CHECK();
    if (int rc = gemm_nn(dP, Tp * B, M, M, WpT, E, E, nullptr, dH, E, true, st)) return rc;
    if (int rc = gemm_tn(ws, Hatt, E, dP, M, Tp * B, E, M, grad_of(m, grads, at + "/preprocess.W"), M, false, st)) return rc;
    if (int rc = colsum(dP, Tp * B, M, M, grad_of(m, grads, at + "/preprocess.b"), false, st)) return rc;
  }

  // ---- encoder: reverse-time scans and their GEMMs, top layer first ----------------------------
  const float* dout = dH;
  for (int l = c.num_layers - 1; l >= 0; --l) {
    LayerTape& tp = tape[l];
    const int D = tp.D, rows = tp.T * B;
    float* hr = ws.f32((size_t)rows * 2 * D);
    float* dh0 = ws.f32((size_t)2 * B * D);
    LVSR_CHECK(hr && dh0, "out of device memory (encoder backward)");
    BiGruBwdArgs a = {};
    a.tape = tp.pre; a.hext = tp.hext; a.mask = mask; a.mask_tstride = tp.mstride; a.dout = dout;
    const std::string bf = enc_base(l, 0), bb = enc_base(l, 1);
    a.Wg_f = m->P(bf + "/gatedrecurrent.state_to_gates"); a.Ws_f = m->P(bf + "/gatedrecurrent.state_to_state");
    a.Wg_b = m->P(bb + "/gatedrecurrent.state_to_gates"); a.Ws_b = m->P(bb + "/gatedrecurrent.state_to_state");
    a.hr_out = hr; a.dh0 = dh0; a.T = tp.T; a.B = B; a.D = D; a.subsample = tp.k;
    if (int rc = bigru_layer_backward(a, st)) return rc;
    // fork: dWcat = X^T dPre, dbcat = colsum(dPre); columns per direction [inputs D | gate_inputs 2D]
    {
      const size_t mark = ws.off;
      float* dWcat = ws.f32((size_t)tp.Din * 6 * D);
      float* dbcat = ws.f32((size_t)6 * D);
      LVSR_CHECK(dWcat && dbcat, "out of device memory (fork gradients)");
      // tensor-core path: every operand transposed once into K-major tf32 hi/lo pairs (the contraction runs over the
      // T*B rows), then five split-K tcgen05 products share them; FFMA tiles for small problems / LVSR_NO_TC_GEMM
      const bool tc = m->use_tc && rows >= 2048 && D % 128 == 0;
      TcOperand dPreT, XT, hrT, hpT[2];
      if (tc) {
        if (int rc = make_tc_operand(ws, tp.pre, rows, 6 * D, 6 * D, &dPreT, st)) return rc;
        if (int rc = make_tc_operand(ws, tp.X, rows, tp.Din, tp.Din, &XT, st)) return rc;
        if (int rc = make_tc_operand(ws, hr, rows, 2 * D, 2 * D, &hrT, st)) return rc;
        for (int dir = 0; dir < 2; ++dir) {
          const float* hprev = tp.hext + (size_t)(dir ? 2 : 0) * B * 2 * D + dir * D;
          if (int rc = make_tc_operand(ws, hprev, rows, D, 2 * D, &hpT[dir], st)) return rc;
        }
        if (int rc = gemm_tn_tc(ws, XT, 0, tp.Din, dPreT, 0, 6 * D, dWcat, 6 * D, false, st)) return rc;
      } else {
        if (int rc = gemm_tn(ws, tp.X, tp.Din, tp.pre, 6 * D, rows, tp.Din, 6 * D, dWcat, 6 * D, false, st)) return rc;
      }
      if (int rc = colsum(tp.pre, rows, 6 * D, 6 * D, dbcat, false, st)) return rc;
      for (int dir = 0; dir < 2; ++dir) {
        const std::string b = enc_base(l, dir);
        const int c0 = dir * 3 * D;
        if (int rc = copy2d(grad_of(m, grads, b + "/fork/fork_inputs.W"), D, dWcat + c0, 6 * D, tp.Din, D, st)) return rc;
        if (int rc = copy2d(grad_of(m, grads, b + "/fork/fork_gate_inputs.W"), 2 * D, dWcat + c0 + D, 6 * D, tp.Din, 2 * D, st)) return rc;
        if (int rc = copy2d(grad_of(m, grads, b + "/fork/fork_inputs.b"), D, dbcat + c0, 6 * D, 1, D, st)) return rc;
        if (int rc = copy2d(grad_of(m, grads, b + "/fork/fork_gate_inputs.b"), 2 * D, dbcat + c0 + D, 6 * D, 1, 2 * D, st)) return rc;
        // recurrent weights: state_to_state = (h*r)^T dA ; state_to_gates = H_prev^T [dGz|dGr]
        float* gWs = grad_of(m, grads, b + "/gatedrecurrent.state_to_state");
        float* gWg = grad_of(m, grads, b + "/gatedrecurrent.state_to_gates");
        if (tc) {
          if (int rc = gemm_tn_tc(ws, hrT, dir * D, D, dPreT, c0, D, gWs, D, false, st)) return rc;
          if (int rc = gemm_tn_tc(ws, hpT[dir], 0, D, dPreT, c0 + D, 2 * D, gWg, 2 * D, false, st)) return rc;
        } else {
          if (int rc = gemm_tn(ws, hr + dir * D, 2 * D, tp.pre + c0, 6 * D, rows, D, D, gWs, D, false, st)) return rc;
          const float* hprev = tp.hext + (size_t)(dir ? 2 : 0) * B * 2 * D + dir * D;     // slot t (forward) / t+2 (backward)
          if (int rc = gemm_tn(ws, hprev, 2 * D, tp.pre + c0 + D, 6 * D, rows, D, 2 * D, gWg, 2 * D, false, st)) return rc;
        }
        if (int rc = colsum(dh0 + (size_t)dir * B * D, B, D, D, grad_of(m, grads, b + "/gatedrecurrent.initial_state"), false, st)) return rc;
      }
      if (ws.off <= ws.cap) ws.off = mark;
    }
    // gradient of the layer input = gradient of the (subsampled) output of the layer below
    if (l > 0) {
      float* dX = ws.f32((size_t)rows * tp.Din);
      float* WcatT = ws.f32((size_t)6 * D * tp.Din);
      LVSR_CHECK(dX && WcatT, "out of device memory (dX)");
      if (m->use_tc && gemm_tc_supported(rows, tp.Din, 6 * D) && (6 * D) % 32 == 0) {
        // dX = dPre . Wcat^T: the K-major form of the right-hand side [N = Din, K = 6D] is Wcat itself
        const size_t mark = ws.off;
        float* a_hi = ws.f32((size_t)rows * 6 * D);
        float* a_lo = ws.f32((size_t)rows * 6 * D);
        float* w_hi = ws.f32((size_t)tp.Din * 6 * D);
        float* w_lo = ws.f32((size_t)tp.Din * 6 * D);
        LVSR_CHECK(a_hi && a_lo && w_hi && w_lo, "out of device memory (dX operands)");
        if (int rc = split_tf32(m->Wcat[l], w_hi, w_lo, (long long)tp.Din * 6 * D, st)) return rc;
        if (int rc = gemm_tc(tp.pre, a_hi, a_lo, rows, 6 * D, w_hi, w_lo, tp.Din, nullptr, dX, tp.Din, st)) return rc;
        if (ws.off <= ws.cap) ws.off = mark;
      } else {
        if (int rc = transpose(m->Wcat[l], WcatT, tp.Din, 6 * D, st)) return rc;
        if (int rc = gemm_nn(tp.pre, rows, 6 * D, 6 * D, WcatT, tp.Din, tp.Din, nullptr, dX, tp.Din, false, st)) return rc;
      }
      dout = dX;
    }
  }
  return 0;
}


static bool is_weight_name(const std::string& name) {
  const std::string leaf = name.substr(name.rfind('.') + 1);
  return leaf == "W" || leaf == "state_to_state" || leaf == "state_to_gates" || leaf == "filters";
}

int lvsr_train_apply_updates(lvsr_model* m, float* grads, float gscale, const lvsr_train_config* tc, void* stream) {
  DeviceGuard device_guard(m);
  LVSR_CHECK(m && grads && tc, "train_apply_updates: null argument");
  LVSR_CHECK(!(tc->decay_rate < 0.f || tc->decay_rate > 1.f), "decay rate needs to be in [0, 1]");   // B/algorithms/__init__.py:481-482
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = m->flat_count;
  const int np = (int)m->params.size();
  if (!m->opt_desc) {
    std::vector<ParamDesc> h(np);
    for (int i = 0; i < np; ++i) {
      h[i].offset = m->params[i].offset;
      h[i].rows = (int)m->params[i].shape[0];
      h[i].cols = (int)(m->params[i].ndim == 2 ? m->params[i].shape[1] : 1);
      h[i].is_weight = is_weight_name(m->params[i].name) ? 1 : 0;
    }
    LVSR_CUDA_OK(cudaMalloc(&m->opt_desc, sizeof(ParamDesc) * np));
    LVSR_CUDA_OK(cudaMemcpy(m->opt_desc, h.data(), sizeof(ParamDesc) * np, cudaMemcpyHostToDevice));
    LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&m->opt_scratch), 1032 * sizeof(float)));
  }
  auto lazy = [&](float** p) -> int {
    if (*p) return 0;
    LVSR_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), (size_t)n * sizeof(float)));
    LVSR_CUDA_OK(cudaMemsetAsync(*p, 0, (size_t)n * sizeof(float), st));
    return 0;
  };
  if (tc->use_momentum) if (int rc = lazy(&m->opt_velocity)) return rc;
  if (tc->use_adadelta) { if (int rc = lazy(&m->opt_ms_step)) return rc; if (int rc = lazy(&m->opt_ms_dx)) return rc; }
  const ParamDesc* desc = static_cast<const ParamDesc*>(m->opt_desc);
  if (tc->decay > 0.f) {
    dim3 grid(64, np);
    add_decay_kernel<<<grid, 256, 0, st>>>(grads, m->flat, desc, np, 2.f * tc->decay / gscale);
    LVSR_LAUNCH_CHECK();
  }
  float* part = m->opt_scratch;
  float* norm = m->opt_scratch + 1024;
  const int nparts = (int)std::min<long long>(1024, std::max<long long>(1, n / 4096));
  sqnorm_partial_kernel<<<nparts, 256, 0, st>>>(grads, n, part);
  LVSR_LAUNCH_CHECK();
  sqnorm_final_kernel<<<1, 32, 0, st>>>(part, nparts, gscale, norm);
  LVSR_LAUNCH_CHECK();
  StepArgs a = {};
  a.grads = grads; a.params = m->flat; a.velocity = m->opt_velocity; a.ms_step = m->opt_ms_step; a.ms_dx = m->opt_ms_dx;
  a.norm = norm; a.n = n; a.gscale = gscale; a.decay = tc->decay; a.threshold = tc->gradient_threshold;
  a.use_momentum = tc->use_momentum; a.learning_rate = tc->scale; a.momentum = tc->momentum;
  a.use_adadelta = tc->use_adadelta; a.decay_rate = tc->decay_rate; a.epsilon = tc->epsilon;
  step_rules_kernel<<<grid1d(n, 256, 1184), 256, 0, st>>>(a);
  LVSR_LAUNCH_CHECK();
  if (tc->max_norm > 0.f) {
    dim3 grid(32, np);
    max_norm_kernel<<<grid, 256, 0, st>>>(grads, m->flat, desc, tc->max_norm);
    LVSR_LAUNCH_CHECK();
  }
  float burn_mult = 1.f;
  if (tc->burn_in_steps > 0) {                         // lvsr/algorithms.py:35-43
    if (m->burn_in_left < 0) m->burn_in_left = tc->burn_in_steps;
    burn_mult = m->burn_in_left <= 0 ? 1.f : 0.f;
    m->burn_in_left = std::max<long long>(0, m->burn_in_left - 1);
  }
  apply_update_kernel<<<np, 256, 0, st>>>(m->flat, grads, desc, burn_mult);
  LVSR_LAUNCH_CHECK();
  return finalize_on_stream(m, st, false);             // re-pack the kernel-side weights from the new parameters
}

int lvsr_train_gradient_norm(lvsr_model* m, float* norm_host) {
  LVSR_CHECK(m && norm_host && m->opt_scratch, "train_gradient_norm: no update has run yet");
  DeviceGuard device_guard(m);
  LVSR_CUDA_OK(cudaMemcpy(norm_host, m->opt_scratch + 1024, sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int lvsr_train_reset(lvsr_model* m) {
  LVSR_CHECK(m, "null model");
  DeviceGuard device_guard(m);
  const size_t bytes = (size_t)m->flat_count * sizeof(float);
  if (m->opt_velocity) LVSR_CUDA_OK(cudaMemset(m->opt_velocity, 0, bytes));
  if (m->opt_ms_step) LVSR_CUDA_OK(cudaMemset(m->opt_ms_step, 0, bytes));
  if (m->opt_ms_dx) LVSR_CUDA_OK(cudaMemset(m->opt_ms_dx, 0, bytes));
  m->burn_in_left = -1;
  return 0;
}

}  // extern "C"
