// Internal model object behind the opaque lvsr_model handle of include/lvsr_b200.h, shared by the
// inference orchestration (api.cu) and the training step (train.cu).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "kernels.h"
#include "lvsr_b200.h"

namespace lvsr {

// Stack-style device workspace.  Top-level API calls bump-allocate from one block; when
// the block is too small the overflow is served by separate cudaMallocs and the block is
// regrown at the end of the call, so a steady-state workload never allocates.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, overflow_bytes = 0;
  int depth = 0;
  std::vector<void*> overflow;
  // the arena is rewound when a call returns while its kernels may still be in flight: safe only if the next call
  // is enqueued on the SAME stream.  A call on another stream first waits for everything the previous one enqueued.
  cudaStream_t last_stream = nullptr;
  bool used = false;
  void bind_stream(cudaStream_t st) {
    if (depth == 0 && used && st != last_stream) cudaStreamSynchronize(last_stream);
    last_stream = st;
    used = true;
  }

  void* alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (off + bytes <= cap) {
      void* p = base + off;
      off += bytes;
      return p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    overflow.push_back(p);
    overflow_bytes += bytes;
    return p;
  }
  float* f32(size_t n) { return static_cast<float*>(alloc(n * sizeof(float))); }
  long long* i64(size_t n) { return static_cast<long long*>(alloc(n * sizeof(long long))); }
  int* i32(size_t n) { return static_cast<int*>(alloc(n * sizeof(int))); }

  // Grow the block up front (only legal while nothing is allocated from it).
  void reserve(size_t bytes, cudaStream_t stream) {
    if (off != 0 || bytes <= cap) return;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return;
    if (base) cudaFree(base - shift);
    base = nullptr;
    cap = 0;
    shift = getenv("LVSR_WS_SHIFT_KB") ? (size_t)atoll(getenv("LVSR_WS_SHIFT_KB")) * 1024 : 0;   // placement experiments
    if (cudaMalloc(reinterpret_cast<void**>(&base), bytes + shift) == cudaSuccess) { cap = bytes; base += shift; }
    else cudaGetLastError();
  }
  size_t shift = 0;
  void enter() { depth++; }
  // returns non-zero on CUDA failure
  int leave(cudaStream_t stream) {
    depth--;
    if (depth > 0) return 0;
    const size_t used = off;
    off = 0;
    if (!overflow.empty()) {
      if (cudaStreamSynchronize(stream) != cudaSuccess) return 1;
      for (void* p : overflow) cudaFree(p);
      overflow.clear();
      if (base) cudaFree(base - shift);
      base = nullptr;
      const size_t want = (size_t)((used + overflow_bytes) * 1.25) + (1 << 20);
      overflow_bytes = 0;
      cap = 0;
      if (cudaMalloc(reinterpret_cast<void**>(&base), want + shift) == cudaSuccess) { cap = want; base += shift; }
      else cudaGetLastError();
    }
    return 0;
  }
  void destroy() {
    for (void* p : overflow) cudaFree(p);
    overflow.clear();
    if (base) cudaFree(base - shift);
    base = nullptr;
    cap = off = 0;
  }
};

struct Param {
  std::string name;
  int64_t shape[2];
  int ndim;
  int64_t count;
  int64_t offset;          // position in the flat parameter / gradient / optimizer-state buffers (floats)
  float* dev;              // = lvsr_model::flat + offset
};

}  // namespace lvsr

using namespace lvsr;

struct lvsr_model {
  lvsr_config cfg;
  int device = 0;                   // the GPU this handle lives on (current device at lvsr_model_create)
  int E;
  std::vector<Param> params;
  std::map<std::string, int> index;
  // ONE allocation for all parameters, each at a 256-byte aligned offset (padding stays zero): the
  // gradient buffer, the optimizer state and the all-reduce of the training step use the same layout
  float* flat = nullptr;
  int64_t flat_count = 0;
  // packed, kernel-side weights (rebuilt by finalize)
  std::vector<float*> Wcat, bcat;   // per encoder layer: [Din, 6D], [6D]
  float* Wd_cat = nullptr;          // [E, 3C] = [distribute gate_inputs (2C) | distribute inputs (C)]
  float* Wb1 = nullptr;             // [E+C, 3C] = Wd_cat stacked on [state_to_gates | 0] (persistent decoder)
  float* Wff_cat = nullptr;         // [Cfb, 3C] = [fork gate_inputs | fork inputs]
  float* bff_cat = nullptr;         // [3C]
  float* FF = nullptr;              // [(V+1), 3C] = lookup . Wff_cat + bff_cat
  // K-major tf32 hi/lo splits of the dense-projection weights (tcgen05 path); null = SIMT path
  std::vector<float*> Wcat_hi, Wcat_lo;
  float *Wp_hi = nullptr, *Wp_lo = nullptr;
  bool use_tc = true;
  // fp16 head/tail splits of the same weights for the inference projections whose input is a BiGRU output (layers >= 1,
  // preprocess): K-major [N, Kpad64] halfs + {scale, 1/scale} on the device; null = tf32 path
  std::vector<void*> Wcat_h16_head, Wcat_h16_tail;
  std::vector<float*> Wcat_h16_scale;
  void *Wp_h16_head = nullptr, *Wp_h16_tail = nullptr;
  float* Wp_h16_scale = nullptr;
  bool use_h16 = true;
  bool h16_stale = true;            // the fp16 operands do not match the parameters (re-split at the next use)
  float v_bias = 0.f;               // host copy of energy_comp/linear.b
  unsigned* status = nullptr;       // device word: launch status of the data-flow decoder (common.cuh LVSR_FLOW_*)
  bool force_stepwise = false;      // set while a failed persistent launch is re-run on the step-wise kernels
  long long dec_fallbacks = 0;      // how often that happened
  bool finalized = false;
  Arena ws;
  // ---- training (train.cu) ----
  Arena tws;                        // tape + backward workspace
  float *opt_velocity = nullptr, *opt_ms_step = nullptr, *opt_ms_dx = nullptr;   // flat layout, allocated on first use
  float* opt_scratch = nullptr;     // [1024 partial sums | norm]
  void* opt_desc = nullptr;         // device copy of the per-parameter table (train::ParamDesc)
  long long burn_in_left = -1;      // BurnIn counter (-1: not started)

  float* P(const std::string& n) const {
    auto it = index.find(n);
    return it == index.end() ? nullptr : params[it->second].dev;
  }
};


namespace lvsr {

static const char* const GEN = "/recognizer/generator";
static const char* const TR = "/recognizer/generator/att_trans";
static const char* const ATT = "/recognizer/generator/att_trans/conv_att";

static inline std::string enc_base(int l, int dir) {
  char buf[128];
  snprintf(buf, sizeof(buf), "/recognizer/encoder/bidir%d/%s", l, dir ? "backward" : "forward");
  return buf;
}

static inline PriorParams prior_of(const lvsr_config& c) {
  PriorParams p;
  p.type = c.prior_type;
  p.initial_begin = c.prior_initial_begin;
  p.initial_end = c.prior_initial_end;
  p.min_speed = c.prior_min_speed;
  p.max_speed = c.prior_max_speed;
  p.before = c.prior_before;
  p.after = c.prior_after;
  return p;
}

// Every entry point runs on the handle's own GPU, whatever device the calling thread has current
// (a handle is bound to the device that was current at lvsr_model_create).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(const lvsr_model* m) {
    if (!m) return;
    int cur = 0;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != m->device) {
      prev = cur;
      cudaSetDevice(m->device);
    }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct ArenaScope {
  Arena& ws;
  cudaStream_t st;
  ArenaScope(lvsr_model* mm, cudaStream_t s) : ws(mm->ws), st(s) { ws.bind_stream(s); ws.enter(); }
  ArenaScope(Arena& a, cudaStream_t s) : ws(a), st(s) { ws.bind_stream(s); ws.enter(); }
  ~ArenaScope() { ws.leave(st); }
};

static inline int check_ready(lvsr_model* m) {
  LVSR_CHECK(m != nullptr, "null model");
  if (!m->finalized) return lvsr_model_finalize(m);
  return 0;
}

// shared orchestration pieces (api.cu)
int finalize_on_stream(lvsr_model* m, cudaStream_t st, bool synchronise);
int readout_merged(lvsr_model* m, int R, const float* states, const float* ctx, float* merged, cudaStream_t st);
ReadoutArgs readout_args(lvsr_model* m, int R, const float* merged);
size_t encoder_ws_bytes(const lvsr_model* m, int T, int B);
size_t cost_ws_bytes(const lvsr_model* m, int Tp, int B, int L);

}  // namespace lvsr
