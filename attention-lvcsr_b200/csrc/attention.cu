// Content + location attention step: one launch = take_glimpses for R decoder rows.
//
// Replaces, for every row, SequenceContentAndConvAttention.take_glimpses
// (lvsr/bricks/attention.py:120-183): window cut, Conv1D of the previous alignment
// (lvsr/expressions.py:28-54, true convolution, centre crop :108-111), energies
// (:98-114), normaliser (:191-213), weighted average (B/bricks/attention.py:235-256)
// and the paste back into [R, T'] (:177-181).
//
// B200 mapping: a row's P/H slices (T' x (M+E) floats, ~1 MB) are streamed once per
// step; a thread-block CLUSTER of `cs` CTAs splits the window of one row along time,
// each CTA produces a local (max, sum, weighted partial context) triple and the
// cluster combines them through distributed shared memory (online-softmax merge) --
// one cluster barrier per step, no global round trip.  Warp-shuffle reductions for
// the energy dot products and the softmax statistics.
#include "kernels.h"
#include "lvsr_b200.h"

namespace lvsr {

namespace {

constexpr int ATT_THREADS = 256;
constexpr int TT = 4;   // time positions per warp pass in the energy phase
constexpr int CT = 4;   // time positions per thread task in the conv phase

// ---------------------------------------------------------------------------------
// window kernel
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) window_kernel(WindowArgs a) {
  __shared__ float s_lo[8], s_hi[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tp = a.Tp;
  if (a.prior.type == LVSR_PRIOR_EXPANDING) {
    // lvsr/bricks/attention.py:127-132 -- step[0] decides for the whole batch
    if (tid == 0) {
      const double st = (double)((a.step ? a.step[0] : 0LL) + a.step_offset);
      double begin = a.prior.initial_begin + st * a.prior.min_speed;
      double end = a.prior.initial_end + st * a.prior.max_speed;
      begin = fmax(0.0, fmin((double)(Tp - 1), begin));
      end = fmax(0.0, fmin((double)Tp, end));
      a.win[0] = (int)floor(begin);
      a.win[1] = (int)ceil(end);
    }
    for (int r = tid; r < a.R; r += blockDim.x) {
      a.lohi[2 * r] = -1e30f;
      a.lohi[2 * r + 1] = 1e30f;
    }
    return;
  }
  float my_lo = 1e30f, my_hi = -1e30f;
  const int chunk = (Tp + 31) / 32;
  for (int r = warp; r < a.R; r += 8) {
    const float* w = a.weights + (long long)r * Tp;
    const int i0 = lane * chunk, i1 = min(Tp, i0 + chunk);
    double pos;
    if (a.prior.type == LVSR_PRIOR_WINDOW_MEAN) {
      double acc = 0.0;                                   // :136-137
      for (int i = i0; i < i1; ++i) acc += (double)w[i] * (double)i;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      pos = acc;
    } else {
      // window_around_median, :138-144: first index j* with cumsum >= 0.5; the reference's
      // argmax of the shifted difference gives j* - 1 (0 when j* == 0 or no crossing)
      double part = 0.0;
      for (int i = i0; i < i1; ++i) part += (double)w[i];
      double incl = part;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        double n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      double run = incl - part;
      int cross = 0x7fffffff;
      for (int i = i0; i < i1; ++i) {
        run += (double)w[i];
        if (run - 0.5 >= 0.0) { cross = i; break; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cross = min(cross, __shfl_xor_sync(0xffffffffu, cross, o));
      pos = (cross == 0x7fffffff || cross == 0) ? 0.0 : (double)(cross - 1);
    }
    const float lo = (float)floor(pos - a.prior.before);   // :146-147
    const float hi = (float)ceil(pos + a.prior.after);
    if (lane == 0) {
      a.lohi[2 * r] = lo;
      a.lohi[2 * r + 1] = hi;
    }
    my_lo = fminf(my_lo, lo);
    my_hi = fmaxf(my_hi, hi);
  }
  if (lane == 0) { s_lo[warp] = my_lo; s_hi[warp] = my_hi; }
  __syncthreads();
  if (tid == 0) {
    float lo = s_lo[0], hi = s_hi[0];
    for (int i = 1; i < 8; ++i) { lo = fminf(lo, s_lo[i]); hi = fmaxf(hi, s_hi[i]); }
    a.win[0] = (int)fmaxf(0.f, lo);                         // :149-150
    a.win[1] = (int)fminf((float)Tp, hi);
  }
}

// ---------------------------------------------------------------------------------
// attention step kernel
// ---------------------------------------------------------------------------------
struct AttSmem {
  float *sq, *sv, *sWh, *sfilt, *salpha, *sF, *se, *su, *sred, *xs, *xctx;
};

__host__ __device__ inline size_t att_smem_floats(int M, int E, int K, int n, int tc_cap, int cs, int KP) {
  size_t f = 0;
  f += M;                     // sq
  f += M;                     // sv
  f += (size_t)K * M;         // sWh
  f += (size_t)K * (2 * n + 1);  // sfilt
  f += tc_cap + 2 * n + 8;    // salpha
  f += (size_t)(tc_cap + TT) * KP;  // sF
  f += tc_cap + TT;           // se
  f += tc_cap + TT;           // su
  f += (size_t)(ATT_THREADS / 32) * 4 + 8;   // block reduction scratch
  f += (size_t)8 * E;         // sred: up to 8 column groups of partial context  (reused for xchg source)
  f += (size_t)cs * 4;        // xs: per-rank scalars (lmax, lsum, anyone, pad)
  f += (size_t)cs * E;        // xctx: per-rank partial context (meaningful on rank 0)
  return f + 16;
}

template <int KP>
__global__ void __launch_bounds__(ATT_THREADS, 2) att_step_kernel(AttStepArgs a, int tc_cap) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int row = blockIdx.x / cs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = a.M, E = a.E, K = a.K, n = a.n, w = 2 * n + 1, Tp = a.Tp, U = a.U;
  const int u = a.row_utt ? a.row_utt[row] : row;

  // arrive now, wait just before the first remote write: guarantees every CTA of the
  // cluster is resident without stalling the prologue
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");

  // carve shared memory
  float* p = smem;
  float* sq = p; p += M;
  float* sv = p; p += M;
  float* sWh = p; p += (size_t)K * M;
  float* sfilt = p; p += (size_t)K * w;
  p += (4 - ((p - smem) & 3)) & 3;
  float* salpha = p; p += tc_cap + 2 * n + 8;
  float* sF = p; p += (size_t)(tc_cap + TT) * KP;
  float* se = p; p += tc_cap + TT;
  float* su = p; p += tc_cap + TT;
  float* sblk = p; p += (ATT_THREADS / 32) * 4 + 8;
  p += (4 - ((p - smem) & 3)) & 3;
  float* sred = p; p += (size_t)8 * E;
  float* xs = p; p += (size_t)cs * 4;
  float* xctx = p; p += (size_t)cs * E;

  const int b0 = a.win[0], b1 = a.win[1];
  const int Tw = max(0, b1 - b0);
  const int tc = (Tw + cs - 1) / cs;
  const int t0 = min(Tw, rank * tc), t1 = min(Tw, t0 + tc);
  const int nt = t1 - t0;             // positions owned by this CTA (window-relative t0..t1)

  // ---- stage constants + the slice of the previous alignment --------------------
  for (int i = tid; i < M; i += ATT_THREADS) {
    sq[i] = a.q[(long long)row * M + i];
    sv[i] = a.v[i];
  }
  for (int i = tid; i < K * M; i += ATT_THREADS) sWh[i] = a.Wh[i];
  for (int i = tid; i < K * w; i += ATT_THREADS) sfilt[i] = a.filt[i];
  {
    const float* wp = a.w_prev + (long long)row * Tp;
    const int len = nt + 2 * n + 8;
    for (int i = tid; i < len; i += ATT_THREADS) {
      const int prel = t0 - n + i;            // window-relative position; zero padding is
      salpha[i] = (prel >= 0 && prel < Tw) ? wp[b0 + prel] : 0.f;   // relative to the CUT (quirk 10)
    }
  }
  __syncthreads();

  // ---- location features: F[t][k] = sum_j alpha_cut[t + n - j] * filt[k][j] ------
  {
    const int ngrp = (nt + CT - 1) / CT;
    const int ntask = ngrp * K;
    for (int task = tid; task < ntask; task += ATT_THREADS) {
      const int k = task / ngrp, tg = task % ngrp;
      const int tb = tg * CT;
      const float* f = sfilt + (size_t)k * w;
      float acc[CT];
#pragma unroll
      for (int i = 0; i < CT; ++i) acc[i] = 0.f;
      int j = 0;
      for (; j + CT <= w; j += CT) {
        // needs salpha[tb + i + 2n - j - jj], i,jj in [0,CT): offsets d = i - jj in (-CT, CT)
        const float* base = salpha + tb + 2 * n - j;
        float wv[2 * CT - 1];
#pragma unroll
        for (int d = 0; d < 2 * CT - 1; ++d) wv[d] = base[d - (CT - 1)];
#pragma unroll
        for (int jj = 0; jj < CT; ++jj) {
          const float fv = f[j + jj];
#pragma unroll
          for (int i = 0; i < CT; ++i) acc[i] = fmaf(wv[i - jj + CT - 1], fv, acc[i]);
        }
      }
      for (; j < w; ++j) {
        const float fv = f[j];
#pragma unroll
        for (int i = 0; i < CT; ++i) acc[i] = fmaf(salpha[tb + i + 2 * n - j], fv, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < CT; ++i) sF[(size_t)(tb + i) * KP + k] = acc[i];   // rows >= nt are scratch
    }
  }
  __syncthreads();

  // ---- energies: e[t] = v . tanh(P[t] + q + F[t] . Wh) ------------------------------
  {
    const int ngrp = (nt + TT - 1) / TT;
    for (int tg = warp; tg < ngrp; tg += ATT_THREADS / 32) {
      const int tb = tg * TT;
      float Fv[TT][KP];
#pragma unroll
      for (int tt = 0; tt < TT; ++tt)
#pragma unroll
        for (int k = 0; k < KP; ++k) Fv[tt][k] = (k < K) ? sF[(size_t)(tb + tt) * KP + k] : 0.f;
      float eacc[TT];
#pragma unroll
      for (int tt = 0; tt < TT; ++tt) eacc[tt] = 0.f;
      const float* prow[TT];
#pragma unroll
      for (int tt = 0; tt < TT; ++tt) {
        const int tabs = b0 + t0 + min(tb + tt, nt - 1);    // clamp: tail lanes recompute the last row
        prow[tt] = a.P + ((long long)tabs * U + u) * M;
      }
      for (int m4 = lane * 4; m4 < M; m4 += 128) {
        float4 pv[TT];
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) pv[tt] = __ldg(reinterpret_cast<const float4*>(prow[tt] + m4));
        const float4 q4 = *reinterpret_cast<const float4*>(sq + m4);
        const float4 v4 = *reinterpret_cast<const float4*>(sv + m4);
        float mt[TT][4];
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) {
          mt[tt][0] = pv[tt].x + q4.x; mt[tt][1] = pv[tt].y + q4.y;
          mt[tt][2] = pv[tt].z + q4.z; mt[tt][3] = pv[tt].w + q4.w;
        }
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          if (k < K) {
            const float4 wh = *reinterpret_cast<const float4*>(sWh + (size_t)k * M + m4);
#pragma unroll
            for (int tt = 0; tt < TT; ++tt) {
              mt[tt][0] = fmaf(Fv[tt][k], wh.x, mt[tt][0]);
              mt[tt][1] = fmaf(Fv[tt][k], wh.y, mt[tt][1]);
              mt[tt][2] = fmaf(Fv[tt][k], wh.z, mt[tt][2]);
              mt[tt][3] = fmaf(Fv[tt][k], wh.w, mt[tt][3]);
            }
          }
        }
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) {
          eacc[tt] = fmaf(v4.x, fast_tanh(mt[tt][0]), eacc[tt]);
          eacc[tt] = fmaf(v4.y, fast_tanh(mt[tt][1]), eacc[tt]);
          eacc[tt] = fmaf(v4.z, fast_tanh(mt[tt][2]), eacc[tt]);
          eacc[tt] = fmaf(v4.w, fast_tanh(mt[tt][3]), eacc[tt]);
        }
      }
#pragma unroll
      for (int tt = 0; tt < TT; ++tt) {
        const float e = warp_sum(eacc[tt]) + a.v_bias;
        if (lane == 0) se[tb + tt] = e;     // entries >= nt are scratch
      }
    }
  }
  __syncthreads();

  // ---- local normaliser statistics ---------------------------------------------------
  const float lo = a.lohi[2 * row], hi = a.lohi[2 * row + 1];
  float lmax = -INFINITY;
  if (a.normalizer == LVSR_NORM_SOFTMAX) {
    for (int t = tid; t < nt; t += ATT_THREADS) lmax = fmaxf(lmax, se[t]);   // max over masked positions too
    lmax = warp_max(lmax);
    if (lane == 0) sblk[warp] = lmax;
    __syncthreads();
    lmax = sblk[0];
#pragma unroll
    for (int i = 1; i < ATT_THREADS / 32; ++i) lmax = fmaxf(lmax, sblk[i]);
    __syncthreads();
  } else {
    lmax = 0.f;
  }
  float lsum = 0.f, anyone = 0.f;
  for (int t = tid; t < nt; t += ATT_THREADS) {
    const int tabs = b0 + t0 + t;
    const float pos = (float)tabs;
    float mval = a.maskH[(long long)tabs * U + u];
    mval *= (pos > lo && pos < hi) ? 1.f : 0.f;          // strict inequalities, :156-157
    float uv;
    const float e = se[t];
    if (a.normalizer == LVSR_NORM_SOFTMAX) uv = __expf(e - lmax);
    else if (a.normalizer == LVSR_NORM_LOGISTIC) uv = sigmoidf_acc(e);
    else uv = fmaxf(e / 1000.f, 0.f);
    uv *= mval;
    su[t] = uv;
    lsum += uv;
    if (mval == 1.f) anyone = 1.f;
  }
  lsum = warp_sum(lsum);
  anyone = warp_max(anyone);
  if (lane == 0) { sblk[8 + warp] = lsum; sblk[16 + warp] = anyone; }
  __syncthreads();
  lsum = 0.f; anyone = 0.f;
#pragma unroll
  for (int i = 0; i < ATT_THREADS / 32; ++i) { lsum += sblk[8 + i]; anyone = fmaxf(anyone, sblk[16 + i]); }

  // ---- partial weighted average with the LOCAL weights ------------------------------
  const int ncol4 = E / 4;
  const int ng = max(1, min(8, ATT_THREADS / ncol4));
  {
    const int c4 = tid % ncol4, g = tid / ncol4;
    if (g < ng) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* hbase = a.H + ((long long)(b0 + t0) * U + u) * E + c4 * 4;
      const long long hstride = (long long)U * E;
      int t = g;
      for (; t + 3 * ng < nt; t += 4 * ng) {
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(hbase + (long long)t * hstride));
        const float4 h1 = __ldg(reinterpret_cast<const float4*>(hbase + (long long)(t + ng) * hstride));
        const float4 h2 = __ldg(reinterpret_cast<const float4*>(hbase + (long long)(t + 2 * ng) * hstride));
        const float4 h3 = __ldg(reinterpret_cast<const float4*>(hbase + (long long)(t + 3 * ng) * hstride));
        const float w0 = su[t], w1 = su[t + ng], w2 = su[t + 2 * ng], w3 = su[t + 3 * ng];
        acc.x = fmaf(w0, h0.x, acc.x); acc.y = fmaf(w0, h0.y, acc.y); acc.z = fmaf(w0, h0.z, acc.z); acc.w = fmaf(w0, h0.w, acc.w);
        acc.x = fmaf(w1, h1.x, acc.x); acc.y = fmaf(w1, h1.y, acc.y); acc.z = fmaf(w1, h1.z, acc.z); acc.w = fmaf(w1, h1.w, acc.w);
        acc.x = fmaf(w2, h2.x, acc.x); acc.y = fmaf(w2, h2.y, acc.y); acc.z = fmaf(w2, h2.z, acc.z); acc.w = fmaf(w2, h2.w, acc.w);
        acc.x = fmaf(w3, h3.x, acc.x); acc.y = fmaf(w3, h3.y, acc.y); acc.z = fmaf(w3, h3.z, acc.z); acc.w = fmaf(w3, h3.w, acc.w);
      }
      for (; t < nt; t += ng) {
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(hbase + (long long)t * hstride));
        const float w0 = su[t];
        acc.x = fmaf(w0, h0.x, acc.x); acc.y = fmaf(w0, h0.y, acc.y); acc.z = fmaf(w0, h0.z, acc.z); acc.w = fmaf(w0, h0.w, acc.w);
      }
      *reinterpret_cast<float4*>(sred + (size_t)g * E + c4 * 4) = acc;
    }
  }
  __syncthreads();

  // ---- exchange through distributed shared memory ------------------------------------
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");   // pairs with the early arrive
  if (tid < cs) {
    float* dst = cluster.map_shared_rank(xs, tid);
    dst[rank * 4 + 0] = lmax;
    dst[rank * 4 + 1] = lsum;
    dst[rank * 4 + 2] = anyone;
  }
  {
    float* dst0 = cluster.map_shared_rank(xctx, 0);
    for (int e = tid; e < E; e += ATT_THREADS) {
      float s = 0.f;
      for (int g = 0; g < ng; ++g) s += sred[(size_t)g * E + e];
      dst0[(size_t)rank * E + e] = s;
    }
  }
  cluster.sync();

  // ---- combine ---------------------------------------------------------------------
  float gmax = -INFINITY;
  for (int r = 0; r < cs; ++r) gmax = fmaxf(gmax, xs[r * 4 + 0]);
  float gsum = 0.f, gany = 0.f, myscale = 0.f;
  for (int r = 0; r < cs; ++r) {
    const float ls = xs[r * 4 + 1];
    float sc = 0.f;
    if (a.normalizer == LVSR_NORM_SOFTMAX) sc = (ls > 0.f) ? __expf(xs[r * 4 + 0] - gmax) : 0.f;
    else sc = 1.f;
    gsum += sc * ls;
    gany = fmaxf(gany, xs[r * 4 + 2]);
    if (r == rank) myscale = sc;
  }
  const float norm = gsum + (gany > 0.f ? 0.f : 1.f);     // +1 when no position has mask == 1, :211-212
  const float inv = 1.f / norm;

  float* wrow = a.w_out + (long long)row * Tp;
  float* erow = a.e_out + (long long)row * Tp;
  for (int t = tid; t < nt; t += ATT_THREADS) {
    wrow[b0 + t0 + t] = su[t] * myscale * inv;
    erow[b0 + t0 + t] = se[t];
  }
  // zero outside the window (paste into zeros, :177-181); ranks interleave the work
  for (int pidx = rank * ATT_THREADS + tid; pidx < Tp; pidx += cs * ATT_THREADS) {
    if (pidx < b0 || pidx >= b0 + Tw) { wrow[pidx] = 0.f; erow[pidx] = 0.f; }
  }
  if (rank == 0) {
    for (int e = tid; e < E; e += ATT_THREADS) {
      float s = 0.f;
      for (int r = 0; r < cs; ++r) {
        float sc = 1.f;
        if (a.normalizer == LVSR_NORM_SOFTMAX) sc = (xs[r * 4 + 1] > 0.f) ? __expf(xs[r * 4 + 0] - gmax) : 0.f;
        s = fmaf(sc, xctx[(size_t)r * E + e], s);
      }
      a.ctx[(long long)row * E + e] = s * inv;
    }
  }
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int KP>
int launch_att(const AttStepArgs& a, int cs, cudaStream_t stream) {
  const int tc_cap = ceil_div(a.Tp, cs);
  const size_t smem = att_smem_floats(a.M, a.E, a.K, a.n, tc_cap, cs, KP) * sizeof(float);
  LVSR_CHECK(smem <= 227 * 1024, "attention_step: shared memory %zu B exceeds 227 KB (Tp=%d, cs=%d)", smem, a.Tp, cs);
  static size_t configured = 0;
  if (smem > configured) {
    LVSR_CUDA_OK(cudaFuncSetAttribute(att_step_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.R * cs);
  cfg.blockDim = dim3(ATT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LVSR_CUDA_OK(cudaLaunchKernelEx(&cfg, att_step_kernel<KP>, a, tc_cap));
  g_launch_count++;
  return 0;
}

}  // namespace

int attention_window(const WindowArgs& a, cudaStream_t stream) {
  ProfScope prof("window", stream);
  window_kernel<<<1, 256, 0, stream>>>(a);
  LVSR_LAUNCH_CHECK();
  return 0;
}

int attention_step(const AttStepArgs& a, cudaStream_t stream) {
  ProfScope prof("attention", stream);
  LVSR_CHECK(a.M % 4 == 0 && a.E % 4 == 0, "attention_step: dim_matcher and encoded dim must be multiples of 4");
  LVSR_CHECK(a.E / 4 <= ATT_THREADS, "attention_step: encoded dim %d > 1024 unsupported", a.E);
  LVSR_CHECK(a.K >= 1 && a.K <= 16, "attention_step: conv_num_filters %d not in [1,16]", a.K);
  if (a.R <= 0) return 0;
  // cluster size: split a row's window over as many CTAs as keeps R*cs within one wave
  int cs = 1;
  const int sms = num_sms();
  while (cs < 8 && a.R * cs * 2 <= 2 * sms && ceil_div(a.Tp, cs * 2) >= 16) cs *= 2;   // two CTAs per SM
  if (a.K == 10) return launch_att<10>(a, cs, stream);
  if (a.K <= 4) return launch_att<4>(a, cs, stream);
  if (a.K <= 8) return launch_att<8>(a, cs, stream);
  return launch_att<16>(a, cs, stream);
}

}  // namespace lvsr
