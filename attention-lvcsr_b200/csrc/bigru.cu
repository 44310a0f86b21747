// Bidirectional GatedRecurrent scan of one encoder layer -- one persistent,
// cluster-resident kernel for BOTH directions of the layer.
//
// Replaces theano.scan over GatedRecurrent.apply for the forward and the backward
// child of Bidirectional (B/bricks/recurrent.py:224-231, 608-620, 655-663) plus the
// x[::k] subsampling of Encoder.apply (lvsr/bricks/__init__.py:75-77).
//
// B200 mapping (the recurrence is latency-bound: two DEPENDENT [rows,D]x[D,*]
// products per step, T sequential steps):
//   * batch rows are independent -> a thread-block CLUSTER of CS CTAs owns RB rows of
//     one direction; clusters never talk to each other (no grid-wide barrier).
//   * inside a cluster the hidden units are split: CTA `rank` owns UC = D/CS units and
//     keeps its slice of state_to_gates / state_to_state IN REGISTERS for the whole
//     sequence (D=256: 96 weights per thread) -- weights are read from HBM once.
//   * per step: gates for the owned units (needs all of h), all-gather of h*r through
//     distributed shared memory, candidate for the owned units, all-gather of h'.
//     Two cluster barriers per step; h never leaves the chip.
//   * each warp splits K over its 32 lanes and finishes with a halving
//     reduce-scatter (62 shuffles instead of 320 butterflies for 64 partial sums).
//   * the fork pre-activations of step t are prefetched into registers one step ahead.
#include "kernels.h"

namespace lvsr {

namespace {

constexpr int RB = 8;        // batch rows per cluster
constexpr int NWARP = 8;     // warps per CTA

__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void st_cluster_f32(uint32_t local_addr, int rank, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(remote), "f"(v) : "memory");
}

// D: hidden units per direction; CS: CTAs per cluster.
template <int D, int CS>
__global__ void __launch_bounds__(NWARP * 32, 1)
bigru_kernel(BiGruArgs a) {
  constexpr int UC = D / CS;          // units owned by this CTA
  constexpr int KPL = D / 32;         // k values per lane
  constexpr int NC2 = UC / NWARP;     // units per warp (candidate columns)
  constexpr int NC1 = 2 * NC2;        // gate columns per warp: [z units | r units]
  static_assert(D % (CS * NWARP) == 0 && D % 32 == 0, "unsupported D / cluster size");
  static_assert(KPL % 4 == 0, "KPL must allow float4 loads");
  constexpr int N1 = RB * NC1, N2 = RB * NC2;

  __shared__ __align__(16) float hbuf[RB][D];    // current state, all units
  __shared__ __align__(16) float hrbuf[RB][D];   // h * reset, all units
  __shared__ float zbuf[RB][UC];                 // update gates of the owned units

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cluster_id = blockIdx.x / CS;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  const int dir = cluster_id & 1;             // 0 forward, 1 backward
  const int row0 = (cluster_id >> 1) * RB;    // first batch row of this cluster
  const int u_warp = rank * UC + warp * NC2;  // first unit owned by this warp

  const float* Wg = dir ? a.Wg_b : a.Wg_f;    // [D, 2D]  cols [update | reset]
  const float* Ws = dir ? a.Ws_b : a.Ws_f;    // [D, D]
  const float* h0 = dir ? a.h0_b : a.h0_f;    // [D]

  // ---- weights -> registers (once) -------------------------------------------------
  float w1[NC1][KPL], w2[NC2][KPL];
#pragma unroll
  for (int c = 0; c < NC1; ++c) {
    const int col = (c < NC2) ? (u_warp + c) : (D + u_warp + (c - NC2));
#pragma unroll
    for (int kk = 0; kk < KPL; ++kk) w1[c][kk] = Wg[(long long)(lane * KPL + kk) * (2 * D) + col];
  }
#pragma unroll
  for (int c = 0; c < NC2; ++c)
#pragma unroll
    for (int kk = 0; kk < KPL; ++kk) w2[c][kk] = Ws[(long long)(lane * KPL + kk) * D + u_warp + c];

  for (int i = tid; i < RB * D; i += NWARP * 32) hbuf[i / D][i % D] = h0[i % D];

  // which reduced outputs this lane ends up owning (static per kernel)
  // phase 1: N1 values -> N1/32 (>=1) per lane; flattened index = row * NC1 + c
  // phase 2: N2 values -> N2/32 (>=1) per lane; flattened index = row * NC2 + c
  constexpr int O1 = (N1 >= 32) ? N1 / 32 : 1;
  constexpr int O2 = (N2 >= 32) ? N2 / 32 : 1;
  constexpr int DUP1 = (N1 >= 32) ? 1 : 32 / N1;   // lanes holding the same value
  constexpr int DUP2 = (N2 >= 32) ? 1 : 32 / N2;
  const int base1 = rs_base<N1>(lane), base2 = rs_base<N2>(lane);
  const bool act1 = (lane % DUP1) == 0, act2 = (lane % DUP2) == 0;

  const int T = a.T, B = a.B;
  const long long pre_ld = 6LL * D;                       // [A | Gz | Gr] per direction
  const float* pre_dir = a.pre + (long long)dir * 3 * D;

  // prefetch registers for the current step
  float pg[O1], pa[O2], pm[O2];
  auto prefetch = [&](int t) {
#pragma unroll
    for (int j = 0; j < O1; ++j) {
      const int idx = base1 + j, row = idx / NC1, c = idx % NC1;
      const int b = row0 + row;
      const int col = (c < NC2) ? (D + u_warp + c) : (2 * D + u_warp + (c - NC2));
      pg[j] = (b < B) ? __ldg(pre_dir + ((long long)t * B + b) * pre_ld + col) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < O2; ++j) {
      const int idx = base2 + j, row = idx / NC2, c = idx % NC2;
      const int b = row0 + row;
      pa[j] = (b < B) ? __ldg(pre_dir + ((long long)t * B + b) * pre_ld + u_warp + c) : 0.f;
      pm[j] = (b < B && a.mask) ? __ldg(a.mask + (long long)t * a.mask_tstride + b) : 1.f;
    }
  };

  // all CTAs of the cluster must be resident before any remote shared-memory write
  __syncthreads();
  cluster_arrive();
  cluster_wait();

  const uint32_t hr_base = smem_u32(&hrbuf[0][0]);
  const uint32_t h_base = smem_u32(&hbuf[0][0]);

  int t = dir ? (T - 1) : 0;
  const int dt = dir ? -1 : 1;
  prefetch(t);

  for (int s = 0; s < T; ++s, t += dt) {
    float g_cur[O1], a_cur[O2], m_cur[O2];
#pragma unroll
    for (int j = 0; j < O1; ++j) g_cur[j] = pg[j];
#pragma unroll
    for (int j = 0; j < O2; ++j) { a_cur[j] = pa[j]; m_cur[j] = pm[j]; }
    if (s + 1 < T) prefetch(t + dt);

    // ---- phase 1: gates of the owned units -----------------------------------------
    float acc1[N1];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float hv[KPL];
#pragma unroll
      for (int q = 0; q < KPL / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(&hbuf[r][lane * KPL + q * 4]);
        hv[q * 4 + 0] = v.x; hv[q * 4 + 1] = v.y; hv[q * 4 + 2] = v.z; hv[q * 4 + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < NC1; ++c) {
        float sacc = 0.f;
#pragma unroll
        for (int kk = 0; kk < KPL; ++kk) sacc = fmaf(hv[kk], w1[c][kk], sacc);
        acc1[r * NC1 + c] = sacc;
      }
    }
    warp_reduce_scatter<N1>(acc1, lane);
    if (act1) {
#pragma unroll
      for (int j = 0; j < O1; ++j) {
        const int idx = base1 + j, row = idx / NC1, c = idx % NC1;
        const float gate = sigmoidf_acc(acc1[j] + g_cur[j]);
        if (c < NC2) {
          zbuf[row][warp * NC2 + c] = gate;
        } else {
          const int u = u_warp + (c - NC2);
          const float hr = hbuf[row][u] * gate;
          const uint32_t addr = hr_base + (uint32_t)((row * D + u) * sizeof(float));
#pragma unroll
          for (int pr = 0; pr < CS; ++pr) st_cluster_f32(addr, pr, hr);
        }
      }
    }
    cluster_arrive();
    cluster_wait();

    // ---- phase 2: candidate + blend for the owned units ----------------------------
    float acc2[N2];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float hv[KPL];
#pragma unroll
      for (int q = 0; q < KPL / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(&hrbuf[r][lane * KPL + q * 4]);
        hv[q * 4 + 0] = v.x; hv[q * 4 + 1] = v.y; hv[q * 4 + 2] = v.z; hv[q * 4 + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < NC2; ++c) {
        float sacc = 0.f;
#pragma unroll
        for (int kk = 0; kk < KPL; ++kk) sacc = fmaf(hv[kk], w2[c][kk], sacc);
        acc2[r * NC2 + c] = sacc;
      }
    }
    warp_reduce_scatter<N2>(acc2, lane);
    if (act2) {
#pragma unroll
      for (int j = 0; j < O2; ++j) {
        const int idx = base2 + j, row = idx / NC2, c = idx % NC2;
        const int u = u_warp + c;
        const float cand = tanhf_acc(acc2[j] + a_cur[j]);
        const float z = zbuf[row][warp * NC2 + c];
        const float hold = hbuf[row][u];
        float hn = cand * z + hold * (1.f - z);
        hn = m_cur[j] * hn + (1.f - m_cur[j]) * hold;
        const uint32_t addr = h_base + (uint32_t)((row * D + u) * sizeof(float));
#pragma unroll
        for (int pr = 0; pr < CS; ++pr) st_cluster_f32(addr, pr, hn);
        const int b = row0 + row;
        if (b < B && (t % a.subsample) == 0) {
          a.out[((long long)(t / a.subsample) * B + b) * (2 * D) + dir * D + u] = hn;
        }
      }
    }
    cluster_arrive();
    cluster_wait();
  }
}

template <int D, int CS>
int launch_bigru(const BiGruArgs& a, cudaStream_t stream) {
  const int groups = ceil_div(a.B, RB);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS * groups * 2);
  cfg.blockDim = dim3(NWARP * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LVSR_CUDA_OK(cudaLaunchKernelEx(&cfg, bigru_kernel<D, CS>, a));
  g_launch_count++;
  return 0;
}

}  // namespace

bool bigru_supported(int D) { return D == 128 || D == 256; }

int bigru_layer(const BiGruArgs& a, cudaStream_t stream) {
  ProfScope prof("bigru", stream);
  if (a.T <= 0 || a.B <= 0) return 0;
  switch (a.D) {
    case 128: return launch_bigru<128, 4>(a, stream);
    case 256: return launch_bigru<256, 8>(a, stream);
    default:
      return set_error("bigru: unsupported hidden size %d (supported: 128, 256)", a.D);
  }
}

}  // namespace lvsr
