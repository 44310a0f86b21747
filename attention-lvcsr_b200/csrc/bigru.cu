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
//     sequence -- weights are read from HBM once per layer.
//   * per step: gates for the owned units (needs all of h), all-gather of h*r, candidate
//     for the owned units, all-gather of h'.  Each all-gather is CS bulk DSMEM copies
//     (cp.async.bulk shared::cta -> shared::cluster, 1 KB each) that complete on the
//     RECEIVER's mbarrier (complete_tx): no fence, no cluster barrier in the loop; the
//     consumer simply waits for RB*D*4 bytes to land.  h never leaves the chip.
//   * each warp splits K over its 32 lanes and finishes with a halving reduce-scatter.
//   * the fork pre-activations of step t+1 are prefetched into registers during step t.
// r1a -> r1b: 4-byte remote stores + barrier.cluster.arrive.release cost 30 % of the
// kernel in the fence (profiles/r1a_summary.md); replaced by the scheme above.
#include "kernels.h"

namespace lvsr {

namespace {

// Two CTAs are co-resident per SM (256 threads, <=128 registers each): they belong to
// different clusters, i.e. independent recurrences, so one chain's exchange latency is
// covered by the other chain's arithmetic.
constexpr int RB = 4;        // batch rows per cluster
constexpr int NWARP = 8;     // warps per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, int rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(local_addr), "r"(rank));
  return remote;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned long long spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1ull << 24)) __trap();   // a lost transfer must fail the launch, not hang the GPU
  }
}
// local shared memory -> shared memory of CTA `rank`, completes on that CTA's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_local, uint32_t src_local, uint32_t bytes,
                                                uint32_t bar_local, int rank) {
  const uint32_t dst = map_to_rank(dst_local, rank);
  const uint32_t bar = map_to_rank(bar_local, rank);
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
      "r"(src_local), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

// D: hidden units per direction; CS: CTAs per cluster.
//
// Work split inside a warp (the kernel is bound by instruction issue, see profiles/r1f_summary.md,
// so the layout minimises instructions that are not FFMAs): lane = kg * CG + cg.
//   kg (KL = 8 groups)  owns the k range [kg * D/8, +D/8) of both products -- exactly the slice of
//                       h that one or half a peer CTA delivers, 128 contiguous bytes per row;
//   cg (CG = 4 groups)  owns NC1/4 of the warp's gate columns and NC2/4 of its candidate columns.
// A lane therefore accumulates RB x 2 gate sums and RB x 1 candidate sums over its D/8 k values and
// the cross-lane reduction runs over the 8 kg-lanes only: 7 + 4 exchanges per step instead of
// 31 + 15 over all 32 lanes.
template <int D, int CS>
__global__ void __launch_bounds__(NWARP * 32, 2)
bigru_kernel(BiGruArgs a) {
  constexpr int UC = D / CS;          // units owned by this CTA
  constexpr int KL = 8;               // lanes that split k
  constexpr int CG = 32 / KL;         // lanes that split the warp's columns
  constexpr int KPG = D / KL;         // k values per lane
  constexpr int KQ = KPG / 4;         // float4 groups per lane
  constexpr int NC2 = UC / NWARP;     // units per warp (candidate columns)
  constexpr int NC1 = 2 * NC2;        // gate columns per warp: [z units | r units]
  constexpr int CPL1 = NC1 / CG;      // gate columns per lane
  constexpr int CPL2 = NC2 / CG;      // candidate columns per lane
  static_assert(D % (CS * NWARP) == 0 && D % (4 * KL) == 0, "unsupported D / cluster size");
  static_assert(NC2 % CG == 0 && CPL2 == 1, "one candidate column per lane");
  static_assert(UC % KPG == 0 || KPG % UC == 0, "a lane's k range must not straddle peer slices unevenly");
  static_assert(KPG <= UC, "a lane's k range sits inside one peer slice");
  constexpr int N1 = RB * CPL1, N2 = RB * CPL2;        // per-lane partial sums of the two phases
  constexpr uint32_t SLICE_BYTES = RB * UC * sizeof(float);
  constexpr uint32_t FULL_BYTES = RB * D * sizeof(float);

  // peer-major: slot p holds the [RB][UC] slice owned by CTA p -> one bulk copy per peer.  Slots
  // are padded by 16 bytes: the 8 kg lanes of a warp read 8 different slots at the same offset,
  // which without the pad is an 8-way bank conflict on every load (measured: 2x the step time).
  constexpr int SLOT = RB * UC + 4;
  __shared__ __align__(128) float hbuf[CS][SLOT];       // h, all units
  __shared__ __align__(128) float hrbuf[CS][SLOT];      // h * reset, all units
  __shared__ __align__(128) float stage_h[RB][UC];      // own slice of h (source of the copies)
  __shared__ __align__(128) float stage_hr[RB][UC];     // own slice of h * reset
  __shared__ float zbuf[RB][UC];                        // update gates of the owned units
  // state_to_state slice: [warp][q][lane][4 k] so a lane fetches four k of its column as one
  // vector; read in the candidate loop (the gate slice lives in registers for the whole sequence)
  __shared__ __align__(16) float w2s[NWARP][KQ][32][4];
  __shared__ __align__(8) unsigned long long mbar[2];   // [0]: h arrivals, [1]: h*r arrivals

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kg = lane / CG, cg = lane % CG;
  const int cluster_id = blockIdx.x / CS;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  const int dir = cluster_id & 1;             // 0 forward, 1 backward
  const int row0 = (cluster_id >> 1) * RB;    // first batch row of this cluster
  const int ul_warp = warp * NC2;             // first owned unit of this warp, local index
  const int u_warp = rank * UC + ul_warp;     // ... global unit index
  const int kpeer = (kg * KPG) / UC, koff = (kg * KPG) % UC;   // where this lane's k range sits in hbuf

  const float* Wg = dir ? a.Wg_b : a.Wg_f;    // [D, 2D]  cols [update | reset]
  const float* Ws = dir ? a.Ws_b : a.Ws_f;    // [D, D]
  const float* h0 = dir ? a.h0_b : a.h0_f;    // [D]

  // ---- weights -> registers / shared memory (once) -----------------------------------
  // gate column cl of the warp (0..NC1): cl < NC2 -> update gate of unit u_warp + cl, else reset gate
  float w1[CPL1][KPG];
#pragma unroll
  for (int j = 0; j < CPL1; ++j) {
    const int cl = cg * CPL1 + j;
    const int col = (cl < NC2) ? (u_warp + cl) : (D + u_warp + (cl - NC2));
#pragma unroll
    for (int kk = 0; kk < KPG; ++kk) w1[j][kk] = Wg[(long long)(kg * KPG + kk) * (2 * D) + col];
  }
#pragma unroll
  for (int q = 0; q < KQ; ++q)
#pragma unroll
    for (int i = 0; i < 4; ++i) w2s[warp][q][lane][i] = Ws[(long long)(kg * KPG + q * 4 + i) * D + u_warp + cg];

  for (int i = tid; i < RB * D; i += NWARP * 32) {
    const int r = i / D, u = i % D;
    hbuf[u / UC][r * UC + u % UC] = h0[u];
  }
  for (int i = tid; i < RB * UC; i += NWARP * 32) stage_h[i / UC][i % UC] = h0[rank * UC + (i % UC)];

  const uint32_t bar_h = smem_u32(&mbar[0]), bar_hr = smem_u32(&mbar[1]);
  if (tid == 0) {
    mbar_init(bar_h, 1);
    mbar_init(bar_hr, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }

  // after the reduce-scatter over the kg lanes this lane holds partial-sum index base1 / base2 of
  // its column group: phase 1 idx = row * CPL1 + j, phase 2 idx = row (duplicated over kg bit 0)
  const int base1 = rs_base<N1, CG>(lane), base2 = rs_base<N2, CG>(lane);
  const int row1 = base1 / CPL1, cl1 = cg * CPL1 + base1 % CPL1;      // gate output of this lane
  const bool is_z = cl1 < NC2;
  const int ul1 = ul_warp + (is_z ? cl1 : cl1 - NC2);
  const int row2 = base2, ul2 = ul_warp + cg;                          // candidate output of this lane
  constexpr int DUP2 = KL / N2;                                        // lanes holding the same candidate sum
  const bool act2 = DUP2 <= 1 || (kg % DUP2) == 0;
  static_assert(N1 == KL && N2 <= KL, "one gate sum per lane");

  const int T = a.T, B = a.B;
  const long long pre_ld = 6LL * D;                       // [A | Gz | Gr] per direction
  const float* pre_dir = a.pre + (long long)dir * 3 * D;
  const int dt = dir ? -1 : 1;
  int t = dir ? (T - 1) : 0;
  // per-lane read pointers into the fork pre-activations, bumped by one time step per iteration
  const bool ok1 = row0 + row1 < B, ok2 = act2 && (row0 + row2 < B);
  const float* pg_ptr = pre_dir + ((long long)t * B + row0 + row1) * pre_ld + (is_z ? D : 2 * D) + (ul1 - ul_warp) + u_warp;
  const float* pa_ptr = pre_dir + ((long long)t * B + row0 + row2) * pre_ld + u_warp + cg;
  const float* pm_ptr = a.mask ? a.mask + (long long)t * a.mask_tstride + row0 + row2 : nullptr;
  const long long pre_step = (long long)dt * B * pre_ld, mask_step = (long long)dt * a.mask_tstride;
  float pg = 0.f, pa = 0.f, pm = 1.f;
  auto prefetch = [&]() {
    pg = ok1 ? __ldg(pg_ptr) : 0.f;
    pa = ok2 ? __ldg(pa_ptr) : 0.f;
    pm = (ok2 && pm_ptr) ? __ldg(pm_ptr) : 1.f;
    pg_ptr += pre_step; pa_ptr += pre_step;
    if (pm_ptr) pm_ptr += mask_step;
  };

  // every CTA of the cluster must be resident (and its mbarriers initialised) before any
  // remote copy is issued
  __syncthreads();
  cluster_sync_all();

  const uint32_t hbuf_mine = smem_u32(&hbuf[rank][0]);     // same offset in every peer: slot `rank`
  const uint32_t hrbuf_mine = smem_u32(&hrbuf[rank][0]);
  const uint32_t stage_h_a = smem_u32(&stage_h[0][0]), stage_hr_a = smem_u32(&stage_hr[0][0]);

  int sub_phase = dir ? ((T - 1) % a.subsample) : 0;   // t % subsample, maintained incrementally
  int t_out = t / a.subsample;
  prefetch();

  for (int s = 0; s < T; ++s, t += dt) {
    const float g_cur = pg, a_cur = pa, m_cur = pm;
    if (s + 1 < T) prefetch();

    // h(s-1) from all peers has landed (step 0 uses the locally initialised h0)
    if (s > 0) mbar_wait(bar_h, (uint32_t)((s - 1) & 1));
    if (tid == 0) {
      mbar_arm(bar_h, FULL_BYTES);    // arrivals of h'(s)
      mbar_arm(bar_hr, FULL_BYTES);   // arrivals of (h*r)(s)
    }

    // ---- phase 1: gates of the owned units -----------------------------------------
    float acc1[N1];
#pragma unroll
    for (int i = 0; i < N1; ++i) acc1[i] = 0.f;
#pragma unroll
    for (int q = 0; q < KQ; ++q) {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(&hbuf[kpeer][r * UC + koff + q * 4]);
#pragma unroll
        for (int j = 0; j < CPL1; ++j) {
          float sacc = acc1[r * CPL1 + j];
          sacc = fmaf(v.x, w1[j][q * 4 + 0], sacc);
          sacc = fmaf(v.y, w1[j][q * 4 + 1], sacc);
          sacc = fmaf(v.z, w1[j][q * 4 + 2], sacc);
          sacc = fmaf(v.w, w1[j][q * 4 + 3], sacc);
          acc1[r * CPL1 + j] = sacc;
        }
      }
    }
    warp_reduce_scatter<N1, CG>(acc1, lane);
    {
      const float gate = fast_sigmoid(acc1[0] + g_cur);
      if (is_z) zbuf[row1][ul1] = gate;
      else stage_hr[row1][ul1] = stage_h[row1][ul1] * gate;
    }
    fence_async_smem();          // generic-proxy writes of stage_hr -> visible to the bulk-copy engine
    __syncthreads();
    if (warp == 0 && lane < CS) dsmem_bulk_copy(hrbuf_mine, stage_hr_a, SLICE_BYTES, bar_hr, lane);

    // ---- phase 2: candidate + blend for the owned units ----------------------------
    mbar_wait(bar_hr, (uint32_t)(s & 1));
    float acc2[N2];
#pragma unroll
    for (int i = 0; i < N2; ++i) acc2[i] = 0.f;
#pragma unroll
    for (int q = 0; q < KQ; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(&w2s[warp][q][lane][0]);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(&hrbuf[kpeer][r * UC + koff + q * 4]);
        float sacc = acc2[r];
        sacc = fmaf(v.x, w.x, sacc);
        sacc = fmaf(v.y, w.y, sacc);
        sacc = fmaf(v.z, w.z, sacc);
        sacc = fmaf(v.w, w.w, sacc);
        acc2[r] = sacc;
      }
    }
    warp_reduce_scatter<N2, CG>(acc2, lane);
    if (act2) {
      const float cand = fast_tanh(acc2[0] + a_cur);
      const float z = zbuf[row2][ul2];
      const float hold = stage_h[row2][ul2];
      float hn = cand * z + hold * (1.f - z);
      hn = m_cur * hn + (1.f - m_cur) * hold;
      stage_h[row2][ul2] = hn;
    }
    fence_async_smem();
    __syncthreads();
    if (warp == 0 && lane < CS) dsmem_bulk_copy(hbuf_mine, stage_h_a, SLICE_BYTES, bar_h, lane);
    if (sub_phase == 0 && warp == 1) {
      // coalesced store of the owned slice: RB rows x UC floats (128 B per row)
      constexpr int F4 = RB * UC / 4;
      for (int i = lane; i < F4; i += 32) {
        const int row = i / (UC / 4), c4 = i % (UC / 4);
        const int b = row0 + row;
        if (b < B) {
          const float4 v = *reinterpret_cast<const float4*>(&stage_h[row][c4 * 4]);
          *reinterpret_cast<float4*>(a.out + ((long long)t_out * B + b) * (2 * D) + dir * D + rank * UC + c4 * 4) = v;
        }
      }
    }
    // advance t % subsample and t / subsample without dividing
    if (dir == 0) {
      if (++sub_phase == a.subsample) { sub_phase = 0; ++t_out; }
    } else {
      if (sub_phase == 0) { sub_phase = a.subsample - 1; --t_out; } else { --sub_phase; }
    }
  }
  // drain: the last h' copies must have landed everywhere before any CTA may exit
  mbar_wait(bar_h, (uint32_t)((T - 1) & 1));
  cluster_sync_all();
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
