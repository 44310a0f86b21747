// Tensor-core path for the dense projections: C[M,N] = A[M,K] . W[K,N] + bias on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulator in TMEM, operands staged
// by TMA into 128B-swizzled shared memory).
//
// Replaces the whole-sequence tensor.dot of Fork(Linear) in RecurrentWithFork
// (lvsr/bricks/__init__.py:39-43) and attention.preprocess (lvsr/bricks/attention.py:228-230):
// the only genuinely dense contractions of the path (260 GFLOP per batch).
//
// Precision: the 1e-4 gate against the float64 oracle rules out single-pass tf32 (2^-11 per
// product).  Every operand is split EXACTLY into two tf32 numbers, x = hi + lo
// (hi = x with the low 13 mantissa bits cleared, lo = tf32(x - hi)), and three products
// are accumulated in fp32:  lo.hi + hi.lo + hi.hi  (the dropped lo.lo term is 2^-22).
// The split of A is one streaming pass (split_tf32_kernel); W is split once at
// lvsr_model_finalize and kept K-major ([N,K]) so both operands use the K-major SWIZZLE_128B
// canonical layout.
//
// Kernel shape: one 128x128 output tile per CTA, BK = 32 floats (one 128-byte swizzle row),
// 3-stage TMA->MMA mbarrier pipeline (4 operand tiles = 64 KB per stage), warp 0 = TMA
// producer, warp 1 = MMA issuer (single elected thread) + TMEM owner, warps 2..5 = epilogue
// (tcgen05.ld 32 lanes x 32 columns, bias add, 128-byte row stores).
#include <cuda.h>
#include <cuda_fp16.h>

#include "kernels.h"

namespace lvsr {

namespace {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32;      // TC_BN: the granularity N must be a multiple of
constexpr int TC_THREADS = 192;
constexpr uint32_t TC_TILE_BYTES = TC_BM * TC_BK * sizeof(float);          // 16 KB: one 128-row operand tile
// Two tile shapes: 128 x 128 (3 stages) and 128 x 256 (2 stages).  The wide tile moves 25 % fewer operand bytes per
// MAC -- the kernel is bound by L2 -> SM operand traffic (every value is a hi AND a lo fp32), not by the tensor pipe.
template <int BN> struct TcShape {
  static constexpr int STAGES = BN == 256 ? 2 : 3;
  static constexpr uint32_t B_TILE_BYTES = (uint32_t)BN * TC_BK * sizeof(float);
  static constexpr uint32_t STAGE_BYTES = 2 * TC_TILE_BYTES + 2 * B_TILE_BYTES;            // A_hi, A_lo, B_hi, B_lo
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  // kind::tf32, fp32 accumulate, both operands K-major, M = 128, N = BN
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  // kind::f16 with fp16 operands (a_format = b_format = 0), fp32 accumulate
  static constexpr uint32_t IDESC_H16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
};
constexpr int TC_BK_H16 = 64;      // fp16 operands: one 128-byte swizzle row holds 64 k values

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
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
    if (++spins > (1ull << 24)) __trap();   // a broken pipeline must fail the launch, not hang the GPU
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B canonical layout: rows are 128 B, 8-row groups are 1024 B apart
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: version 1, layout_type 2).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address, 16 B units
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

struct TcGemmParams {
  float* C;
  const float* bias;
  int M, N, K, ldc;
  int kb_per_split;            // K blocks handled by one blockIdx.z (split-K: partial products, summed by the caller)
  long long c_split_stride;    // elements between the partial outputs of consecutive splits
  const float* out_scale;      // H16: device pointer to the factor that undoes the power-of-two weight scaling (or null)
};

// H16: operands are fp16 heads and fp16 tails scaled by 2^11 (x = head + tail / 2048); head.head goes to one TMEM
// accumulator, tail.head + head.tail to a second one, the epilogue adds main + cross / 2048.  Same 2^-22 error class as the
// 3xTF32 split at half the operand bytes and half the tensor time (kind::f16 issues twice the MACs of kind::tf32) --
// for operands of known range only (the BiGRU outputs, |h| <= 1, against weights scaled below 2^14).
template <int BN, bool H16>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B needs 1024-byte aligned tiles
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int TC_STAGES = TcShape<BN>::STAGES;
  constexpr uint32_t TC_STAGE_BYTES = TcShape<BN>::STAGE_BYTES, B_TILE = TcShape<BN>::B_TILE_BYTES;
  constexpr uint32_t TC_IDESC = H16 ? TcShape<BN>::IDESC_H16 : TcShape<BN>::IDESC;
  constexpr int BKE = H16 ? TC_BK_H16 : TC_BK;          // k values per 128-byte row
  constexpr uint32_t TMEM_COLS = H16 ? 2 * BN : BN;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(tiles + (size_t)TC_STAGES * TC_STAGE_BYTES);
  // bars[0..S): full, bars[S..2S): empty, bars[2S]: accumulator ready; then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TC_BM;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int nkb = min(p.kb_per_split, p.K / BKE - kb0);
  p.C += (long long)blockIdx.z * p.c_split_stride;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      bar_init(smem_addr(&bars[s]), 1);
      bar_init(smem_addr(&bars[TC_STAGES + s]), 1);
    }
    bar_init(smem_addr(&bars[2 * TC_STAGES]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_addr(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (uint32_t)((kb / TC_STAGES) & 1);
        bar_wait(smem_addr(&bars[TC_STAGES + s]), ph ^ 1u);       // slot free (first round passes immediately)
        const uint32_t full = smem_addr(&bars[s]);
        bar_expect_tx(full, TC_STAGE_BYTES);
        const uint32_t base = smem_addr(tiles + (size_t)s * TC_STAGE_BYTES);
        tma_load_2d(base + 0 * TC_TILE_BYTES, &map_a_hi, (kb0 + kb) * BKE, m0, full);
        tma_load_2d(base + 1 * TC_TILE_BYTES, &map_a_lo, (kb0 + kb) * BKE, m0, full);
        tma_load_2d(base + 2 * TC_TILE_BYTES, &map_b_hi, (kb0 + kb) * BKE, n0, full);
        tma_load_2d(base + 2 * TC_TILE_BYTES + B_TILE, &map_b_lo, (kb0 + kb) * BKE, n0, full);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (uint32_t)((kb / TC_STAGES) & 1);
        bar_wait(smem_addr(&bars[s]), ph);
        tc_fence_after();
        const uint32_t base = smem_addr(tiles + (size_t)s * TC_STAGE_BYTES);
        const uint64_t da_hi = make_smem_desc(base + 0 * TC_TILE_BYTES), da_lo = make_smem_desc(base + 1 * TC_TILE_BYTES);
        const uint64_t db_hi = make_smem_desc(base + 2 * TC_TILE_BYTES), db_lo = make_smem_desc(base + 2 * TC_TILE_BYTES + B_TILE);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {           // UMMA K = 8 tf32 / 16 fp16 values = 32 bytes = +2 in 16-byte units
          const uint64_t adv = (uint64_t)(k * 2);
          if constexpr (H16) {
            tc_mma_f16(tmem_base + BN, da_lo + adv, db_hi + adv, TC_IDESC, (kb | k) ? 1u : 0u);   // cross terms
            tc_mma_f16(tmem_base + BN, da_hi + adv, db_lo + adv, TC_IDESC, 1u);
            tc_mma_f16(tmem_base, da_hi + adv, db_hi + adv, TC_IDESC, (kb | k) ? 1u : 0u);
          } else {
            tc_mma_tf32(tmem_base, da_lo + adv, db_hi + adv, TC_IDESC, (kb | k) ? 1u : 0u);   // small terms first
            tc_mma_tf32(tmem_base, da_hi + adv, db_lo + adv, TC_IDESC, 1u);
            tc_mma_tf32(tmem_base, da_hi + adv, db_hi + adv, TC_IDESC, 1u);
          }
        }
        tc_commit(smem_addr(&bars[TC_STAGES + s]));     // smem slot reusable once these MMAs retire
      }
      tc_commit(smem_addr(&bars[2 * TC_STAGES]));       // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global (+bias) =====
    bar_wait(smem_addr(&bars[2 * TC_STAGES]), 0);
    tc_fence_after();
    const int q = warp & 3;                              // TMEM lane quarter this warp may touch
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      if constexpr (H16) {
        uint32_t x[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]),
              "=r"(x[8]), "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15]),
              "=r"(x[16]), "=r"(x[17]), "=r"(x[18]), "=r"(x[19]), "=r"(x[20]), "=r"(x[21]), "=r"(x[22]), "=r"(x[23]),
              "=r"(x[24]), "=r"(x[25]), "=r"(x[26]), "=r"(x[27]), "=r"(x[28]), "=r"(x[29]), "=r"(x[30]), "=r"(x[31])
            : "r"(taddr + (uint32_t)BN));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        const float os = p.out_scale ? __ldg(p.out_scale) : 1.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = __float_as_uint(fmaf(__uint_as_float(x[j]), 1.f / 2048.f, __uint_as_float(r[j])) * os);
      } else {
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      }
      if (row < p.M) {
        float* crow = p.C + (long long)row * p.ldc + n0 + c;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                 __uint_as_float(r[j + 3]));
          if (p.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j));
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          }
          *reinterpret_cast<float4*>(crow + j) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// x = hi + lo with both parts exactly representable in tf32
__global__ void split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo,
                                  long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    float4 h, l;
    auto split = [](float a, float& ah, float& al) {
      ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
      const float r = a - ah;
      uint32_t t;
      asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(t) : "f"(r));
      al = __uint_as_float(t);
    };
    split(v.x, h.x, l.x); split(v.y, h.y, l.y); split(v.z, h.z, l.z); split(v.w, h.w, l.w);
    hi[i] = h;
    lo[i] = l;
  }
}

// [K, N] row-major -> K-major [N, K] hi/lo pair (weights, once per finalize)
__global__ void transpose_split_kernel(const float* __restrict__ W, float* __restrict__ hi, float* __restrict__ lo,
                                       int K, int N, int Kpad, int ldw) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? W[(long long)k * ldw + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < Kpad) {                      // k in [K, Kpad): zero padding of the contraction dimension
      const float a = tile[threadIdx.x][i];
      const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
      uint32_t t;
      asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(t) : "f"(a - ah));
      hi[(long long)n * Kpad + k] = ah;
      lo[(long long)n * Kpad + k] = __uint_as_float(t);
    }
  }
}

// split + zero-pad the contraction dimension in one pass: x [M, K] -> hi / lo [M, Kpad] (layer 0: K = 40 -> 64)
__global__ void split_pad_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                      long long M, int K, int Kpad) {
  const long long total = M * Kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Kpad;
    const int k = (int)(i % Kpad);
    const float a = k < K ? x[r * K + k] : 0.f;
    const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(t) : "f"(a - ah));
    hi[i] = ah;
    lo[i] = __uint_as_float(t);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int get_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  LVSR_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  LVSR_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

// 2-D fp32 tensor [rows, K] (K contiguous), box = [128 rows, 32 floats], 128-byte swizzle
int make_map(CUtensorMap* map, const float* ptr, long long rows, int K, int box_rows = TC_BM) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LVSR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

}  // namespace

// the contraction dimension is zero-padded to a multiple of the 32-float TMA box (layer 0: K = 40 -> 64)
int gemm_tc_kpad(int K) { return ceil_div(K, TC_BK) * TC_BK; }

bool gemm_tc_supported(int M, int N, int K) {
  return M >= 1 && N % TC_BN == 0 && K >= 4 && K % 4 == 0;
}

// Wt_hi / Wt_lo: [N, gemm_tc_kpad(K)]
int split_weight_tf32(const float* W, int K, int N, float* Wt_hi, float* Wt_lo, cudaStream_t stream) {
  const int Kpad = gemm_tc_kpad(K);
  dim3 grid(ceil_div(N, 32), ceil_div(Kpad, 32)), block(32, 8);
  transpose_split_kernel<<<grid, block, 0, stream>>>(W, Wt_hi, Wt_lo, K, N, Kpad, N);
  LVSR_LAUNCH_CHECK();
  return 0;
}

// [K rows, N columns, leading dimension ldw] -> K-major hi/lo [N, gemm_tc_kpad(K)] (operands of the TN product)
int transpose_split_tf32(const float* W, int K, int N, int ldw, float* hi, float* lo, cudaStream_t stream) {
  const int Kpad = gemm_tc_kpad(K);
  dim3 grid(ceil_div(N, 32), ceil_div(Kpad, 32)), block(32, 8);
  LVSR_CHECK(grid.y <= 65535, "transpose_split: too many rows (%d)", K);
  transpose_split_kernel<<<grid, block, 0, stream>>>(W, hi, lo, K, N, Kpad, ldw);
  LVSR_LAUNCH_CHECK();
  return 0;
}

// element-wise exact split x = hi + lo (n % 4 == 0)
int split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t stream) {
  const long long n4 = n / 4;
  split_tf32_kernel<<<(int)std::min<long long>(4096, (n4 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo), n4);
  LVSR_LAUNCH_CHECK();
  return 0;
}

int gemm_tc_splits_launched(int Kpad, int splits) {
  const int total_kb = Kpad / TC_BK;
  splits = std::max(1, std::min(splits, total_kb));
  return ceil_div(total_kb, ceil_div(total_kb, splits));
}

// C[M,N] (+ bias) = A . B^T with both operands K-major hi/lo pairs: A [M, Kpad], B [N, Kpad].  splits > 1: split-K, partial
// result z goes to C + z * split_stride (the caller adds them up).
int gemm_tc_presplit(const float* A_hi, const float* A_lo, int M, const float* B_hi, const float* B_lo, int N, int Kpad,
                     const float* bias, float* C, int ldc, int splits, long long split_stride, cudaStream_t stream) {
  ProfScope prof("gemm", stream);
  LVSR_CHECK(M >= 1 && N % TC_BN == 0 && Kpad % TC_BK == 0 && Kpad >= TC_BK && splits >= 1, "gemm_tc_presplit: unsupported shape M=%d N=%d K=%d", M, N, Kpad);
  if (int rc = get_encode()) return rc;
  const bool wide = (N % 256 == 0) && getenv("LVSR_TC_NARROW") == nullptr;
  const int BN = wide ? 256 : 128;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  if (int rc = make_map(&ma_hi, A_hi, M, Kpad)) return rc;
  if (int rc = make_map(&ma_lo, A_lo, M, Kpad)) return rc;
  if (int rc = make_map(&mb_hi, B_hi, N, Kpad, BN)) return rc;
  if (int rc = make_map(&mb_lo, B_lo, N, Kpad, BN)) return rc;
  static bool configured[LVSR_MAX_DEVICES] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    LVSR_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcShape<128>::SMEM));
    LVSR_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcShape<256>::SMEM));
    configured[dev] = true;
  }
  const int total_kb = Kpad / TC_BK;
  splits = std::min(splits, total_kb);
  TcGemmParams p;
  p.C = C; p.bias = bias; p.M = M; p.N = N; p.K = Kpad; p.ldc = ldc;
  p.kb_per_split = ceil_div(total_kb, splits);
  p.c_split_stride = split_stride;
  p.out_scale = nullptr;
  dim3 grid(N / BN, ceil_div(M, TC_BM), ceil_div(total_kb, p.kb_per_split));
  if (wide) gemm_tc_kernel<256, false><<<grid, TC_THREADS, TcShape<256>::SMEM, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  else gemm_tc_kernel<128, false><<<grid, TC_THREADS, TcShape<128>::SMEM, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  LVSR_LAUNCH_CHECK();
  return 0;
}

// C[M,N] = A[M,K] . W + bias with W given as the K-major hi/lo pair produced by split_weight_tf32.
// A_hi / A_lo: caller-provided scratch of M * gemm_tc_kpad(K) floats each.
int gemm_tc(const float* A, float* A_hi, float* A_lo, int M, int K, const float* Wt_hi, const float* Wt_lo, int N,
            const float* bias, float* C, int ldc, cudaStream_t stream) {
  LVSR_CHECK(gemm_tc_supported(M, N, K), "gemm_tc: unsupported shape M=%d N=%d K=%d", M, N, K);
  if (int rc = get_encode()) return rc;
  const int Kpad = gemm_tc_kpad(K);
  if (Kpad == K) {
    const long long n4 = (long long)M * K / 4;
    split_tf32_kernel<<<(int)std::min<long long>(4096, (n4 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(A), reinterpret_cast<float4*>(A_hi), reinterpret_cast<float4*>(A_lo), n4);
  } else {
    const long long n = (long long)M * Kpad;
    split_pad_tf32_kernel<<<(int)std::min<long long>(4096, (n + 255) / 256), 256, 0, stream>>>(A, A_hi, A_lo, M, K, Kpad);
  }
  LVSR_LAUNCH_CHECK();
  return gemm_tc_presplit(A_hi, A_lo, M, Wt_hi, Wt_lo, N, Kpad, bias, C, ldc, 1, 0, stream);
}


// ---- fp16 head/tail variant (inference projections whose input is a BiGRU output) ------------------------------------
namespace {

__device__ __forceinline__ void split_h16(float x, float y, __half2& head, __half2& tail) {
  head = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(head);
  tail = __floats2half2_rn((x - hf.x) * 2048.f, (y - hf.y) * 2048.f);
}

// x [M, K] fp32 -> heads / scaled tails [M, Kpad] fp16 (zero padding of the contraction dimension)
__global__ void split_h16_kernel(const float* __restrict__ x, __half2* __restrict__ head, __half2* __restrict__ tail, long long M,
                                 int K, int Kpad) {
  const long long total = M * (Kpad / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (Kpad / 2);
    const int k = (int)(i % (Kpad / 2)) * 2;
    float a = 0.f, b = 0.f;
    if (k + 1 < K) {
      const float2 v = *reinterpret_cast<const float2*>(x + r * K + k);      // K % 2 == 0, rows 8-byte aligned
      a = v.x; b = v.y;
    } else if (k < K) {
      a = x[r * K + k];
    }
    __half2 h, t;
    split_h16(a, b, h, t);
    head[i] = h;
    tail[i] = t;
  }
}

// scale2[0] = power of two that brings max|W| below 2^14 (1 for every sane model), scale2[1] = its inverse
__global__ void weight_scale_kernel(const float* __restrict__ W, long long n, float* __restrict__ scale2) {
  __shared__ float red[32];
  float mx = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, fabsf(W[i]));
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) {
      float sc = 1.f, inv = 1.f;
      if (mx > 16384.f && mx < 3.0e38f) {
        const int e = ((__float_as_int(mx) >> 23) & 0xff) - 127;
        sc = __int_as_float((127 - (e - 13)) << 23);
        inv = __int_as_float((127 + (e - 13)) << 23);
      }
      scale2[0] = sc;
      scale2[1] = inv;
    }
  }
}

// [K, N] row-major -> K-major [N, Kpad] heads / scaled tails of W * scale2[0]
__global__ void transpose_split_h16_kernel(const float* __restrict__ W, __half* __restrict__ head, __half* __restrict__ tail,
                                           int K, int N, int Kpad, int ldw, const float* __restrict__ scale2) {
  __shared__ float tile[32][33];
  const float sc = scale2[0];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? W[(long long)k * ldw + n] * sc : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < Kpad) {
      const float a = tile[threadIdx.x][i];
      const __half h = __float2half_rn(a);
      head[(long long)n * Kpad + k] = h;
      tail[(long long)n * Kpad + k] = __float2half_rn((a - __half2float(h)) * 2048.f);
    }
  }
}

// 2-D fp16 tensor [rows, Kpad] (K contiguous), box = [box_rows, 64 halfs], 128-byte swizzle
int make_map_h16(CUtensorMap* map, const void* ptr, long long rows, int Kpad, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kpad * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK_H16, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LVSR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp16) failed (%d)", (int)r);
  return 0;
}

}  // namespace

int gemm_tc_kpad_h16(int K) { return ceil_div(K, TC_BK_H16) * TC_BK_H16; }
bool gemm_tc_h16_supported(int M, int N, int K) { return M >= 1 && N % TC_BN == 0 && K >= 2 && K % 2 == 0; }

// head / tail: [N, gemm_tc_kpad_h16(K)] halfs; scale2: 2 floats on the device
int split_weight_h16(const float* W, int K, int N, void* head, void* tail, float* scale2, cudaStream_t stream) {
  const int Kpad = gemm_tc_kpad_h16(K);
  weight_scale_kernel<<<1, 1024, 0, stream>>>(W, (long long)K * N, scale2);
  LVSR_LAUNCH_CHECK();
  dim3 grid(ceil_div(N, 32), ceil_div(Kpad, 32)), block(32, 8);
  transpose_split_h16_kernel<<<grid, block, 0, stream>>>(W, static_cast<__half*>(head), static_cast<__half*>(tail), K, N, Kpad, N,
                                                          scale2);
  LVSR_LAUNCH_CHECK();
  return 0;
}

// C[M,N] = A[M,K] . W + bias with W given as the K-major head/tail pair of split_weight_h16.  |A| must stay inside the
// fp16 range (the callers pass BiGRU outputs).  A_head / A_tail: scratch of M * gemm_tc_kpad_h16(K) halfs each.
int gemm_tc_h16(const float* A, void* A_head, void* A_tail, int M, int K, const void* Wt_head, const void* Wt_tail,
                const float* scale2, int N, const float* bias, float* C, int ldc, cudaStream_t stream) {
  ProfScope prof("gemm", stream);
  LVSR_CHECK(gemm_tc_h16_supported(M, N, K), "gemm_tc_h16: unsupported shape M=%d N=%d K=%d", M, N, K);
  if (int rc = get_encode()) return rc;
  const int Kpad = gemm_tc_kpad_h16(K);
  {
    const long long n = (long long)M * (Kpad / 2);
    split_h16_kernel<<<(int)std::min<long long>(4096, (n + 255) / 256), 256, 0, stream>>>(
        A, static_cast<__half2*>(A_head), static_cast<__half2*>(A_tail), M, K, Kpad);
    LVSR_LAUNCH_CHECK();
  }
  const bool wide = (N % 256 == 0) && getenv("LVSR_TC_NARROW") == nullptr;
  const int BN = wide ? 256 : 128;
  CUtensorMap ma_h, ma_t, mb_h, mb_t;
  if (int rc = make_map_h16(&ma_h, A_head, M, Kpad, TC_BM)) return rc;
  if (int rc = make_map_h16(&ma_t, A_tail, M, Kpad, TC_BM)) return rc;
  if (int rc = make_map_h16(&mb_h, Wt_head, N, Kpad, BN)) return rc;
  if (int rc = make_map_h16(&mb_t, Wt_tail, N, Kpad, BN)) return rc;
  static bool configured[LVSR_MAX_DEVICES] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    LVSR_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcShape<128>::SMEM));
    LVSR_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcShape<256>::SMEM));
    configured[dev] = true;
  }
  TcGemmParams p;
  p.C = C; p.bias = bias; p.M = M; p.N = N; p.K = Kpad; p.ldc = ldc;
  p.kb_per_split = Kpad / TC_BK_H16;
  p.c_split_stride = 0;
  p.out_scale = scale2 ? scale2 + 1 : nullptr;
  dim3 grid(N / BN, ceil_div(M, TC_BM), 1);
  if (wide) gemm_tc_kernel<256, true><<<grid, TC_THREADS, TcShape<256>::SMEM, stream>>>(ma_h, ma_t, mb_h, mb_t, p);
  else gemm_tc_kernel<128, true><<<grid, TC_THREADS, TcShape<128>::SMEM, stream>>>(ma_h, ma_t, mb_h, mb_t, p);
  LVSR_LAUNCH_CHECK();
  return 0;
}

}  // namespace lvsr
