// TF32 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared
// memory -> tcgen05.mma.kind::tf32 with the fp32 accumulator in TMEM -> tcgen05.ld epilogue.
//
//   C[M,N] = epi( A . B ),  fp32 in memory, TF32 multiply, fp32 accumulate.
//   A is K-major ([M][K], transA = 0) or MN-major ([K][M], transA = 1);
//   B is K-major ([N][K], transB = 1, an nn.Linear weight) or MN-major ([K][N], transB = 0).
// All three products of a linear layer (forward, data gradient, weight gradient) therefore
// run on the same kernel without materialising a transpose.
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA
// issuer, warps 2..5 = epilogue (each owns the 32 TMEM lanes of its warp_id % 4 quarter).
// Tile 128 x BN x 32 (one 128-byte swizzle span of fp32 along K per stage, four UMMA_K=8
// instructions per stage), STAGES-deep mbarrier ring.  One tile per CTA; split-K over
// gridDim.z writes raw partials that rtk::k_splitk_reduce folds deterministically.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_kernels.cuh"

namespace rttc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;   // floats = 128 bytes
constexpr int UMMA_K = 8;     // tf32
constexpr int NUM_THREADS = 192;
// persistent GEMM: 8 epilogue warps (two per TMEM lane quarter, alternating 32-column chunks): an
// epilogue warp is alone on its scheduler and latency-bound, so the drain time halves
constexpr int P_EPI_WARPS = 8;
constexpr int NUM_THREADS_P = 64 + 32 * P_EPI_WARPS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 zero-fills the chunk.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once all prior cp.async of this thread landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout).
//   layout 2 = SWIZZLE_128B          (16-byte chunks XOR row%8; K-major operands)
//   layout 1 = SWIZZLE_128B_BASE32B  (32-byte chunks XOR row%4; the only swizzled layout the
//              tensor core accepts for MN-major 32-bit (tf32) operands — filled by TMA with
//              CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  d |= (uint64_t)layout << 61;
  return d;
}


// Epilogue operands shared by the GEMM and implicit-conv kernels.
struct EpiArgs {
  float* C;            // output rows (or split-K partial slab), row pitch ldc floats
  int ldc;
  int M, N;
  float alpha;
  const float* bias;
  const float* bias2;
  int relu;
  const float* mask;
  int ldmask;
  int accumulate;
  int round_tf32;
  int raw;             // 1: store the raw accumulator (split-K partial), no epilogue ops
};

// Per-warp, per-tile epilogue state: everything that does not depend on the chunk is computed
// once (the first version re-derived pointers, predicates and parameter loads per chunk and
// spent ~3900 cycles per 32x32 chunk in a ~400-instruction branchy sequence).
struct EpiWarp {
  float* dst;            // C + (row0 + lane/8) * ldc + n0 + 4 * (lane%8)
  const float* msk;      // same position in the mask (or null)
  size_t row_step;       // 4 * ldc floats
  size_t mrow_step;      // 4 * ldmask floats
  int n0;                // first column of the tile
  int rows_left;         // M - (row0 + lane/8): row i*4 of this lane is valid iff i*4 < rows_left
  bool fast;             // aligned, full-width tile: 16-byte path without column predicates
  bool simple;           // fast, and only bias (+ReLU): the short instruction sequence
};

__device__ __forceinline__ EpiWarp epi_begin(const EpiArgs& e, int lane, int row0, int n0, int bn) {
  EpiWarp w;
  const int r = lane >> 3, cg = lane & 7;
  w.dst = e.C + (size_t)(row0 + r) * e.ldc + n0 + 4 * cg;
  w.msk = e.mask ? e.mask + (size_t)(row0 + r) * e.ldmask + n0 + 4 * cg : nullptr;
  w.row_step = (size_t)4 * e.ldc;
  w.mrow_step = (size_t)4 * e.ldmask;
  w.n0 = n0;
  w.rows_left = e.M - (row0 + r);
  w.fast = ((e.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.C) & 15) == 0) && (n0 + bn <= e.N) &&
           ((n0 & 3) == 0) &&
           (e.raw || ((!e.mask || (((e.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.mask) & 15) == 0))) &&
                      (!e.bias || ((reinterpret_cast<uintptr_t>(e.bias) & 15) == 0)) &&
                      (!e.bias2 || ((reinterpret_cast<uintptr_t>(e.bias2) & 15) == 0))));
  w.simple = w.fast && !e.raw && !e.mask && !e.accumulate && e.alpha == 1.f && !e.bias2 &&
             w.rows_left >= 32;
  return w;
}

// One warp moves its 32-row x 32-column accumulator chunk (lane = row, straight out of
// tcgen05.ld) to global memory through a padded shared-memory transpose, so that every store
// instruction writes four complete 128-byte row segments.  `stage` = this warp's private
// 32 x 36 float scratch; chunk column offset `coff` = 32 * c.
__device__ __forceinline__ void epilogue_chunk(const EpiArgs& e, const EpiWarp& w, const uint32_t* v, float* stage,
                                               int lane, int row0, int nb) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(stage + lane * 36 + 4 * q) =
        make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                    __uint_as_float(v[4 * q + 3]));
  __syncwarp();
  const int cg = lane & 7;
  const int n = nb + 4 * cg;
  if (w.simple) {
    // ---- bias (+ReLU) only, all 32 rows valid: the epilogue warps are alone on their schedulers,
    // so the chunk time is the length of this dependent instruction sequence
    float4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = *reinterpret_cast<const float4*>(stage + (i * 4 + (lane >> 3)) * 36 + 4 * cg);
    float* dptr = w.dst + (nb - w.n0);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias) b = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    if (e.relu && e.round_tf32) {
      // output feeds the next tensor-core product: store the nearest TF32 value (rtk::rna_tf32)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(dptr + i * w.row_step) =
            make_float4(rtk::rna_tf32(fmaxf(t[i].x + b.x, 0.f)), rtk::rna_tf32(fmaxf(t[i].y + b.y, 0.f)),
                        rtk::rna_tf32(fmaxf(t[i].z + b.z, 0.f)), rtk::rna_tf32(fmaxf(t[i].w + b.w, 0.f)));
    } else if (e.round_tf32) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(dptr + i * w.row_step) =
            make_float4(rtk::rna_tf32(t[i].x + b.x), rtk::rna_tf32(t[i].y + b.y), rtk::rna_tf32(t[i].z + b.z),
                        rtk::rna_tf32(t[i].w + b.w));
    } else if (e.relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(dptr + i * w.row_step) =
            make_float4(fmaxf(t[i].x + b.x, 0.f), fmaxf(t[i].y + b.y, 0.f), fmaxf(t[i].z + b.z, 0.f),
                        fmaxf(t[i].w + b.w, 0.f));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(dptr + i * w.row_step) =
            make_float4(t[i].x + b.x, t[i].y + b.y, t[i].z + b.z, t[i].w + b.w);
    }
    __syncwarp();
    return;
  }
  if (w.fast) {
    // ---- straight-line 16-byte path: loads first (LDS + optional mask / old C), then math, then stores
    float4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = *reinterpret_cast<const float4*>(stage + (i * 4 + (lane >> 3)) * 36 + 4 * cg);
    const int chunk_off = nb - w.n0;                  // w.dst / w.msk already point at column n0 + 4*cg
    float* dptr = w.dst + chunk_off;
    if (e.raw) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i * 4 < w.rows_left) *reinterpret_cast<float4*>(dptr + i * w.row_step) = t[i];
    } else {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e.bias) b = __ldg(reinterpret_cast<const float4*>(e.bias + n));
      if (e.bias2) {
        float4 b2 = __ldg(reinterpret_cast<const float4*>(e.bias2 + n));
        b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
      }
      float4 mk[8], old[8];
      if (e.mask) {
        const float* mp = w.msk + chunk_off;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          mk[i] = (i * 4 < w.rows_left) ? *reinterpret_cast<const float4*>(mp + i * w.mrow_step)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (e.accumulate) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          old[i] = (i * 4 < w.rows_left) ? *reinterpret_cast<const float4*>(dptr + i * w.row_step)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x[4] = {t[i].x * e.alpha + b.x, t[i].y * e.alpha + b.y, t[i].z * e.alpha + b.z, t[i].w * e.alpha + b.w};
        if (e.relu) {
#pragma unroll
          for (int k = 0; k < 4; ++k) x[k] = fmaxf(x[k], 0.f);
        }
        if (e.mask) {
          x[0] = mk[i].x > 0.f ? x[0] : 0.f; x[1] = mk[i].y > 0.f ? x[1] : 0.f;
          x[2] = mk[i].z > 0.f ? x[2] : 0.f; x[3] = mk[i].w > 0.f ? x[3] : 0.f;
        }
        if (e.accumulate) { x[0] += old[i].x; x[1] += old[i].y; x[2] += old[i].z; x[3] += old[i].w; }
        if (e.round_tf32) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t tt;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tt) : "f"(x[k]));
            x[k] = __uint_as_float(tt);
          }
        }
        if (i * 4 < w.rows_left) *reinterpret_cast<float4*>(dptr + i * w.row_step) = make_float4(x[0], x[1], x[2], x[3]);
      }
    }
    __syncwarp();
    return;
  }
  // ---- general path: column tails / unaligned operands
  float b[4] = {0.f, 0.f, 0.f, 0.f};
  if (!e.raw) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (n + k < e.N) {
        if (e.bias) b[k] += __ldg(e.bias + n + k);
        if (e.bias2) b[k] += __ldg(e.bias2 + n + k);
      }
  }
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3);
    const int m = row0 + r;
    float4 t4 = *reinterpret_cast<const float4*>(stage + r * 36 + 4 * cg);
    if (m >= e.M || n >= e.N) continue;
    float x[4] = {t4.x, t4.y, t4.z, t4.w};
    float* dst = e.C + (size_t)m * e.ldc + n;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (n + k >= e.N) continue;
      float y = x[k];
      if (!e.raw) {
        y = y * e.alpha + b[k];
        if (e.relu) y = fmaxf(y, 0.f);
        if (e.mask) y = e.mask[(size_t)m * e.ldmask + n + k] > 0.f ? y : 0.f;
        if (e.accumulate) y += dst[k];
        if (e.round_tf32) {
          uint32_t tt;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tt) : "f"(y));
          y = __uint_as_float(tt);
        }
      }
      dst[k] = y;
    }
  }
  __syncwarp();
}

// Extra epilogue switch on top of rtk::GemmArgs
struct TcArgs {
  rtk::GemmArgs g;
  int num_kb_total;   // ceil(K / BLOCK_K)
  int kb_per_split;
  int round_tf32;     // round outputs to TF32 (RN) so the consumer GEMM multiplies exact values
  long long* dbg;     // optional clock64 phase stamps of CTA (0,0,0) (tuning aid)
};

template <int BN, int A_MN, int B_MN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;  // 16 KiB
  static constexpr int B_BYTES = BN * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

template <int BN, int A_MN, int B_MN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ TcArgs a) {
  using L = SmemLayout<BN, A_MN, B_MN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const rtk::GemmArgs& g = a.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * BN;
  const int kb0 = blockIdx.z * a.kb_per_split;
  int kb1 = kb0 + a.kb_per_split;
  if (kb1 > a.num_kb_total) kb1 = a.num_kb_total;
  const int num_kb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const bool dbg_cta = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (dbg_cta && threadIdx.x == 0) a.dbg[0] = clock64();   // setup done

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      for (int i = 0; i < num_kb; ++i) {
        int s = i % STAGES;
        uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        int k0 = (kb0 + i) * BLOCK_K;
        if (A_MN) {
          // A stored [K][M]: 32-float MN atoms, each a (BLOCK_K rows x 128 B) block
#pragma unroll
          for (int j = 0; j < BLOCK_M / 32; ++j)
            tma_load_2d(&tmA, &full_bar[s], sa + j * (BLOCK_K * 128), m0 + 32 * j, k0);
        } else {
          tma_load_2d(&tmA, &full_bar[s], sa, k0, m0);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j)
            tma_load_2d(&tmB, &full_bar[s], sb + j * (BLOCK_K * 128), n0 + 32 * j, k0);
        } else {
          tma_load_2d(&tmB, &full_bar[s], sb, k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer (one thread)
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)A_MN << 15) |
                             ((uint32_t)B_MN << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BLOCK_M >> 4) << 24);
      for (int i = 0; i < num_kb; ++i) {
        int s = i % STAGES;
        uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        if (dbg_cta && i == 0) a.dbg[1] = clock64();           // first operands landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // K-major: 8-row groups 1024 B apart (SBO), +32 B per UMMA_K inside the swizzle span.
          // MN-major: 32-float atoms BLOCK_K*128 B apart (LBO); inside an atom the K rows are
          // 128 B apart in groups of 4 (SBO = 512 B); one UMMA_K = 8 rows = 1024 B.
          uint64_t ad = A_MN ? make_smem_desc(sa + k * 1024, BLOCK_K * 128, 512, 1)
                             : make_smem_desc(sa + k * (UMMA_K * 4), 16, 1024, 2);
          uint64_t bd = B_MN ? make_smem_desc(sb + k * 1024, BLOCK_K * 128, 512, 1)
                             : make_smem_desc(sb + k * (UMMA_K * 4), 16, 1024, 2);
          umma_tf32(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs retire
      }
      umma_commit(tmem_full);         // accumulator complete
      if (dbg_cta) a.dbg[2] = clock64();                        // all MMAs issued
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> global
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    mbar_wait(tmem_full, 0);
    if (dbg_cta && threadIdx.x == 64) a.dbg[3] = clock64();     // accumulator ready
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // all MMAs have retired, so the pipeline stages are free: reuse them as transpose scratch
    float* stage = reinterpret_cast<float*>(smem) + q * (32 * 36);
    EpiArgs e;
    const bool split = gridDim.z > 1;
    e.C = split ? g.ws + (size_t)blockIdx.z * g.M * g.N : g.C;
    e.ldc = split ? g.N : g.ldc;
    e.M = g.M; e.N = g.N; e.alpha = g.alpha; e.bias = g.bias; e.bias2 = g.bias2; e.relu = g.relu;
    e.mask = g.mask; e.ldmask = g.ldmask; e.accumulate = g.accumulate; e.round_tf32 = a.round_tf32;
    e.raw = split ? 1 : 0;
    const EpiWarp ew = epi_begin(e, lane, m0 + q * 32, n0, BN);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      if (n0 + c * 32 < g.N) epilogue_chunk(e, ew, v, stage, lane, m0 + q * 32, n0 + c * 32);
    }
    if (dbg_cta && threadIdx.x == 64) a.dbg[4] = clock64();     // epilogue done
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
}


// ---------------------------------------------------------------------- persistent GEMM
// Same math and operand handling as k_gemm_tc, restructured so that a tile's epilogue overlaps
// the next tile's main loop: one CTA per SM loops over tiles (tile = blockIdx.x + i*gridDim.x,
// N-tiles of one M-row adjacent so co-running CTAs share the A rows in L2), the accumulator is
// double-buffered in TMEM (2 x BN columns) behind tmem_full / tmem_empty mbarriers, and the
// epilogue warps drain accumulator `t & 1` while the MMA warp is already filling the other one.
// (Measured on the one-tile-per-CTA kernel: main loop 5.7 us + epilogue 7.9 us per 128x128x512
// tile, serialised and in lock-step across the chip.)
template <int BN, int A_MN, int B_MN, int STAGES>
struct SmemLayoutP {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;
  static constexpr int B_BYTES = BN * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SCRATCH_OFFSET = STAGES * STAGE_BYTES;          // P_EPI_WARPS x 32 x 36 floats
  static constexpr int BAR_OFFSET = SCRATCH_OFFSET + P_EPI_WARPS * 32 * 36 * 4;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;
};

template <int BN, int A_MN, int B_MN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS_P)
k_gemm_tc_p(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ TcArgs a) {
  using L = SmemLayoutP<BN, A_MN, B_MN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const rtk::GemmArgs& g = a.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (g.N + BN - 1) / BN;
  const int tiles_m = (g.M + BLOCK_M - 1) / BLOCK_M;
  const int splits = (a.num_kb_total + a.kb_per_split - 1) / a.kb_per_split;
  const int total_tiles = tiles_n * tiles_m * splits;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], P_EPI_WARPS);     // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (m block, n block, k split); n fastest
  auto decode = [&](int tile, int& mb, int& nb, int& z) {
    nb = tile % tiles_n;
    int r = tile / tiles_n;
    mb = r % tiles_m;
    z = r / tiles_m;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;   // global k-block counter: stage = it % STAGES
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mb, nb, z;
        decode(tile, mb, nb, z);
        const int m0 = mb * BLOCK_M, n0 = nb * BN;
        const int kb0 = z * a.kb_per_split;
        int kb1 = kb0 + a.kb_per_split;
        if (kb1 > a.num_kb_total) kb1 = a.num_kb_total;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          int s = it % STAGES;
          uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
          int k0 = kb * BLOCK_K;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 32; ++j)
              tma_load_2d(&tmA, &full_bar[s], sa + j * (BLOCK_K * 128), m0 + 32 * j, k0);
          } else {
            tma_load_2d(&tmA, &full_bar[s], sa, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j)
              tma_load_2d(&tmB, &full_bar[s], sb + j * (BLOCK_K * 128), n0 + 32 * j, k0);
          } else {
            tma_load_2d(&tmB, &full_bar[s], sb, k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)A_MN << 15) |
                             ((uint32_t)B_MN << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BLOCK_M >> 4) << 24);
      uint32_t it = 0;
      int lt = 0;        // local tile counter: accumulator = lt & 1
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        int mb, nb, z;
        decode(tile, mb, nb, z);
        const int kb0 = z * a.kb_per_split;
        int kb1 = kb0 + a.kb_per_split;
        if (kb1 > a.num_kb_total) kb1 = a.num_kb_total;
        const int acc = lt & 1;
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          int s = it % STAGES;
          uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            uint64_t ad = A_MN ? make_smem_desc(sa + k * 1024, BLOCK_K * 128, 512, 1)
                               : make_smem_desc(sa + k * (UMMA_K * 4), 16, 1024, 2);
            uint64_t bd = B_MN ? make_smem_desc(sb + k * 1024, BLOCK_K * 128, 512, 1)
                               : make_smem_desc(sb + k * (UMMA_K * 4), 16, 1024, 2);
            umma_tf32(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;       // which of the quarter's two warps: chunk parity
    float* stage = reinterpret_cast<float*>(smem + L::SCRATCH_OFFSET) + (warp - 2) * (32 * 36);
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      int mb, nb, z;
      decode(tile, mb, nb, z);
      const int m0 = mb * BLOCK_M, n0 = nb * BN;
      const int acc = lt & 1;
      EpiArgs e;
      const bool split = splits > 1;
      e.C = split ? g.ws + (size_t)z * g.M * g.N : g.C;
      e.ldc = split ? g.N : g.ldc;
      e.M = g.M; e.N = g.N; e.alpha = g.alpha; e.bias = g.bias; e.bias2 = g.bias2; e.relu = g.relu;
      e.mask = g.mask; e.ldmask = g.ldmask; e.accumulate = g.accumulate; e.round_tf32 = a.round_tf32;
      e.raw = split ? 1 : 0;
      const EpiWarp ew = epi_begin(e, lane, m0 + q * 32, n0, BN);
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      constexpr int CHUNKS = BN / 32;
      constexpr int CSTEP = CHUNKS >= 2 ? 2 : 1;      // BN = 32: both warps would own the one chunk
      const int c_first = CHUNKS >= 2 ? half : 0;
      const int c_last = CHUNKS >= 2 ? CHUNKS - 2 + half : 0;
      if (CHUNKS < 2 && half == 1) {
        // nothing to drain: still hand the accumulator back
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[acc])) : "memory");
        continue;
      }
#pragma unroll 1
      for (int c = c_first; c < CHUNKS; c += CSTEP) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
        if (c == c_last) {
          // every column of this accumulator is now in registers: hand it back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[acc])) : "memory");
        }
        if (n0 + c * 32 < g.N) epilogue_chunk(e, ew, v, stage, lane, m0 + q * 32, n0 + c * 32);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ CTA-pair persistent GEMM
// k_gemm_tc_p<256> is bound by shared-memory bandwidth, not by the tensor pipe: a 128x256x8 TF32 MMA reads
// 12 KB of operands from shared memory in the 176 cycles it takes at peak while TMA writes the next 12 KB
// (140 B/clk against the SM's 128 B/clk; ncu: the MMA warp never waits for data or for an accumulator, the
// producer and the epilogue warps wait for IT, tensor pipe 52 % active).  Here two CTAs of a cluster (the two
// SMs of a TPC) share one 256 x 256 tile: tcgen05.mma.cta_group::2 with M = 256, each CTA holds its 128 rows of
// A and HALF of the B tile (its 128 of the 256 columns) -- 32 KB per k-block and SM instead of 48 KB -- and the
// accumulator rows of its half in its own TMEM.  K-major A and B only (the hidden-layer products), no split-K.
//   full[s]       leader CTA's barrier: its producer arms it for both CTAs' bytes, both CTAs' TMA complete on it
//   empty[s]      one per CTA, released by the leader's commit (multicast to both)
//   tmem_full[a]  one per CTA (multicast commit); tmem_empty[a]: leader's, 2 x P_EPI_WARPS arrivals (the peer's
//                 epilogue warps arrive remotely)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

constexpr int PAIR_BN = 256;          // columns of the pair's tile (each CTA stages 128 of them)
template <int STAGES>
struct SmemLayoutPair {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;
  static constexpr int B_BYTES = (PAIR_BN / 2) * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SCRATCH_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = SCRATCH_OFFSET + P_EPI_WARPS * 32 * 36 * 4;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;
};

template <int STAGES>
__global__ void __launch_bounds__(NUM_THREADS_P, 1)
k_gemm_tc_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ TcArgs a) {
  using L = SmemLayoutPair<STAGES>;
  constexpr int BN = PAIR_BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const rtk::GemmArgs& g = a.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tiles_n = g.N / BN;
  const int tiles_m = (g.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int total_tiles = tiles_n * tiles_m;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * P_EPI_WARPS);     // one arrive per epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();      // both CTAs' barriers exist before anything arrives on them from the peer
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        const int nb = tile % tiles_n, mb = tile / tiles_n;
        const int m0 = mb * (2 * BLOCK_M) + (int)rank * BLOCK_M;
        const int n0 = nb * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < a.num_kb_total; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
          const int k0 = kb * BLOCK_K;
          tma_load_2d_pair(&tmA, bar, sa, k0, m0);
          tma_load_2d_pair(&tmB, bar, sb, k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
      uint32_t it = 0;
      int lt = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs, ++lt) {
        const int acc = lt & 1;
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);   // both epilogues have drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < a.num_kb_total; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t ad = make_smem_desc(sa + k * (UMMA_K * 4), 16, 1024, 2);
            const uint64_t bd = make_smem_desc(sb + k * (UMMA_K * 4), 16, 1024, 2);
            umma_tf32_pair(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[s]);
        }
        umma_commit_pair(&tmem_full[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stage = reinterpret_cast<float*>(smem + L::SCRATCH_OFFSET) + (warp - 2) * (32 * 36);
    int lt = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++lt) {
      const int nb = tile % tiles_n, mb = tile / tiles_n;
      const int m0 = mb * (2 * BLOCK_M) + (int)rank * BLOCK_M, n0 = nb * BN;
      const int acc = lt & 1;
      EpiArgs e;
      e.C = g.C; e.ldc = g.ldc;
      e.M = g.M; e.N = g.N; e.alpha = g.alpha; e.bias = g.bias; e.bias2 = g.bias2; e.relu = g.relu;
      e.mask = g.mask; e.ldmask = g.ldmask; e.accumulate = g.accumulate; e.round_tf32 = a.round_tf32;
      e.raw = 0;
      const EpiWarp ew = epi_begin(e, lane, m0 + q * 32, n0, BN);
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      constexpr int CHUNKS = BN / 32;
      const int c_last = CHUNKS - 2 + half;
#pragma unroll 1
      for (int c = half; c < CHUNKS; c += 2) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
        if (c == c_last) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
        }
        if (m0 + q * 32 < g.M) epilogue_chunk(e, ew, v, stage, lane, m0 + q * 32, n0 + c * 32);
      }
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();      // the peer may still read this CTA's operands / signal its barriers until here
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------- implicit-GEMM conv
// Forward convolution without an im2col buffer: out[(img,oh,ow), f] = relu(sum_k A[..,k] W[f,k] + b)
// where the A tile (128 output pixels x 32 taps) is gathered straight from the input by four
// producer warps into the 128B-swizzled K-major stage the tensor core reads:
//   IN_U8 = 1: input uint8 NCHW frames (replay batch), k = (c, kh, kw), scaled by 1/255 on load
//              (rltime/models/torch/modules/cnn.py:44-45) — 28 KB read per frame instead of a
//              400 KB fp32 im2col round trip;
//   IN_U8 = 0: input fp32 NHWC, k = (kh, kw, c), C % 32 == 0: a k-block is one 128-byte run.
// Warps 0-3: A producers, then epilogue; warp 4: TMA producer for the filter tile; warp 5: TMEM
// allocator + MMA issuer.
struct ConvArgs {
  const void* in;      // uint8 NCHW or float NHWC
  float* out;          // [M][N] (NHWC)
  const float* bias;
  int rows;            // images
  int C, H, W, KH, S, OH, OW;
  int M, N, K;         // M = rows*OH*OW, N = filters, K = C*KH*KH (multiple of 32)
  float scale;
  int round_tf32;
  // Two networks in one product (k_conv_tc_p, BN = 64 = 2 x 32 filters): the filter tile is [W_a ; W_b], the
  // 32-column accumulator chunk 0 goes to `out` (+bias) for every row, chunk 1 to `out2` (+bias2) for the
  // rows >= row_shift2 only, stored at row - row_shift2 (the second network reads a suffix of the same
  // frames).  row_shift2 must be a multiple of the 128-row tile.
  float* out2;
  const float* bias2;
  int split2;
  int row_shift2;
};

template <int BN, int IN_U8, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS)
k_conv_tc(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvArgs a) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;
  constexpr int B_BYTES = BN * BLOCK_K * 4;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * BN;
  const int num_kb = a.K / BLOCK_K;

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 128 + 1);   // 128 gather threads + the TMA thread
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ---------------- A gather: thread t owns 16-byte chunk j = t % 8 of rows i*16 + t/8
    const int t = threadIdx.x;
    const int j = t & 7;
    long long base[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int gm = m0 + i * 16 + (t >> 3);
      if (gm < a.M) {
        int ow = gm % a.OW;
        int oh = (gm / a.OW) % a.OH;
        long long img = gm / (a.OW * a.OH);
        base[i] = IN_U8 ? (img * a.C * a.H + (long long)oh * a.S) * a.W + (long long)ow * a.S
                        : ((img * a.H + (long long)oh * a.S) * a.W + (long long)ow * a.S) * a.C;
      } else {
        base[i] = -1;
      }
    }
    auto gather = [&](int kb, float4* v) {
      long long koff;
      if (IN_U8) {
        int k = kb * BLOCK_K + 4 * j;
        int kw = k % a.KH, kh = (k / a.KH) % a.KH, c = k / (a.KH * a.KH);
        koff = ((long long)c * a.H + kh) * a.W + kw;
      } else {
        int k = kb * BLOCK_K;
        int c0 = k % a.C, tap = k / a.C;
        int kw = tap % a.KH, kh = tap / a.KH;
        koff = ((long long)kh * a.W + kw) * a.C + c0 + 4 * j;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (base[i] < 0) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (IN_U8) {
          uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(a.in) + base[i] + koff));
          v[i] = make_float4(__fmul_rn((float)(u & 0xff), a.scale), __fmul_rn((float)((u >> 8) & 0xff), a.scale),
                             __fmul_rn((float)((u >> 16) & 0xff), a.scale), __fmul_rn((float)(u >> 24), a.scale));
        } else {
          v[i] = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(a.in) + base[i] + koff));
        }
      }
    };
    if (!IN_U8) {
      // fp32 NHWC input: a chunk is a plain 16-byte copy, so the producers issue LDGSTS straight
      // into the swizzled stage and never wait for the data themselves: every free stage is in
      // flight at once (the register-staged version kept ONE k-block per thread in flight and a
      // 128 x 32 tile of conv1 took ~14 us, nearly all of it load latency).  The MMA thread
      // orders the landed generic-proxy writes before its async-proxy reads.
      const float* in = static_cast<const float*>(a.in);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int k = kb * BLOCK_K;
        const int c0 = k % a.C, tap = k / a.C;
        const int kw = tap % a.KH, kh = tap / a.KH;
        const long long koff = ((long long)kh * a.W + kw) * a.C + c0 + 4 * j;
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 16 + (t >> 3);
          const bool ok = base[i] >= 0;
          cp_async16(sa + r * 128 + ((j ^ (r & 7)) << 4), ok ? in + base[i] + koff : in, ok ? 16u : 0u);
        }
        cp_async_arrive(&full_bar[s]);
      }
    } else {
    // software pipeline: the loads of k-block kb+1 are in flight while kb waits for its stage
    float4 cur[8], nxt[8];
    gather(0, cur);
    for (int kb = 0; kb < num_kb; ++kb) {
      if (kb + 1 < num_kb) gather(kb + 1, nxt);
      int s = kb % STAGES;
      uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* sa = smem + s * STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int r = i * 16 + (t >> 3);
        *reinterpret_cast<float4*>(sa + r * 128 + ((j ^ (r & 7)) << 4)) = cur[i];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic -> async proxy
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
    }
    }
    // ---------------- epilogue (same warps: TMEM lane quarter == warp)
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);
    EpiArgs e;
    e.C = a.out; e.ldc = a.N; e.M = a.M; e.N = a.N; e.alpha = 1.f; e.bias = a.bias; e.bias2 = nullptr;
    e.relu = 1; e.mask = nullptr; e.ldmask = 0; e.accumulate = 0; e.round_tf32 = a.round_tf32; e.raw = 0;
    const EpiWarp ew = epi_begin(e, lane, m0 + warp * 32, n0, BN);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
      if (n0 + c * 32 < a.N) epilogue_chunk(e, ew, v, stage, lane, m0 + warp * 32, n0 + c * 32);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (warp == 4) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        int s = kb % STAGES;
        uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], B_BYTES);
        tma_load_2d(&tmB, &full_bar[s], smem + s * STAGE_BYTES + A_BYTES, kb * BLOCK_K, n0);
      }
    }
  } else {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BLOCK_M >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        int s = kb % STAGES;
        uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        if (!IN_U8) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async writes -> MMA reads
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
          umma_tf32(tmem_base, make_smem_desc(sa + k * (UMMA_K * 4), 16, 1024, 2),
                    make_smem_desc(sb + k * (UMMA_K * 4), 16, 1024, 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full);
    }
  }
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
}


// Persistent implicit-GEMM forward convolution: same gather / math as k_conv_tc, restructured
// like k_gemm_tc_p.  One CTA per SM loops over M tiles; EIGHT producer warps gather (4 x 16 B
// per thread per k-block, twice the loads in flight of the one-tile kernel), the accumulator is
// double-buffered in TMEM and four dedicated epilogue warps drain tile i while tile i+1 is
// being gathered and multiplied.  (Measured on k_conv_tc: conv1 spends ~14 us per 128x32 tile,
// almost all of it serialised gather latency and per-tile prologue.)
constexpr int CONV_P_THREADS = 448;   // 8 gather + 1 TMA + 1 MMA + 4 epilogue warps

template <int BN, int IN_U8, int STAGES>
__global__ void __launch_bounds__(CONV_P_THREADS)
k_conv_tc_p(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvArgs a) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;
  constexpr int B_BYTES = BN * BLOCK_K * 4;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int SCRATCH_OFFSET = STAGES * STAGE_BYTES;
  constexpr int BAR_OFFSET = SCRATCH_OFFSET + 4 * 32 * 36 * 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = a.K / BLOCK_K;
  const int tiles_n = (a.N + BN - 1) / BN;
  const int tiles_m = (a.M + BLOCK_M - 1) / BLOCK_M;
  const int total_tiles = tiles_n * tiles_m;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;

  if (warp == 8 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 256 + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ---------------- A gather: thread t owns 16-byte chunk j = t % 8 of rows i*32 + t/8, i < 4
    const int t = threadIdx.x;
    const int j = t & 7;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BLOCK_M;
      long long base[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int gm = m0 + i * 32 + (t >> 3);
        if (gm < a.M) {
          int ow = gm % a.OW;
          int oh = (gm / a.OW) % a.OH;
          long long img = gm / (a.OW * a.OH);
          base[i] = IN_U8 ? (img * a.C * a.H + (long long)oh * a.S) * a.W + (long long)ow * a.S
                          : ((img * a.H + (long long)oh * a.S) * a.W + (long long)ow * a.S) * a.C;
        } else {
          base[i] = -1;
        }
      }
      auto gather = [&](int kb, float4* v) {
        long long koff;
        if (IN_U8) {
          int k = kb * BLOCK_K + 4 * j;
          int kw = k % a.KH, kh = (k / a.KH) % a.KH, c = k / (a.KH * a.KH);
          koff = ((long long)c * a.H + kh) * a.W + kw;
        } else {
          int k = kb * BLOCK_K;
          int c0 = k % a.C, tap = k / a.C;
          int kw = tap % a.KH, kh = tap / a.KH;
          koff = ((long long)kh * a.W + kw) * a.C + c0 + 4 * j;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (base[i] < 0) {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          } else if (IN_U8) {
            uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(a.in) + base[i] + koff));
            v[i] = make_float4(__fmul_rn((float)(u & 0xff), a.scale), __fmul_rn((float)((u >> 8) & 0xff), a.scale),
                               __fmul_rn((float)((u >> 16) & 0xff), a.scale), __fmul_rn((float)(u >> 24), a.scale));
          } else {
            v[i] = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(a.in) + base[i] + koff));
          }
        }
      };
      if (!IN_U8) {
        // fp32 NHWC: LDGSTS straight into the swizzled stage, all free stages in flight
        const float* in = static_cast<const float*>(a.in);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int k = kb * BLOCK_K;
          const int cc = k % a.C, tap = k / a.C;
          const int kw = tap % a.KH, kh = tap / a.KH;
          const long long koff = ((long long)kh * a.W + kw) * a.C + cc + 4 * j;
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 32 + (t >> 3);
            const bool ok = base[i] >= 0;
            cp_async16(sa + r * 128 + ((j ^ (r & 7)) << 4), ok ? in + base[i] + koff : in, ok ? 16u : 0u);
          }
          cp_async_arrive(&full_bar[s]);
        }
        continue;
      }
      // two k-blocks of loads in flight ahead of the one being stored
      float4 c0[4], c1[4], c2[4];
      gather(0, c0);
      if (num_kb > 1) gather(1, c1);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        if (kb + 2 < num_kb) gather(kb + 2, c2);
        int s = it % STAGES;
        uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int r = i * 32 + (t >> 3);
          *reinterpret_cast<float4*>(sa + r * 128 + ((j ^ (r & 7)) << 4)) = c0[i];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
#pragma unroll
        for (int i = 0; i < 4; ++i) { c0[i] = c1[i]; c1[i] = c2[i]; }
      }
        }
  } else if (warp == 8) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          int s = it % STAGES;
          uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], B_BYTES);
          tma_load_2d(&tmB, &full_bar[s], smem + s * STAGE_BYTES + A_BYTES, kb * BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BLOCK_M >> 4) << 24);
      uint32_t it = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          int s = it % STAGES;
          uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          if (!IN_U8) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async writes -> MMA reads
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_tf32(d_tmem, make_smem_desc(sa + k * (UMMA_K * 4), 16, 1024, 2),
                      make_smem_desc(sb + k * (UMMA_K * 4), 16, 1024, 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ---------------- epilogue warps 10..13: TMEM lane quarter = warp % 4
    const int q = warp & 3;
    float* stage = reinterpret_cast<float*>(smem + SCRATCH_OFFSET) + q * (32 * 36);
    EpiArgs e;
    e.C = a.out; e.ldc = a.N; e.M = a.M; e.N = a.N; e.alpha = 1.f; e.bias = a.bias; e.bias2 = nullptr;
    e.relu = 1; e.mask = nullptr; e.ldmask = 0; e.accumulate = 0; e.round_tf32 = a.round_tf32; e.raw = 0;
    EpiArgs e2 = e;           // second network of a split product: its own output rows / bias
    if (BN == 64 && a.split2) {
      e.ldc = 32; e.N = 32;
      e2.C = a.out2; e2.ldc = 32; e2.N = 32; e2.M = a.M - a.row_shift2; e2.bias = a.bias2;
    }
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BN;
      const int acc = lt & 1;
      const bool split = BN == 64 && a.split2;
      const EpiWarp ew = epi_begin(e, lane, m0 + q * 32, n0, split ? 32 : BN);
      const EpiWarp ew2 = epi_begin(e2, lane, m0 - a.row_shift2 + q * 32, 0, 32);
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
        if (c == BN / 32 - 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[acc])) : "memory");
        }
        if (split && c == 1) {
          if (m0 >= a.row_shift2) epilogue_chunk(e2, ew2, v, stage, lane, m0 - a.row_shift2 + q * 32, 0);
        } else if (n0 + c * 32 < a.N) {
          epilogue_chunk(e, ew, v, stage, lane, m0 + q * 32, n0 + c * 32);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ----------------------------------------------------- implicit-GEMM conv data gradient
// dx[img,h,w,c] = relu'(act) * sum_{kh,kw,f} dy[img,(h-kh)/S,(w-kw)/S,f] W[f,kh,kw,c] without the
// dcol buffer / col2im scatter.  Input pixels are split into S*S parity classes (h%S, w%S): inside a
// class only the taps kh = S*dh + h%S, kw = S*dw + w%S contribute, so each class is a dense
// stride-1 "full" correlation of dy with (KH/S)^2 taps:
//   A row = class pixel (img,hh,ww), k = (dh,dw,f): dy[img, hh-dh, ww-dw, f..]  (zero outside dy)
//   B     = Wt[class][c][(dh,dw,f)]  (re-laid copy of the filters, rtk::k_conv_wT)
// Persistent CTA as k_conv_tc_p: 8 LDGSTS producer warps (zero-fill for the border taps), TMA for
// the filter tile, double-buffered TMEM accumulator; the epilogue applies the ReLU mask of the
// forward activation and writes each pixel's C channels as one run.
struct ConvDxArgs {
  const float* dy;     // [rows][OH][OW][F]
  const float* act;    // forward input of the layer (post-ReLU), [rows][H][W][C]: mask = act > 0
  float* dx;           // [rows][H][W][C]
  int rows, C, H, W, KH, S, OH, OW, F;
  int Hc, Wc;          // class grid: ceil(H/S), ceil(W/S)
  int KD;              // taps per dimension inside a class = KH / S
  int Kc;              // KD*KD*F
  int Mc;              // rows*Hc*Wc
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(CONV_P_THREADS)
k_convdx_tc(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvDxArgs a) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;
  constexpr int B_BYTES = BN * BLOCK_K * 4;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = a.Kc / BLOCK_K;
  const int tiles_m = (a.Mc + BLOCK_M - 1) / BLOCK_M;
  const int total_tiles = a.S * a.S * tiles_m;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;

  if (warp == 8 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 256 + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ---------------- A gather: thread t owns 16-byte chunk j = t % 8 of rows i*32 + t/8, i < 4
    const int t = threadIdx.x;
    const int j = t & 7;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile % tiles_m) * BLOCK_M;
      int hh[4], ww[4];
      long long ibase[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gm = m0 + i * 32 + (t >> 3);
        if (gm < a.Mc) {
          ww[i] = gm % a.Wc;
          hh[i] = (gm / a.Wc) % a.Hc;
          ibase[i] = (long long)(gm / (a.Wc * a.Hc)) * a.OH * a.OW;
        } else {
          ww[i] = hh[i] = -(1 << 20);   // every tap falls outside dy
          ibase[i] = 0;
        }
      }
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int k = kb * BLOCK_K;
        const int f0 = k % a.F, tapi = k / a.F;
        const int dw = tapi % a.KD, dh = tapi / a.KD;
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 32 + (t >> 3);
          const int oh = hh[i] - dh, ow = ww[i] - dw;
          const bool ok = oh >= 0 && oh < a.OH && ow >= 0 && ow < a.OW;
          const float* src = ok ? a.dy + ((ibase[i] + (long long)oh * a.OW + ow) * a.F + f0 + 4 * j) : a.dy;
          cp_async16(sa + r * 128 + ((j ^ (r & 7)) << 4), src, ok ? 16u : 0u);
        }
        cp_async_arrive(&full_bar[s]);
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int cls = tile / tiles_m;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          int s = it % STAGES;
          uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], B_BYTES);
          tma_load_2d(&tmB, &full_bar[s], smem + s * STAGE_BYTES + A_BYTES, kb * BLOCK_K, cls * a.C);
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BLOCK_M >> 4) << 24);
      uint32_t it = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          int s = it % STAGES;
          uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async writes -> MMA reads
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_tf32(d_tmem, make_smem_desc(sa + k * (UMMA_K * 4), 16, 1024, 2),
                      make_smem_desc(sb + k * (UMMA_K * 4), 16, 1024, 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ---------------- epilogue warps 10..13: TMEM lane = class pixel; mask + one run per pixel
    const int q = warp & 3;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int cls = tile / tiles_m;
      const int php = cls / a.S, pwp = cls % a.S;
      const int gm = (tile % tiles_m) * BLOCK_M + q * 32 + lane;
      long long off = -1;
      if (gm < a.Mc) {
        const int w = (gm % a.Wc) * a.S + pwp;
        const int h = ((gm / a.Wc) % a.Hc) * a.S + php;
        const long long img = gm / (a.Wc * a.Hc);
        if (h < a.H && w < a.W) off = ((img * a.H + h) * a.W + w) * a.C;
      }
      const int acc = lt & 1;
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
        if (c == BN / 32 - 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[acc])) : "memory");
        }
        if (off >= 0) {
          const float4* mk = reinterpret_cast<const float4*>(a.act + off + c * 32);
          float4* dst = reinterpret_cast<float4*>(a.dx + off + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 m4 = __ldg(mk + i);
            dst[i] = make_float4(m4.x > 0.f ? __uint_as_float(v[4 * i]) : 0.f,
                                 m4.y > 0.f ? __uint_as_float(v[4 * i + 1]) : 0.f,
                                 m4.z > 0.f ? __uint_as_float(v[4 * i + 2]) : 0.f,
                                 m4.w > 0.f ? __uint_as_float(v[4 * i + 3]) : 0.f);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// --------------------------------------------------------- implicit-GEMM conv weight gradient
// dW[f][k] = sum_p dy[p][f] * col[p][k] over a slab of output pixels p, without an im2col
// buffer.  A = dy^T: MN-major from memory (TMA, 128B_ATOM_32B swizzle); B = col: MN-major,
// gathered by four producer warps into the same swizzled atom layout (32 k-values x 32 pixels
// per atom).  grid = (K / BN tiles, 1, pixel slabs); raw partials go to the split-K workspace.
struct ConvDwArgs {
  const void* in;      // layer input: uint8 NCHW or float NHWC
  float* ws;           // [slabs][F][K] raw partials
  int C, H, W, KH, S, OH, OW;
  int P;               // output pixels = rows*OH*OW (contraction length)
  int F, K;            // filters, taps
  int kb_per_split;    // 32-pixel blocks per slab
  float scale;
};

constexpr int CONVDW_THREADS = 320;   // 8 gather (first 4 also epilogue) + 1 TMA + 1 MMA warps

template <int BN, int IN_U8, int STAGES>
__global__ void __launch_bounds__(CONVDW_THREADS)
k_convdw_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ ConvDwArgs a) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;          // 4 atoms of dy^T (only ceil(F/32) filled)
  constexpr int B_BYTES = BN * BLOCK_K * 4;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int total_kb = (a.P + BLOCK_K - 1) / BLOCK_K;
  const int kb0 = blockIdx.z * a.kb_per_split;
  int kb1 = kb0 + a.kb_per_split;
  if (kb1 > total_kb) kb1 = total_kb;
  const int num_kb = kb1 - kb0;
  const int a_atoms = (a.F + 31) / 32;

  if (warp == 8 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 256 + 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows of the A tile that no TMA box fills must still be finite (their products land in
  // accumulator rows that are never stored, but NaN garbage is poor hygiene): zero once.
  for (int i = threadIdx.x; i < STAGES * (A_BYTES / 16); i += blockDim.x) {
    int s = i / (A_BYTES / 16), o = i - s * (A_BYTES / 16);
    if (o >= a_atoms * (BLOCK_K * 128 / 16))
      *reinterpret_cast<float4*>(smem + s * STAGE_BYTES + o * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ---------------- B gather (8 warps): thread t owns 16-byte chunk j = t % 8 of pixel row
    // t/8 in each of the BN/32 atoms
    const int t = threadIdx.x;
    const int j = t & 7;
    constexpr int NAT = BN / 32;
    // tap offsets of this thread's chunk in every atom are fixed for the whole kernel
    long long koff[NAT];
    bool kvalid[NAT];
#pragma unroll
    for (int at = 0; at < NAT; ++at) {
      int k = n0 + 32 * at + 4 * j;
      kvalid[at] = k < a.K;
      if (IN_U8) {
        int kw = k % a.KH, kh = (k / a.KH) % a.KH, c = k / (a.KH * a.KH);
        koff[at] = ((long long)c * a.H + kh) * a.W + kw;
      } else {
        int c = k % a.C, tap = k / a.C;
        int kw = tap % a.KH, kh = tap / a.KH;
        koff[at] = ((long long)kh * a.W + kw) * a.C + c;
      }
    }
    auto gather = [&](int kb, float4* v) {
      {
        const int half = 0;
        int p = (kb0 + kb) * BLOCK_K + (t >> 3);
        long long base = -1;
        if (p < a.P) {
          int ow = p % a.OW;
          int oh = (p / a.OW) % a.OH;
          long long img = p / (a.OW * a.OH);
          base = IN_U8 ? (img * a.C * a.H + (long long)oh * a.S) * a.W + (long long)ow * a.S
                       : ((img * a.H + (long long)oh * a.S) * a.W + (long long)ow * a.S) * a.C;
        }
#pragma unroll
        for (int at = 0; at < NAT; ++at) {
          float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
          if (base >= 0 && kvalid[at]) {
            if (IN_U8) {
              uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(a.in) + base + koff[at]));
              r = make_float4(__fmul_rn((float)(u & 0xff), a.scale), __fmul_rn((float)((u >> 8) & 0xff), a.scale),
                              __fmul_rn((float)((u >> 16) & 0xff), a.scale), __fmul_rn((float)(u >> 24), a.scale));
            } else {
              r = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(a.in) + base + koff[at]));
            }
          }
          v[half * NAT + at] = r;
        }
      }
    };
    if (!IN_U8) {
      // fp32 NHWC input: LDGSTS straight into the swizzled atoms, all free stages in flight
      const float* in = static_cast<const float*>(a.in);
      const int r = t >> 3;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int p = (kb0 + kb) * BLOCK_K + r;
        long long base = -1;
        if (p < a.P) {
          int ow = p % a.OW;
          int oh = (p / a.OW) % a.OH;
          long long img = p / (a.OW * a.OH);
          base = ((img * a.H + (long long)oh * a.S) * a.W + (long long)ow * a.S) * a.C;
        }
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const uint32_t sb = smem_u32(smem + s * STAGE_BYTES + A_BYTES);
#pragma unroll
        for (int at = 0; at < NAT; ++at) {
          const bool ok = base >= 0 && kvalid[at];
          // 128B_BASE32B swizzle: 32-byte chunk index XOR (row % 4), 16-byte half preserved
          cp_async16(sb + at * (BLOCK_K * 128) + r * 128 + ((((j >> 1) ^ (r & 3)) << 5) | ((j & 1) << 4)),
                     ok ? in + base + koff[at] : in, ok ? 16u : 0u);
        }
        cp_async_arrive(&full_bar[s]);
      }
    } else {
    float4 cur[NAT], nx1[NAT], nx2[NAT];
    if (num_kb > 0) gather(0, cur);
    if (num_kb > 1) gather(1, nx1);
    for (int kb = 0; kb < num_kb; ++kb) {
      if (kb + 2 < num_kb) gather(kb + 2, nx2);
      int s = kb % STAGES;
      uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* sb = smem + s * STAGE_BYTES + A_BYTES;
      {
        const int half = 0;
        int r = (t >> 3);
#pragma unroll
        for (int at = 0; at < NAT; ++at)
          // 128B_BASE32B swizzle: 32-byte chunk index XOR (row % 4), 16-byte half preserved
          *reinterpret_cast<float4*>(sb + at * (BLOCK_K * 128) + r * 128 + ((((j >> 1) ^ (r & 3)) << 5) | ((j & 1) << 4))) =
              cur[half * NAT + at];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
#pragma unroll
      for (int q = 0; q < NAT; ++q) { cur[q] = nx1[q]; nx1[q] = nx2[q]; }
    }
    }
    if (warp < 4) {
    // ---------------- epilogue: rows < F of the accumulator -> raw split-K partial
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* stage = reinterpret_cast<float*>(smem) + warp * (32 * 36);
    EpiArgs e;
    e.C = a.ws + (size_t)blockIdx.z * a.F * a.K; e.ldc = a.K; e.M = a.F; e.N = a.K; e.alpha = 1.f;
    e.bias = nullptr; e.bias2 = nullptr; e.relu = 0; e.mask = nullptr; e.ldmask = 0; e.accumulate = 0;
    e.round_tf32 = 0; e.raw = 1;
    const EpiWarp ew = epi_begin(e, lane, warp * 32, n0, BN);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
      if (warp * 32 < a.F && n0 + c * 32 < a.K) epilogue_chunk(e, ew, v, stage, lane, warp * 32, n0 + c * 32);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  } else if (warp == 8) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        int s = kb % STAGES;
        uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], a_atoms * (BLOCK_K * 128));
        for (int at = 0; at < a_atoms; ++at)
          tma_load_2d(&tmA, &full_bar[s], smem + s * STAGE_BYTES + at * (BLOCK_K * 128), 32 * at,
                      (kb0 + kb) * BLOCK_K);
      }
    }
  } else {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        int s = kb % STAGES;
        uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        if (!IN_U8) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async writes -> MMA reads
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
          umma_tf32(tmem_base, make_smem_desc(sa + k * 1024, BLOCK_K * 128, 512, 1),
                    make_smem_desc(sb + k * 1024, BLOCK_K * 128, 512, 1), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full);
    }
  }
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
}

}  // namespace rttc
