// Persistent tensor-core LSTM recurrence for sm_100a (rltime/models/torch/modules/lstm.py:84-116).
//
// One cooperative launch runs EVERY recurrence of a learner update side by side: up to two
// weight sets ("groups": target net, online net), up to four sequences of B <= 32 rows per group
// (the double-Q selection pass and the training pass share the online weights).  The three
// 20-step recurrences of one update are therefore 20 dependent steps, not 60.
//
// Per group the 4U gate columns are split over U/UPC CTAs.  A CTA keeps its [4*UPC x U] slice of
// W_hh resident in 128B-swizzled shared memory for the whole launch (B operand, K-major) and per
// step pulls the masked carry-in h_{t-1} of all its sequences into shared memory as the A
// operand: sequence q occupies rows [32q, 32q+B) of the 128-row tile, so its gate sums land in
// TMEM lane quarter q and epilogue warp q owns them:
//   D[128 x 4*UPC] (TMEM, fp32) = hprev_t[128 x U] . W_slice^T        tcgen05.mma.kind::tf32
//   lane b: gates = act(xg_t[b] + D[b]), c = f*c_prev*keep + i*g, h = o*tanh(c)
// and writes h_t * keep_{t+1} into the exchange block of step t+1.
//
// Exchange buffer (global, L2-resident): X[group][step][k-block][row][32 floats], stored ALREADY
// in the 128B-swizzled byte order the MMA descriptor expects, so a step's operand is one
// contiguous region and the producer fetches it with plain bulk copies (cp.async.bulk, one per
// k-block).  Measured on B200: tiled-mode TMA boxes of 32 rows x 128 B at a 2 KB pitch took
// ~6400 cycles for 64 KB when 64 CTAs pull the same rows; contiguous bulk copies of the same
// bytes land in under 2000.
// The K loop rotates over NCH independent TMEM accumulators: back-to-back tcgen05.mma into ONE
// accumulator retire every ~116 cycles at N = 32 (dependent-accumulate latency), rotating hides it;
// the epilogue adds the partial sums.
// Steps are separated by a release/acquire counter per group (red.release + ld.acquire, no
// returning atomic); the producer orders the acquired generic-proxy writes before its
// async-proxy reads with fence.proxy.async.
#pragma once
#include "rt_gemm_tc.cuh"

namespace rttc {

constexpr int LSTM_MAX_SEQ = 4;
constexpr int LSTM_TC_THREADS = 192;
constexpr int LSTM_NCH = 4;      // independent accumulator chains (= UMMA_K sub-steps per k-block)

struct LstmSeq {
  const float* xg;        // (T*B, 4U) x W_ih^T + b_ih + b_hh, row t*B + b
  const float* hx;        // (B, U) stored state of step 0
  const float* cx;
  const float* initials;  // (T*B)
  float* h_all;           // (T*B, U)
  float* gates;           // (T*B, 4U) activated gates, or null (only BPTT needs them)
  float* c_all;           // (T*B, U) or null
  float* cprev;           // (T*B, U) masked carry-in cell state, or null
  float* hprev;           // (T*B, U) masked carry-in hidden state (row-major, for dW_hh), or null
};

struct LstmTcArgs {
  LstmSeq seq[2][LSTM_MAX_SEQ];
  int nseq[2];
  int T, B, U;
  int exp;                 // tuning experiments: 1 = skip the MMAs, 2 = skip the per-step loads
  int rn;                  // round every h that becomes an MMA operand to the nearest TF32 value
  int arows;               // 32 * max sequences per group: rows per k-block of the A operand
  float* xchg;             // exchange buffer: [2 groups][T][U/32][arows][32] floats, swizzled
  unsigned int* counters;  // one per group, 32 words apart, zeroed before the launch
  long long* dbg;          // optional clock64 stamps of the first CTA of each group
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// 1 - 2/(1+e^{2x}): absolute error ~1e-7 over the whole range (saturates cleanly at +-1)
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Shared memory: resident W slice [U/32][4*UPC rows][128 B], then the WHOLE h_{t-1} operand of one
// step, [U/32][arows][128 B], so every copy of a step is in flight at once.  The MMA reads 128-row
// tiles: rows >= arows alias the next k-block / the tail pad and only feed accumulator lanes
// nobody reads.
template <int UPC>
struct LstmSmem {
  static constexpr int N = 4 * UPC;
  static int total(int U, int arows) {
    return U * N * 4 + (U / BLOCK_K) * arows * 128 + (BLOCK_M - arows) * 128 + (U / BLOCK_K + 2) * 8 + 16 + 1024;
  }
};

// float offset of element (row r of the 128-row tile, hidden unit k) inside one step's block
__device__ __forceinline__ size_t xchg_off(int arows, int r, int k) {
  const int kb = k >> 5, c = (k & 31) >> 2;
  return (size_t)kb * arows * 32 + (size_t)r * 32 + (size_t)((c ^ (r & 7)) << 2) + (k & 3);
}

template <int UPC>
__global__ void __launch_bounds__(LSTM_TC_THREADS)
k_lstm_seq_tc(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ LstmTcArgs a) {
  constexpr int N = 4 * UPC;
  constexpr int TCOLS = LSTM_NCH * N;                   // 128 or 256 TMEM columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int U = a.U, T = a.T, B = a.B;
  const int KB = U / BLOCK_K;
  uint8_t* wsm = smem;                                  // [KB][N rows][128 B]
  const int ablk = a.arows * 128;                       // bytes of one k-block of the A operand
  uint8_t* asm_ = smem + (size_t)U * N * 4;             // [KB][arows][128 B] + tail pad
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(asm_ + (size_t)KB * ablk + (BLOCK_M - a.arows) * 128);
  uint64_t* w_full = full_bar + KB;
  uint64_t* tmem_full = w_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpg = U / UPC;                              // CTAs per group
  const int grp = blockIdx.x / cpg;
  const int u0 = (blockIdx.x - grp * cpg) * UPC;
  const int nseq = a.nseq[grp];
  long long* const dbg = a.dbg ? a.dbg + grp * 512 : nullptr;
  unsigned int* counter = a.counters + grp * 32;
  const unsigned int per_step = (unsigned int)(cpg * nseq);
  const bool dbg_cta = a.dbg && u0 == 0;   // first CTA of each group stamps its own 8 x 64 block
  const size_t step_floats = (size_t)KB * a.arows * 32;
  float* const xg_ = a.xchg + (size_t)grp * T * step_floats;   // this group's exchange blocks

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < KB; ++s) mbar_init(&full_bar[s], 1);
    mbar_init(w_full, 1);
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- producer: resident W slice once, then the h_{t-1} block of every step
      const CUtensorMap* tmW = grp == 0 ? &tmW0 : &tmW1;
      mbar_expect_tx(w_full, (uint32_t)(U * N * 4));
      for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int g = 0; g < 4; ++g)
          tma_load_2d(tmW, w_full, wsm + (size_t)kb * (N * 128) + g * (UPC * 128), kb * BLOCK_K, g * U + u0);
      const uint32_t cp_bytes = (uint32_t)(nseq * 32 * 128);
      for (int t = 0; t < T; ++t) {
        // publication t+1 (publication 1 = the step-0 block written in this kernel's prologue)
        const unsigned int want = (unsigned int)(t + 1) * per_step;
        unsigned int v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < want);
        // the block was written through the generic proxy by other SMs: order those writes before
        // this SM's async-proxy reads.  Every MMA of step t-1 has retired (its epilogue ran
        // before the publication), so the operand buffers are free.
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (dbg_cta) dbg[8 * t + 0] = clock64();
        const float* src = xg_ + (size_t)t * step_floats;
        for (int kb = 0; kb < KB; ++kb) {
          if (a.exp == 2) {
            mbar_expect_tx(&full_bar[kb], 0u);
            continue;
          }
          mbar_expect_tx(&full_bar[kb], cp_bytes);
          bulk_load(asm_ + (size_t)kb * ablk, src + (size_t)kb * a.arows * 32, cp_bytes, &full_bar[kb]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(BLOCK_M >> 4) << 24);
      mbar_wait(w_full, 0);
      const uint32_t wbase = smem_u32(wsm);
      const uint32_t abase = smem_u32(asm_);
      for (int t = 0; t < T; ++t) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full_bar[kb], (uint32_t)(t & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (a.exp == 1) continue;
          const uint32_t sa = abase + (uint32_t)(kb * ablk);
          const uint32_t sb = wbase + (uint32_t)kb * (N * 128);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            uint64_t ad = make_smem_desc(sa + k * (UMMA_K * 4), 16, 1024, 2);
            uint64_t bd = make_smem_desc(sb + k * (UMMA_K * 4), 16, 1024, 2);
            // sub-step k accumulates into its own column block: consecutive MMAs are independent
            umma_tf32(tmem_base + (uint32_t)(k * N), ad, bd, idesc, kb > 0 ? 1u : 0u);
          }
        }
        umma_commit(tmem_full);
        if (dbg_cta) dbg[8 * t + 1] = clock64();
      }
    }
  } else {
    // ---------------- epilogue: warp quarter q owns sequence q (TMEM lanes [32q, 32q+32))
    const int q = warp & 3;
    if (q < nseq) {
      const LstmSeq sq = a.seq[grp][q];
      const int b = lane;
      const int r = q * 32 + b;                    // row of the A tile
      const bool act = b < B;
      float c[UPC];
      {
        // publication 1: step-0 block = stored h * keep_0 (lstm.py:67-70, 95-98)
        const float keep0 = act ? 1.f - sq.initials[b] : 0.f;
        if (act) {
#pragma unroll
          for (int j = 0; j < UPC; j += 4) {
            float4 h4 = *reinterpret_cast<const float4*>(sq.hx + (size_t)b * U + u0 + j);
            float4 c4 = *reinterpret_cast<const float4*>(sq.cx + (size_t)b * U + u0 + j);
            h4 = make_float4(h4.x * keep0, h4.y * keep0, h4.z * keep0, h4.w * keep0);
            if (a.rn) h4 = make_float4(rtk::rna_tf32(h4.x), rtk::rna_tf32(h4.y), rtk::rna_tf32(h4.z), rtk::rna_tf32(h4.w));
            *reinterpret_cast<float4*>(xg_ + xchg_off(a.arows, r, u0 + j)) = h4;
            if (sq.hprev) *reinterpret_cast<float4*>(sq.hprev + (size_t)b * U + u0 + j) = h4;
            c[j] = c4.x; c[j + 1] = c4.y; c[j + 2] = c4.z; c[j + 3] = c4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < UPC; ++j) c[j] = 0.f;
        }
        __syncwarp();
        // release: cumulative over the warp's stores ordered by the __syncwarp above
        if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
      }
      for (int t = 0; t < T; ++t) {
        const size_t row = (size_t)t * B + b;
        // issue this step's input-gate loads before waiting on the tensor core
        float xin[N];
        float keep = 0.f, keep_next = 0.f;
        if (act) {
          const float* xr = sq.xg + row * 4 * U + u0;
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int j = 0; j < UPC; j += 4) {
              float4 x4 = __ldg(reinterpret_cast<const float4*>(xr + (size_t)g * U + j));
              xin[g * UPC + j] = x4.x; xin[g * UPC + j + 1] = x4.y; xin[g * UPC + j + 2] = x4.z;
              xin[g * UPC + j + 3] = x4.w;
            }
          keep = 1.f - sq.initials[row];
          if (t + 1 < T) keep_next = 1.f - sq.initials[row + B];
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) xin[i] = 0.f;
        }
        mbar_wait(tmem_full, (uint32_t)(t & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg_cta && q == 0 && lane == 0) dbg[8 * t + 2] = clock64();
        // gate pre-activations = xin + sum of the NCH partial accumulators
#pragma unroll
        for (int ch = 0; ch < LSTM_NCH; ++ch)
#pragma unroll
          for (int cc = 0; cc < N / 32; ++cc) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * N + cc * 32), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) xin[cc * 32 + i] += __uint_as_float(v[i]);
          }
        // the accumulators are in registers: the next step's MMAs may overwrite them
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (dbg_cta && q == 0 && lane == 0) dbg[8 * t + 5] = clock64();
        float hv[UPC], cpv[UPC];
        if (act) {
#pragma unroll
          for (int j = 0; j < UPC; ++j) {
            const float gi = fast_sigmoid(xin[j]);
            const float gf = fast_sigmoid(xin[UPC + j]);
            const float gg = fast_tanh(xin[2 * UPC + j]);
            const float go = fast_sigmoid(xin[3 * UPC + j]);
            const float cp = c[j] * keep;
            const float cn = gf * cp + gi * gg;
            c[j] = cn;
            cpv[j] = cp;
            hv[j] = go * fast_tanh(cn);
            if (a.rn) hv[j] = rtk::rna_tf32(hv[j]);   // h_t feeds the next step's MMA and the heads GEMMs
            xin[j] = gi; xin[UPC + j] = gf; xin[2 * UPC + j] = gg; xin[3 * UPC + j] = go;
          }
          // exchange block of step t+1 first: it is on the critical path of every CTA
          if (t + 1 < T) {
            float* xn = xg_ + (size_t)(t + 1) * step_floats;
#pragma unroll
            for (int j = 0; j < UPC; j += 4)
              *reinterpret_cast<float4*>(xn + xchg_off(a.arows, r, u0 + j)) =
                  make_float4(hv[j] * keep_next, hv[j + 1] * keep_next, hv[j + 2] * keep_next, hv[j + 3] * keep_next);
          }
        }
        if (dbg_cta && q == 0 && lane == 0) dbg[8 * t + 6] = clock64();
        if (t + 1 < T) {
          __syncwarp();
          if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        }
        if (dbg_cta && q == 0 && lane == 0) dbg[8 * t + 4] = clock64();
        // everything else is off the critical path
        if (act) {
          float* ho = sq.h_all + row * U + u0;
#pragma unroll
          for (int j = 0; j < UPC; j += 4)
            *reinterpret_cast<float4*>(ho + j) = make_float4(hv[j], hv[j + 1], hv[j + 2], hv[j + 3]);
          if (sq.gates) {
            float* gr = sq.gates + row * 4 * U + u0;
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int j = 0; j < UPC; j += 4)
                *reinterpret_cast<float4*>(gr + (size_t)g * U + j) =
                    make_float4(xin[g * UPC + j], xin[g * UPC + j + 1], xin[g * UPC + j + 2], xin[g * UPC + j + 3]);
            float* co = sq.c_all + row * U + u0;
            float* cpo = sq.cprev + row * U + u0;
#pragma unroll
            for (int j = 0; j < UPC; j += 4) {
              *reinterpret_cast<float4*>(co + j) = make_float4(c[j], c[j + 1], c[j + 2], c[j + 3]);
              *reinterpret_cast<float4*>(cpo + j) = make_float4(cpv[j], cpv[j + 1], cpv[j + 2], cpv[j + 3]);
            }
            if (t + 1 < T) {
              float* hp = sq.hprev + (row + B) * U + u0;
#pragma unroll
              for (int j = 0; j < UPC; j += 4)
                *reinterpret_cast<float4*>(hp + j) = make_float4(hv[j] * keep_next, hv[j + 1] * keep_next,
                                                                 hv[j + 2] * keep_next, hv[j + 3] * keep_next);
            }
          }
        }
        if (dbg_cta && q == 0 && lane == 0) dbg[8 * t + 3] = clock64();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TCOLS)
                 : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// Same recurrence, recurrent product on mma.sync (m16n8k8 TF32, fp32 accumulate) instead of tcgen05.
// Why: with M = 32..64 batch rows per CTA the tcgen05 version is bound by instruction issue, not by
// math -- 64 dependent 128 x 32 x 8 MMAs per step at the ~115-cycle SS-mode floor = ~7400 cycles per
// step, 3/4 of the 128 accumulator rows unused.  Here the [4*UPC x U] slice of W_hh lives in the
// REGISTERS of eight compute warps as mma B fragments (warp w owns k in [64 w, 64 w + 64)), the staged
// h_{t-1} tile (same pre-swizzled exchange blocks, same bulk copies) supplies the A fragments with
// conflict-free LDS, and the eight K-partials are folded through shared memory in a fixed order by the
// thread that then runs the cell for those outputs:
//   producer warp (1)  : waits for the step counter, bulk-loads the h_{t-1} blocks          [as above]
//   compute warps (8)  : 16*MT x 32 x 64 partial product each -> smem; thread (row r, unit pair p) folds
//                        its 2 units x 4 gates, updates c / h, writes the exchange block of t+1
// UPC = 8 (N = 32: n-tile == gate), up to two sequences per weight group (MT = 2 or 4 m-tiles).
constexpr int LSTM_MMA_COMPUTE_WARPS = 8;
constexpr int LSTM_MMA_THREADS = 32 * (1 + LSTM_MMA_COMPUTE_WARPS);
constexpr int LSTM_FIN_PITCH = 36;     // padded row pitch (floats) of the partial tiles

struct LstmMmaSmem {
  static int total(int U, int arows) {
    return (U / BLOCK_K) * arows * 128                                   // h tile
           + LSTM_MMA_COMPUTE_WARPS * arows * LSTM_FIN_PITCH * 4         // K-partials
           + (U / BLOCK_K + 2) * 8 + 16 + 1024;
  }
};

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int MT>   // m-tiles of 16 rows: 2 (one sequence per group) or 4 (two)
__global__ void __launch_bounds__(LSTM_MMA_THREADS, 1)
k_lstm_seq_mma(const __grid_constant__ LstmTcArgs a, const float* __restrict__ whh0, const float* __restrict__ whh1) {
  constexpr int UPC = 8;
  constexpr int AROWS = 16 * MT;
  constexpr int NCW = LSTM_MMA_COMPUTE_WARPS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int U = a.U, T = a.T, B = a.B;
  const int KB = U / BLOCK_K;
  float* asm_ = reinterpret_cast<float*>(smem);                               // [KB][AROWS][32] swizzled
  float* part = asm_ + (size_t)KB * AROWS * 32;                               // [8][AROWS][36]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(part + (size_t)NCW * AROWS * LSTM_FIN_PITCH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpg = U / UPC;                              // CTAs per group
  const int grp = blockIdx.x / cpg;
  const int u0 = (blockIdx.x - grp * cpg) * UPC;
  const int nseq = a.nseq[grp];
  unsigned int* counter = a.counters + grp * 32;
  const unsigned int per_step = (unsigned int)(cpg * NCW);      // one publication per compute warp
  const size_t step_floats = (size_t)KB * a.arows * 32;
  float* const xg_ = a.xchg + (size_t)grp * T * step_floats;   // this group's exchange blocks
  const int ablk = a.arows * 128;                       // bytes of one k-block of the exchange layout

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < KB; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // rows of the tile no copy ever fills (a group with fewer sequences than the other) feed accumulator rows
  // nobody reads; zero them once so they stay finite
  for (int i = threadIdx.x; i < KB * AROWS * 8; i += blockDim.x) {
    const int rr = (i >> 3) % AROWS;
    if (rr >= nseq * 32) reinterpret_cast<float4*>(asm_)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- producer: the h_{t-1} block of every step: the rows of THIS group's sequences of
      // every k-block (the exchange traffic -- every CTA pulls its group's whole h -- is what bounds a step:
      // 128 CTAs x 128 KB = 16 MB through L2 if both groups loaded two sequences' worth)
      const uint32_t cp_bytes = (uint32_t)(nseq * 32 * 128);
      for (int t = 0; t < T; ++t) {
        const unsigned int want = (unsigned int)(t + 1) * per_step;
        unsigned int v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < want);
        asm volatile("fence.proxy.async.global;" ::: "memory");
        const uint8_t* src = reinterpret_cast<const uint8_t*>(xg_ + (size_t)t * step_floats);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_expect_tx(&full_bar[kb], cp_bytes);
          bulk_load(reinterpret_cast<uint8_t*>(asm_) + (size_t)kb * (AROWS * 128), src + (size_t)kb * ablk, cp_bytes,
                    &full_bar[kb]);
        }
      }
    }
    return;
  }
  // ---------------- compute warps: k in [64 w, 64 w + 64)
  const int w = warp - 1;
  const int g = lane >> 2, tq = lane & 3;
  const float* whh = grp == 0 ? whh0 : whh1;
  // resident B fragments: k-step s (8 per warp), n-tile nt == gate nt: W[gate nt, unit u0 + g][k]
  uint32_t wb[8][4][2];
#pragma unroll
  for (int s = 0; s < 8; ++s)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float* wr = whh + (size_t)(nt * U + u0 + g) * U + 64 * w + 8 * s + tq;
      wb[s][nt][0] = __float_as_uint(__ldg(wr));
      wb[s][nt][1] = __float_as_uint(__ldg(wr + 4));
    }
  // cell ownership: thread ti of the 256 -> row r = ti / 4 (+ 64 for the second half when AROWS > 64 never
  // happens: AROWS <= 64), unit pair p = ti % 4 (units u0 + 2p, u0 + 2p + 1)
  const int ti = w * 32 + lane;
  const int r = ti >> 2, p2 = (ti & 3) * 2;
  const int q = r >> 5, b = r & 31;
  const bool act = r < AROWS && q < nseq && b < B;
  const LstmSeq sq = a.seq[grp][q < LSTM_MAX_SEQ ? q : 0];
  float c0 = 0.f, c1 = 0.f;
  {
    // publication 1: step-0 block = stored h * keep_0 (lstm.py:67-70, 95-98)
    if (act) {
      const float keep0 = 1.f - sq.initials[b];
      float2 h2 = *reinterpret_cast<const float2*>(sq.hx + (size_t)b * U + u0 + p2);
      const float2 cc = *reinterpret_cast<const float2*>(sq.cx + (size_t)b * U + u0 + p2);
      h2 = make_float2(h2.x * keep0, h2.y * keep0);
      if (a.rn) h2 = make_float2(rtk::rna_tf32(h2.x), rtk::rna_tf32(h2.y));
      *reinterpret_cast<float2*>(xg_ + xchg_off(a.arows, r, u0 + p2)) = h2;
      if (sq.hprev) *reinterpret_cast<float2*>(sq.hprev + (size_t)b * U + u0 + p2) = h2;
      c0 = cc.x; c1 = cc.y;
    }
    __syncwarp();
    if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
  }
  for (int t = 0; t < T; ++t) {
    const size_t row = (size_t)t * B + b;
    // this step's input gates and masks: issued before waiting for the operands
    float2 xin[4];
    float keep = 0.f, keep_next = 0.f;
    if (act) {
      const float* xr = sq.xg + row * 4 * U + u0 + p2;
#pragma unroll
      for (int gt = 0; gt < 4; ++gt) xin[gt] = __ldg(reinterpret_cast<const float2*>(xr + (size_t)gt * U));
      keep = 1.f - sq.initials[row];
      if (t + 1 < T) keep_next = 1.f - sq.initials[row + B];
    } else {
#pragma unroll
      for (int gt = 0; gt < 4; ++gt) xin[gt] = make_float2(0.f, 0.f);
    }
    float acc[MT][4][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
    for (int kbl = 0; kbl < 2; ++kbl) {
      const int kb = 2 * w + kbl;
      mbar_wait(&full_bar[kb], (uint32_t)(t & 1));
      const float* ablock = asm_ + (size_t)kb * AROWS * 32;
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {          // k-step inside the k-block: k = 8 s4 + tq (+4)
        const int s = 4 * kbl + s4;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (mt >= 2 * nseq) continue;             // m-tiles of sequences this group does not have
          const int r0 = 16 * mt + g, r1 = r0 + 8;
          uint32_t af[4];
          // 128B swizzle of the exchange layout: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
          af[0] = __float_as_uint(ablock[r0 * 32 + (((2 * s4) ^ (r0 & 7)) << 2) + tq]);
          af[1] = __float_as_uint(ablock[r1 * 32 + (((2 * s4) ^ (r1 & 7)) << 2) + tq]);
          af[2] = __float_as_uint(ablock[r0 * 32 + (((2 * s4 + 1) ^ (r0 & 7)) << 2) + tq]);
          af[3] = __float_as_uint(ablock[r1 * 32 + (((2 * s4 + 1) ^ (r1 & 7)) << 2) + tq]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32_16x8x8(acc[mt][nt], af, wb[s][nt][0], wb[s][nt][1]);
        }
      }
    }
    // every compute warp is done folding the previous step's partials before they are overwritten
    asm volatile("bar.sync 1, %0;" ::"n"(32 * NCW) : "memory");
    {
      float* pw = part + (size_t)w * AROWS * LSTM_FIN_PITCH;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int r0 = 16 * mt + g, cc = 8 * nt + 2 * tq;
          *reinterpret_cast<float2*>(pw + r0 * LSTM_FIN_PITCH + cc) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
          *reinterpret_cast<float2*>(pw + (r0 + 8) * LSTM_FIN_PITCH + cc) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
        }
    }
    asm volatile("bar.sync 2, %0;" ::"n"(32 * NCW) : "memory");
    if (r < AROWS) {
      // fold the 8 K-partials of (row r, gate gt, units p2, p2+1) in warp order, add the input gates
#pragma unroll
      for (int gt = 0; gt < 4; ++gt) {
        float2 sum = *reinterpret_cast<const float2*>(part + r * LSTM_FIN_PITCH + 8 * gt + p2);
#pragma unroll
        for (int ww = 1; ww < NCW; ++ww) {
          const float2 v = *reinterpret_cast<const float2*>(part + ((size_t)ww * AROWS + r) * LSTM_FIN_PITCH + 8 * gt + p2);
          sum.x += v.x; sum.y += v.y;
        }
        xin[gt].x += sum.x; xin[gt].y += sum.y;
      }
    }
    float h0 = 0.f, h1 = 0.f, cp0 = 0.f, cp1 = 0.f;
    if (act) {
      const float gi0 = fast_sigmoid(xin[0].x), gi1 = fast_sigmoid(xin[0].y);
      const float gf0 = fast_sigmoid(xin[1].x), gf1 = fast_sigmoid(xin[1].y);
      const float gg0 = fast_tanh(xin[2].x), gg1 = fast_tanh(xin[2].y);
      const float go0 = fast_sigmoid(xin[3].x), go1 = fast_sigmoid(xin[3].y);
      cp0 = c0 * keep; cp1 = c1 * keep;
      c0 = gf0 * cp0 + gi0 * gg0;
      c1 = gf1 * cp1 + gi1 * gg1;
      h0 = go0 * fast_tanh(c0);
      h1 = go1 * fast_tanh(c1);
      if (a.rn) { h0 = rtk::rna_tf32(h0); h1 = rtk::rna_tf32(h1); }
      xin[0] = make_float2(gi0, gi1); xin[1] = make_float2(gf0, gf1);
      xin[2] = make_float2(gg0, gg1); xin[3] = make_float2(go0, go1);
      // exchange block of step t+1 first: it is on the critical path of every CTA
      if (t + 1 < T)
        *reinterpret_cast<float2*>(xg_ + (size_t)(t + 1) * step_floats + xchg_off(a.arows, r, u0 + p2)) =
            make_float2(h0 * keep_next, h1 * keep_next);
    }
    if (t + 1 < T) {
      __syncwarp();
      if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
    }
    // everything else is off the critical path
    if (act) {
      *reinterpret_cast<float2*>(sq.h_all + row * U + u0 + p2) = make_float2(h0, h1);
      if (sq.gates) {
        float* gr = sq.gates + row * 4 * U + u0 + p2;
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) *reinterpret_cast<float2*>(gr + (size_t)gt * U) = xin[gt];
        *reinterpret_cast<float2*>(sq.c_all + row * U + u0 + p2) = make_float2(c0, c1);
        *reinterpret_cast<float2*>(sq.cprev + row * U + u0 + p2) = make_float2(cp0, cp1);
        if (t + 1 < T)
          *reinterpret_cast<float2*>(sq.hprev + (row + B) * U + u0 + p2) = make_float2(h0 * keep_next, h1 * keep_next);
      }
    }
  }
}

}  // namespace rttc
