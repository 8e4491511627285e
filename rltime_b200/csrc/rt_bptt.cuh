// One-launch BPTT recurrence (rltime/models/torch/modules/lstm.py:84-116 backward).
//
// Replaces the 2 x (T-1) launches of the stepwise path (k_lstm_cell_bwd + split-K GEMM per step)
// with ONE cooperative launch.  Per step t = T-1 .. 0:
//     dh_t     = dout_t + (dgates_{t+1} . W_hh) * (1 - initial_{t+1})
//     dgates_t = cell backward(dh_t, dc carry, gates_t, c_t, cprev_t)
// The recurrent product [B x 4U] . [4U x U] is split 2-D over KS x NS CTAs:
//     K-slice i  = the 4 gate rows of the units [64 i, 64 i + 64)   (256 rows of W_hh)
//     N-slice j  = the units [32 j, 32 j + 32)                       (32 columns of W_hh)
// CTA (i, j) keeps its 256 x 32 block of W_hh in REGISTERS for the whole sequence as the B
// fragments of 32 mma.sync.m16n8k8 (TF32, fp32 accumulate) per warp: warp w owns the 16 x 8 output
// tile (batch rows 16 (w & 1).., columns 8 (w >> 1)..) over all 256 k.  Each step it
//   (cell)  finishes dh for its (B x 32/KS) share of N-slice j from the KS partials of the previous
//           product (fixed order), runs the cell backward keeping dc in a register, writes dgates_t;
//   (gemm)  stages dgates_t[:, K-slice i] (32 KB) in shared memory, multiplies, writes its
//           partial [B x 32] to part[i].
// Synchronisation is by producer group, not by grid: the dgates of K-slice i are written by the 2 KS
// CTAs (*, 2i) and (*, 2i+1) -> counter cell_done[i]; the partials of N-slice j by the 8 CTAs (*, j) ->
// counter part_done[j] (release / acquire on global counters 128 B apart, monotone over the steps).
// Everything a cell needs that does not depend on the recurrence (gates, c, c_prev, dout, masks) is
// loaded BEFORE it waits for the partials.  Partial buffers alternate by step parity, so a CTA two
// phases ahead never overwrites partials a slower CTA still reads.
// B == 32, U in {256, 512}; grid = (U/64) * (U/32) CTAs, all co-resident (cooperative launch).
#pragma once
#include <cstdint>

namespace rtbptt {

constexpr int THREADS = 256;
constexpr int KROWS = 256;             // gate rows per K-slice (4 gates x 64 units)
// The dgates tile is staged in two halves of 128 k (2 gates x 64 units): 16.9 KB of shared memory instead of
// 33.3 KB, so that a CTA of this latency-bound kernel fits on an SM NEXT TO two CTAs of the weight-gradient
// GEMMs of the side branch (2 x 97 KB) -- otherwise the cooperative launch waits for that branch's first
// GEMM to drain (measured: 75 us on the critical path).
constexpr int KHALF = KROWS / 2;
constexpr int DG_PITCH = KHALF + 4;    // padded row pitch of the staged half tile: conflict-free fragment loads
constexpr int CTR_STRIDE = 32;         // counters 128 B apart

struct Args {
  const float* dout;       // (T*B, U)   d(loss)/d(h_t)
  const float* gates;      // (T*B, 4U)  activated gates, column = gate * U + unit
  const float* c_all;      // (T*B, U)
  const float* cprev;      // (T*B, U)   masked carry-in cell state
  const float* initials;   // (T*B)
  const float* whh;        // (4U, U) row-major
  float* dgates;           // (T*B, 4U)  out
  float* part;             // (2, KS, B, U) scratch: partial products, alternating by step parity
  unsigned int* counter;   // [(KS + NS) * CTR_STRIDE] zeroed before the launch: cell_done[KS], part_done[NS]
  int T, B, U;
};

__device__ __forceinline__ void signal(unsigned int* ctr) {
  // every thread's prior global stores -> visible before the increment (CTA barrier + cumulative release)
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
}
__device__ __forceinline__ void wait_for(const unsigned int* ctr, unsigned int target) {
  if (threadIdx.x == 0) {
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KS>   // K-slices = U / 64 (8 at U = 512, 4 at U = 256)
__global__ void __launch_bounds__(THREADS, 2) k_lstm_bptt_p(const Args a) {
  extern __shared__ __align__(16) float dg_s[];   // [32][DG_PITCH]
  const int U = a.U, B = a.B, T = a.T;
  const int i = blockIdx.x % KS;          // K-slice
  const int j = blockIdx.x / KS;          // N-slice
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int U4 = 4 * U;
  unsigned int* cell_done = a.counter;                          // [KS]
  unsigned int* part_done = a.counter + KS * CTR_STRIDE;        // [NS]
  const size_t part_words = (size_t)KS * B * U;

  // ---- resident W block as mma B fragments: k-step s covers slice rows kk = 8 s .. 8 s + 7
  const int mi = warp & 1, ni = warp >> 1;
  const int ncol = 32 * j + 8 * ni + (lane >> 2);           // this lane's output column (unit)
  uint32_t wb[32][2];
#pragma unroll
  for (int s = 0; s < 32; ++s) {
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int kk = 8 * s + (lane & 3) + 4 * h2;           // row inside the slice
      const int krow = (kk >> 6) * U + 64 * i + (kk & 63);  // gate * U + unit
      wb[s][h2] = __float_as_uint(a.whh[(size_t)krow * U + ncol]);
    }
  }

  // ---- cell ownership: (B x 32/KS) elements of N-slice j; consecutive threads = consecutive units
  constexpr int UPC = 32 / KS;                               // units per CTA in the cell phase
  const bool cell = tid < 32 * UPC;
  const int cb = tid / UPC;                                  // batch row
  const int cu = 32 * j + UPC * i + (tid % UPC);             // unit
  float dc_carry = 0.f;

  for (int t = T - 1; t >= 0; --t) {
    const size_t ro = (size_t)t * B;
    const int step = T - 1 - t;                              // 0, 1, ...
    // ---------------- cell backward of step t
    float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f, dh = 0.f, cv = 0.f, cpv = 0.f, keep_n = 0.f, keep = 0.f;
    size_t g0 = 0;
    if (cell) {
      // everything that does not depend on the recurrence: issued before the wait
      const size_t e = (ro + cb) * U + cu;
      g0 = (ro + cb) * U4 + cu;
      gi = __ldg(a.gates + g0); gf = __ldg(a.gates + g0 + U); gg = __ldg(a.gates + g0 + 2 * U); go = __ldg(a.gates + g0 + 3 * U);
      dh = __ldg(a.dout + e);
      cv = __ldg(a.c_all + e);
      cpv = __ldg(a.cprev + e);
      keep = 1.f - __ldg(a.initials + ro + cb);
      if (t < T - 1) keep_n = 1.f - __ldg(a.initials + ro + B + cb);
    }
    if (t < T - 1) {
      wait_for(part_done + j * CTR_STRIDE, (unsigned int)(KS * step));      // partials of step t+1 complete
      if (cell) {
        const float* pp = a.part + (size_t)((step - 1) & 1) * part_words + (size_t)cb * U + cu;
        float pv[KS];
#pragma unroll
        for (int p = 0; p < KS; ++p) pv[p] = __ldcg(pp + (size_t)p * B * U);
        float carry = 0.f;
#pragma unroll
        for (int p = 0; p < KS; ++p) carry += pv[p];
        dh += carry * keep_n;
      }
    }
    if (cell) {
      const float tc = tanhf(cv);
      const float dc = dc_carry + dh * go * (1.f - tc * tc);
      a.dgates[g0] = dc * gg * gi * (1.f - gi);
      a.dgates[g0 + U] = dc * cpv * gf * (1.f - gf);
      a.dgates[g0 + 2 * U] = dc * gi * (1.f - gg * gg);
      a.dgates[g0 + 3 * U] = dh * tc * go * (1.f - go);
      dc_carry = dc * gf * keep;
    }
    if (t == 0) break;
    signal(cell_done + (j >> 1) * CTR_STRIDE);                               // my share of K-slice j/2 is written
    wait_for(cell_done + i * CTR_STRIDE, (unsigned int)(2 * KS * (step + 1)));   // dgates_t[:, K-slice i] complete

    // ---------------- partial product: dgates_t[:, K-slice i] . W block, in two halves of 128 k
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    const float* arow0 = dg_s + (16 * mi + (lane >> 2)) * DG_PITCH + (lane & 3);
    const float* arow1 = arow0 + 8 * DG_PITCH;
    // both halves' loads are issued up front (registers), the second half is stored once the first is consumed
    float4 stage[2][(32 * (KHALF / 4)) / THREADS];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf)
#pragma unroll
      for (int it = 0; it < (32 * (KHALF / 4)) / THREADS; ++it) {
        const int v = tid + it * THREADS;
        const int b = v / (KHALF / 4), kk = hf * KHALF + (v % (KHALF / 4)) * 4;
        stage[hf][it] = __ldcg(reinterpret_cast<const float4*>(
            a.dgates + (ro + b) * U4 + (size_t)(kk >> 6) * U + 64 * i + (kk & 63)));
      }
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      if (hf) __syncthreads();                 // every warp is done reading the first half
#pragma unroll
      for (int it = 0; it < (32 * (KHALF / 4)) / THREADS; ++it) {
        const int v = tid + it * THREADS;
        const int b = v / (KHALF / 4), kk = (v % (KHALF / 4)) * 4;
        *reinterpret_cast<float4*>(dg_s + b * DG_PITCH + kk) = stage[hf][it];
      }
      __syncthreads();
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        uint32_t af[4];
        af[0] = __float_as_uint(arow0[8 * s]);
        af[1] = __float_as_uint(arow1[8 * s]);
        af[2] = __float_as_uint(arow0[8 * s + 4]);
        af[3] = __float_as_uint(arow1[8 * s + 4]);
        mma_tf32(d, af, wb[16 * hf + s][0], wb[16 * hf + s][1]);
      }
    }
    {
      const int b0 = 16 * mi + (lane >> 2);
      const int n0 = 32 * j + 8 * ni + 2 * (lane & 3);
      float* p0 = a.part + (size_t)(step & 1) * part_words + ((size_t)i * B + b0) * U + n0;
      *reinterpret_cast<float2*>(p0) = make_float2(d[0], d[1]);
      *reinterpret_cast<float2*>(p0 + (size_t)8 * U) = make_float2(d[2], d[3]);
    }
    signal(part_done + j * CTR_STRIDE);                                      // my partial of N-slice j is written
  }
}

}  // namespace rtbptt
