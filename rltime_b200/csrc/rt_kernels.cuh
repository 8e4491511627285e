// Learner kernels for sm_100a: fp32 GEMM (SIMT path), im2col / col2im, LSTM cell, IQN
// quantile embedding, dueling heads, double-Q target + value rescaling, fused quantile-Huber
// loss forward+backward, column reductions, fused grad-norm + clip + Adam.
//
// Activations are NHWC ((row, pixel), channel) so every convolution / linear layer is a
// GEMM C[M,N] = A[M,K] . W[N,K]^T on K-major operands — the layout the tcgen05 path in
// rt_gemm_tc.cuh consumes directly.  Reference call sites are cited per kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtk {

// fp32 -> nearest TF32 value (ties away), still an fp32 bit pattern.  tcgen05.mma.kind::tf32 reads
// only the upper 19 bits of its operands, i.e. TRUNCATES them (a systematic shrink of every product
// by ~3.5e-4 per operand); operands rounded here at their producer make the multiply an unbiased
// round-to-nearest TF32 product instead (RT_GEMM_TF32_RN).
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
  return __uint_as_float(t);
}

// ------------------------------------------------------------------------------- GEMM
// C[M,N] = epi( alpha * sum_k Aop[m,k] * Bop[k,n] )
//   Aop[m,k] = transA ? A[k*lda+m] : A[m*lda+k]
//   Bop[k,n] = transB ? B[n*ldb+k] : B[k*ldb+n]      (transB=1: nn.Linear weight [N,K])
// epi(v)[m,n] = (accumulate ? C_old : 0) + mask( relu( v + bias[n] + bias2[n] ) )
//   mask: multiplies by (mask[m*ldmask+n] > 0), i.e. the ReLU derivative of a saved output.
struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  int lda, ldb, ldc;
  int transA, transB;
  float alpha;
  const float* bias;
  const float* bias2;
  int relu;
  const float* mask;
  int ldmask;
  int accumulate;
  float* ws;      // split-K workspace [splits][M][N] (raw partial sums)
  int kchunk;     // K range per z-slice
  int round_out;  // store the nearest TF32 value (the output is an operand of a tensor-core product)
};

__device__ __forceinline__ float gemm_epilogue(const GemmArgs& g, int m, int n, float acc) {
  float v = acc * g.alpha;
  if (g.bias) v += g.bias[n];
  if (g.bias2) v += g.bias2[n];
  if (g.relu) v = fmaxf(v, 0.f);
  if (g.mask) v = (g.mask[(size_t)m * g.ldmask + n] > 0.f) ? v : 0.f;
  if (g.accumulate) v += g.C[(size_t)m * g.ldc + n];
  if (g.round_out) v = rna_tf32(v);
  return v;
}

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) k_sgemm(const __grid_constant__ GemmArgs g) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.kchunk;
  const int kend = min(g.K, kbeg + g.kchunk);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / NT; ++i) {
      int e = tid + i * NT;
      int kk, mm;
      if (!g.transA) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < g.M && gk < kend)
        v = g.transA ? g.A[(size_t)gk * g.lda + gm] : g.A[(size_t)gm * g.lda + gk];
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / NT; ++i) {
      int e = tid + i * NT;
      int kk, nn;
      if (g.transB) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < g.N && gk < kend)
        v = g.transB ? g.B[(size_t)gn * g.ldb + gk] : g.B[(size_t)gk * g.ldb + gn];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      if (gridDim.z > 1)
        g.ws[((size_t)blockIdx.z * g.M + m) * g.N + n] = acc[i][j];
      else
        g.C[(size_t)m * g.ldc + n] = gemm_epilogue(g, m, n, acc[i][j]);
    }
  }
}

// Split-K reduction for SMALL outputs with MANY splits (conv weight gradients: 8k..37k outputs, up to 148
// splits): a CTA of 32 x 8 threads folds 32 outputs, the 8 row-lanes striding over the splits (fixed order),
// then a fixed 8-way tree.
__global__ void k_splitk_reduce_small(const __grid_constant__ GemmArgs g, int splits) {
  __shared__ float s[8][33];
  size_t total = (size_t)g.M * g.N;
  size_t idx = (size_t)blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (idx < total)
    for (int sp = threadIdx.y; sp < splits; sp += 8) acc += g.ws[(size_t)sp * total + idx];
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && idx < total) {
    float t = ((s[0][threadIdx.x] + s[1][threadIdx.x]) + (s[2][threadIdx.x] + s[3][threadIdx.x])) +
              ((s[4][threadIdx.x] + s[5][threadIdx.x]) + (s[6][threadIdx.x] + s[7][threadIdx.x]));
    int m = (int)(idx / g.N), n = (int)(idx - (size_t)m * g.N);
    g.C[(size_t)m * g.ldc + n] = gemm_epilogue(g, m, n, t);
  }
}

// ... and for LARGE outputs with few splits (hidden-layer weight gradient: 512k outputs, 4 splits): a thread
// owns four consecutive outputs and folds the splits in index order with 16-byte loads (the CTA-per-32-outputs
// form above is 16k tiny CTAs here: 44 us next to the BPTT kernel for 10 MB of traffic; this one 5.5 us).
__global__ void __launch_bounds__(256) k_splitk_reduce(const __grid_constant__ GemmArgs g, int splits) {
  const size_t total = (size_t)g.M * g.N;
  if ((g.N & 3) == 0) {
    const size_t total4 = total >> 2;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (size_t)gridDim.x * blockDim.x) {
      const float4* p = reinterpret_cast<const float4*>(g.ws) + q;
      float4 acc = p[0];
      for (int sp = 1; sp < splits; ++sp) {
        const float4 v = p[(size_t)sp * total4];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      const size_t idx = q << 2;
      const int m = (int)(idx / g.N), n = (int)(idx - (size_t)m * g.N);
      float* c = g.C + (size_t)m * g.ldc + n;
      c[0] = gemm_epilogue(g, m, n, acc.x);
      c[1] = gemm_epilogue(g, m, n + 1, acc.y);
      c[2] = gemm_epilogue(g, m, n + 2, acc.z);
      c[3] = gemm_epilogue(g, m, n + 3, acc.w);
    }
    return;
  }
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    float acc = g.ws[idx];
    for (int sp = 1; sp < splits; ++sp) acc += g.ws[(size_t)sp * total + idx];
    const int m = (int)(idx / g.N), n = (int)(idx - (size_t)m * g.N);
    g.C[(size_t)m * g.ldc + n] = gemm_epilogue(g, m, n, acc);
  }
}

// ------------------------------------------------------------------------ frame layout
// uint8 NCHW frames (as acted / stored in the replay buffer) -> fp32 NHWC scaled by 1/255
// (rltime/models/torch/modules/cnn.py:44-45).  One thread per pixel: the C channel planes are read
// coalesced along W, the pixel's channels are written as one run (16 bytes at C = 4).
__global__ void k_u8_nchw_to_f32_nhwc(const uint8_t* __restrict__ x, float* __restrict__ xf, size_t pixels,
                                      int C, int HW, float scale, int rn) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (size_t)gridDim.x * blockDim.x) {
    size_t img = p / HW, hw = p - img * HW;
    const uint8_t* src = x + img * (size_t)C * HW + hw;
    float* dst = xf + p * C;
    if (C == 4) {
      float4 v = make_float4(__fmul_rn((float)src[0], scale), __fmul_rn((float)src[HW], scale),
                             __fmul_rn((float)src[2 * (size_t)HW], scale), __fmul_rn((float)src[3 * (size_t)HW], scale));
      if (rn) v = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
      *reinterpret_cast<float4*>(dst) = v;
    } else {
      for (int c = 0; c < C; ++c) {
        float v = __fmul_rn((float)src[(size_t)c * HW], scale);
        dst[c] = rn ? rna_tf32(v) : v;
      }
    }
  }
}

// Filters [f][kh][kw][c] (internal layout) -> Wt[class (ph,pw)][c][(dh,dw,f)] with kh = S*dh + ph,
// kw = S*dw + pw: the K-major B operand of the data-gradient implicit GEMM (rttc::k_convdx_tc).
__global__ void k_conv_wT(const float* __restrict__ w, float* __restrict__ wt, int F, int KH, int S, int C) {
  const int KD = KH / S;
  const size_t total = (size_t)F * KH * KH * C;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int f = (int)(r % F); r /= F;
    const int dw = (int)(r % KD); r /= KD;
    const int dh = (int)(r % KD); r /= KD;
    const int c = (int)(r % C); r /= C;
    const int cls = (int)r;
    const int kh = S * dh + cls / S, kw = S * dw + cls % S;
    wt[idx] = w[(((size_t)f * KH + kh) * KH + kw) * C + c];
  }
}

// conv input NHWC float.  col[(m,oh,ow), (kh,kw,c)].  VEC = 4 moves float4 (C % 4 == 0).
template <int VEC>
__global__ void k_im2col_f32_nhwc(const float* __restrict__ x, float* __restrict__ col, int rows,
                                  int C, int H, int W, int KH, int S, int OH, int OW) {
  const int K = C * KH * KH;
  const int KV = K / VEC;
  size_t total = (size_t)rows * OH * OW * KV;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int k = (int)(idx % KV) * VEC;
    size_t r = idx / KV;
    int ow = (int)(r % OW);
    int oh = (int)((r / OW) % OH);
    size_t m = r / ((size_t)OW * OH);
    int c = k % C, kw = (k / C) % KH, kh = k / (C * KH);
    const float* src = x + ((m * H + (oh * S + kh)) * W + (ow * S + kw)) * C + c;
    if (VEC == 4)
      *reinterpret_cast<float4*>(col + r * K + k) = *reinterpret_cast<const float4*>(src);
    else
      col[r * K + k] = src[0];
  }
}

// Transposed convolution data-gradient, gather form (deterministic, no atomics):
// dx[m,ih,iw,c] = sum over (kh,kw) with oh*S+kh == ih, ow*S+kw == iw of dcol[(m,oh,ow),(kh,kw,c)].
__global__ void k_col2im_nhwc(const float* __restrict__ dcol, float* __restrict__ dx, int rows, int C,
                              int H, int W, int KH, int S, int OH, int OW) {
  const int K = C * KH * KH;
  const int C4 = C >> 2;   // C % 4 == 0 (checked by the launcher)
  size_t total = (size_t)rows * H * W * C4;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % C4) * 4;
    size_t r = idx / C4;
    int iw = (int)(r % W);
    int ih = (int)((r / W) % H);
    size_t m = r / ((size_t)W * H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kh = 0; kh < KH; ++kh) {
      int t = ih - kh;
      if (t < 0 || t % S) continue;
      int oh = t / S;
      if (oh >= OH) continue;
      for (int kw = 0; kw < KH; ++kw) {
        int u = iw - kw;
        if (u < 0 || u % S) continue;
        int ow = u / S;
        if (ow >= OW) continue;
        float4 v = *reinterpret_cast<const float4*>(dcol + ((m * OH + oh) * OW + ow) * K + (kh * KH + kw) * C + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    *reinterpret_cast<float4*>(dx + r * C + c) = acc;
  }
}

// scalar fallback for C % 4 != 0
__global__ void k_col2im_nhwc_s(const float* __restrict__ dcol, float* __restrict__ dx, int rows, int C,
                                int H, int W, int KH, int S, int OH, int OW) {
  const int K = C * KH * KH;
  size_t total = (size_t)rows * H * W * C;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    size_t r = idx / C;
    int iw = (int)(r % W);
    int ih = (int)((r / W) % H);
    size_t m = r / ((size_t)W * H);
    float acc = 0.f;
    for (int kh = 0; kh < KH; ++kh) {
      int t = ih - kh;
      if (t < 0 || t % S) continue;
      int oh = t / S;
      if (oh >= OH) continue;
      for (int kw = 0; kw < KH; ++kw) {
        int u = iw - kw;
        if (u < 0 || u % S) continue;
        int ow = u / S;
        if (ow >= OW) continue;
        acc += dcol[((m * OH + oh) * OW + ow) * K + (kh * KH + kw) * C + c];
      }
    }
    dx[idx] = acc;
  }
}

// y *= (ref > 0)   (ReLU derivative through a saved post-activation output)
__global__ void k_relu_bwd_inplace(float* __restrict__ dy, const float* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    if (!(y[i] > 0.f)) dy[i] = 0.f;
}

// ------------------------------------------------------------------------------- LSTM
// rltime/models/torch/modules/lstm.py:84-116 (single-sample case), time-major rows t*B+b.
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// t = 0: hprev/cprev[0] = stored state * (1 - initials[0])   (lstm.py:67-70, 95-98)
__global__ void k_lstm_init(const float* __restrict__ hx, const float* __restrict__ cx,
                            const float* __restrict__ initials, float* __restrict__ hprev,
                            float* __restrict__ cprev, int B, int U) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * U) return;
  int b = i / U;
  float keep = 1.f - initials[b];
  hprev[i] = hx[i] * keep;
  cprev[i] = cx[i] * keep;
}

// gates = xg[t] (x W_ih^T + b_ih + b_hh) + hg (h W_hh^T); PyTorch gate order i, f, g, o.
// Writes activated gates (for BPTT), c, h, and the masked carry-in of step t+1.
__global__ void k_lstm_cell(const float* __restrict__ xg, const float* __restrict__ hg,
                            const float* __restrict__ cprev, float* __restrict__ gates,
                            float* __restrict__ c_out, float* __restrict__ h_out,
                            const float* __restrict__ next_initials, float* __restrict__ hprev_next,
                            float* __restrict__ cprev_next, int B, int U) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * U) return;
  int b = i / U, u = i - b * U;
  size_t g0 = (size_t)b * 4 * U + u;
  float gi = sigmoidf_(xg[g0] + hg[g0]);
  float gf = sigmoidf_(xg[g0 + U] + hg[g0 + U]);
  float gg = tanhf(xg[g0 + 2 * U] + hg[g0 + 2 * U]);
  float go = sigmoidf_(xg[g0 + 3 * U] + hg[g0 + 3 * U]);
  float c = gf * cprev[i] + gi * gg;
  float h = go * tanhf(c);
  gates[g0] = gi; gates[g0 + U] = gf; gates[g0 + 2 * U] = gg; gates[g0 + 3 * U] = go;
  c_out[i] = c;
  h_out[i] = h;
  if (hprev_next) {
    float keep = 1.f - next_initials[b];
    hprev_next[i] = h * keep;
    cprev_next[i] = c * keep;
  }
}

// BPTT cell: dh = dout[t] + dh_carry; produces pre-activation gate grads and the carries.
// dh_carry = (dgates[t+1] . W_hh) arrives as `nparts` raw split-K partials (`parts`, part_stride
// floats apart; nparts = 1: the finished product) that are folded here in a fixed order, and
// passes through step t+1's episode-reset mask (`next_initials`).
__global__ void k_lstm_cell_bwd(const float* __restrict__ dout, const float* __restrict__ parts, int nparts,
                                size_t part_stride, float* __restrict__ dc_carry,
                                const float* __restrict__ gates, const float* __restrict__ c,
                                const float* __restrict__ cprev, const float* __restrict__ initials,
                                const float* __restrict__ next_initials, float* __restrict__ dgates, int B,
                                int U) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * U) return;
  int b = i / U, u = i - b * U;
  size_t g0 = (size_t)b * 4 * U + u;
  float gi = gates[g0], gf = gates[g0 + U], gg = gates[g0 + 2 * U], go = gates[g0 + 3 * U];
  float carry = 0.f;
  for (int p = 0; p < nparts; ++p) carry += parts[(size_t)p * part_stride + i];
  float dh = dout[i] + (nparts ? carry * (1.f - next_initials[b]) : 0.f);
  float tc = tanhf(c[i]);
  float dc = dc_carry[i] + dh * go * (1.f - tc * tc);
  dgates[g0] = dc * gg * gi * (1.f - gi);
  dgates[g0 + U] = dc * cprev[i] * gf * (1.f - gf);
  dgates[g0 + 2 * U] = dc * gi * (1.f - gg * gg);
  dgates[g0 + 3 * U] = dh * tc * go * (1.f - go);
  // carry into step t-1 passes through this step's episode-reset mask
  dc_carry[i] = dc * gf * (1.f - initials[b]);
}

// ------------------------------------------------------------------------------- IQN
// cos(pi * i * tau), i = 1..E   (rltime/policies/torch/iqn.py:91-93)
__global__ void k_cos_features(const float* __restrict__ tau, float* __restrict__ cf, int rows, int E, int rn) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * E) return;
  int i = (int)(idx % E);
  float t = tau[idx / E];
  float v = cosf(__fmul_rn(__fmul_rn((float)(i + 1), 3.14159274101257324f), t));
  cf[idx] = rn ? rna_tf32(v) : v;
}

// xq[r,d] = x[r / Nq, d] * phi[r,d]      (iqn.py:84, 99-100); D % 4 == 0
__global__ void k_quantile_mul(const float* __restrict__ x, const float* __restrict__ phi,
                               float* __restrict__ xq, size_t rowsq, int D, int Nq, int rn) {
  const int D4 = D >> 2;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rowsq * D4) return;
  size_t r = idx / D4;
  int d4 = (int)(idx - r * D4);
  float4 a = reinterpret_cast<const float4*>(x + (r / Nq) * D)[d4];
  float4 p = reinterpret_cast<const float4*>(phi + r * D)[d4];
  float4 v = make_float4(a.x * p.x, a.y * p.y, a.z * p.z, a.w * p.w);
  if (rn) v = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
  reinterpret_cast<float4*>(xq + r * D)[d4] = v;
}

// dphi_pre[r,d] = dxq[r,d] * x[m,d] * (phi > 0);  dx[m,d] = sum_q dxq[m*Nq+q, d] * phi[m*Nq+q, d]
__global__ void k_quantile_mul_bwd(const float* __restrict__ dxq, const float* __restrict__ x,
                                   const float* __restrict__ phi, float* __restrict__ dphi,
                                   float* __restrict__ dx, int M, int D, int Nq) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * D) return;
  int m = (int)(idx / D), d = (int)(idx - (size_t)m * D);
  float xv = x[idx], acc = 0.f;
  for (int q = 0; q < Nq; ++q) {
    size_t r = ((size_t)m * Nq + q) * D + d;
    float p = phi[r], g = dxq[r];
    acc += g * p;
    dphi[r] = (p > 0.f) ? g * xv : 0.f;
  }
  dx[idx] = acc;
}

// qmean[r,a] = mean_q q[r,q,a]   (IQNPolicy._actor_predict_postprocess, policies/torch/iqn.py:124-131)
__global__ void k_quantile_mean(const float* __restrict__ q, float* __restrict__ out, int rows, int Nq, int A) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * A) return;
  int r = i / A, a = i - r * A;
  float s = 0.f;
  for (int k = 0; k < Nq; ++k) s += q[((size_t)r * Nq + k) * A + a];
  out[i] = s / (float)Nq;
}

// ------------------------------------------------------------------------------ target
// IQN._get_bootstrap_target_value (rltime/training/torch/iqn.py:37-52) +
// TorchTrainer.calc_target_values / _vf_unscale / _vf_scale (torch_trainer.py:46-78,144-147).
// One warp per row m: the Nq x A selection values are staged through shared memory with
// coalesced loads, lane a averages action a over the quantiles in index order, lane q then forms
// target quantile q.  Block = 128 threads (4 rows); dynamic smem = 4 * Nq * A floats.
__global__ void __launch_bounds__(128) k_iqn_target(const float* __restrict__ tq, const float* __restrict__ sq,
                             const double* __restrict__ returns, const double* __restrict__ masks,
                             const long long* __restrict__ nsteps, float* __restrict__ targets, int M,
                             int Nq, int A, float gamma, double vf_eps) {
  extern __shared__ float s_sel[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + warp;
  if (m >= M) return;
  float* ss = s_sel + (size_t)warp * Nq * A;
  const float* src = sq + (size_t)m * Nq * A;
  for (int i = lane; i < Nq * A; i += 32) ss[i] = src[i];
  __syncwarp();
  int best = 0;
  float best_v = -INFINITY;
  for (int a0 = 0; a0 < A; a0 += 32) {
    const int a = a0 + lane;
    float sv = -INFINITY;
    if (a < A) {
      float t = 0.f;
      for (int q = 0; q < Nq; ++q) t += ss[q * A + a];
      sv = t / (float)Nq;
    }
    // first maximum in action order (strict >), as a sequential scan would pick it
    for (int l = 0; l < 32 && a0 + l < A; ++l) {
      float v = __shfl_sync(0xffffffffu, sv, l);
      if (v > best_v) { best_v = v; best = a0 + l; }
    }
  }
  const float ret = (float)returns[m], mask = (float)masks[m];
  const float disc = powf(gamma, (float)nsteps[m]);
  for (int q = lane; q < Nq; q += 32) {
    float boot = tq[((size_t)m * Nq + q) * A + best];
    if (vf_eps > 0.0) {
      double sx = (double)boot, a = fabs(sx), e = vf_eps;
      double x = a / e - ((1.0 / (2.0 * (e * e))) * sqrt(4.0 * e * a + (2.0 * e + 1.0) * (2.0 * e + 1.0))) +
                 (2.0 * e + 1.0) / (2.0 * (e * e));
      x *= (sx > 0.0) - (sx < 0.0);
      boot = (float)x;
    }
    float y = ret + disc * boot * mask;
    if (vf_eps > 0.0) {
      float sg = (float)((y > 0.f) - (y < 0.f));
      y = sg * (sqrtf(fabsf(y) + 1.f) - 1.f) + (float)vf_eps * y;
    }
    targets[(size_t)m * Nq + q] = y;
  }
}

// -------------------------------------------------------------------------------- loss
// IQN._compute_grads (rltime/training/torch/iqn.py:72-125) forward AND backward, one CTA per
// row m.  theta_j = q[m,j,a_m]; delta_ij = y_i - theta_j; rho_ij = |tau_j - 1[delta<0]|
// row_loss = (1/Nq') sum_ij rho*huber/kappa ; report = mean_ij |delta|
// dtheta_j = -(w_m * scale) * (1/Nq') sum_i rho * huber'(delta) / kappa
__global__ void k_iqn_loss(const float* __restrict__ q, const float* __restrict__ targets,
                           const float* __restrict__ tau, const long long* __restrict__ actions,
                           const double* __restrict__ weights, float* __restrict__ dtheta,
                           float* __restrict__ row_loss, float* __restrict__ report, int Nq, int A,
                           float kappa, float grad_scale) {
  extern __shared__ float sm[];
  float* s_y = sm;            // Nq
  float* s_th = sm + Nq;      // Nq
  float* s_tau = sm + 2 * Nq; // Nq
  float* s_red = sm + 3 * Nq; // 2 * blockDim
  int m = blockIdx.x;
  int act = (int)actions[m];
  for (int j = threadIdx.x; j < Nq; j += blockDim.x) {
    s_y[j] = targets[(size_t)m * Nq + j];
    s_th[j] = q[((size_t)m * Nq + j) * A + act];
    s_tau[j] = tau[(size_t)m * Nq + j];
  }
  __syncthreads();
  float w = weights ? (float)weights[m] : 1.f;
  float lsum = 0.f, asum = 0.f;
  // thread j owns online quantile j (column), loops over target samples i
  for (int j = threadIdx.x; j < Nq; j += blockDim.x) {
    float th = s_th[j], tj = s_tau[j], g = 0.f;
    for (int i = 0; i < Nq; ++i) {
      float d = s_y[i] - th;
      float a = fabsf(d);
      float hub = (a <= kappa) ? 0.5f * d * d : kappa * (a - 0.5f * kappa);
      float dh = (a <= kappa) ? d : (d > 0.f ? kappa : -kappa);
      float rho = fabsf(tj - (d < 0.f ? 1.f : 0.f));
      lsum += rho * hub / kappa;
      asum += a;
      g += rho * dh / kappa;
    }
    dtheta[(size_t)m * Nq + j] = -(w * grad_scale) * g / (float)Nq;
  }
  s_red[threadIdx.x] = lsum;
  s_red[blockDim.x + threadIdx.x] = asum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, a = 0.f;
    for (int t = 0; t < blockDim.x; ++t) { l += s_red[t]; a += s_red[blockDim.x + t]; }
    row_loss[m] = (l / (float)Nq) * w;            // weighted per-row loss (dqn.py:83-95)
    report[m] = a / (float)(Nq * Nq);             // iqn.py:112
  }
}

// stats[0] = aggregate(row_loss), stats[1] = mean(report) = td_mean (iqn.py:127-129)
// DQN._compute_grads (rltime/training/torch/dqn.py:126-160): td = Q(s,a) - y, huber / mse
// (dqn.py:96-110), importance weights; reports the SIGNED td error (dqn.py:73-81 hands it to the
// history buffer, which takes abs) and the chosen q-value (logged as "qvalue").
__global__ void k_dqn_loss(const float* __restrict__ q, const float* __restrict__ targets,
                           const long long* __restrict__ actions, const double* __restrict__ weights,
                           float* __restrict__ dtheta, float* __restrict__ row_loss,
                           float* __restrict__ report, float* __restrict__ row_q, int M, int A, float kappa,
                           int mse, float grad_scale) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float qa = q[(size_t)m * A + (int)actions[m]];
  float td = qa - targets[m];
  float w = weights ? (float)weights[m] : 1.f;
  float a = fabsf(td), l, dl;
  if (mse) {
    l = td * td;
    dl = 2.f * td;
  } else {
    l = (a <= kappa) ? 0.5f * td * td : kappa * (a - 0.5f * kappa);
    dl = (a <= kappa) ? td : (td > 0.f ? kappa : -kappa);
  }
  row_loss[m] = l * w;
  report[m] = td;
  row_q[m] = qa;
  dtheta[m] = w * grad_scale * dl;
}

// stats[0] = aggregated loss (sum x scale: the mean / sum / per-timestep combinations of
// dqn.py:116-124 are all one scale factor), stats[1] = mean of `aux` (td_mean or qvalue)
__global__ void k_loss_stats(const float* __restrict__ row_loss, const float* __restrict__ aux,
                             float* __restrict__ stats, int M, float scale) {
  __shared__ double s[2][256];
  double l = 0.0, r = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) { l += row_loss[i]; r += aux[i]; }
  s[0][threadIdx.x] = l; s[1][threadIdx.x] = r;
  __syncthreads();
  if (threadIdx.x == 0) {
    double L = 0, R = 0;
    for (int t = 0; t < blockDim.x; ++t) { L += s[0][t]; R += s[1][t]; }
    stats[0] = (float)(L * scale);
    stats[1] = (float)(R / M);
  }
}

// -------------------------------------------------------------------- column reductions
// out[n] = sum_m x[m, n]; two deterministic stages.  Stage 1: a CTA of 32 x 8 threads owns 32
// columns and a slab of rows; a warp reads 128 contiguous bytes per row.
__global__ void k_colsum_partial(const float* __restrict__ x, float* __restrict__ part, size_t rows,
                                 int N, int rows_per_block) {
  __shared__ float s[8][33];
  int n = blockIdx.x * 32 + threadIdx.x;
  size_t r0 = (size_t)blockIdx.y * rows_per_block;
  size_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc = 0.f;
  if (n < N)
    for (size_t r = r0 + threadIdx.y; r < r1; r += 8) acc += x[r * N + n];
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += s[j][threadIdx.x];
    part[(size_t)blockIdx.y * N + n] = t;
  }
}
__global__ void k_colsum_final(const float* __restrict__ part, float* __restrict__ out, int parts,
                               int N, int accumulate, int pstride) {
  __shared__ float s[8][33];
  int n = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (n < N)
    for (int p = threadIdx.y; p < parts; p += 8) acc += part[(size_t)p * pstride + n];
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = ((s[0][threadIdx.x] + s[1][threadIdx.x]) + (s[2][threadIdx.x] + s[3][threadIdx.x])) +
              ((s[4][threadIdx.x] + s[5][threadIdx.x]) + (s[6][threadIdx.x] + s[7][threadIdx.x]));
    out[n] = accumulate ? out[n] + t : t;
  }
}

// ------------------------------------------------------------------------- small heads
// out layer (A actions) + dueling value layer + combine in one pass, one warp per row
// (rltime/policies/torch/dqn.py:78-112): adv = h1 Wout^T + b, v = v1 Wv^T + bv,
// q = v + adv - mean_a adv.  A <= 32.
// FI > 0: F == 128 * FI known at compile time, so the column loop unrolls fully and every data load
// of a row group is in flight before the first FMA (the generic loop waits one memory latency per
// 128 columns).
template <int MAXA, int RPW = 4, int FI = 0>   // RPW rows per warp: each weight vector read from L1 serves RPW rows
__global__ void __launch_bounds__(256) k_heads_out(const float* __restrict__ h1, const float* __restrict__ v1,
                            const float* __restrict__ Wout, const float* __restrict__ bout,
                            const float* __restrict__ Wv, const float* __restrict__ bv,
                            float* __restrict__ adv, float* __restrict__ vout, float* __restrict__ q,
                            size_t rows, int F, int A, int ldh) {
  int lane = threadIdx.x & 31;
  // grid-stride over groups of RPW rows: the launch sizes the grid to the resident warps, so the
  // last wave is as full as the others
  const size_t warps_total = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r0 = (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW; r0 < rows;
       r0 += warps_total * RPW) {
  float acc[RPW][MAXA], accv[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    accv[i] = 0.f;
#pragma unroll
    for (int a = 0; a < MAXA; ++a) acc[i][a] = 0.f;
  }
  // rows past the end re-read the last row (results discarded): keeps the loop branch-free
  const float* hr[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) hr[i] = h1 + (r0 + i < rows ? r0 + i : rows - 1) * ldh;
  const ptrdiff_t voff = v1 ? v1 - h1 : 0;
  auto fma_cols = [&](int f, const float4 (&hv)[RPW], const float4 (&vv)[RPW]) {
#pragma unroll
    for (int a = 0; a < MAXA; ++a)
      if (a < A) {
        float4 w = __ldg(reinterpret_cast<const float4*>(Wout + (size_t)a * F + f));
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          acc[i][a] = fmaf(hv[i].x, w.x, acc[i][a]); acc[i][a] = fmaf(hv[i].y, w.y, acc[i][a]);
          acc[i][a] = fmaf(hv[i].z, w.z, acc[i][a]); acc[i][a] = fmaf(hv[i].w, w.w, acc[i][a]);
        }
      }
    if (v1) {
      float4 w = __ldg(reinterpret_cast<const float4*>(Wv + f));
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        accv[i] = fmaf(vv[i].x, w.x, accv[i]); accv[i] = fmaf(vv[i].y, w.y, accv[i]);
        accv[i] = fmaf(vv[i].z, w.z, accv[i]); accv[i] = fmaf(vv[i].w, w.w, accv[i]);
      }
    }
  };
  if (FI > 0) {
    float4 hv[FI > 0 ? FI : 1][RPW], vv[FI > 0 ? FI : 1][RPW];
#pragma unroll
    for (int j = 0; j < FI; ++j)
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        hv[j][i] = __ldcs(reinterpret_cast<const float4*>(hr[i] + lane * 4 + 128 * j));
        vv[j][i] = v1 ? __ldcs(reinterpret_cast<const float4*>(hr[i] + voff + lane * 4 + 128 * j))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int j = 0; j < FI; ++j) fma_cols(lane * 4 + 128 * j, hv[j], vv[j]);
  } else {
    // F % 4 == 0: 16-byte loads, RPW independent rows of loads in flight per lane
    for (int f = lane * 4; f < F; f += 128) {
      float4 hv[RPW], vv[RPW];
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        hv[i] = __ldcs(reinterpret_cast<const float4*>(hr[i] + f));
        vv[i] = v1 ? __ldcs(reinterpret_cast<const float4*>(hr[i] + voff + f)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      fma_cols(f, hv, vv);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
#pragma unroll
      for (int a = 0; a < MAXA; ++a) acc[i][a] += __shfl_xor_sync(0xffffffffu, acc[i][a], o);
      accv[i] += __shfl_xor_sync(0xffffffffu, accv[i], o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      size_t r = r0 + i;
      if (r >= rows) break;
      float mean = 0.f;
#pragma unroll
      for (int a = 0; a < MAXA; ++a)
        if (a < A) {
          acc[i][a] += bout[a];
          mean += acc[i][a];
        }
      mean /= (float)A;
      float vv = v1 ? accv[i] + bv[0] : 0.f;
      if (v1) vout[r] = vv;
#pragma unroll
      for (int a = 0; a < MAXA; ++a)
        if (a < A) {
          adv[r * A + a] = acc[i][a];
          q[r * A + a] = v1 ? (vv + acc[i][a] - mean) : acc[i][a];
        }
    }
  }
  }
}

// Backward of the two small layers in ONE pass over the hidden activations (rows x C virtual
// columns, C = F out-layer inputs [+ F value-layer inputs]):
//   data gradients with the ReLU masks of their inputs
//     dh1[r,f] = g_r (Wout[act_r,f] - dueling * mean_a Wout[a,f]) * (h1 > 0),  dv1[r,f] = g_r Wv[f] * (v1 > 0)
//   and, per slab of rows, the partial sums the weight / bias gradients are made of
//     part[slab][a][c]  = sum_{r: act_r = a} g_r h1[r,c]   (c < F, a < A)
//     part[slab][0][c]  = sum_r g_r v1[r,c-F]              (c >= F)
//     part[slab][A][c]  = sum_r dh1|dv1[r,c]               (bias gradient of the hidden layers)
// k_heads_bwd_final folds the slabs in a fixed order.  Grid (ceil(C/128), slabs), block (32, 8);
// a thread owns 4 consecutive columns; warps stride the slab's rows.
template <int MAXA>
__global__ void __launch_bounds__(256) k_heads_bwd_fused(
    const float* __restrict__ dtheta, const long long* __restrict__ actions,
    const float* __restrict__ Wout, const float* __restrict__ Wv, const float* __restrict__ h1,
    const float* __restrict__ v1, float* __restrict__ dh1, float* __restrict__ dv1,
    float* __restrict__ part, float* __restrict__ partb, size_t rows, int F, int A, int Nq,
    int dueling, int ldh, int rows_per_block) {
  __shared__ __align__(16) float s_w[MAXA][128];
  __shared__ __align__(16) float s_red[8][128];
  const int C = F * (dueling ? 2 : 1);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 128 + tx * 4;
  const bool valid = c < C;
  const bool isv = c >= F;
  const int f = isv ? c - F : c;
  for (int i = ty * 32 + tx; i < MAXA * 128; i += 256) {
    int a = i >> 7, col = blockIdx.x * 128 + (i & 127);
    float val = 0.f;
    if (a < A && col < F) {
      val = Wout[(size_t)a * F + col];
      if (dueling) {
        float mean = 0.f;
        for (int b = 0; b < A; ++b) mean += Wout[(size_t)b * F + col];
        val -= mean / (float)A;
      }
    }
    s_w[a][i & 127] = val;
  }
  __syncthreads();
  float acc[MAXA][4], cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < MAXA; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
  float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid && isv) wv = *reinterpret_cast<const float4*>(Wv + f);
  const size_t r0 = (size_t)blockIdx.y * rows_per_block;
  size_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  const float* src = isv ? v1 : h1;
  float* dst = isv ? dv1 : dh1;
  // bias partials of the out / value layers (column block 0 only, one lane per warp):
  //   accb[a] = sum_{r: act = a} g_r,  accb[MAXA] = sum_r g_r
  // (g and act are warp-uniform per row: lane a keeps action a's sum, every lane the total)
  float accb = 0.f, accg = 0.f;
  const bool bias_cta = blockIdx.x == 0;
  if (valid) {
    constexpr int RIF = 4;   // rows in flight per thread: 4 independent 16-byte loads
    for (size_t rb = r0 + ty; rb < r1; rb += 8 * RIF) {
      float4 xs[RIF];
      float gs[RIF];
      int acts[RIF];
#pragma unroll
      for (int i = 0; i < RIF; ++i) {
        const size_t r = rb + (size_t)8 * i;
        const bool in = r < r1;
        const size_t rr = in ? r : rb;
        xs[i] = __ldcs(reinterpret_cast<const float4*>(src + rr * ldh + f));
        gs[i] = in ? dtheta[rr] : 0.f;
        acts[i] = (int)actions[rr / Nq];
      }
#pragma unroll
      for (int i = 0; i < RIF; ++i) {
        const size_t r = rb + (size_t)8 * i;
        if (r >= r1) break;
        const float g = gs[i];
        const int act = acts[i];
        const float4 x = xs[i];
        float4 d;
        if (!isv) {
          const float4 w = *reinterpret_cast<const float4*>(&s_w[act][tx * 4]);
          d.x = x.x > 0.f ? g * w.x : 0.f; d.y = x.y > 0.f ? g * w.y : 0.f;
          d.z = x.z > 0.f ? g * w.z : 0.f; d.w = x.w > 0.f ? g * w.w : 0.f;
          const float gx = g * x.x, gy = g * x.y, gz = g * x.z, gw = g * x.w;
#pragma unroll
          for (int a = 0; a < MAXA; ++a) {
            const bool hit = a == act;
            acc[a][0] += hit ? gx : 0.f; acc[a][1] += hit ? gy : 0.f;
            acc[a][2] += hit ? gz : 0.f; acc[a][3] += hit ? gw : 0.f;
          }
        } else {
          d.x = x.x > 0.f ? g * wv.x : 0.f; d.y = x.y > 0.f ? g * wv.y : 0.f;
          d.z = x.z > 0.f ? g * wv.z : 0.f; d.w = x.w > 0.f ? g * wv.w : 0.f;
          acc[0][0] = fmaf(g, x.x, acc[0][0]); acc[0][1] = fmaf(g, x.y, acc[0][1]);
          acc[0][2] = fmaf(g, x.z, acc[0][2]); acc[0][3] = fmaf(g, x.w, acc[0][3]);
        }
        __stcs(reinterpret_cast<float4*>(dst + r * ldh + f), d);
        cs[0] += d.x; cs[1] += d.y; cs[2] += d.z; cs[3] += d.w;
        if (bias_cta) {
          accb += (tx == act) ? g : 0.f;
          accg += g;
        }
      }
    }
  }
  // fold the 8 warps (fixed order) and emit this slab's partials
  const int NP = A + 1;
#pragma unroll
  for (int o = 0; o <= MAXA; ++o) {
    if (o > A) break;
    float4 v4 = (o == A) ? make_float4(cs[0], cs[1], cs[2], cs[3])
                         : make_float4(acc[o < MAXA ? o : 0][0], acc[o < MAXA ? o : 0][1],
                                       acc[o < MAXA ? o : 0][2], acc[o < MAXA ? o : 0][3]);
    __syncthreads();
    *reinterpret_cast<float4*>(&s_red[ty][tx * 4]) = v4;
    __syncthreads();
    int col = ty * 32 + tx;      // 256 threads: the first 128 fold one column each
    if (col < 128 && blockIdx.x * 128 + col < C) {
      float t = ((s_red[0][col] + s_red[1][col]) + (s_red[2][col] + s_red[3][col])) +
                ((s_red[4][col] + s_red[5][col]) + (s_red[6][col] + s_red[7][col]));
      part[((size_t)blockIdx.y * NP + o) * C + blockIdx.x * 128 + col] = t;
    }
  }
  if (blockIdx.x == 0) {
    __syncthreads();
    if (tx < MAXA) s_red[ty][tx] = accb;
    if (tx == 0) s_red[ty][MAXA] = accg;
    __syncthreads();
    const int a = ty * 32 + tx;
    if (a <= MAXA) {
      float t = ((s_red[0][a] + s_red[1][a]) + (s_red[2][a] + s_red[3][a])) +
                ((s_red[4][a] + s_red[5][a]) + (s_red[6][a] + s_red[7][a]));
      partb[(size_t)blockIdx.y * (MAXA + 1) + a] = t;
    }
  }
}

// Folds the slab partials of k_heads_bwd_fused into the gradients of the out layer (Wout, bout),
// the value layer (Wv, bv) and the hidden-layer biases.  Blocks [0, ceil(C/32)) own 32 columns each
// (threads (32, 8): ty strides the slabs); the last block reduces the two bias vectors from dtheta.
template <int MAXA>
__global__ void __launch_bounds__(256) k_heads_bwd_final(
    const float* __restrict__ part, const float* __restrict__ partb, int slabs,
    float* __restrict__ g_outw, float* __restrict__ g_outb, float* __restrict__ g_vw,
    float* __restrict__ g_vb, float* __restrict__ g_fcb, float* __restrict__ g_vhb, int F, int A,
    int dueling) {
  const int C = F * (dueling ? 2 : 1);
  const int NP = A + 1;
  const int col_blocks = (C + 31) / 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  __shared__ float s[8][MAXA + 1][33];
  if ((int)blockIdx.x < col_blocks) {
    const int c = blockIdx.x * 32 + tx;
    float acc[MAXA + 1];
#pragma unroll
    for (int o = 0; o <= MAXA; ++o) acc[o] = 0.f;
    if (c < C)
      for (int p = ty; p < slabs; p += 8) {
#pragma unroll
        for (int o = 0; o <= MAXA; ++o)
          if (o <= A) acc[o] += part[((size_t)p * NP + o) * C + c];
      }
#pragma unroll
    for (int o = 0; o <= MAXA; ++o) s[ty][o][tx] = acc[o];
    __syncthreads();
    if (ty == 0 && c < C) {
      float t[MAXA + 1];
#pragma unroll
      for (int o = 0; o <= MAXA; ++o)
        t[o] = ((s[0][o][tx] + s[1][o][tx]) + (s[2][o][tx] + s[3][o][tx])) +
               ((s[4][o][tx] + s[5][o][tx]) + (s[6][o][tx] + s[7][o][tx]));
      if (c < F) {
        float tot = 0.f;
#pragma unroll
        for (int a = 0; a < MAXA; ++a)
          if (a < A) tot += t[a];
        const float sub = dueling ? tot / (float)A : 0.f;
#pragma unroll
        for (int a = 0; a < MAXA; ++a)
          if (a < A) g_outw[(size_t)a * F + c] = t[a] - sub;
#pragma unroll
        for (int o = 0; o <= MAXA; ++o)
          if (o == A) g_fcb[c] = t[o];
      } else {
        g_vw[c - F] = t[0];
#pragma unroll
        for (int o = 0; o <= MAXA; ++o)
          if (o == A) g_vhb[c - F] = t[o];
      }
    }
    return;
  }
  // bias gradients: bout[a] = sum_{r: act = a} g_r - dueling * (sum_r g_r) / A;  bv = sum_r g_r
  // (slab partials from k_heads_bwd_fused; thread a folds column a in slab order)
  const int tid = ty * 32 + tx;
  float* sb = &s[0][0][0];
  if (tid <= MAXA) {
    float t = 0.f;
    for (int p = 0; p < slabs; ++p) t += partb[(size_t)p * (MAXA + 1) + tid];
    sb[tid] = t;
  }
  __syncthreads();
  if (tid < A) g_outb[tid] = sb[tid] - (dueling ? sb[MAXA] / (float)A : 0.f);
  if (tid == 0 && dueling) g_vb[0] = sb[MAXA];
}

// --------------------------------------------------------------------------- optimiser
// stage 1: per-block sum of squares of the flat gradient (deterministic)
__global__ void __launch_bounds__(256) k_sumsq_partial(const float* __restrict__ g, double* __restrict__ part, size_t n) {
  __shared__ double s[256];
  double acc = 0.0;
  const size_t n4 = n >> 2;     // the flat buffer is 256-byte aligned
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 v = g4[i];
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    double v = g[(n4 << 2) + threadIdx.x];
    acc += v * v;
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}
// stage 2: norm, clip coefficient (torch.nn.utils.clip_grad_norm_: coef = min(1, c/(norm+1e-6)))
// stats[2] = grad_norm, stats[3] = clip coefficient applied
__global__ void k_gradnorm_final(const double* __restrict__ part, int parts, float* __restrict__ stats,
                                 float clip, float grad_scale, float dyn_alpha) {
  // one warp: lane-strided partial sums, then a fixed butterfly
  double t = 0.0;
  for (int i = threadIdx.x; i < parts; i += 32) t += part[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float norm = (float)sqrt(t) * grad_scale;
    stats[2] = norm;
    float coef = 1.f;
    if (clip > 0.f) {
      float cv = clip;
      if (dyn_alpha >= 0.f) {
        // dynamic clipping (torch_trainer.py:153-175): clip to clip x EMA of the gradient norm;
        // stats[4] = moving average, stats[5] = initialised flag
        float ma = stats[5] > 0.5f ? stats[4] * dyn_alpha + norm * (1.f - dyn_alpha) : norm;
        stats[4] = ma;
        stats[5] = 1.f;
        cv = ma * clip;
      }
      coef = cv / (norm + 1e-6f);
      if (coef > 1.f) coef = 1.f;
    }
    stats[3] = coef;
  }
}
// torch.optim.Adam single-tensor math (torch_trainer.py:82-83,199): one pass over the flat
// parameter / gradient / moment buffers: 16 B read + 12 B written per parameter.
// The flat buffers are 256-byte aligned and padded to 64 floats, so the pass runs on float4s.
// `shadow` (optional): the copy of the parameters the tensor-core kernels read: tensors flagged in
// `wflag` (one byte per 64-float block: 1 = operand of a GEMM-shaped kernel) rounded to the nearest
// TF32 value, everything else (biases, the SIMT head layers) verbatim.
__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, float coef, float step_size,
                                          float b1, float b2, float eps, float bc2_sqrt) {
  float gi = g * coef;
  float mi = m + (gi - m) * (1.f - b1);       // exp_avg.lerp_(grad, 1 - beta1)
  float vi = v * b2 + (1.f - b2) * gi * gi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p = p - step_size * (mi / denom);
  m = mi;
  v = vi;
  return p;
}
__global__ void __launch_bounds__(256)
k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
       float* __restrict__ v, size_t n, const float* __restrict__ stats, float lr,
       float b1, float b2, float eps, float bc1, float bc2_sqrt, float grad_scale,
       float* __restrict__ shadow, const uint8_t* __restrict__ wflag) {
  const float coef = stats[3] * grad_scale;
  const float step_size = lr / bc1;
  const size_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, coef, step_size, b1, b2, eps, bc2_sqrt);
    adam_one(pp.y, gg.y, mm.y, vv.y, coef, step_size, b1, b2, eps, bc2_sqrt);
    adam_one(pp.z, gg.z, mm.z, vv.z, coef, step_size, b1, b2, eps, bc2_sqrt);
    adam_one(pp.w, gg.w, mm.w, vv.w, coef, step_size, b1, b2, eps, bc2_sqrt);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
    if (shadow) {
      if (wflag[i >> 4]) pp = make_float4(rna_tf32(pp.x), rna_tf32(pp.y), rna_tf32(pp.z), rna_tf32(pp.w));
      reinterpret_cast<float4*>(shadow)[i] = pp;
    }
  }
}
// shadow = wflag ? rna_tf32(p) : p over the whole flat buffer (parameter load / target sync)
__global__ void k_shadow_params(const float* __restrict__ p, float* __restrict__ shadow,
                                const uint8_t* __restrict__ wflag, size_t n) {
  const size_t n4 = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<const float4*>(p)[i];
    if (wflag[i >> 4]) pp = make_float4(rna_tf32(pp.x), rna_tf32(pp.y), rna_tf32(pp.z), rna_tf32(pp.w));
    reinterpret_cast<float4*>(shadow)[i] = pp;
  }
}


// --------------------------------------------------------------------- persistent LSTM
// Whole recurrence of lstm.py:84-116 in ONE launch: grid = U / 4 CTAs (co-resident, one per
// SM), each CTA keeps the 16 rows of W_hh of its 4 hidden units (4 gates x 4 units x U
// floats = 32 KiB at U = 512) in shared memory for all time-steps, stages the masked h_{t-1}
// (B x U) through shared memory each step, and keeps every cell state c in the register of
// the thread that owns (batch b = lane (+32 j), unit = warp).  Steps are separated by a
// grid-wide barrier on a monotone global counter.
namespace lstm_seq {
constexpr int UPB = 4;        // hidden units per CTA == warps per CTA
constexpr int HPAD = 4;       // row padding (floats) of the staged h tile: conflict-free LDS.128

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}
}  // namespace lstm_seq

// xg: (T*B, 4U) = x W_ih^T + b_ih + b_hh;  W_hh: (4U, U);  hx, cx: (B, U) state of step 0;
// initials: (T*B).  Outputs as in the stepwise path: gates (T*B,4U) activated, c_all, h_all,
// hprev / cprev (masked carry-ins, needed by BPTT), all (T*B, U).  B <= 32, U = 32 * KPL.
//
// Warps w and w+4 own hidden unit blockIdx.x*4 + (w & 3), each one half of K.  Their slice
// of W_hh (4 gate rows x U/2) lives in REGISTERS for the whole sequence (the shared-memory
// version was bound by 5 LDS.128 per 16 FMA).  Per step a lane forms the 4 partial dot products
// of every batch row over its K slice, a 124-shuffle reduce-scatter butterfly leaves lane b with
// the gate sums of batch row b over that warp's K half, the upper half hands its four values
// over through shared memory, and the lower half finishes (activations, cell update, outputs)
// keeping c in a register.
// h_{t-1} is exchanged through `hrep`: REP copies of the (B x U) state at distinct addresses,
// double-buffered by step parity.  Every CTA writes its 4 units into all copies and reads the
// whole state from copy blockIdx.x % REP — 128 CTAs pulling the same 64 KB at the same moment
// otherwise serialise on a few L2 slices (measured: ~8 B/clk/SM).
constexpr int LSTM_REP = 1;   // replicas did not help (the 8 B/clk/SM was the LSU path, not an L2 hot spot)

template <int KPL>   // K elements per lane per half: U = 2 * 32 * KPL
__global__ void __launch_bounds__(256)
k_lstm_seq_fwd(const float* __restrict__ xg, const float* __restrict__ Whh, const float* __restrict__ hx,
               const float* __restrict__ cx, const float* __restrict__ initials,
               float* __restrict__ gates, float* __restrict__ c_all, float* __restrict__ h_all,
               float* __restrict__ hprev, float* __restrict__ cprev, float* __restrict__ hrep, int T, int B,
               int U, unsigned int* __restrict__ barrier, long long* __restrict__ dbg) {
  using namespace lstm_seq;
  extern __shared__ __align__(16) float hs[];   // [32][U + HPAD] h_{t-1} (rows >= B zero), [4][32][4] handover, mbarrier
  const int HS = U + HPAD;
  float* red = hs + 32 * HS;
  unsigned long long* tma_bar = reinterpret_cast<unsigned long long*>(red + 4 * 32 * 4);
  const unsigned bar_addr = static_cast<unsigned>(__cvta_generic_to_shared(tma_bar));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int uw = warp & (UPB - 1), khalf = warp >> 2;
  const int unit = blockIdx.x * UPB + uw;
  const int b = lane;                       // batch row this lane finishes
  const int koff = khalf * (U / 2);
  // lane l owns the 16-byte chunks {c*128 + 4l .. +3} of its K half: conflict-free LDS.128
  float w[4][KPL];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int i = 0; i < KPL; i += 4) {
      float4 t4 = *reinterpret_cast<const float4*>(Whh + ((size_t)g * U + unit) * U + koff + (i / 4) * 128 + lane * 4);
      w[g][i] = t4.x; w[g][i + 1] = t4.y; w[g][i + 2] = t4.z; w[g][i + 3] = t4.w;
    }
  for (int i = threadIdx.x; i < 32 * HS; i += blockDim.x) hs[i] = 0.f;
  float c_reg = (khalf == 0 && b < B) ? cx[(size_t)b * U + unit] : 0.f;
  const size_t rep_stride = (size_t)B * U;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    // step 0 reads the stored state; later steps read this CTA's replica of h_{t-1}
    const float* hsrc = t == 0 ? hx : hrep + ((size_t)((t - 1) & 1) * LSTM_REP + (blockIdx.x % LSTM_REP)) * rep_stride;
    const float* ini = initials + (size_t)t * B;
    float xin[4] = {0.f, 0.f, 0.f, 0.f};
    const float keep_b = b < B ? 1.f - ini[b] : 0.f;
    if (khalf == 0 && b < B) {
      const float* xr = xg + ((size_t)t * B + b) * 4 * U + unit;
      xin[0] = __ldg(xr); xin[1] = __ldg(xr + U); xin[2] = __ldg(xr + 2 * U); xin[3] = __ldg(xr + 3 * U);
    }
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[8 * t + 0] = clock64();
    // stage h_{t-1}: one thread issues B bulk copies (one 4U-byte row each) that the TMA engine
    // streams from L2 into the padded rows; everybody waits on the mbarrier.  (LDG.128 through the
    // LSU managed ~8 B/clk/SM here; the copy is unmasked: (h * keep) . w == keep * (h . w).)
    if (warp == 0) {
      // lane 0 posts the expected byte count, then every lane issues the copy of one batch row
      // (one thread issuing all 32 took ~2.7k cycles)
      // the rows were written by other SMs through the generic proxy and published by the grid
      // barrier; order them before this SM's async-proxy (TMA) reads of global memory
      asm volatile("fence.proxy.async.global;" ::: "memory");
      if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
                     "r"((unsigned)(B * U * 4))
                     : "memory");
      __syncwarp();
      if (lane < B) {
        unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(hs + (size_t)lane * HS));
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
            "l"(hsrc + (size_t)lane * U), "r"((unsigned)(U * 4)), "r"(bar_addr)
            : "memory");
      }
    }
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[8 * t + 1] = clock64();
    {
      unsigned ok = 0;
      const unsigned parity = (unsigned)(t & 1);
      while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity)
            : "memory");
      }
    }
    __syncthreads();
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[8 * t + 2] = clock64();
    // partial dot products: part[bb*4 + g] = sum_{k in this lane's slice} h[bb][k] * W[g][k]
    float part[128];
#pragma unroll
    for (int bb = 0; bb < 32; ++bb) {
      float hv[KPL];
#pragma unroll
      for (int i = 0; i < KPL; i += 4) {
        float4 t4 = *reinterpret_cast<const float4*>(hs + (size_t)bb * HS + koff + (i / 4) * 128 + lane * 4);
        hv[i] = t4.x; hv[i + 1] = t4.y; hv[i + 2] = t4.z; hv[i + 3] = t4.w;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < KPL; i += 2) {
          a0 = fmaf(hv[i], w[g][i], a0);
          a1 = fmaf(hv[i + 1], w[g][i + 1], a1);
        }
        part[bb * 4 + g] = a0 + a1;
      }
    }
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[8 * t + 3] = clock64();
    // reduce-scatter butterfly over the 32 lanes: lane l ends with indices 4l .. 4l+3
#pragma unroll
    for (int o = 16, n = 128; o >= 1; o >>= 1, n >>= 1) {
      const int half = n >> 1;
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        float send = up ? part[i] : part[i + half];
        float keepv = up ? part[i + half] : part[i];
        part[i] = keepv + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    if (khalf == 1) *reinterpret_cast<float4*>(red + (uw * 32 + lane) * 4) = make_float4(part[0], part[1], part[2], part[3]);
    __syncthreads();
    if (khalf == 0 && b < B) {
      float4 hi = *reinterpret_cast<const float4*>(red + (uw * 32 + lane) * 4);
      size_t row = (size_t)t * B + b;
      float gi = sigmoidf_(xin[0] + keep_b * (part[0] + hi.x));
      float gf = sigmoidf_(xin[1] + keep_b * (part[1] + hi.y));
      float gg = tanhf(xin[2] + keep_b * (part[2] + hi.z));
      float go = sigmoidf_(xin[3] + keep_b * (part[3] + hi.w));
      float cp = c_reg * keep_b;
      float c = gf * cp + gi * gg;
      float h = go * tanhf(c);
      c_reg = c;
      float* gr = gates + row * 4 * U + unit;
      gr[0] = gi; gr[U] = gf; gr[2 * U] = gg; gr[3 * U] = go;
      c_all[row * U + unit] = c;
      h_all[row * U + unit] = h;
      cprev[row * U + unit] = cp;
      hprev[row * U + unit] = hs[(size_t)b * HS + unit] * keep_b;
      if (t + 1 < T) {
        float* hr = hrep + (size_t)(t & 1) * LSTM_REP * rep_stride + (size_t)b * U + unit;
#pragma unroll
        for (int r = 0; r < LSTM_REP; ++r) hr[r * rep_stride] = h;
      }
    }
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[8 * t + 4] = clock64();
    if (t + 1 < T) lstm_seq::grid_barrier(barrier, (unsigned int)(t + 1) * gridDim.x);
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[8 * t + 5] = clock64();
  }
}

}  // namespace rtk
