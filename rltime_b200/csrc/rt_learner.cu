// IQN / R2D2-style learner for B200 (sm_100a): nature-CNN -> [LSTM] -> FC -> IQN quantile +
// dueling heads, double-Q n-step targets with value rescaling, quantile-Huber loss, global
// grad-norm clip and Adam — every op a hand-written kernel (rt_kernels.cuh), no cuDNN/cuBLAS.
//
// Replaces, behind the C ABI of include/rltime_b200.h:
//   rltime/models/torch/{sequential,modules/cnn,modules/lstm,modules/fc}.py   (forward)
//   rltime/policies/torch/{dqn,iqn}.py                                        (heads)
//   rltime/training/torch/{torch_trainer,dqn,iqn}.py                          (target, loss, step)
//   rltime/training/multi_step_trainer.py:90-131                              (burn-in)
//
// Parameters live in one flat fp32 buffer per network (online / target) with matching flat
// gradient and Adam-moment buffers, so clip + Adam + target sync are single passes.  Internal
// weight layouts are permuted once at load time (conv filters to (f,kh,kw,c), CNN-feature
// columns to (h,w,c)) so that activations can stay NHWC and every layer is a K-major GEMM.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "rt_common.h"
#include "rt_kernels.cuh"
#include "rt_gemm_tc.cuh"
#include "rt_lstm_tc.cuh"
#include "rt_bptt.cuh"

namespace {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmapKey {
  const void* ptr;
  uint64_t d0, d1, pitch;
  uint32_t b0, b1;
  int mn;
  bool operator<(const TmapKey& o) const {
    return std::tie(ptr, d0, d1, pitch, b0, b1, mn) <
           std::tie(o.ptr, o.d0, o.d1, o.pitch, o.b0, o.b1, o.mn);
  }
};

// Everything a GEMM launch needs besides its operands.
struct GemmCtx {
  float* ws = nullptr;        // split-K workspace
  size_t ws_floats = 0;
  int mode = 0;               // 0 = fp32 SIMT, 1 = TF32 tcgen05 where eligible
  int round_tf32 = 0;
  PFN_tmapEncodeTiled encode = nullptr;
  std::map<TmapKey, CUtensorMap> tmaps;
  long long tc_launches = 0, simt_launches = 0;
  int force_bn = 0, force_stages = 0;   // tuning overrides (RT_TC_BN / RT_TC_STAGES, rt_gemm_bench)
  long long* dbg = nullptr;             // clock64 stamps of CTA 0 (rt_gemm_bench)
  int persistent = 1;                   // overlap epilogue with the next tile (k_gemm_tc_p)
  int pair = 1;                         // CTA-pair (cta_group::2) kernel for the wide K-major products (RT_TC_PAIR=0: off)
  // split-K partials left in `ws` for a fused consumer (BPTT cell kernel) instead of k_splitk_reduce
  int defer_reduce = 0;
  int last_splits = 1;                  // splits of the last GEMM (1: the result is in C)
  int split_min_kb = 16;                // fewest k-blocks a tcgen05 split may get
  int num_sms = 148;
  // live timing of every GEMM-shaped launch (bench.py roofline): CUDA events on the launch stream
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_used = 0;
  double prof_flops = 0;
  std::vector<double> prof_fl;          // flops of each timed launch
  std::vector<int> prof_shape;          // 6 ints per timed launch: kind (0 gemm, 1 conv fwd, 2 conv dW, 3 conv dX), M, N, K, transA, transB
};

struct ConvL {
  int cin, hin, win, f, k, s, hout, wout, K;
  size_t w, b;  // offsets in the flat parameter buffer
};

enum PermKind { PERM_NONE = 0, PERM_CONV = 1, PERM_FEAT_COLS = 2, PERM_FEAT_ROWS = 3, PERM_WIH_EXTRA = 4 };

struct PreL {   // linear + ReLU layer in front of the LSTM / the last FC module
  int in, out;
  size_t w, b;
};

struct PInfo {
  std::string name;
  std::vector<int> shape;
  size_t off, count;
  int perm;
  int conv_c = 0, conv_k = 0;  // PERM_CONV geometry
  bool gemm_w = false;         // weight operand of a GEMM-shaped (tensor-core) kernel: TF32-rounded in the shadow copy
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
int grid1d(size_t n, int threads = 256);
// fixed-order fold of split-K partials: vector form for large outputs, CTA-per-32-outputs form for small ones
inline void launch_splitk_reduce(cudaStream_t st, const rtk::GemmArgs& g, int splits) {
  const size_t total = (size_t)g.M * g.N;
  if (total >= ((size_t)1 << 18) && (g.N & 3) == 0)
    rtk::k_splitk_reduce<<<grid1d(total / 4 + 1), 256, 0, st>>>(g, splits);
  else
    rtk::k_splitk_reduce_small<<<(unsigned)((total + 31) / 32), dim3(32, 8), 0, st>>>(g, splits);
}

#define RT_TRY(x)                   \
  do {                              \
    int rc__ = (x);                 \
    if (rc__ != RT_OK) return rc__; \
  } while (0)

}  // namespace

struct rt_learner {
  rt_model_desc md;
  rt_train_desc td;
  int device = 0;
  std::vector<ConvL> conv;
  int featC = 0, featHW = 0, cfeat = 0;  // last conv output (no CNN: the raw float observation, featHW = 1)
  int feat = 0;                          // trunk features that reach the LSTM / the last FC module
  std::vector<PreL> pre;                 // FC layers between the two
  std::vector<float*> pre_out, pre_out2, d_pre;   // activations (primary / second set), gradients
  float* d_cfeat = nullptr;              // gradient w.r.t. the conv output when FC layers follow it
  bool wih_perm = false;                 // W_ih columns follow the conv feature permutation
  int X = 0;                             // extra 1-D features of a tuple observation, fed to the LSTM
  size_t o_wihx = 0;                     // [4U][X] block of W_ih (stored behind the [4U][feat] block)
  int U = 0, D = 0, F = 0, A = 0, Nq = 0, E = 0;
  bool dueling = false;
  bool dqn = false;            // plain DQNPolicy: no quantile layer, one output per action
  float loss_scale = 1.f;      // aggregation of the per-row losses (dqn.py:116-124) as one factor
  float* row_q = nullptr;      // chosen q-values of the last step (DQN "qvalue" log)
  bool fused_hidden = false;   // FC + value-hidden layers stored / executed as one [2F x D] layer
  int ldh = 0;                 // row pitch of h1 / v1 / dh1 / dv1 (F, or 2F when fused)
  std::vector<PInfo> pinfo;
  size_t nparams = 0;
  // parameter offsets
  size_t o_wih = 0, o_whh = 0, o_bih = 0, o_bhh = 0, o_fcw = 0, o_fcb = 0, o_outw = 0, o_outb = 0,
         o_vhw = 0, o_vhb = 0, o_vw = 0, o_vb = 0, o_qw = 0, o_qb = 0;
  float* p[2] = {nullptr, nullptr};  // 0 online, 1 target
  // RT_GEMM_TF32_RN: what the kernels read.  pr[i] == p[i] unless rn, in which case it is a shadow of
  // p[i] with every GEMM weight rounded to the nearest TF32 value (biases / SIMT head layers verbatim),
  // refreshed by the Adam pass, parameter loads and target syncs.
  float* pr[2] = {nullptr, nullptr};
  int rn = 0;
  uint8_t* wflag = nullptr;          // one byte per 64-float block of the flat buffer: 1 = GEMM weight
  float* grad = nullptr;
  float* adam_m = nullptr;
  float* adam_v = nullptr;
  long long adam_t = 0;
  float lr = 1e-3f;

  // geometry
  int B = 0, T = 0, P = 0, n = 0, S = 0;
  int R = 0;          // rnn_steps_train: LSTM sequence length of the target / training passes (divides T)
  int max_rows = 0;   // trunk rows per pass
  int M = 0, MQ = 0;  // head rows (T*B) and quantile-expanded rows
  int chunk_rows = 128;

  // activations (one set; passes are sequential)
  std::vector<float*> c_out;   // conv outputs
  std::vector<float*> d_c;     // conv output grads
  float *col = nullptr, *dcol = nullptr;
  float* xf = nullptr;         // frames as fp32 NHWC * (1/255): every conv layer reads NHWC runs
  // one fp32-frame buffer per replay batch slot (keyed by the batch's frame pointer), so that the uint8 ->
  // fp32 NHWC conversion of draw k+1 can run on the replay stream while update k still reads its own
  // (rt_learner_prefetch); xf0 = the default buffer (hand-built batches, burn-in, acting)
  float* xf0 = nullptr;
  size_t xf_floats = 0;
  std::map<const void*, float*> xf_slots;
  const void* prefetched_x = nullptr;
  cudaEvent_t ev_prefetch = nullptr;
  float *xg = nullptr, *hg = nullptr, *hprev = nullptr, *cprev = nullptr, *gates = nullptr,
        *c_all = nullptr, *h_all = nullptr;
  float *tau = nullptr, *cf = nullptr, *phi = nullptr, *xq = nullptr, *h1 = nullptr, *v1 = nullptr,
        *adv = nullptr, *v = nullptr, *q = nullptr;
  float *tq = nullptr, *sq = nullptr, *targets = nullptr;
  float *dtheta = nullptr, *row_loss = nullptr, *report = nullptr, *stats = nullptr;
  float *dh1 = nullptr, *dv1 = nullptr, *dxq = nullptr, *dphi = nullptr,
        *dfeatq = nullptr, *dgates = nullptr, *dh_carry = nullptr, *dc_carry = nullptr, *dfeat = nullptr;
  GemmCtx gx;
  float* colsum_part = nullptr;
  float* hb_part = nullptr;   // slab partials of k_heads_bwd_fused: [hb_slabs][A + 1][F or 2F]
  float* hb_partb = nullptr;  // ... and of the out / value bias gradients: [hb_slabs][33]
  int hb_slabs = 256;
  unsigned int* grid_barrier = nullptr;
  unsigned int* bptt_counters = nullptr;   // producer-group counters of the one-launch BPTT kernel
  long long* lstm_dbg = nullptr;
  float* lstm_hrep = nullptr;   // replicated h exchange buffer of the persistent LSTM kernel
  int lstm_persistent = 1;
  int lstm_tc = 1;              // tensor-core multi-sequence recurrence (TF32 mode)
  int lstm_mma = 1;             // ... its mma.sync variant with register-resident W_hh (RT_LSTM_MMA=0: tcgen05 kernel)
  float* lstm_xchg = nullptr;   // swizzled h exchange blocks of that kernel
  int lstm_tcap = 0;            // time-steps the exchange buffer holds
  int lstm_upc = 8;             // preferred hidden units per CTA of that kernel (8 or 16)
  float *xg2 = nullptr;         // input gates of the target network's pass
  float *h_all2 = nullptr, *h_all3 = nullptr;   // LSTM outputs of the target / selection passes
  int conv_implicit = 1;
  int conv_implicit_bwd = 1;
  int conv_persistent = 1;
  int bptt_persistent = 1;      // one-launch BPTT recurrence (rt_bptt.cuh); RT_BPTT_PERSISTENT=0: stepwise path
  int conv_shallow = 1;         // shallow conv rings (2 CTAs/SM) inside the multi-branch phases (RT_CONV_SHALLOW; 2: backward too)
  int conv_dx_implicit = 1;
  int conv1_pair = 1;           // conv1 of the online + target networks as one N = 64 product (RT_CONV1_PAIR=0: separate)
  float* wcat = nullptr;        // [2 x f][K] staging of the two conv1 filter banks
  std::vector<float*> conv_wt;  // re-laid filters for the data-gradient implicit GEMM (per layer, null for conv1)
  float* dcol_full = nullptr;
  int num_sms = 148;
  double* sumsq_part = nullptr;
  float* tau_stage = nullptr;  // device staging for injected taus (5 segments)
  unsigned long long rng_counter = 0;
  cudaEvent_t ev_loss = nullptr;   // recorded once the losses / |td| of a step are final (before the backward pass)
  // CUDA-graph replay of the update (the ~120 launches of a step leave ~2 us of idle GPU between
  // consecutive kernels when issued one by one).  Graphs cannot be captured on the legacy default
  // stream, so the update runs on the learner's own stream, forked from / joined to the caller's.
  struct StepGraphs {
    const void* key[12] = {};
    cudaGraphExec_t fwd = nullptr, bwd = nullptr, bwd2 = nullptr;
    long long n_fwd = 0, n_bwd = 0, n_bwd2 = 0;      // launches each replay stands for (rt_launch_count)
  };
  std::vector<StepGraphs> graphs;
  int graphs_enabled = 1;
  // acting step (rt_learner_act) replayed from a graph once the same buffers were seen three times in a row
  struct ActGraph {
    const void* key[10] = {};
    cudaGraphExec_t g = nullptr;
    long long n = 0;
    int seen = 0;
  };
  std::vector<ActGraph> act_graphs;
  int act_graphs_enabled = 1;       // RT_ACT_GRAPH=0: issue the ~20 launches of an acting step one by one
  long long steps_done = 0;
  cudaStream_t own = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_late = nullptr;    // data-parallel: gradients of every non-conv parameter are final (before the conv backward)
  size_t conv_param_end = 0;        // flat index where the non-conv parameters start
  // Two-branch backward pass: the data-gradient chain (heads dX -> BPTT -> conv dX) is the critical
  // path and most of it is latency-bound (20 dependent BPTT steps on a mostly idle GPU); every
  // weight / bias gradient hangs off it as a leaf, so those run on a side stream (a parallel branch
  // of the captured graph) with their own split-K workspace and column-sum scratch.
  int overlap_bwd = 1;
  int dp_split = 0;                 // data-parallel backward as two graphs with an event in between (RT_DP_SPLIT=1)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_side[8] = {};
  int ev_side_next = 0;
  GemmCtx gx2;
  float* colsum_part2 = nullptr;
  bool side_active = false;    // inside a two-branch phase
  // second activation set: lets the TARGET network's passes (CNN + input gates, heads) run on the
  // side branch of the forward graph next to the online network's
  std::vector<float*> c_out2;
  float *cf2 = nullptr, *phi2 = nullptr, *xq2 = nullptr, *h1b = nullptr, *v1b = nullptr, *adv2 = nullptr,
        *vb2 = nullptr;
  int overlap_fwd = 1;
  // third branch (forward only): the selection heads pass, with a third activation set
  cudaStream_t side_b = nullptr;
  cudaEvent_t ev_side_b[2] = {};     // [0] fork, [1] join
  cudaEvent_t ev_phi = nullptr;      // early quantile embeddings of the training pass are done
  int phi_early = 1;                 // RT_PHI_EARLY=0: embeddings inside the heads passes, after the recurrence
  GemmCtx gx3;
  float *cf3 = nullptr, *phi3 = nullptr, *xq3 = nullptr, *h1c = nullptr, *v1c = nullptr, *adv3 = nullptr,
        *vb3 = nullptr;
  float* h_stats = nullptr;        // pinned read-back of stats[0..3]
  std::map<std::string, std::pair<void*, long long>> debug;
  std::vector<void*> allocs;
  // data parallelism inside the library (rt_comm_*): NCCL communicator + communication stream
  ncclComm_t comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  int reserve_sms = 0;               // SMs the persistent conv-backward kernels leave to NCCL (world > 1)
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_comm = nullptr;
};

namespace {

template <typename T>
int dalloc(rt_learner* h, T** p, size_t count, const char* name = nullptr) {
  RT_CUDA(cudaMalloc(reinterpret_cast<void**>(p), (count ? count : 1) * sizeof(T)));
  RT_CUDA(cudaMemset(*p, 0, (count ? count : 1) * sizeof(T)));
  h->allocs.push_back(*p);
  if (name) h->debug[name] = std::make_pair((void*)*p, (long long)count);
  return RT_OK;
}

size_t reserve_params(rt_learner* h, size_t count) {
  size_t off = (h->nparams + 63) / 64 * 64;
  h->nparams = off + count;
  return off;
}

size_t add_param(rt_learner* h, const std::string& name, std::vector<int> shape, int perm,
                 int conv_c = 0, int conv_k = 0, long long fixed_off = -1) {
  PInfo pi;
  pi.name = name;
  pi.shape = shape;
  pi.count = 1;
  for (int d : shape) pi.count *= (size_t)d;
  if (fixed_off >= 0) {   // placed inside a region reserved earlier (registration order is unchanged)
    pi.off = (size_t)fixed_off;
    pi.perm = perm;
    pi.conv_c = conv_c;
    pi.conv_k = conv_k;
    h->pinfo.push_back(pi);
    return pi.off;
  }
  pi.off = (h->nparams + 63) / 64 * 64;  // 256-byte aligned tensors
  pi.perm = perm;
  pi.conv_c = conv_c;
  pi.conv_k = conv_k;
  h->nparams = pi.off + pi.count;
  h->pinfo.push_back(pi);
  return pi.off;
}

// reference flat index -> internal flat index for one tensor
size_t perm_index(const rt_learner* h, const PInfo& pi, size_t j) {
  switch (pi.perm) {
    case PERM_CONV: {  // [f][c][kh][kw] -> [f][kh][kw][c]
      int C = pi.conv_c, K = pi.conv_k;
      size_t per_f = (size_t)C * K * K;
      size_t f = j / per_f, r = j % per_f;
      int c = (int)(r / (K * K)), kh = (int)((r / K) % K), kw = (int)(r % K);
      return f * per_f + ((size_t)kh * K + kw) * C + c;
    }
    case PERM_FEAT_COLS: {  // [rows][c*HW + hw] -> [rows][hw*C + c]
      size_t cols = (size_t)h->cfeat;
      size_t row = j / cols, col = j % cols;
      int c = (int)(col / h->featHW), hw = (int)(col % h->featHW);
      return row * cols + (size_t)hw * h->featC + c;
    }
    case PERM_WIH_EXTRA: {  // [4U][feat + X] -> [4U][feat] block (feature columns permuted) + [4U][X] block
      size_t cols = (size_t)h->feat + h->X;
      size_t row = j / cols, col = j % cols;
      if (col >= (size_t)h->feat) return (size_t)4 * h->U * h->feat + row * h->X + (col - h->feat);
      if (!h->wih_perm) return row * h->feat + col;
      int c = (int)(col / h->featHW), hw = (int)(col % h->featHW);
      return row * h->feat + (size_t)hw * h->featC + c;
    }
    case PERM_FEAT_ROWS: {  // [c*HW + hw][inner] -> [hw*C + c][inner]
      size_t inner = pi.count / (size_t)h->cfeat;
      size_t row = j / inner, in = j % inner;
      int c = (int)(row / h->featHW), hw = (int)(row % h->featHW);
      return ((size_t)hw * h->featC + c) * inner + in;
    }
    default:
      return j;
  }
}

int gemm_simt(GemmCtx& cx, cudaStream_t st, rtk::GemmArgs g) {
  constexpr int BM = 128, BN = 64, BK = 16;
  int tm = cdiv(g.M, BM), tn = cdiv(g.N, BN);
  long long tiles = (long long)tm * tn;
  int splits = 1;
  if (tiles < 148 && g.K >= 2048) {
    splits = (int)((296 + tiles - 1) / tiles);
    int maxs = g.K / 512;
    if (splits > maxs) splits = maxs;
    size_t per = (size_t)g.M * g.N;
    if ((size_t)splits * per > cx.ws_floats) splits = (int)(cx.ws_floats / per);
    if (splits < 1) splits = 1;
  }
  int kchunk = cdiv(g.K, splits);
  kchunk = cdiv(kchunk, BK) * BK;
  splits = cdiv(g.K, kchunk);
  g.kchunk = kchunk;
  g.ws = cx.ws;
  dim3 grid(tn, tm, splits);
  rtk::k_sgemm<BM, BN, BK, 8, 4><<<grid, 256, 0, st>>>(g);
  RT_LAUNCH_CHECK();
  cx.simt_launches++;
  cx.last_splits = splits;
  if (splits > 1 && !cx.defer_reduce) {
    launch_splitk_reduce(st, g, splits);
    RT_LAUNCH_CHECK();
  }
  return RT_OK;
}

int get_tmap(GemmCtx& cx, const float* ptr, uint64_t d0, uint64_t d1, uint64_t pitch_elems, uint32_t b0,
             uint32_t b1, int mn_major, const CUtensorMap** out) {
  TmapKey key{ptr, d0, d1, pitch_elems, b0, b1, mn_major};
  auto it = cx.tmaps.find(key);
  if (it == cx.tmaps.end()) {
    if (!cx.encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      RT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (!fn || qres != cudaDriverEntryPointSuccess)
        return rt::fail(RT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
      cx.encode = (PFN_tmapEncodeTiled)fn;
    }
    CUtensorMap tm;
    cuuint64_t gdim[2] = {d0, d1};
    cuuint64_t gstr[1] = {pitch_elems * sizeof(float)};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cx.encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return rt::fail(RT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu) pitch=%llu box=(%u,%u)",
                      (int)r, (unsigned long long)d0, (unsigned long long)d1,
                      (unsigned long long)pitch_elems, b0, b1);
    it = cx.tmaps.emplace(key, tm).first;
  }
  *out = &it->second;
  return RT_OK;
}

template <int BN, int A_MN, int B_MN, int STAGES>
int launch_tc(const CUtensorMap* ta, const CUtensorMap* tb, const rttc::TcArgs& a, dim3 grid, cudaStream_t st) {
  using L = rttc::SmemLayout<BN, A_MN, B_MN, STAGES>;
  auto kern = rttc::k_gemm_tc<BN, A_MN, B_MN, STAGES>;
  static bool configured = false;
  if (!configured) {
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  kern<<<grid, rttc::NUM_THREADS, L::TOTAL, st>>>(*ta, *tb, a);
  RT_LAUNCH_CHECK();
  return RT_OK;
}

template <int BN, int A_MN, int B_MN, int STAGES>
int launch_tc_p(const CUtensorMap* ta, const CUtensorMap* tb, const rttc::TcArgs& a, int ctas, cudaStream_t st) {
  using L = rttc::SmemLayoutP<BN, A_MN, B_MN, STAGES>;
  auto kern = rttc::k_gemm_tc_p<BN, A_MN, B_MN, STAGES>;
  static bool configured = false;
  if (!configured) {
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  kern<<<ctas, rttc::NUM_THREADS_P, L::TOTAL, st>>>(*ta, *tb, a);
  RT_LAUNCH_CHECK();
  return RT_OK;
}

// CTA-pair variant (rttc::k_gemm_tc_pair): clusters of 2, as many pairs as the device can co-schedule
int launch_tc_pair(const CUtensorMap* ta, const CUtensorMap* tb, const rttc::TcArgs& a, int ctas, long long pair_tiles,
                   cudaStream_t st, bool* launched) {
  constexpr int STAGES = 5;
  using L = rttc::SmemLayoutPair<STAGES>;
  auto kern = rttc::k_gemm_tc_pair<STAGES>;
  static int max_pairs = 0;      // 0: not asked yet; -1: clusters of 2 with this footprint are not available
  *launched = false;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(rttc::NUM_THREADS_P);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!max_pairs) {
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    cfg.gridDim = dim3(2);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) cudaGetLastError();
    max_pairs = (e == cudaSuccess && n > 0) ? n : -1;
  }
  if (max_pairs < 0) return RT_OK;
  long long pairs = ctas / 2;
  if (pairs > max_pairs) pairs = max_pairs;
  if (pairs > pair_tiles) pairs = pair_tiles;
  if (pairs < 1) return RT_OK;
  // the same number of waves on the fewest pairs (320 tiles: 5 waves on 64 pairs instead of 74): the last
  // wave is full, and the SMs left over run the kernels of the other branches
  const long long waves = (pair_tiles + pairs - 1) / pairs;
  pairs = (pair_tiles + waves - 1) / waves;
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  RT_CUDA(cudaLaunchKernelEx(&cfg, kern, *ta, *tb, a));
  rt::launch_counter()++;
  *launched = true;
  return RT_OK;
}

// persistent variant: deepest pipeline that fits one CTA per SM next to the epilogue scratch
template <int BN, int A_MN, int B_MN>
int launch_tc_persistent(const CUtensorMap* ta, const CUtensorMap* tb, const rttc::TcArgs& a, int ctas,
                         cudaStream_t st) {
  constexpr int STAGE_BYTES = rttc::BLOCK_M * rttc::BLOCK_K * 4 + BN * rttc::BLOCK_K * 4;
  constexpr int STAGES = (180 * 1024) / STAGE_BYTES >= 6 ? 6 : ((180 * 1024) / STAGE_BYTES >= 5 ? 5 : (180 * 1024) / STAGE_BYTES);
  return launch_tc_p<BN, A_MN, B_MN, STAGES>(ta, tb, a, ctas, st);
}

template <int BN, int A_MN, int B_MN>
int launch_tc_stages(int stages, const CUtensorMap* ta, const CUtensorMap* tb, const rttc::TcArgs& a, dim3 grid,
                     cudaStream_t st) {
  constexpr int STAGE_BYTES = rttc::BLOCK_M * rttc::BLOCK_K * 4 + BN * rttc::BLOCK_K * 4;
  if (stages >= 6 && 6 * STAGE_BYTES <= 200 * 1024) return launch_tc<BN, A_MN, B_MN, 6>(ta, tb, a, grid, st);
  if (stages >= 4 && 4 * STAGE_BYTES <= 200 * 1024) return launch_tc<BN, A_MN, B_MN, 4>(ta, tb, a, grid, st);
  return launch_tc<BN, A_MN, B_MN, 3>(ta, tb, a, grid, st);
}

bool tc_eligible(const rtk::GemmArgs& g) {
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (g.transA && g.transB) return false;            // (MN-major A, K-major B) is never needed
  if (g.K < 32 || g.N < 8 || g.M < 1) return false;
  if ((g.lda & 3) || (g.ldb & 3) || !al16(g.A) || !al16(g.B)) return false;
  if (!g.transA && g.K > g.lda) return false;
  if (g.transA && g.M > g.lda) return false;
  return true;
}

int gemm_tc(GemmCtx& cx, cudaStream_t st, rtk::GemmArgs g) {
  const int A_MN = g.transA ? 1 : 0, B_MN = g.transB ? 0 : 1;
  int BN = g.N <= 32 ? 32 : (g.N <= 64 ? 64 : 128);
  int tm = cdiv(g.M, rttc::BLOCK_M);
  // few row tiles AND a short K loop (e.g. the recurrent step, M = B): narrower N tiles put
  // more SMs to work; long-K products get their parallelism from split-K instead
  while (BN > 32 && (long long)tm * cdiv(g.N, BN) < 74 && (cdiv(g.K, rttc::BLOCK_K) < 64 || tm == 1)) BN >>= 1;
  // wide-N products with plenty of row tiles: 128x256 tiles re-read A half as often (the fc-size
  // GEMMs move ~330 MB through L2 at 128x128 and run at the L2 rate, not the tensor rate)
  // (not for short-K products: those are epilogue-bound and balance better over twice as many tiles)
  if (BN == 128 && g.N % 256 == 0 && (long long)tm * (g.N / 256) >= cx.num_sms && cdiv(g.K, rttc::BLOCK_K) >= 8) BN = 256;
  int stages = BN == 128 ? 3 : 4;
  if (cx.force_bn) BN = cx.force_bn;
  if (cx.force_stages) stages = cx.force_stages;
  int tn = cdiv(g.N, BN);
  long long tiles = (long long)tm * tn;
  int num_kb = cdiv(g.K, rttc::BLOCK_K);
  int splits = 1;
  if (tiles < 74 && num_kb >= 64) {
    // as many k-slices as fit ONE wave of CTAs (rounding up left a second, nearly empty round)
    splits = (int)(cx.num_sms / tiles);
    if (cx.defer_reduce) splits = (int)((cx.num_sms + tiles - 1) / tiles);   // BPTT: tuned with the cell kernel's fold
    if (splits > num_kb / cx.split_min_kb) splits = num_kb / cx.split_min_kb;
    if (splits > 32) splits = 32;
    size_t per = (size_t)g.M * g.N;
    if ((size_t)splits * per > cx.ws_floats) splits = (int)(cx.ws_floats / per);
    if (splits < 1) splits = 1;
  }
  int kbps = cdiv(num_kb, splits);
  splits = cdiv(num_kb, kbps);
  const CUtensorMap *ta = nullptr, *tb = nullptr;
  if (A_MN) RT_TRY(get_tmap(cx, g.A, g.M, g.K, g.lda, 32, rttc::BLOCK_K, 1, &ta));
  else      RT_TRY(get_tmap(cx, g.A, g.K, g.M, g.lda, rttc::BLOCK_K, rttc::BLOCK_M, 0, &ta));
  if (B_MN) RT_TRY(get_tmap(cx, g.B, g.N, g.K, g.ldb, 32, rttc::BLOCK_K, 1, &tb));
  else      RT_TRY(get_tmap(cx, g.B, g.K, g.N, g.ldb, rttc::BLOCK_K, BN, 0, &tb));
  rttc::TcArgs a;
  g.ws = cx.ws;
  g.kchunk = kbps * rttc::BLOCK_K;
  a.g = g;
  a.num_kb_total = num_kb;
  a.kb_per_split = kbps;
  a.round_tf32 = cx.round_tf32 || g.round_out;
  a.dbg = cx.dbg;
  dim3 grid(tn, tm, splits);
  int rc;
  const long long total_tiles = (long long)tn * tm * splits;
  const bool persistent = cx.persistent && total_tiles > cx.num_sms;
  const int ctas = (int)(total_tiles < cx.num_sms ? total_tiles : cx.num_sms);
  if (cx.pair && persistent && BN == 256 && !A_MN && !B_MN && splits == 1 && g.N % 256 == 0) {
    // the wide K-major products: two SMs per 256 x 256 tile, each staging half of B
    const CUtensorMap* tb2 = nullptr;
    RT_TRY(get_tmap(cx, g.B, g.K, g.N, g.ldb, rttc::BLOCK_K, rttc::PAIR_BN / 2, 0, &tb2));
    bool launched = false;
    RT_TRY(launch_tc_pair(ta, tb2, a, ctas, (long long)cdiv(g.M, 2 * rttc::BLOCK_M) * (g.N / 256), st, &launched));
    if (launched) {
      cx.tc_launches++;
      cx.last_splits = 1;
      return RT_OK;
    }
  }
#define RT_TC_CASE(bn, am, bm)                                                                  \
  if (BN == bn && A_MN == am && B_MN == bm)                                                     \
    rc = persistent ? launch_tc_persistent<bn, am, bm>(ta, tb, a, ctas, st)                     \
                    : launch_tc_stages<bn, am, bm>(stages, ta, tb, a, grid, st);                \
  else
  RT_TC_CASE(32, 0, 0) RT_TC_CASE(64, 0, 0) RT_TC_CASE(128, 0, 0) RT_TC_CASE(256, 0, 0)
  RT_TC_CASE(32, 0, 1) RT_TC_CASE(64, 0, 1) RT_TC_CASE(128, 0, 1) RT_TC_CASE(256, 0, 1)
  RT_TC_CASE(32, 1, 1) RT_TC_CASE(64, 1, 1) RT_TC_CASE(128, 1, 1) RT_TC_CASE(256, 1, 1)
  rc = rt::fail(RT_ERR_INVALID, "no tcgen05 GEMM instantiation for BN=%d A_MN=%d B_MN=%d", BN, A_MN, B_MN);
#undef RT_TC_CASE
  if (rc != RT_OK) return rc;
  cx.tc_launches++;
  cx.last_splits = splits;
  if (splits > 1 && !cx.defer_reduce) {
    launch_splitk_reduce(st, a.g, splits);
    RT_LAUNCH_CHECK();
  }
  return RT_OK;
}

struct ProfScope {   // brackets one GEMM-shaped launch with events when profiling is on
  GemmCtx& cx;
  cudaStream_t st;
  bool on;
  ProfScope(GemmCtx& c, cudaStream_t s, double flops, int kind = 0, long long M = 0, long long N = 0,
            long long K = 0, int tA = 0, int tB = 0) : cx(c), st(s) {
    on = cx.profile && cx.prof_used + 2 <= cx.prof_ev.size();
    if (on) {
      cudaEventRecord(cx.prof_ev[cx.prof_used], st);
      cx.prof_flops += flops;
      cx.prof_fl.push_back(flops);
      const int rec[6] = {kind, (int)M, (int)N, (int)K, tA, tB};
      cx.prof_shape.insert(cx.prof_shape.end(), rec, rec + 6);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(cx.prof_ev[cx.prof_used + 1], st);
      cx.prof_used += 2;
    }
  }
};

int gemm(GemmCtx& cx, cudaStream_t st, rtk::GemmArgs g) {
  if (g.M <= 0 || g.N <= 0) return RT_OK;
  ProfScope ps(cx, st, 2.0 * g.M * g.N * g.K, 0, g.M, g.N, g.K, g.transA, g.transB);
  if (cx.mode == 1 && tc_eligible(g)) return gemm_tc(cx, st, g);
  return gemm_simt(cx, st, g);
}

rtk::GemmArgs mk(const float* A, int lda, int transA, const float* B, int ldb, int transB, float* C,
                 int ldc, int M, int N, int K) {
  rtk::GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.transA = transA; g.transB = transB;
  g.alpha = 1.f;
  return g;
}

int colsum(rt_learner* h, cudaStream_t st, const float* x, size_t rows, int N, float* out,
           int accumulate, float* scratch = nullptr) {
  float* part = scratch ? scratch : h->colsum_part;
  // enough row slabs to fill the GPU even when N is a single 32-column block
  int col_blocks = cdiv(N, 32);
  int parts = cdiv(592, col_blocks);
  if ((size_t)parts * 64 > rows) parts = cdiv(rows, 64);
  if (parts > 2048) parts = 2048;
  if (parts < 1) parts = 1;
  int rpb = cdiv(rows, parts);
  parts = cdiv(rows, rpb);
  dim3 grid(col_blocks, parts);
  rtk::k_colsum_partial<<<grid, dim3(32, 8), 0, st>>>(x, part, rows, N, rpb);
  RT_LAUNCH_CHECK();
  rtk::k_colsum_final<<<cdiv(N, 32), dim3(32, 8), 0, st>>>(part, out, parts, N, accumulate, N);
  RT_LAUNCH_CHECK();
  return RT_OK;
}

int grid1d(size_t n, int threads) {
  size_t b = (n + threads - 1) / threads;
  if (b > 148 * 32) b = 148 * 32;
  if (b < 1) b = 1;
  return (int)b;
}

// uint8 NCHW frames (the replay batch) -> fp32 NHWC * (1/255) (cnn.py:44-45), once per pass
int launch_frames_to_nhwc(cudaStream_t st, const uint8_t* x, float* xf, int rows, int C, int H, int W, float scale,
                          int rn) {
  size_t pixels = (size_t)rows * H * W;
  rtk::k_u8_nchw_to_f32_nhwc<<<grid1d(pixels), 256, 0, st>>>(x, xf, pixels, C, H * W, scale, rn);
  RT_LAUNCH_CHECK();
  return RT_OK;
}

int launch_im2col_f32(cudaStream_t st, const float* xin, float* col, int rc, const ConvL& L) {
  size_t n = (size_t)rc * L.hout * L.wout * L.K;
  if (L.cin % 4 == 0)
    rtk::k_im2col_f32_nhwc<4><<<grid1d(n / 4), 256, 0, st>>>(xin, col, rc, L.cin, L.hin, L.win, L.k, L.s,
                                                             L.hout, L.wout);
  else
    rtk::k_im2col_f32_nhwc<1><<<grid1d(n), 256, 0, st>>>(xin, col, rc, L.cin, L.hin, L.win, L.k, L.s,
                                                         L.hout, L.wout);
  RT_LAUNCH_CHECK();
  return RT_OK;
}

template <int BN, int IN_U8>
int launch_conv_tc(const CUtensorMap* tb, const rttc::ConvArgs& a, cudaStream_t st) {
  constexpr int STAGES = 4;
  constexpr int SMEM = STAGES * (rttc::BLOCK_M * rttc::BLOCK_K * 4 + BN * rttc::BLOCK_K * 4) +
                       (2 * STAGES + 1) * 8 + 16 + 1024;
  auto kern = rttc::k_conv_tc<BN, IN_U8, STAGES>;
  static bool configured = false;
  if (!configured) {
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dim3 grid(cdiv(a.N, BN), cdiv(a.M, rttc::BLOCK_M));
  kern<<<grid, rttc::NUM_THREADS, SMEM, st>>>(*tb, a);
  RT_LAUNCH_CHECK();
  return RT_OK;
}

// SHALLOW: a ring small enough (<= ~92 KB with the epilogue scratch) for two of these CTAs per SM, so
// the target and online networks' convolutions of the two-branch forward pass share every SM
template <int BN, int IN_U8, bool SHALLOW = false>
int launch_conv_tc_p(const CUtensorMap* tb, const rttc::ConvArgs& a, int ctas, cudaStream_t st) {
  constexpr int STAGE_BYTES = rttc::BLOCK_M * rttc::BLOCK_K * 4 + BN * rttc::BLOCK_K * 4;
  constexpr int DEEP = (160 * 1024) / STAGE_BYTES >= 6 ? 6 : (160 * 1024) / STAGE_BYTES;
  constexpr int STAGES = SHALLOW ? ((72 * 1024) / STAGE_BYTES >= 4 ? 4 : (72 * 1024) / STAGE_BYTES) : DEEP;
  constexpr int SMEM = STAGES * STAGE_BYTES + 4 * 32 * 36 * 4 + (2 * STAGES + 4) * 8 + 16 + 1024;
  auto kern = rttc::k_conv_tc_p<BN, IN_U8, STAGES>;
  static bool configured = false;
  if (!configured) {
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  kern<<<ctas, rttc::CONV_P_THREADS, SMEM, st>>>(*tb, a);
  RT_LAUNCH_CHECK();
  return RT_OK;
}

// Implicit-GEMM forward convolution (no im2col buffer) when the layer shape allows it.
bool conv_tc_eligible(const rt_learner* h, size_t i, const void* xin) {
  const ConvL& L = h->conv[i];
  if (h->gx.mode != 1 || !h->conv_implicit) return false;
  if (L.K % 32 || L.f % 4 || L.f > 128) return false;
  // fp32 NHWC input: every 32-float k-block must be one contiguous run inside a patch row
  // (kw, c): whole channel blocks (C % 32 == 0) or whole patch rows (conv1: 8 pixels x 4 channels)
  return L.cin % 4 == 0 && (L.cin % 32 == 0 || (L.k * L.cin) % 32 == 0 && 32 % L.cin == 0) &&
         ((uintptr_t)xin & 15) == 0;
}

int conv_forward_tc(rt_learner* h, GemmCtx& cx, cudaStream_t st, const float* net, size_t i, const void* xin,
                    float* out, int rows) {
  const ConvL& L = h->conv[i];
  rttc::ConvArgs a;
  a.in = xin; a.out = out; a.bias = net + L.b; a.rows = rows;
  a.C = L.cin; a.H = L.hin; a.W = L.win; a.KH = L.k; a.S = L.s; a.OH = L.hout; a.OW = L.wout;
  a.M = rows * L.hout * L.wout; a.N = L.f; a.K = L.K;
  a.scale = (float)(1.0 / 255.0);
  a.round_tf32 = cx.round_tf32 || h->rn;   // the output is the A operand of the next layer's product
  a.out2 = nullptr; a.bias2 = nullptr; a.split2 = 0; a.row_shift2 = 0;
  const int BN = L.f <= 32 ? 32 : (L.f <= 64 ? 64 : 128);
  const CUtensorMap* tb = nullptr;
  RT_TRY(get_tmap(cx, net + L.w, L.K, L.f, L.K, rttc::BLOCK_K, BN, 0, &tb));
  cx.tc_launches++;
  ProfScope ps(cx, st, 2.0 * a.M * a.N * a.K, 1, a.M, a.N, a.K);
  const long long tiles = (long long)cdiv(a.M, rttc::BLOCK_M) * cdiv(a.N, BN);
  // measured: for the gather-bound convolutions 2-3 one-tile CTAs per SM beat one persistent
  // CTA (conv1 103 vs 120 us, conv2/3 34 vs 50 us), so the persistent form is opt-in
  if (h->conv_persistent && tiles > h->num_sms) {
    const int ctas = h->num_sms;
    const bool shallow = h->conv_shallow && h->side_active;
    if (shallow && BN == 32) return launch_conv_tc_p<32, 0, true>(tb, a, ctas, st);
    if (shallow && BN == 64) return launch_conv_tc_p<64, 0, true>(tb, a, ctas, st);
    if (BN == 32) return launch_conv_tc_p<32, 0>(tb, a, ctas, st);
    if (BN == 64) return launch_conv_tc_p<64, 0>(tb, a, ctas, st);
    return launch_conv_tc_p<128, 0>(tb, a, ctas, st);
  }
  if (BN == 32) return launch_conv_tc<32, 0>(tb, a, st);
  if (BN == 64) return launch_conv_tc<64, 0>(tb, a, st);
  return launch_conv_tc<128, 0>(tb, a, st);
}

// conv1 of the online AND the target network in ONE implicit GEMM (N = 2 x 32 filters) over the shared
// fp32 frames: the online network needs rows [0, rows), the target network the suffix [shift_rows, rows)
// (target_states = states shifted by n steps).  The 4x-overlapping patch gather -- the L2 traffic that
// bounds this layer -- is done once instead of twice.
bool conv1_pair_eligible(const rt_learner* h, int shift_rows) {
  if (h->conv.empty() || !h->conv1_pair || !h->wcat) return false;
  const ConvL& L = h->conv[0];
  const long long shift = (long long)shift_rows * L.hout * L.wout;
  return L.f == 32 && conv_tc_eligible(h, 0, h->xf) && shift % rttc::BLOCK_M == 0 && h->conv_persistent;
}
int conv1_pair_forward(rt_learner* h, cudaStream_t st, const float* net_a, const float* net_b, const float* xf,
                       int rows, int shift_rows) {
  const ConvL& L = h->conv[0];
  const size_t nw = (size_t)L.f * L.K;
  RT_CUDA(cudaMemcpyAsync(h->wcat, net_a + L.w, nw * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RT_CUDA(cudaMemcpyAsync(h->wcat + nw, net_b + L.w, nw * sizeof(float), cudaMemcpyDeviceToDevice, st));
  rttc::ConvArgs a;
  a.in = xf; a.out = h->c_out[0]; a.bias = net_a + L.b; a.rows = rows;
  a.C = L.cin; a.H = L.hin; a.W = L.win; a.KH = L.k; a.S = L.s; a.OH = L.hout; a.OW = L.wout;
  a.M = rows * L.hout * L.wout; a.N = 2 * L.f; a.K = L.K;
  a.scale = (float)(1.0 / 255.0);
  a.round_tf32 = h->rn;
  a.out2 = h->c_out2[0]; a.bias2 = net_b + L.b; a.split2 = 1; a.row_shift2 = shift_rows * L.hout * L.wout;
  const CUtensorMap* tb = nullptr;
  RT_TRY(get_tmap(h->gx, h->wcat, L.K, 2 * L.f, L.K, rttc::BLOCK_K, 64, 0, &tb));
  h->gx.tc_launches++;
  ProfScope ps(h->gx, st, 2.0 * a.M * a.N * a.K, 1, a.M, a.N, a.K);
  const long long tiles = cdiv(a.M, rttc::BLOCK_M);
  const int ctas = (int)(tiles < h->num_sms ? tiles : h->num_sms);
  return launch_conv_tc_p<64, 0>(tb, a, ctas, st);
}

template <int BN, int IN_U8, int STAGES = 4>
int launch_convdw_tc(const CUtensorMap* ta, const rttc::ConvDwArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int SMEM = STAGES * (rttc::BLOCK_M * rttc::BLOCK_K * 4 + BN * rttc::BLOCK_K * 4) +
                       (2 * STAGES + 1) * 8 + 16 + 1024;
  auto kern = rttc::k_convdw_tc<BN, IN_U8, STAGES>;
  static bool configured = false;
  if (!configured) {
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  kern<<<grid, rttc::CONVDW_THREADS, SMEM, st>>>(*ta, a);
  RT_LAUNCH_CHECK();
  return RT_OK;
}

// Implicit-GEMM weight gradient of conv layer i over all `rows` frames: dW = dy^T . col.
int conv_dw_tc(rt_learner* h, GemmCtx& cx, cudaStream_t st, size_t i, const void* xin, const float* dy, int rows,
               float* dW) {
  const ConvL& L = h->conv[i];
  rttc::ConvDwArgs a;
  a.in = xin; a.ws = cx.ws;
  a.C = L.cin; a.H = L.hin; a.W = L.win; a.KH = L.k; a.S = L.s; a.OH = L.hout; a.OW = L.wout;
  a.P = rows * L.hout * L.wout; a.F = L.f; a.K = L.K;
  a.scale = (float)(1.0 / 255.0);
  const int BN = L.K <= 32 ? 32 : (L.K <= 64 ? 64 : 128);
  int tiles = cdiv(L.K, BN);
  int total_kb = cdiv(a.P, rttc::BLOCK_K);
  int splits = cdiv(296, tiles);
  if (splits > total_kb / 8) splits = total_kb / 8;
  size_t per = (size_t)L.f * L.K;
  if ((size_t)splits * per > cx.ws_floats) splits = (int)(cx.ws_floats / per);
  if (splits < 1) splits = 1;
  a.kb_per_split = cdiv(total_kb, splits);
  splits = cdiv(total_kb, a.kb_per_split);
  const CUtensorMap* ta = nullptr;
  RT_TRY(get_tmap(cx, dy, L.f, a.P, L.f, 32, rttc::BLOCK_K, 1, &ta));
  dim3 grid(tiles, 1, splits);
  cx.tc_launches++;
  ProfScope ps(cx, st, 2.0 * a.P * (double)L.f * L.K, 2, L.f, L.K, a.P);
  const bool shallow = h->conv_shallow >= 2 && h->side_active;
  if (BN == 32) RT_TRY((launch_convdw_tc<32, 0>(ta, a, grid, st)));
  else if (BN == 64) RT_TRY((launch_convdw_tc<64, 0>(ta, a, grid, st)));
  else if (shallow) RT_TRY((launch_convdw_tc<128, 0, 3>(ta, a, grid, st)));
  else RT_TRY((launch_convdw_tc<128, 0>(ta, a, grid, st)));
  rtk::GemmArgs g = mk(nullptr, 0, 0, nullptr, 0, 0, dW, L.K, L.f, L.K, a.P);
  g.ws = cx.ws;
  launch_splitk_reduce(st, g, splits);
  RT_LAUNCH_CHECK();
  return RT_OK;
}

template <int BN, bool SHALLOW = false>
int launch_convdx_tc(const CUtensorMap* tb, const rttc::ConvDxArgs& a, int ctas, cudaStream_t st) {
  constexpr int STAGE_BYTES = rttc::BLOCK_M * rttc::BLOCK_K * 4 + BN * rttc::BLOCK_K * 4;
  constexpr int DEEP = (160 * 1024) / STAGE_BYTES >= 6 ? 6 : (160 * 1024) / STAGE_BYTES;
  constexpr int STAGES = SHALLOW ? ((80 * 1024) / STAGE_BYTES >= 4 ? 4 : (80 * 1024) / STAGE_BYTES) : DEEP;
  constexpr int SMEM = STAGES * STAGE_BYTES + (2 * STAGES + 4) * 8 + 16 + 1024;
  auto kern = rttc::k_convdx_tc<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  kern<<<ctas, rttc::CONV_P_THREADS, SMEM, st>>>(*tb, a);
  RT_LAUNCH_CHECK();
  return RT_OK;
}

bool conv_dx_tc_eligible(const rt_learner* h, int i) {
  const ConvL& L = h->conv[i];
  return h->gx.mode == 1 && h->conv_dx_implicit && i > 0 && L.k % L.s == 0 && L.f % 32 == 0 &&
         (L.cin == 32 || L.cin == 64 || L.cin == 128) && h->conv_wt[i] != nullptr;
}

// Data gradient of conv layer i (i > 0), ReLU mask of the previous layer fused:
// d_c[i-1] = relu'(c_out[i-1]) * conv_transpose(dy, W_i)
int conv_dx_tc(rt_learner* h, cudaStream_t st, const float* net, int i, const float* dy, int rows) {
  const ConvL& L = h->conv[i];
  const size_t nw = (size_t)L.f * L.K;
  rtk::k_conv_wT<<<grid1d(nw), 256, 0, st>>>(net + L.w, h->conv_wt[i], L.f, L.k, L.s, L.cin);
  RT_LAUNCH_CHECK();
  rttc::ConvDxArgs a;
  a.dy = dy; a.act = h->c_out[i - 1]; a.dx = h->d_c[i - 1];
  a.rows = rows; a.C = L.cin; a.H = L.hin; a.W = L.win; a.KH = L.k; a.S = L.s; a.OH = L.hout; a.OW = L.wout;
  a.F = L.f;
  a.Hc = cdiv(L.hin, L.s); a.Wc = cdiv(L.win, L.s);
  a.KD = L.k / L.s;
  a.Kc = a.KD * a.KD * L.f;
  a.Mc = rows * a.Hc * a.Wc;
  const int BN = L.cin;
  const CUtensorMap* tb = nullptr;
  RT_TRY(get_tmap(h->gx, h->conv_wt[i], a.Kc, (uint64_t)L.s * L.s * L.cin, a.Kc, rttc::BLOCK_K, BN, 0, &tb));
  const long long tiles = (long long)L.s * L.s * cdiv(a.Mc, rttc::BLOCK_M);
  // data-parallel: this persistent kernel runs while NCCL reduces the non-conv gradient bucket; CTAs that hold
  // every SM for the whole kernel would keep NCCL's CTAs waiting, so it leaves them room
  const int avail = h->num_sms - h->reserve_sms;
  const int ctas = (int)(tiles < avail ? tiles : avail);
  h->gx.tc_launches++;
  ProfScope ps(h->gx, st, 2.0 * rows * L.hout * L.wout * (double)L.f * L.K, 3, (long long)rows * L.hin * L.win, L.cin, (long long)L.f * L.k * L.k);
  const bool shallow = h->conv_shallow >= 2 && h->side_active;
  if (shallow && BN == 32) return launch_convdx_tc<32, true>(tb, a, ctas, st);
  if (shallow && BN == 64) return launch_convdx_tc<64, true>(tb, a, ctas, st);
  if (BN == 32) return launch_convdx_tc<32>(tb, a, ctas, st);
  if (BN == 64) return launch_convdx_tc<64>(tb, a, ctas, st);
  return launch_convdx_tc<128>(tb, a, ctas, st);
}

// CNN forward for `rows` frames: frames to fp32 NHWC once, then every layer is an implicit GEMM
// over NHWC runs (fallback: im2col + GEMM, chunked so the im2col buffers stay L2-resident).
// `xf_pre`: the frames of this pass already converted (a slice of h->xf written by the caller).
// `second`: write the conv outputs to the second activation set with the side branch's GEMM context
// (only the implicit-GEMM path; callers check cnn_all_implicit first).
bool cnn_all_implicit(const rt_learner* h) {
  bool all = true;
  for (size_t i = 0; i < h->conv.size(); ++i)
    all = all && conv_tc_eligible(h, i, i == 0 ? (const void*)h->xf : (const void*)h->c_out[i - 1]);
  return all;
}
int cnn_forward(rt_learner* h, cudaStream_t st, const float* net, const uint8_t* x, int rows,
                const float* xf_pre = nullptr, bool second = false, int first_layer = 0) {
  // a pass that converts its own frames (burn-in, the non-shared target / training passes, acting) uses the
  // DEFAULT frame buffer: h->xf may be a batch slot's private buffer holding frames that rt_learner_prefetch
  // already converted for the training window of this very update
  if (!xf_pre)
    RT_TRY(launch_frames_to_nhwc(st, x, h->xf0, rows, h->md.in_c, h->md.in_h, h->md.in_w, (float)(1.0 / 255.0), h->rn));
  const float* xf = xf_pre ? xf_pre : h->xf0;
  {
    bool all = true;
    for (size_t i = 0; i < h->conv.size(); ++i)
      all = all && conv_tc_eligible(h, i, i == 0 ? (const void*)xf : (const void*)h->c_out[i - 1]);
    if (all) {
      std::vector<float*>& out = second ? h->c_out2 : h->c_out;
      GemmCtx& cx = second ? h->gx2 : h->gx;
      for (size_t i = first_layer; i < h->conv.size(); ++i)
        RT_TRY(conv_forward_tc(h, cx, st, net, i, i == 0 ? (const void*)xf : (const void*)out[i - 1], out[i], rows));
      return RT_OK;
    }
  }
  RT_REQUIRE(!second && first_layer == 0, "second activation set needs the implicit-GEMM conv path");
  for (int r0 = 0; r0 < rows; r0 += h->chunk_rows) {
    int rc = rows - r0 < h->chunk_rows ? rows - r0 : h->chunk_rows;
    for (size_t i = 0; i < h->conv.size(); ++i) {
      const ConvL& L = h->conv[i];
      size_t opix = (size_t)L.hout * L.wout;
      const float* xin = (i == 0 ? xf : h->c_out[i - 1]) + (size_t)r0 * L.hin * L.win * L.cin;
      RT_TRY(launch_im2col_f32(st, xin, h->col, rc, L));
      float* out = h->c_out[i] + (size_t)r0 * opix * L.f;
      rtk::GemmArgs g = mk(h->col, L.K, 0, net + L.w, L.K, 1, out, L.f, (int)(rc * opix), L.f, L.K);
      g.bias = net + L.b;
      g.relu = 1;
      RT_TRY(gemm(h->gx, st, g));
    }
  }
  return RT_OK;
}

// Feature extractor of one pass: CNN (if any) + the FC layers in front of the LSTM / last FC module
// (fc.py:29-36: linear + ReLU).  *out = (rows, h->feat) features.  Without a CNN the observation rows
// are float32 vectors and feed the first FC layer (or the LSTM / heads) directly.
int feature_forward(rt_learner* h, cudaStream_t st, const float* net, const uint8_t* x, int rows,
                    const float** out, const float* xf_pre = nullptr, bool second = false, int first_conv = 0) {
  const float* f = reinterpret_cast<const float*>(x);
  if (!h->conv.empty()) {
    RT_TRY(cnn_forward(h, st, net, x, rows, xf_pre, second, first_conv));
    f = (second ? h->c_out2 : h->c_out).back();
  }
  GemmCtx& cx = second ? h->gx2 : h->gx;
  for (size_t k = 0; k < h->pre.size(); ++k) {
    const PreL& L = h->pre[k];
    float* o = (second ? h->pre_out2 : h->pre_out)[k];
    rtk::GemmArgs g = mk(f, L.in, 0, net + L.w, L.in, 1, o, L.out, rows, L.out, L.in);
    g.bias = net + L.b;
    g.relu = 1;
    g.round_out = h->rn;
    RT_TRY(gemm(cx, st, g));
    f = o;
  }
  *out = f;
  return RT_OK;
}

// One LSTM recurrence (lstm.py:50-122) over time-major rows t*Beff + b.
struct SeqDesc {
  const float* net;       // parameter buffer (online / target)
  const float* xg;        // (timesteps*Beff, 4U) input gates incl. both biases
  const float* hx;        // stored state of step 0
  const float* cx;
  const float* initials;  // (timesteps*Beff)
  float* h_all;           // output (timesteps*Beff, U)
  int slot;               // exchange-buffer slot: hprev rows [slot*max_rows, ...)
  bool bptt;              // keep gates / c_all / cprev for the backward pass
};

// xg = [feat | extra] . W_ih^T + b_ih + b_hh for `rows` rows (the extra features of a tuple observation
// are concatenated at the LSTM input, sequential.py:146-165: a second skinny product into the same sums)
int lstm_xgates(rt_learner* h, cudaStream_t st, const float* net, const float* feat, int rows, float* xg,
                const float* extra, GemmCtx* cx = nullptr) {
  int U = h->U;
  rtk::GemmArgs g = mk(feat, h->feat, 0, net + h->o_wih, h->feat, 1, xg, 4 * U, rows, 4 * U, h->feat);
  g.bias = net + h->o_bih;
  g.bias2 = net + h->o_bhh;
  RT_TRY(gemm(cx ? *cx : h->gx, st, g));
  if (h->X) {
    RT_REQUIRE(extra, "the model takes %d extra features but the batch has none (rt_learner_io.field_extra)", h->X);
    rtk::GemmArgs g2 = mk(extra, h->X, 0, net + h->o_wihx, h->X, 1, xg, 4 * U, rows, 4 * U, h->X);
    g2.accumulate = 1;
    RT_TRY(gemm_simt(cx ? *cx : h->gx, st, g2));
  }
  return RT_OK;
}

// fp32 recurrence of one sequence: persistent SIMT kernel when it fits, else one GEMM + cell
// kernel per step.  gates / c_all / cprev always land in the shared BPTT buffers.
int lstm_recur_one(rt_learner* h, cudaStream_t st, const SeqDesc& q, int timesteps, int Beff) {
  int U = h->U;
  float* hprev = h->hprev + (size_t)q.slot * h->max_rows * U;
  int nb = cdiv((size_t)Beff * U, 256);
  {
    const int ctas = U / rtk::lstm_seq::UPB;
    size_t smem = ((size_t)32 * (U + rtk::lstm_seq::HPAD) + 4 * 32 * 4) * sizeof(float) + 16;
    if (h->lstm_persistent && timesteps > 1 && (U == 512 || U == 256) && Beff <= 32 &&
        ctas <= h->num_sms) {
      void (*kern)(const float*, const float*, const float*, const float*, const float*, float*, float*, float*,
                   float*, float*, float*, int, int, int, unsigned int*, long long*) =
          U == 512 ? rtk::k_lstm_seq_fwd<8> : rtk::k_lstm_seq_fwd<4>;
      static size_t configured[2] = {0, 0};
      if (configured[U == 512] < smem) {
        RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[U == 512] = smem;
      }
      RT_CUDA(cudaMemsetAsync(h->grid_barrier, 0, sizeof(unsigned int), st));
      const float* whh = q.net + h->o_whh;
      void* args[] = {(void*)&q.xg, (void*)&whh, (void*)&q.hx, (void*)&q.cx, (void*)&q.initials,
                      (void*)&h->gates, (void*)&h->c_all, (void*)&q.h_all, (void*)&hprev,
                      (void*)&h->cprev, (void*)&h->lstm_hrep, (void*)&timesteps, (void*)&Beff, (void*)&U,
                      (void*)&h->grid_barrier, (void*)&h->lstm_dbg};
      // cooperative launch: the runtime guarantees all CTAs are co-resident (grid barrier)
      RT_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(ctas), dim3(256), args, smem, st));
      rt::launch_counter()++;
      return RT_OK;
    }
  }
  rtk::k_lstm_init<<<nb, 256, 0, st>>>(q.hx, q.cx, q.initials, hprev, h->cprev, Beff, U);
  RT_LAUNCH_CHECK();
  for (int t = 0; t < timesteps; ++t) {
    size_t ro = (size_t)t * Beff;
    rtk::GemmArgs gh = mk(hprev + ro * U, U, 0, q.net + h->o_whh, U, 1, h->hg, 4 * U, Beff, 4 * U, U);
    RT_TRY(gemm(h->gx, st, gh));
    bool last = t == timesteps - 1;
    rtk::k_lstm_cell<<<nb, 256, 0, st>>>(
        q.xg + ro * 4 * U, h->hg, h->cprev + ro * U, h->gates + ro * 4 * U, h->c_all + ro * U,
        q.h_all + ro * U, last ? nullptr : q.initials + ro + Beff,
        last ? nullptr : hprev + (ro + Beff) * U, last ? nullptr : h->cprev + (ro + Beff) * U, Beff, U);
    RT_LAUNCH_CHECK();
  }
  return RT_OK;
}

template <int UPC>
int launch_lstm_tc(rt_learner* h, cudaStream_t st, const CUtensorMap* w0, const CUtensorMap* w1,
                   const rttc::LstmTcArgs& a, int ctas) {
  auto kern = rttc::k_lstm_seq_tc<UPC>;
  const int smem = rttc::LstmSmem<UPC>::total(a.U, a.arows);
  static int configured = 0;
  if (configured < smem) {
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  RT_CUDA(cudaMemsetAsync(h->grid_barrier, 0, 64 * sizeof(unsigned int), st));
  void* args[] = {(void*)w0, (void*)w1, (void*)&a};
  // cooperative launch: all CTAs co-resident (they wait on each other's step counters)
  RT_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(ctas), dim3(rttc::LSTM_TC_THREADS), args, (size_t)smem, st));
  rt::launch_counter()++;
  return RT_OK;
}

template <int MT>
int launch_lstm_mma(rt_learner* h, cudaStream_t st, const rttc::LstmTcArgs& a, const float* whh0, const float* whh1,
                    int ctas) {
  auto kern = rttc::k_lstm_seq_mma<MT>;
  const int smem = rttc::LstmMmaSmem::total(a.U, 16 * MT);
  static int configured = 0;
  if (configured < smem) {
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  RT_CUDA(cudaMemsetAsync(h->grid_barrier, 0, 64 * sizeof(unsigned int), st));
  void* args[] = {(void*)&a, (void*)&whh0, (void*)&whh1};
  RT_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(ctas), dim3(rttc::LSTM_MMA_THREADS), args, (size_t)smem, st));
  rt::launch_counter()++;
  return RT_OK;
}

// All recurrences of one phase.  TF32 path: ONE tensor-core launch runs them side by side
// (sequences of the same network share a CTA group, see rt_lstm_tc.cuh); otherwise one after
// the other on the fp32 path (the BPTT sequence last: it owns gates / c_all / cprev).
int lstm_run(rt_learner* h, cudaStream_t st, const SeqDesc* seqs, int nseq, int timesteps, int Beff) {
  const int U = h->U;
  const float* nets[2] = {nullptr, nullptr};
  int groups = 0;
  bool ok = h->gx.mode == 1 && h->lstm_tc && timesteps > 1 && Beff <= 32 && U % 32 == 0 && U <= 512 &&
            timesteps <= h->lstm_tcap;
  rttc::LstmTcArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < nseq && ok; ++i) {
    int g = 0;
    while (g < groups && nets[g] != seqs[i].net) ++g;
    if (g == groups) {
      if (groups == 2) { ok = false; break; }
      nets[groups++] = seqs[i].net;
    }
    if (a.nseq[g] == rttc::LSTM_MAX_SEQ) { ok = false; break; }
    rttc::LstmSeq& q = a.seq[g][a.nseq[g]++];
    q.xg = seqs[i].xg; q.hx = seqs[i].hx; q.cx = seqs[i].cx; q.initials = seqs[i].initials;
    q.h_all = seqs[i].h_all;
    q.gates = seqs[i].bptt ? h->gates : nullptr;
    q.c_all = seqs[i].bptt ? h->c_all : nullptr;
    q.cprev = seqs[i].bptt ? h->cprev : nullptr;
    q.hprev = seqs[i].bptt ? h->hprev + (size_t)seqs[i].slot * h->max_rows * U : nullptr;
  }
  // narrowest CTA slice (most CTAs, shortest MMA + epilogue per step) that is co-resident and
  // whose W slice + one step of h fit in shared memory
  int upc = 0;
  if (ok) {
    a.arows = 32 * (a.nseq[0] > a.nseq[1] ? a.nseq[0] : a.nseq[1]);
    const int cand[2] = {h->lstm_upc == 16 ? 16 : 8, 16};
    for (int i = 0; i < 2 && !upc; ++i) {
      const int c = cand[i];
      const int smem = c == 8 ? rttc::LstmSmem<8>::total(U, a.arows) : rttc::LstmSmem<16>::total(U, a.arows);
      if (U % c == 0 && groups * (U / c) <= h->num_sms && smem <= 227 * 1024) upc = c;
    }
    if (!upc) ok = false;
  }
  if (ok) {
    a.T = timesteps; a.B = Beff; a.U = U;
    if (const char* e = getenv("RT_LSTM_EXP")) a.exp = atoi(e);
    a.rn = h->rn;
    a.xchg = h->lstm_xchg;
    a.counters = h->grid_barrier;
    a.dbg = h->lstm_dbg;
    // mma.sync variant (register-resident W_hh slice, rt_lstm_tc.cuh): U = 512, at most two sequences per
    // weight group, 8 units per CTA
    const int maxseq = a.nseq[0] > a.nseq[1] ? a.nseq[0] : a.nseq[1];
    if (h->lstm_mma && U == 512 && maxseq <= 2 && groups * (U / 8) <= h->num_sms &&
        rttc::LstmMmaSmem::total(U, 32 * maxseq) <= 227 * 1024) {
      const int ctas = groups * (U / 8);
      const float* whh0 = nets[0] + h->o_whh;
      const float* whh1 = nets[groups - 1] + h->o_whh;
      if (maxseq == 1) return launch_lstm_mma<2>(h, st, a, whh0, whh1, ctas);
      return launch_lstm_mma<4>(h, st, a, whh0, whh1, ctas);
    }
    const CUtensorMap *w0 = nullptr, *w1 = nullptr;
    RT_TRY(get_tmap(h->gx, nets[0] + h->o_whh, U, 4 * U, U, rttc::BLOCK_K, upc, 0, &w0));
    RT_TRY(get_tmap(h->gx, nets[groups - 1] + h->o_whh, U, 4 * U, U, rttc::BLOCK_K, upc, 0, &w1));
    const int ctas = groups * (U / upc);
    if (upc == 8) return launch_lstm_tc<8>(h, st, w0, w1, a, ctas);
    return launch_lstm_tc<16>(h, st, w0, w1, a, ctas);
  }
  for (int i = 0; i < nseq; ++i) RT_TRY(lstm_recur_one(h, st, seqs[i], timesteps, Beff));
  return RT_OK;
}

// LSTM forward of one sequence set: input gates + recurrence, output in h->h_all (slot 0).
int lstm_forward(rt_learner* h, cudaStream_t st, const float* net, const float* feat, int rows,
                 int timesteps, const float* hx, const float* cx, const float* initials, const float* extra) {
  RT_TRY(lstm_xgates(h, st, net, feat, rows, h->xg, extra));
  SeqDesc q{net, h->xg, hx, cx, initials, h->h_all, 0, true};
  return lstm_run(h, st, &q, 1, timesteps, rows / timesteps);
}

struct StateView {  // device pointers to the (rows, ...) leaves of a batch slice
  const uint8_t* x;       // uint8 frames; float32 vectors when the model has no CNN
  float* hx;
  float* cx;
  const float* initials;
  const float* extra;     // (rows, X) extra features of a tuple observation, or null
};

// CNN (+LSTM); returns the feature pointer [rows, D-or-feat] that feeds the heads.
int trunk_forward(rt_learner* h, cudaStream_t st, const float* net, const StateView& sv, int rows,
                  int timesteps, const float** feat_out) {
  const float* feat = nullptr;
  RT_TRY(feature_forward(h, st, net, sv.x, rows, &feat));
  if (h->U) {
    RT_TRY(lstm_forward(h, st, net, feat, rows, timesteps, sv.hx, sv.cx, sv.initials, sv.extra));
    feat = h->h_all;
  }
  *feat_out = feat;
  return RT_OK;
}

// IQN quantile layer + FC + out (+ dueling) on M rows (iqn.py:67-122, dqn.py:74-112).
// Scratch / outputs of one heads pass.  The training pass must use the primary set (its backward
// reads cf, phi, xq, h1, v1); q_out is where the (rows, A) action values land.
struct HeadSet {
  float *cf, *phi, *xq, *h1, *v1, *adv, *v, *q_out;
  GemmCtx* gx;
};
HeadSet primary_set(rt_learner* h, float* q_out) {
  return HeadSet{h->cf, h->phi, h->xq, h->h1, h->v1, h->adv, h->v, q_out, &h->gx};
}
HeadSet second_set(rt_learner* h, float* q_out) {
  return HeadSet{h->cf2, h->phi2, h->xq2, h->h1b, h->v1b, h->adv2, h->vb2, q_out, &h->gx2};
}
HeadSet third_set(rt_learner* h, float* q_out) {
  return HeadSet{h->cf3, h->phi3, h->xq3, h->h1c, h->v1c, h->adv3, h->vb3, q_out, &h->gx3};
}

// phase 0: the whole pass; 1: only the quantile embedding phi = relu(cos(pi i tau) Wq + bq), which does not depend
// on the trunk (it can run next to the LSTM recurrence); 2: everything after it
int heads_forward(rt_learner* h, cudaStream_t st, const float* net, const float* feat, int M,
                  const float* tau, const HeadSet* set = nullptr, int phase = 0) {
  const HeadSet hs = set ? *set : primary_set(h, h->q);
  int Nq = h->Nq, D = h->D, F = h->F, A = h->A, E = h->E;
  size_t MQ = (size_t)M * Nq;
  rtk::GemmArgs g;
  const float* xq = feat;     // DQN: the heads read the trunk output directly
  if (!h->dqn && phase != 2) {
    rtk::k_cos_features<<<cdiv(MQ * E, 256), 256, 0, st>>>(tau, hs.cf, (int)MQ, E, h->rn);
    RT_LAUNCH_CHECK();
    g = mk(hs.cf, E, 0, net + h->o_qw, E, 1, hs.phi, D, (int)MQ, D, E);
    g.bias = net + h->o_qb;
    g.relu = 1;
    RT_TRY(gemm(*hs.gx, st, g));
  }
  if (phase == 1) return RT_OK;
  if (!h->dqn) {
    rtk::k_quantile_mul<<<cdiv(MQ * (D / 4), 256), 256, 0, st>>>(feat, hs.phi, hs.xq, MQ, D, Nq, h->rn);
    RT_LAUNCH_CHECK();
    xq = hs.xq;
  }
  const int ldh = h->ldh;
  if (h->fused_hidden) {
    // [h1 | v1] = relu(xq . [Wfc ; Wvh]^T + [bfc | bvh]): one GEMM, the A operand is read once
    g = mk(xq, D, 0, net + h->o_fcw, D, 1, hs.h1, ldh, (int)MQ, 2 * F, D);
    g.bias = net + h->o_fcb;
    g.relu = 1;
    RT_TRY(gemm(*hs.gx, st, g));
  } else {
    g = mk(xq, D, 0, net + h->o_fcw, D, 1, hs.h1, F, (int)MQ, F, D);
    g.bias = net + h->o_fcb;
    g.relu = 1;
    RT_TRY(gemm(*hs.gx, st, g));
    if (h->dueling) {
      g = mk(xq, D, 0, net + h->o_vhw, D, 1, hs.v1, F, (int)MQ, F, D);
      g.bias = net + h->o_vhb;
      g.relu = 1;
      RT_TRY(gemm(*hs.gx, st, g));
    }
  }
  // out layer + value layer + dueling combine: one warp per 2 rows (A <= 8) / per row, grid sized
  // to the resident warps (grid-stride inside)
  {
    const float* v1 = h->dueling ? hs.v1 : nullptr;
    if (A <= 8 && F == 512) {
      static int occ = 0;
      if (!occ) RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rtk::k_heads_out<8, 2, 4>, 256, 0));
      int blocks = cdiv(cdiv(MQ, 2) * 32, 256);
      int resident = h->num_sms * (occ > 0 ? occ : 1);
      if (blocks > resident) blocks = resident;
      rtk::k_heads_out<8, 2, 4><<<blocks, 256, 0, st>>>(hs.h1, v1, net + h->o_outw, net + h->o_outb,
                                                       net + h->o_vw, net + h->o_vb, hs.adv, hs.v, hs.q_out, MQ, F, A, ldh);
    } else if (A <= 8) {
      static int occ = 0;
      if (!occ) RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rtk::k_heads_out<8, 2>, 256, 0));
      int blocks = cdiv(cdiv(MQ, 2) * 32, 256);
      int resident = h->num_sms * (occ > 0 ? occ : 1);
      if (blocks > resident) blocks = resident;
      rtk::k_heads_out<8, 2><<<blocks, 256, 0, st>>>(hs.h1, v1, net + h->o_outw, net + h->o_outb,
                                                    net + h->o_vw, net + h->o_vb, hs.adv, hs.v, hs.q_out, MQ, F, A, ldh);
    } else {
      int blocks = cdiv(MQ * 32, 256);
      if (blocks > h->num_sms * 8) blocks = h->num_sms * 8;
      rtk::k_heads_out<32, 1><<<blocks, 256, 0, st>>>(hs.h1, v1, net + h->o_outw, net + h->o_outb,
                                                     net + h->o_vw, net + h->o_vb, hs.adv, hs.v, hs.q_out, MQ, F, A, ldh);
    }
    RT_LAUNCH_CHECK();
  }
  return RT_OK;
}

// ---- two-branch backward plumbing.  side_begin() makes the side stream wait for everything
// enqueued on `st` so far and returns the stream / GEMM context / column-sum scratch a leaf (weight
// or bias gradient) should use; when the second branch is off these are the main ones.
struct SideCtx {
  cudaStream_t st;
  GemmCtx* gx;
  float* colsum_scratch;
};
int side_begin(rt_learner* h, cudaStream_t st, SideCtx* out) {
  if (!h->side_active) {
    *out = SideCtx{st, &h->gx, h->colsum_part};
    return RT_OK;
  }
  cudaEvent_t ev = h->ev_side[h->ev_side_next];
  h->ev_side_next = (h->ev_side_next + 1) % 7;     // [7] is the join event
  RT_CUDA(cudaEventRecord(ev, st));
  RT_CUDA(cudaStreamWaitEvent(h->side, ev, 0));
  *out = SideCtx{h->side, &h->gx2, h->colsum_part2};
  return RT_OK;
}
int side_join(rt_learner* h, cudaStream_t st) {
  if (!h->side_active) return RT_OK;
  RT_CUDA(cudaEventRecord(h->ev_side[7], h->side));
  RT_CUDA(cudaStreamWaitEvent(st, h->ev_side[7], 0));
  return RT_OK;
}

int heads_backward(rt_learner* h, cudaStream_t st, const float* net, const float* feat, int M,
                   const long long* actions) {
  int Nq = h->Nq, D = h->D, F = h->F, A = h->A, E = h->E;
  size_t MQ = (size_t)M * Nq;
  float* G = h->grad;
  const int duel = h->dueling ? 1 : 0;
  const float* xq = h->dqn ? feat : h->xq;
  float* dxq = h->dqn ? h->dfeatq : h->dxq;   // DQN: the data gradient of the heads IS d(loss)/d(feat)
  // small layers (out, value): data gradients (ReLU masks fused), weight / bias gradients and the
  // hidden-layer bias gradients in one pass over [h1 | v1] + one fold launch.  The slab count makes
  // the grid one full wave of resident CTAs.
  {
    const int C = F * (1 + duel);
    const int col_blocks = cdiv(C, 128);
    static int occ8 = 0, occ32 = 0;
    if (!occ8) {
      RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ8, rtk::k_heads_bwd_fused<8>, 256, 0));
      RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ32, rtk::k_heads_bwd_fused<32>, 256, 0));
    }
    int occ = A <= 8 ? occ8 : occ32;
    int slabs = (h->num_sms * (occ > 0 ? occ : 1)) / col_blocks;
    if (slabs > h->hb_slabs) slabs = h->hb_slabs;
    if ((size_t)slabs * 32 > MQ) slabs = cdiv(MQ, 32);
    if (slabs < 1) slabs = 1;
    int rpb = cdiv(MQ, slabs);
    slabs = cdiv(MQ, rpb);
    dim3 grid(col_blocks, slabs);
    float* g_vhb = G + h->o_vhb;
    if (A <= 8) {
      rtk::k_heads_bwd_fused<8><<<grid, dim3(32, 8), 0, st>>>(
          h->dtheta, actions, net + h->o_outw, net + h->o_vw, h->h1, h->v1, h->dh1, h->dv1, h->hb_part,
          h->hb_partb, MQ, F, A, Nq, duel, h->ldh, rpb);
      RT_LAUNCH_CHECK();
      rtk::k_heads_bwd_final<8><<<cdiv(C, 32) + 1, dim3(32, 8), 0, st>>>(
          h->hb_part, h->hb_partb, slabs, G + h->o_outw, G + h->o_outb, G + h->o_vw, G + h->o_vb,
          G + h->o_fcb, g_vhb, F, A, duel);
    } else {
      rtk::k_heads_bwd_fused<32><<<grid, dim3(32, 8), 0, st>>>(
          h->dtheta, actions, net + h->o_outw, net + h->o_vw, h->h1, h->v1, h->dh1, h->dv1, h->hb_part,
          h->hb_partb, MQ, F, A, Nq, duel, h->ldh, rpb);
      RT_LAUNCH_CHECK();
      rtk::k_heads_bwd_final<32><<<cdiv(C, 32) + 1, dim3(32, 8), 0, st>>>(
          h->hb_part, h->hb_partb, slabs, G + h->o_outw, G + h->o_outb, G + h->o_vw, G + h->o_vb,
          G + h->o_fcb, g_vhb, F, A, duel);
    }
    RT_LAUNCH_CHECK();
  }
  // data gradient first (it feeds BPTT); the weight gradients of the hidden and quantile layers are
  // leaves and start on the side branch only once the data-gradient GEMM is done -- two L2-bound
  // 21 GFLOP GEMMs side by side each took twice as long -- so they overlap the latency-bound BPTT
  if (h->fused_hidden) {
    // [dh1 | dv1] against the stacked [Wfc ; Wvh]: data gradient of both hidden layers in one GEMM
    // (bias gradients: k_heads_bwd_final)
    const int F2 = 2 * F;
    RT_TRY(gemm(h->gx, st, mk(h->dh1, F2, 0, net + h->o_fcw, D, 0, dxq, D, (int)MQ, D, F2)));
  } else {
    RT_TRY(gemm(h->gx, st, mk(h->dh1, F, 0, net + h->o_fcw, D, 0, dxq, D, (int)MQ, D, F)));
    if (h->dueling) {
      rtk::GemmArgs g = mk(h->dv1, F, 0, net + h->o_vhw, D, 0, dxq, D, (int)MQ, D, F);
      g.accumulate = 1;
      RT_TRY(gemm(h->gx, st, g));
    }
  }
  if (!h->dqn) {
    rtk::k_quantile_mul_bwd<<<cdiv((size_t)M * D, 256), 256, 0, st>>>(h->dxq, feat, h->phi, h->dphi,
                                                                     h->dfeatq, M, D, Nq);
    RT_LAUNCH_CHECK();
  }
  SideCtx sd;
  RT_TRY(side_begin(h, st, &sd));
  if (h->fused_hidden) {
    const int F2 = 2 * F;
    RT_TRY(gemm(*sd.gx, sd.st, mk(h->dh1, F2, 1, xq, D, 0, G + h->o_fcw, D, F2, D, (int)MQ)));
  } else {
    RT_TRY(gemm(*sd.gx, sd.st, mk(h->dh1, F, 1, xq, D, 0, G + h->o_fcw, D, F, D, (int)MQ)));
    if (h->dueling)
      RT_TRY(gemm(*sd.gx, sd.st, mk(h->dv1, F, 1, xq, D, 0, G + h->o_vhw, D, F, D, (int)MQ)));
  }
  if (h->dqn) return RT_OK;
  RT_TRY(gemm(*sd.gx, sd.st, mk(h->dphi, D, 1, h->cf, E, 0, G + h->o_qw, E, D, E, (int)MQ)));
  RT_TRY(colsum(h, sd.st, h->dphi, MQ, D, G + h->o_qb, 0, sd.colsum_scratch));
  return RT_OK;
}

// BPTT through the LSTM given dfeatq = d(loss)/d(h_all); produces dfeat (CNN features).
int lstm_backward(rt_learner* h, cudaStream_t st, const float* net, const float* feat, int rows,
                  int timesteps, const float* initials, const float* extra) {
  int U = h->U, Beff = rows / timesteps;
  float* G = h->grad;
  int nb = cdiv((size_t)Beff * U, 256);
  const int bptt_ctas = (U / 64) * (U / 32);
  const bool bptt_one_launch = h->bptt_persistent && h->gx.mode == 1 && Beff == 32 && (U == 512 || U == 256) &&
                               timesteps > 1 && bptt_ctas <= h->num_sms &&
                               (size_t)2 * (U / 64) * Beff * U <= h->gx.ws_floats;
  if (bptt_one_launch) {
    // rt_bptt.cuh: the whole recurrence in one cooperative launch (RT_BPTT_PERSISTENT=0: stepwise path)
    rtbptt::Args ba;
    ba.dout = h->dfeatq; ba.gates = h->gates; ba.c_all = h->c_all; ba.cprev = h->cprev;
    ba.initials = initials; ba.whh = net + h->o_whh; ba.dgates = h->dgates; ba.part = h->gx.ws;
    ba.counter = h->bptt_counters; ba.T = timesteps; ba.B = Beff; ba.U = U;
    const int KS = U / 64, NS = U / 32;
    RT_CUDA(cudaMemsetAsync(h->bptt_counters, 0, (size_t)(KS + NS) * rtbptt::CTR_STRIDE * sizeof(unsigned int), st));
    void* args[] = {(void*)&ba};
    const size_t smem = (size_t)32 * rtbptt::DG_PITCH * sizeof(float);
    void* kern = KS == 8 ? (void*)rtbptt::k_lstm_bptt_p<8> : (void*)rtbptt::k_lstm_bptt_p<4>;
    RT_CUDA(cudaLaunchCooperativeKernel(kern, dim3(bptt_ctas), dim3(rtbptt::THREADS), args, smem, st));
    rt::launch_counter()++;
  }
  RT_CUDA(cudaMemsetAsync(h->dc_carry, 0, (size_t)Beff * U * sizeof(float), st));
  // dh_carry = dgates[t+1] . W_hh is a skinny GEMM (M = B rows, K = 4U): it runs split-K over as
  // many CTAs as there are SMs and leaves the raw partials in the workspace; the cell kernel of
  // step t folds them (fixed order) while it forms dh, so no reduce launch sits between the steps
  const float* parts = nullptr;
  int nparts = 0;
  const size_t part_stride = (size_t)Beff * U;
  for (int t = bptt_one_launch ? -1 : timesteps - 1; t >= 0; --t) {
    size_t ro = (size_t)t * Beff;
    rtk::k_lstm_cell_bwd<<<nb, 256, 0, st>>>(
        h->dfeatq + ro * U, parts, nparts, part_stride, h->dc_carry,
        h->gates + ro * 4 * U, h->c_all + ro * U, h->cprev + ro * U, initials + ro,
        t == timesteps - 1 ? nullptr : initials + ro + Beff, h->dgates + ro * 4 * U, Beff, U);
    RT_LAUNCH_CHECK();
    if (t > 0) {
      h->gx.defer_reduce = 1;
      h->gx.split_min_kb = 8;
      int rc = gemm(h->gx, st, mk(h->dgates + ro * 4 * U, 4 * U, 0, net + h->o_whh, U, 0, h->dh_carry, U,
                                   Beff, U, 4 * U));
      h->gx.defer_reduce = 0;
      h->gx.split_min_kb = 16;
      RT_TRY(rc);
      nparts = h->gx.last_splits;
      parts = nparts > 1 ? h->gx.ws : h->dh_carry;
    }
  }
  // weight / bias gradients: side branch; the feature gradient continues the critical path
  SideCtx sd;
  RT_TRY(side_begin(h, st, &sd));
  RT_TRY(gemm(*sd.gx, sd.st, mk(h->dgates, 4 * U, 1, h->hprev, U, 0, G + h->o_whh, U, 4 * U, U, rows)));
  RT_TRY(gemm(*sd.gx, sd.st, mk(h->dgates, 4 * U, 1, feat, h->feat, 0, G + h->o_wih, h->feat, 4 * U, h->feat, rows)));
  if (h->X)
    RT_TRY(gemm_simt(*sd.gx, sd.st, mk(h->dgates, 4 * U, 1, extra, h->X, 0, G + h->o_wihx, h->X, 4 * U, h->X, rows)));
  RT_TRY(colsum(h, sd.st, h->dgates, rows, 4 * U, G + h->o_bih, 0, sd.colsum_scratch));
  RT_CUDA(cudaMemcpyAsync(G + h->o_bhh, G + h->o_bih, (size_t)4 * U * sizeof(float),
                          cudaMemcpyDeviceToDevice, sd.st));
  {
    // gradient w.r.t. the trunk features; those are post-ReLU outputs (last conv / FC layer), so the ReLU
    // derivative is applied in this GEMM's epilogue instead of a separate pass over dfeat
    rtk::GemmArgs g = mk(h->dgates, 4 * U, 0, net + h->o_wih, h->feat, 0, h->dfeat, h->feat, rows, h->feat, 4 * U);
    if (!h->conv.empty() || !h->pre.empty()) {
      g.mask = feat;
      g.ldmask = h->feat;
    }
    RT_TRY(gemm(h->gx, st, g));
  }
  return RT_OK;
}

// Conv stack backward; `dlast` = gradient w.r.t. the (post-ReLU) last conv output.
int cnn_backward(rt_learner* h, cudaStream_t st, const float* net, const uint8_t* x, int rows,
                 float* dlast, bool dlast_masked = false) {
  (void)x;   // the frames of this pass are already in h->xf (fp32 NHWC), see cnn_forward
  float* G = h->grad;
  int nl = (int)h->conv.size();
  if (!dlast_masked) {
    const ConvL& L = h->conv[nl - 1];
    size_t n = (size_t)rows * L.hout * L.wout * L.f;
    rtk::k_relu_bwd_inplace<<<grid1d(n), 256, 0, st>>>(dlast, h->c_out[nl - 1], n);
    RT_LAUNCH_CHECK();
  }
  {
    bool all = h->conv_implicit_bwd != 0;
    for (int i = 0; i < nl; ++i)
      all = all && conv_tc_eligible(h, i, i == 0 ? (const void*)h->xf : (const void*)h->c_out[i - 1]) &&
            h->conv[i].f % 4 == 0;
    if (all) {
      // whole batch at once: implicit-GEMM dW (no im2col), one dcol GEMM + col2im per layer
      for (int i = nl - 1; i >= 0; --i) {
        const ConvL& L = h->conv[i];
        size_t opix = (size_t)L.hout * L.wout;
        float* dy = i == nl - 1 ? dlast : h->d_c[i];
        const void* xin = i == 0 ? (const void*)h->xf : (const void*)h->c_out[i - 1];
        SideCtx sd;
        RT_TRY(side_begin(h, st, &sd));
        RT_TRY(conv_dw_tc(h, *sd.gx, sd.st, i, xin, dy, rows, G + L.w));
        RT_TRY(colsum(h, sd.st, dy, (size_t)rows * opix, L.f, G + L.b, 0, sd.colsum_scratch));
        if (i > 0 && conv_dx_tc_eligible(h, i)) {
          RT_TRY(conv_dx_tc(h, st, net, i, dy, rows));
        } else if (i > 0) {
          RT_TRY(gemm(h->gx, st, mk(dy, L.f, 0, net + L.w, L.K, 0, h->dcol_full, L.K, (int)(rows * opix), L.K, L.f)));
          size_t n_in = (size_t)rows * L.hin * L.win * L.cin;
          rtk::k_col2im_nhwc<<<grid1d(n_in / 4), 256, 0, st>>>(h->dcol_full, h->d_c[i - 1], rows, L.cin, L.hin,
                                                          L.win, L.k, L.s, L.hout, L.wout);
          RT_LAUNCH_CHECK();
          rtk::k_relu_bwd_inplace<<<grid1d(n_in), 256, 0, st>>>(h->d_c[i - 1], h->c_out[i - 1], n_in);
          RT_LAUNCH_CHECK();
        }
      }
      return RT_OK;
    }
  }
  for (int r0 = 0; r0 < rows; r0 += h->chunk_rows) {
    int rc = rows - r0 < h->chunk_rows ? rows - r0 : h->chunk_rows;
    int first = r0 == 0;
    for (int i = nl - 1; i >= 0; --i) {
      const ConvL& L = h->conv[i];
      size_t opix = (size_t)L.hout * L.wout;
      float* dy = (i == nl - 1 ? dlast : h->d_c[i]) + (size_t)r0 * opix * L.f;
      // recompute this layer's im2col input
      {
        const float* xin = (i == 0 ? h->xf : h->c_out[i - 1]) + (size_t)r0 * L.hin * L.win * L.cin;
        RT_TRY(launch_im2col_f32(st, xin, h->col, rc, L));
      }
      rtk::GemmArgs g = mk(dy, L.f, 1, h->col, L.K, 0, G + L.w, L.K, L.f, L.K, (int)(rc * opix));
      g.accumulate = first ? 0 : 1;
      RT_TRY(gemm(h->gx, st, g));
      RT_TRY(colsum(h, st, dy, (size_t)rc * opix, L.f, G + L.b, first ? 0 : 1));
      if (i > 0) {
        const ConvL& Lp = h->conv[i - 1];
        RT_TRY(gemm(h->gx, st, mk(dy, L.f, 0, net + L.w, L.K, 0, h->dcol, L.K, (int)(rc * opix), L.K, L.f)));
        float* dxp = h->d_c[i - 1] + (size_t)r0 * Lp.hout * Lp.wout * Lp.f;
        size_t n_in = (size_t)rc * L.hin * L.win * L.cin;
        if (L.cin % 4 == 0)
          rtk::k_col2im_nhwc<<<grid1d(n_in / 4), 256, 0, st>>>(h->dcol, dxp, rc, L.cin, L.hin, L.win, L.k,
                                                              L.s, L.hout, L.wout);
        else
          rtk::k_col2im_nhwc_s<<<grid1d(n_in), 256, 0, st>>>(h->dcol, dxp, rc, L.cin, L.hin, L.win, L.k,
                                                            L.s, L.hout, L.wout);
        RT_LAUNCH_CHECK();
        rtk::k_relu_bwd_inplace<<<grid1d(n_in), 256, 0, st>>>(
            dxp, h->c_out[i - 1] + (size_t)r0 * Lp.hout * Lp.wout * Lp.f, n_in);
        RT_LAUNCH_CHECK();
      }
    }
  }
  return RT_OK;
}

// Backward of the feature extractor: `dlast` = gradient w.r.t. the (post-ReLU) trunk features of the
// training pass, `x` = that pass's observation rows.  FC layers first (weight / bias gradients on the side
// branch, the data gradient with the previous layer's ReLU mask fused in the GEMM epilogue), then the CNN.
int features_backward(rt_learner* h, cudaStream_t st, const float* net, const uint8_t* x, int rows, float* dlast,
                      bool dlast_masked) {
  float* G = h->grad;
  const int np = (int)h->pre.size();
  if (np) {
    if (!dlast_masked) {
      size_t n = (size_t)rows * h->pre[np - 1].out;
      rtk::k_relu_bwd_inplace<<<grid1d(n), 256, 0, st>>>(dlast, h->pre_out[np - 1], n);
      RT_LAUNCH_CHECK();
    }
    for (int k = np - 1; k >= 0; --k) {
      const PreL& L = h->pre[k];
      const float* in = k > 0 ? h->pre_out[k - 1]
                              : (h->conv.empty() ? reinterpret_cast<const float*>(x) : h->c_out.back());
      float* dz = k == np - 1 ? dlast : h->d_pre[k];
      SideCtx sd;
      RT_TRY(side_begin(h, st, &sd));
      RT_TRY(gemm(*sd.gx, sd.st, mk(dz, L.out, 1, in, L.in, 0, G + L.w, L.in, L.out, L.in, rows)));
      RT_TRY(colsum(h, sd.st, dz, rows, L.out, G + L.b, 0, sd.colsum_scratch));
      if (k > 0) {
        rtk::GemmArgs g = mk(dz, L.out, 0, net + L.w, L.in, 0, h->d_pre[k - 1], L.in, rows, L.in, L.out);
        g.mask = h->pre_out[k - 1];
        g.ldmask = L.in;
        RT_TRY(gemm(h->gx, st, g));
      } else if (!h->conv.empty()) {
        // the ReLU mask of the last conv layer is applied by cnn_backward
        RT_TRY(gemm(h->gx, st, mk(dz, L.out, 0, net + L.w, L.in, 0, h->d_cfeat, L.in, rows, L.in, L.out)));
      }
    }
    dlast = h->d_cfeat;
    dlast_masked = false;
  }
  if (!h->conv.empty()) RT_TRY(cnn_backward(h, st, net, x, rows, dlast, dlast_masked));
  return RT_OK;
}

// grad-norm, clip, Adam (torch_trainer.py:177-199) on the flat buffers.  grad_scale folds the
// 1/world_size of a data-parallel gradient mean into the same pass.
int apply_grads(rt_learner* h, cudaStream_t st, float grad_scale) {

    int parts = 512;
    rtk::k_sumsq_partial<<<parts, 256, 0, st>>>(h->grad, h->sumsq_part, h->nparams);
    RT_LAUNCH_CHECK();
    rtk::k_gradnorm_final<<<1, 32, 0, st>>>(h->sumsq_part, parts, h->stats,
                                           h->td.clip_grad > 0 ? (float)h->td.clip_grad : 0.f, grad_scale,
                                           (float)h->td.clip_grad_dynamic_alpha);
    RT_LAUNCH_CHECK();
    h->adam_t++;
    double b1 = 0.9, b2 = 0.999;
    float bc1 = (float)(1.0 - std::pow(b1, (double)h->adam_t));
    float bc2s = (float)std::sqrt(1.0 - std::pow(b2, (double)h->adam_t));
    rtk::k_adam<<<grid1d(h->nparams / 4), 256, 0, st>>>(h->p[0], h->grad, h->adam_m, h->adam_v, h->nparams,
                                                       h->stats, h->lr, (float)b1, (float)b2,
                                                       (float)h->td.adam_epsilon, bc1, bc2s, grad_scale,
                                                       h->rn ? h->pr[0] : nullptr, h->wflag);
    RT_LAUNCH_CHECK();
    return RT_OK;
}

__global__ void k_uniform(float* out, size_t n, unsigned long long seed, unsigned long long ctr) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (ctr + i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  out[i] = (float)(z >> 40) * (1.0f / 16777216.0f);  // 24-bit mantissa, [0, 1)
}

// masked write-back of the burned-in LSTM state into the batch (multi_step_trainer.py:117-126
// + LSTM.get_state, lstm.py:150-152)
__global__ void k_store_state(const float* __restrict__ h_last, const float* __restrict__ c_last,
                              const float* __restrict__ initials, float* __restrict__ hx,
                              float* __restrict__ cx, int B, int U) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * U) return;
  float keep = 1.f - initials[i / U];
  hx[i] = h_last[i] * keep;
  cx[i] = c_last[i] * keep;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int rt_learner_create(const rt_model_desc* md, const rt_train_desc* td, int32_t device,
                      rt_learner** out) {
  RT_REQUIRE(md && td && out, "null argument");
  RT_REQUIRE(md->num_conv >= 0 && md->num_conv <= RT_MAX_CONV, "num_conv out of range");
  RT_REQUIRE(md->num_pre_fc >= 0 && md->num_pre_fc <= RT_MAX_PRE_FC, "num_pre_fc out of range");
  RT_REQUIRE(md->extra_dim >= 0 && (md->extra_dim == 0 || md->lstm_units > 0),
             "extra features are fed to the LSTM layer: extra_dim needs lstm_units > 0");
  RT_REQUIRE(md->num_quantiles >= 0 && md->num_quantiles <= 256, "num_quantiles out of range");
  RT_REQUIRE(md->num_actions >= 1 && md->num_actions <= 64, "num_actions out of range");
  RT_REQUIRE(td->mbatch >= 1 && td->nstep_train >= 1 && td->burn_in >= 0 && td->nstep_target >= 1,
             "bad batch geometry");
  RT_CUDA(cudaSetDevice(device));
  rt_learner* h = new rt_learner();
  h->md = *md;
  h->td = *td;
  h->device = device;
  h->U = md->lstm_units; h->F = md->fc_size; h->A = md->num_actions;
  h->dqn = md->num_quantiles == 0;
  h->Nq = h->dqn ? 1 : md->num_quantiles;
  h->E = md->embedding_dim; h->dueling = md->dueling != 0;
  h->B = td->mbatch; h->T = td->nstep_train; h->P = td->burn_in; h->n = td->nstep_target;
  h->S = h->T + h->P;
  h->R = td->rnn_steps_train > 0 ? td->rnn_steps_train : h->T;
  RT_REQUIRE(h->T % h->R == 0, "nstep_train (%d) must be divisible by rnn_steps_train (%d)", h->T, h->R);
  h->lr = (float)td->lr;
  {
    // mean / sum over the batch, optionally a different aggregation over time-steps first.  The
    // reference views the losses as (timesteps, -1) with timesteps = rnn_steps_train (dqn.py:120-130)
    const double Md = (double)td->nstep_train * td->mbatch;
    const double Rd = (double)h->R, Bd = Md / Rd;
    if (td->loss_timestep_agg == 0) h->loss_scale = (float)(td->loss_sum ? 1.0 : 1.0 / Md);
    else h->loss_scale = (float)((td->loss_timestep_agg == 1 ? 1.0 / Rd : 1.0) *
                                 (td->loss_sum ? 1.0 : 1.0 / Bd));
  }
  RT_REQUIRE(!(h->P > 0 && h->U == 0), "burn-in only makes sense for recurrent models");

  int c = md->in_c, hh = md->in_h, ww = md->in_w;
  for (int i = 0; i < md->num_conv; ++i) {
    ConvL L;
    L.cin = c; L.hin = hh; L.win = ww; L.f = md->conv_filters[i]; L.k = md->conv_kernel[i];
    L.s = md->conv_stride[i];
    RT_REQUIRE(L.f > 0 && L.k > 0 && L.s > 0 && L.k <= hh && L.k <= ww, "bad conv layer %d", i);
    L.hout = (hh - L.k) / L.s + 1;
    L.wout = (ww - L.k) / L.s + 1;
    L.K = c * L.k * L.k;
    char nm[96];
    snprintf(nm, sizeof(nm), "model.layers.0.layers.%d.weight", i);
    L.w = add_param(h, nm, {L.f, c, L.k, L.k}, PERM_CONV, c, L.k);
    h->pinfo.back().gemm_w = true;
    snprintf(nm, sizeof(nm), "model.layers.0.layers.%d.bias", i);
    L.b = add_param(h, nm, {L.f}, PERM_NONE);
    h->conv.push_back(L);
    c = L.f; hh = L.hout; ww = L.wout;
  }
  if (md->num_conv == 0) {   // raw float observation: no layout permutation (featHW = 1)
    h->featC = c * hh * ww; h->featHW = 1;
  } else {
    h->featC = c; h->featHW = hh * ww;
  }
  h->cfeat = c * hh * ww;
  h->X = md->extra_dim;
  // FC layers in front of the LSTM / last FC module (fc.py:18-24), reference registration order
  int next_module = md->num_conv ? 1 : 0;
  {
    int d = h->cfeat;
    for (int k = 0; k < md->num_pre_fc; ++k) {
      RT_REQUIRE(md->pre_fc_size[k] > 0 && md->pre_fc_module[k] >= next_module - (k ? 1 : 0), "bad FC layer %d", k);
      PreL L;
      L.in = d; L.out = md->pre_fc_size[k];
      char nm[96];
      snprintf(nm, sizeof(nm), "model.layers.%d.layers.%d.0.weight", md->pre_fc_module[k], md->pre_fc_sub[k]);
      L.w = add_param(h, nm, {L.out, L.in}, k == 0 ? PERM_FEAT_COLS : PERM_NONE);
      h->pinfo.back().gemm_w = true;
      snprintf(nm, sizeof(nm), "model.layers.%d.layers.%d.0.bias", md->pre_fc_module[k], md->pre_fc_sub[k]);
      L.b = add_param(h, nm, {L.out}, PERM_NONE);
      h->pre.push_back(L);
      d = L.out;
      next_module = md->pre_fc_module[k] + 1;
    }
    h->feat = d;
  }
  // the gradients of everything registered so far (CNN + these FC layers) are final only at the very end of
  // the backward pass: the data-parallel "late" bucket starts here
  h->conv_param_end = (h->nparams + 63) / 64 * 64;
  const bool trunk_perm = h->pre.empty();   // the consumer of the trunk features sees the conv layout
  int fc_layer = next_module + (h->U ? 1 : 0);
  if (h->U) {
    int U = h->U;
    char nm[96];
    snprintf(nm, sizeof(nm), "model.layers.%d.lstm_cell.weight_ih", next_module);
    h->wih_perm = trunk_perm;
    if (h->X) {
      h->o_wih = add_param(h, nm, {4 * U, h->feat + h->X}, PERM_WIH_EXTRA);
      h->o_wihx = h->o_wih + (size_t)4 * U * h->feat;
    } else {
      h->o_wih = add_param(h, nm, {4 * U, h->feat}, trunk_perm ? PERM_FEAT_COLS : PERM_NONE);
    }
    h->pinfo.back().gemm_w = true;
    snprintf(nm, sizeof(nm), "model.layers.%d.lstm_cell.weight_hh", next_module);
    h->o_whh = add_param(h, nm, {4 * U, U}, PERM_NONE);
    h->pinfo.back().gemm_w = true;
    snprintf(nm, sizeof(nm), "model.layers.%d.lstm_cell.bias_ih", next_module);
    h->o_bih = add_param(h, nm, {4 * U}, PERM_NONE);
    snprintf(nm, sizeof(nm), "model.layers.%d.lstm_cell.bias_hh", next_module);
    h->o_bhh = add_param(h, nm, {4 * U}, PERM_NONE);
    h->D = U;
  } else {
    h->D = h->feat;
  }
  int featperm_cols = (h->U || !trunk_perm) ? PERM_NONE : PERM_FEAT_COLS;
  int featperm_rows = (h->U || !trunk_perm) ? PERM_NONE : PERM_FEAT_ROWS;
  // dueling: the advantage hidden layer (model FC) and the value hidden layer read the same
  // input, so their weights / biases are placed back to back and run as ONE [2F x D] layer
  h->fused_hidden = h->dueling && ((size_t)h->F * h->D) % 4 == 0 && h->F % 4 == 0;
  long long fw = -1, fb = -1;
  if (h->fused_hidden) {
    fw = (long long)reserve_params(h, (size_t)2 * h->F * h->D);
    fb = (long long)reserve_params(h, (size_t)2 * h->F);
  }
  {
    char nm[96];
    snprintf(nm, sizeof(nm), "model.layers.%d.layers.0.0.weight", fc_layer);
    h->o_fcw = add_param(h, nm, {h->F, h->D}, featperm_cols, 0, 0, fw);
    h->pinfo.back().gemm_w = true;
    snprintf(nm, sizeof(nm), "model.layers.%d.layers.0.0.bias", fc_layer);
    h->o_fcb = add_param(h, nm, {h->F}, PERM_NONE, 0, 0, fb);
  }
  h->o_outw = add_param(h, "out_layer.weight", {h->A, h->F}, PERM_NONE);
  h->o_outb = add_param(h, "out_layer.bias", {h->A}, PERM_NONE);
  if (h->dueling) {
    h->o_vhw = add_param(h, "value_hidden_layer.weight", {h->F, h->D}, featperm_cols, 0, 0,
                         h->fused_hidden ? fw + (long long)h->F * h->D : -1);
    h->pinfo.back().gemm_w = true;
    h->o_vhb = add_param(h, "value_hidden_layer.bias", {h->F}, PERM_NONE, 0, 0,
                         h->fused_hidden ? fb + h->F : -1);
    h->o_vw = add_param(h, "value_layer.weight", {1, h->F}, PERM_NONE);
    h->o_vb = add_param(h, "value_layer.bias", {1}, PERM_NONE);
  }
  if (!h->dqn) {
    h->o_qw = add_param(h, "quantile_layer.weight", {h->D, h->E}, featperm_rows);
    h->pinfo.back().gemm_w = true;
    h->o_qb = add_param(h, "quantile_layer.bias", {h->D}, featperm_rows);
  }
  h->nparams = (h->nparams + 63) / 64 * 64;

  RT_TRY(dalloc(h, &h->p[0], h->nparams, "params_online"));
  RT_TRY(dalloc(h, &h->p[1], h->nparams, "params_target"));
  RT_REQUIRE(td->gemm_mode >= RT_GEMM_FP32_SIMT && td->gemm_mode <= RT_GEMM_TF32_RN, "bad gemm_mode %d", td->gemm_mode);
  h->rn = td->gemm_mode == RT_GEMM_TF32_RN;
  h->pr[0] = h->p[0];
  h->pr[1] = h->p[1];
  if (h->rn) {
    RT_TRY(dalloc(h, &h->pr[0], h->nparams, "params_online_tf32"));
    RT_TRY(dalloc(h, &h->pr[1], h->nparams, "params_target_tf32"));
    std::vector<uint8_t> flags(h->nparams / 64, 0);
    for (const PInfo& pi : h->pinfo)
      if (pi.gemm_w)
        for (size_t b = pi.off / 64; b <= (pi.off + pi.count - 1) / 64; ++b) flags[b] = 1;
    RT_TRY(dalloc(h, &h->wflag, flags.size()));
    RT_CUDA(cudaMemcpy(h->wflag, flags.data(), flags.size(), cudaMemcpyHostToDevice));
  }
  RT_TRY(dalloc(h, &h->grad, h->nparams, "grad"));
  RT_TRY(dalloc(h, &h->adam_m, h->nparams, "adam_m"));
  RT_TRY(dalloc(h, &h->adam_v, h->nparams, "adam_v"));

  h->M = h->T * h->B;
  h->MQ = h->M * h->Nq;
  int burn_rows = h->P * h->B;
  int shared_rows = h->M + h->n * h->B;   // online CNN shared by the selection and training passes
  h->max_rows = shared_rows > burn_rows ? shared_rows : burn_rows;
  size_t rows = (size_t)h->max_rows;
  size_t maxcol = 0;
  for (size_t i = 0; i < h->conv.size(); ++i) {
    const ConvL& L = h->conv[i];
    size_t opix = (size_t)L.hout * L.wout;
    float *co = nullptr, *dco = nullptr;
    char nm[32];
    snprintf(nm, sizeof(nm), "c%d", (int)i);
    RT_TRY(dalloc(h, &co, rows * opix * L.f, nm));
    snprintf(nm, sizeof(nm), "dc%d", (int)i);
    RT_TRY(dalloc(h, &dco, (size_t)h->M * opix * L.f, nm));
    h->c_out.push_back(co);
    h->d_c.push_back(dco);
    float* co2 = nullptr;
    RT_TRY(dalloc(h, &co2, (size_t)h->M * opix * L.f));   // target pass: M rows
    h->c_out2.push_back(co2);
    float* wt = nullptr;
    if (i > 0) RT_TRY(dalloc(h, &wt, (size_t)L.f * L.K));
    h->conv_wt.push_back(wt);
    size_t cc = (size_t)h->chunk_rows * opix * L.K;
    if (cc > maxcol) maxcol = cc;
  }
  h->xf_floats = h->conv.empty() ? 1 : rows * (size_t)md->in_c * md->in_h * md->in_w;
  RT_TRY(dalloc(h, &h->xf, h->xf_floats, "xf"));
  h->xf0 = h->xf;
  RT_CUDA(cudaEventCreateWithFlags(&h->ev_prefetch, cudaEventDisableTiming));
  for (size_t k = 0; k < h->pre.size(); ++k) {
    float *a = nullptr, *b2 = nullptr, *d = nullptr;
    char nm[32];
    snprintf(nm, sizeof(nm), "pre%d", (int)k);
    RT_TRY(dalloc(h, &a, rows * h->pre[k].out, nm));
    RT_TRY(dalloc(h, &b2, (size_t)h->M * h->pre[k].out));
    RT_TRY(dalloc(h, &d, (size_t)h->M * h->pre[k].out));
    h->pre_out.push_back(a);
    h->pre_out2.push_back(b2);
    h->d_pre.push_back(d);
  }
  if (!h->pre.empty() && !h->conv.empty()) RT_TRY(dalloc(h, &h->d_cfeat, (size_t)h->M * h->cfeat, "d_cfeat"));
  if (!h->conv.empty()) RT_TRY(dalloc(h, &h->wcat, (size_t)2 * h->conv[0].f * h->conv[0].K));
  RT_TRY(dalloc(h, &h->col, maxcol));
  RT_TRY(dalloc(h, &h->dcol, maxcol));
  {
    size_t mx = 0;
    for (size_t i = 1; i < h->conv.size(); ++i) {
      size_t v = (size_t)h->M * h->conv[i].hout * h->conv[i].wout * h->conv[i].K;
      if (v > mx) mx = v;
    }
    RT_TRY(dalloc(h, &h->dcol_full, mx));
  }
  if (h->U) {
    size_t U = h->U;
    RT_TRY(dalloc(h, &h->xg, rows * 4 * U, "xg"));
    RT_TRY(dalloc(h, &h->hg, rows * 4 * U, "hg"));
    RT_TRY(dalloc(h, &h->hprev, 3 * rows * U, "hprev"));   // exchange slots: train / target / selection
    RT_TRY(dalloc(h, &h->xg2, rows * 4 * U, "xg2"));
    RT_TRY(dalloc(h, &h->h_all2, rows * U, "h_all2"));
    RT_TRY(dalloc(h, &h->h_all3, rows * U, "h_all3"));
    RT_TRY(dalloc(h, &h->cprev, rows * U, "cprev"));
    RT_TRY(dalloc(h, &h->gates, rows * 4 * U, "gates"));
    RT_TRY(dalloc(h, &h->c_all, rows * U, "c_all"));
    RT_TRY(dalloc(h, &h->h_all, rows * U, "h_all"));
    RT_TRY(dalloc(h, &h->dgates, (size_t)h->M * 4 * U, "dgates"));
    RT_TRY(dalloc(h, &h->dh_carry, (size_t)h->M * U));
    RT_TRY(dalloc(h, &h->dc_carry, (size_t)h->M * U));
    RT_TRY(dalloc(h, &h->dfeat, (size_t)h->M * h->feat, "dfeat"));
  }
  size_t MQ = h->MQ, D = h->D, F = h->F, A = h->A;
  RT_TRY(dalloc(h, &h->tau, MQ, "tau"));
  RT_TRY(dalloc(h, &h->tau_stage, MQ * 3));
  RT_TRY(dalloc(h, &h->cf, MQ * h->E, "cf"));
  RT_TRY(dalloc(h, &h->phi, MQ * D, "phi"));
  RT_TRY(dalloc(h, &h->xq, MQ * D, "xq"));
  h->ldh = h->fused_hidden ? 2 * h->F : h->F;
  if (h->fused_hidden) {
    RT_TRY(dalloc(h, &h->h1, MQ * 2 * F, "h1"));
    h->v1 = h->h1 + F;
  } else {
    RT_TRY(dalloc(h, &h->h1, MQ * F, "h1"));
    RT_TRY(dalloc(h, &h->v1, MQ * F, "v1"));
  }
  RT_TRY(dalloc(h, &h->adv, MQ * A, "adv"));
  RT_TRY(dalloc(h, &h->v, MQ, "v"));
  // second heads set (target pass on the side branch of the forward graph)
  RT_TRY(dalloc(h, &h->cf2, MQ * h->E));
  RT_TRY(dalloc(h, &h->phi2, MQ * D));
  RT_TRY(dalloc(h, &h->xq2, MQ * D));
  if (h->fused_hidden) {
    RT_TRY(dalloc(h, &h->h1b, MQ * 2 * F));
    h->v1b = h->h1b + F;
  } else {
    RT_TRY(dalloc(h, &h->h1b, MQ * F));
    RT_TRY(dalloc(h, &h->v1b, MQ * F));
  }
  RT_TRY(dalloc(h, &h->adv2, MQ * A));
  RT_TRY(dalloc(h, &h->vb2, MQ));
  // third heads set (selection pass)
  RT_TRY(dalloc(h, &h->cf3, MQ * h->E));
  RT_TRY(dalloc(h, &h->phi3, MQ * D));
  RT_TRY(dalloc(h, &h->xq3, MQ * D));
  if (h->fused_hidden) {
    RT_TRY(dalloc(h, &h->h1c, MQ * 2 * F));
    h->v1c = h->h1c + F;
  } else {
    RT_TRY(dalloc(h, &h->h1c, MQ * F));
    RT_TRY(dalloc(h, &h->v1c, MQ * F));
  }
  RT_TRY(dalloc(h, &h->adv3, MQ * A));
  RT_TRY(dalloc(h, &h->vb3, MQ));
  RT_TRY(dalloc(h, &h->q, MQ * A, "q"));
  RT_TRY(dalloc(h, &h->tq, MQ * A, "tq"));
  RT_TRY(dalloc(h, &h->sq, MQ * A, "sq"));
  RT_TRY(dalloc(h, &h->targets, MQ, "targets"));
  RT_TRY(dalloc(h, &h->dtheta, MQ, "dtheta"));
  RT_TRY(dalloc(h, &h->row_loss, (size_t)h->M, "row_loss"));
  RT_TRY(dalloc(h, &h->report, (size_t)h->M, "report"));
  RT_TRY(dalloc(h, &h->row_q, (size_t)h->M, "row_q"));
  RT_TRY(dalloc(h, &h->stats, 8, "stats"));
  RT_CUDA(cudaEventCreateWithFlags(&h->ev_loss, cudaEventDisableTiming));
  RT_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  RT_CUDA(cudaEventCreateWithFlags(&h->ev_late, cudaEventDisableTiming));
  RT_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  {
    // the critical chain outranks the weight-gradient branch when both have CTAs ready
    int prio_lo = 0, prio_hi = 0;
    RT_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    RT_CUDA(cudaStreamCreateWithPriority(&h->own, cudaStreamNonBlocking, prio_hi));
    RT_CUDA(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_lo));
    RT_CUDA(cudaStreamCreateWithPriority(&h->side_b, cudaStreamNonBlocking, prio_lo));
    for (auto& e : h->ev_side_b) RT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    RT_CUDA(cudaEventCreateWithFlags(&h->ev_phi, cudaEventDisableTiming));
    if (const char* e = getenv("RT_PHI_EARLY")) h->phi_early = atoi(e);
  }
  if (const char* e = getenv("RT_GRAPHS")) h->graphs_enabled = atoi(e);
  if (const char* e = getenv("RT_ACT_GRAPH")) h->act_graphs_enabled = atoi(e);
  if (const char* e = getenv("RT_OVERLAP_BWD")) h->overlap_bwd = atoi(e);
  if (const char* e = getenv("RT_DP_SPLIT")) h->dp_split = atoi(e);
  if (const char* e = getenv("RT_OVERLAP_FWD")) h->overlap_fwd = atoi(e);
  if (const char* e = getenv("RT_CONV_SHALLOW")) h->conv_shallow = atoi(e);
  if (const char* e = getenv("RT_BPTT_PERSISTENT")) h->bptt_persistent = atoi(e);
  for (auto& e : h->ev_side) RT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  RT_CUDA(cudaMallocHost(&h->h_stats, 8 * sizeof(float)));
  if (h->fused_hidden) {
    RT_TRY(dalloc(h, &h->dh1, MQ * 2 * F, "dh1"));
    h->dv1 = h->dh1 + F;
  } else {
    RT_TRY(dalloc(h, &h->dh1, MQ * F, "dh1"));
    RT_TRY(dalloc(h, &h->dv1, MQ * F, "dv1"));
  }
  RT_TRY(dalloc(h, &h->dxq, MQ * D, "dxq"));
  RT_TRY(dalloc(h, &h->dphi, MQ * D, "dphi"));
  RT_TRY(dalloc(h, &h->dfeatq, (size_t)h->M * D, "dfeatq"));
  h->gx.ws_floats = (size_t)64 << 20;  // 256 MiB split-K workspace
  RT_TRY(dalloc(h, &h->gx.ws, h->gx.ws_floats));
  h->gx2.ws_floats = (size_t)16 << 20;  // 64 MiB: split-K partials of the weight-gradient branch
  RT_TRY(dalloc(h, &h->gx2.ws, h->gx2.ws_floats));
  h->gx3.ws_floats = (size_t)1 << 20;   // heads forward GEMMs never split
  RT_TRY(dalloc(h, &h->gx3.ws, h->gx3.ws_floats));
  h->gx.mode = td->gemm_mode == RT_GEMM_FP32_SIMT ? 0 : 1;
  if (const char* e = getenv("RT_TC_BN")) h->gx.force_bn = atoi(e);
  if (const char* e = getenv("RT_TC_STAGES")) h->gx.force_stages = atoi(e);
  if (const char* e = getenv("RT_TC_PERSISTENT")) h->gx.persistent = atoi(e);
  if (const char* e = getenv("RT_TC_PAIR")) h->gx.pair = atoi(e);
  size_t maxN = 4 * (size_t)(h->U ? h->U : 1);
  if (D > maxN) maxN = D;
  if (F > maxN) maxN = F;
  if (A > maxN) maxN = A;
  for (auto& L : h->conv)
    if ((size_t)L.f > maxN) maxN = L.f;
  for (auto& L : h->pre)
    if ((size_t)L.out > maxN) maxN = L.out;
  RT_TRY(dalloc(h, &h->colsum_part, 2048 * maxN));
  RT_TRY(dalloc(h, &h->colsum_part2, 2048 * maxN));
  RT_TRY(dalloc(h, &h->sumsq_part, 1024));
  RT_REQUIRE(h->A <= 32, "num_actions > 32 not supported by the fused head kernels");
  RT_REQUIRE(h->F % 4 == 0 && h->D % 4 == 0, "fc_size and the quantile-layer width must be multiples of 4");

  RT_TRY(dalloc(h, &h->grid_barrier, 64));
  RT_TRY(dalloc(h, &h->bptt_counters, (size_t)(8 + 16) * rtbptt::CTR_STRIDE));
  if (h->U) RT_TRY(dalloc(h, &h->lstm_hrep, (size_t)2 * rtk::LSTM_REP * 32 * h->U));
  if (h->U && h->U % 32 == 0) {
    h->lstm_tcap = h->T > h->P ? h->T : h->P;
    RT_TRY(dalloc(h, &h->lstm_xchg, (size_t)2 * h->lstm_tcap * (h->U / 32) * 128 * 32));
  }
  if (getenv("RT_DEBUG_TIMELINE")) RT_TRY(dalloc(h, &h->lstm_dbg, 8 * 256, "lstm_dbg"));
  {
    cudaDeviceProp prop;
    RT_CUDA(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    h->gx.num_sms = prop.multiProcessorCount;
    // the weight-gradient branch runs the same GEMM selection with its own workspace
    h->gx2.mode = h->gx.mode; h->gx2.num_sms = h->gx.num_sms; h->gx2.persistent = h->gx.persistent;
    h->gx2.force_bn = h->gx.force_bn; h->gx2.force_stages = h->gx.force_stages;
    h->gx2.round_tf32 = h->gx.round_tf32;
    h->gx3.mode = h->gx.mode; h->gx3.num_sms = h->gx.num_sms; h->gx3.persistent = h->gx.persistent;
    h->gx3.force_bn = h->gx.force_bn; h->gx3.force_stages = h->gx.force_stages;
    h->gx3.round_tf32 = h->gx.round_tf32;
    const char* e = getenv("RT_LSTM_STEPWISE");
    if (e && e[0] == '1') h->lstm_persistent = 0;
    e = getenv("RT_LSTM_TC");
    if (e && e[0] == '0') h->lstm_tc = 0;
    e = getenv("RT_LSTM_MMA");
    if (e) h->lstm_mma = atoi(e);
    e = getenv("RT_LSTM_UPC");
    if (e && atoi(e) == 16) h->lstm_upc = 16;
    e = getenv("RT_CONV_IM2COL");
    if (e && e[0] == '1') h->conv_implicit = 0;
    e = getenv("RT_CONV_PERSISTENT");
    if (e && e[0] == '0') h->conv_persistent = 0;
    e = getenv("RT_CONV_DX_COL2IM");
    if (e && e[0] == '1') h->conv_dx_implicit = 0;
    e = getenv("RT_CONV1_PAIR");
    if (e && e[0] == '0') h->conv1_pair = 0;
    e = getenv("RT_CONV_BWD_IM2COL");
    if (e && e[0] == '1') h->conv_implicit_bwd = 0;
  }
  RT_TRY(dalloc(h, &h->hb_part, (size_t)h->hb_slabs * (A + 1) * 2 * F));
  RT_TRY(dalloc(h, &h->hb_partb, (size_t)h->hb_slabs * 33));
  *out = h;
  return RT_OK;
}

void rt_learner_destroy(rt_learner* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  if (h->ev_loss) cudaEventDestroy(h->ev_loss);
  for (auto& g : h->graphs) {
    if (g.fwd) cudaGraphExecDestroy(g.fwd);
    if (g.bwd) cudaGraphExecDestroy(g.bwd);
    if (g.bwd2) cudaGraphExecDestroy(g.bwd2);
  }
  for (auto& g : h->act_graphs)
    if (g.g) cudaGraphExecDestroy(g.g);
  if (h->ev_late) cudaEventDestroy(h->ev_late);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->own) cudaStreamDestroy(h->own);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->side_b) cudaStreamDestroy(h->side_b);
  for (auto& e : h->ev_side_b) if (e) cudaEventDestroy(e);
  if (h->ev_phi) cudaEventDestroy(h->ev_phi);
  for (auto& e : h->ev_side) if (e) cudaEventDestroy(e);
  if (h->h_stats) cudaFreeHost(h->h_stats);
  if (h->ev_prefetch) cudaEventDestroy(h->ev_prefetch);
  if (h->comm) rt_comm_destroy(h);
  delete h;
}

int32_t rt_learner_num_params(const rt_learner* h) { return h ? (int32_t)h->pinfo.size() : 0; }
int64_t rt_learner_num_weights(const rt_learner* h) {
  int64_t n = 0;
  if (h)
    for (auto& pi : h->pinfo) n += (int64_t)pi.count;
  return n;
}

int rt_learner_param_info(const rt_learner* h, int32_t i, char* name, int32_t name_cap,
                          int64_t* shape4, int32_t* ndim) {
  RT_REQUIRE(h && i >= 0 && i < (int)h->pinfo.size() && name && shape4 && ndim, "bad argument");
  const PInfo& pi = h->pinfo[i];
  snprintf(name, name_cap, "%s", pi.name.c_str());
  *ndim = (int)pi.shape.size();
  for (int d = 0; d < 4; ++d) shape4[d] = d < (int)pi.shape.size() ? pi.shape[d] : 1;
  return RT_OK;
}

static float* which_buffer(rt_learner* h, int which) {
  switch (which) {
    case RT_BUF_ONLINE: return h->p[0];
    case RT_BUF_TARGET: return h->p[1];
    case RT_BUF_GRAD: return h->grad;
    case RT_BUF_ADAM_M: return h->adam_m;
    case RT_BUF_ADAM_V: return h->adam_v;
  }
  return nullptr;
}

int rt_learner_load_params(rt_learner* h, int32_t which, const float* const* tensors) {
  RT_REQUIRE(h && tensors, "null argument");
  float* dst = which_buffer(h, which);
  RT_REQUIRE(dst, "bad buffer selector");
  RT_CUDA(cudaSetDevice(h->device));
  std::vector<float> flat(h->nparams, 0.f);
  for (size_t i = 0; i < h->pinfo.size(); ++i) {
    const PInfo& pi = h->pinfo[i];
    for (size_t j = 0; j < pi.count; ++j) flat[pi.off + perm_index(h, pi, j)] = tensors[i][j];
  }
  RT_CUDA(cudaMemcpy(dst, flat.data(), h->nparams * sizeof(float), cudaMemcpyHostToDevice));
  if (h->rn && (which == RT_BUF_ONLINE || which == RT_BUF_TARGET)) {
    rtk::k_shadow_params<<<grid1d(h->nparams / 4), 256, 0, 0>>>(h->p[which], h->pr[which], h->wflag, h->nparams);
    RT_LAUNCH_CHECK();
    RT_CUDA(cudaDeviceSynchronize());
  }
  return RT_OK;
}

int rt_learner_get_params(rt_learner* h, int32_t which, float* const* tensors) {
  RT_REQUIRE(h && tensors, "null argument");
  float* src = which_buffer(h, which);
  RT_REQUIRE(src, "bad buffer selector");
  RT_CUDA(cudaSetDevice(h->device));
  RT_CUDA(cudaDeviceSynchronize());
  std::vector<float> flat(h->nparams);
  RT_CUDA(cudaMemcpy(flat.data(), src, h->nparams * sizeof(float), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < h->pinfo.size(); ++i) {
    const PInfo& pi = h->pinfo[i];
    for (size_t j = 0; j < pi.count; ++j) tensors[i][j] = flat[pi.off + perm_index(h, pi, j)];
  }
  return RT_OK;
}

int rt_learner_sync_target(rt_learner* h, void* stream) {
  RT_REQUIRE(h, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  // TorchPolicy.copy_from with factor 1.0 (torch_policy.py:61-68): one flat copy
  RT_CUDA(cudaMemcpyAsync(h->p[1], h->p[0], h->nparams * sizeof(float), cudaMemcpyDeviceToDevice,
                          (cudaStream_t)stream));
  if (h->rn)
    RT_CUDA(cudaMemcpyAsync(h->pr[1], h->pr[0], h->nparams * sizeof(float), cudaMemcpyDeviceToDevice,
                            (cudaStream_t)stream));
  return RT_OK;
}

int rt_learner_params_changed(rt_learner* h, void* stream) {
  RT_REQUIRE(h, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  if (!h->rn) return RT_OK;
  for (int w = 0; w < 2; ++w) {
    rtk::k_shadow_params<<<grid1d(h->nparams / 4), 256, 0, (cudaStream_t)stream>>>(h->p[w], h->pr[w], h->wflag,
                                                                                 h->nparams);
    RT_LAUNCH_CHECK();
  }
  return RT_OK;
}

int rt_learner_set_lr(rt_learner* h, double lr) {
  RT_REQUIRE(h, "null argument");
  h->lr = (float)lr;
  return RT_OK;
}

int rt_learner_get_opt_state(rt_learner* h, int64_t* adam_steps, double* lr) {
  RT_REQUIRE(h && adam_steps && lr, "null argument");
  *adam_steps = (int64_t)h->adam_t;
  *lr = (double)h->lr;
  return RT_OK;
}

int rt_learner_set_opt_state(rt_learner* h, int64_t adam_steps, double lr) {
  RT_REQUIRE(h && adam_steps >= 0, "bad argument");
  h->adam_t = (long long)adam_steps;
  h->lr = (float)lr;
  return RT_OK;
}

int rt_learner_get_aux_state(rt_learner* h, uint64_t* rng_counter, float* clip_ema, int32_t* clip_ema_init) {
  RT_REQUIRE(h && rng_counter && clip_ema && clip_ema_init, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  RT_CUDA(cudaDeviceSynchronize());
  float s[2];
  RT_CUDA(cudaMemcpy(s, h->stats + 4, sizeof(s), cudaMemcpyDeviceToHost));
  *rng_counter = h->rng_counter;
  *clip_ema = s[0];
  *clip_ema_init = s[1] > 0.5f ? 1 : 0;
  return RT_OK;
}

int rt_learner_set_aux_state(rt_learner* h, uint64_t rng_counter, float clip_ema, int32_t clip_ema_init) {
  RT_REQUIRE(h, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  RT_CUDA(cudaDeviceSynchronize());
  const float s[2] = {clip_ema, clip_ema_init ? 1.f : 0.f};
  RT_CUDA(cudaMemcpy(h->stats + 4, s, sizeof(s), cudaMemcpyHostToDevice));
  h->rng_counter = rng_counter;
  return RT_OK;
}

}  // extern "C"

namespace {

// Records `body`'s launches on `st` into an executable graph (nothing runs).
template <class F>
int capture_graph(cudaStream_t st, F&& body, cudaGraphExec_t* out, long long* launches) {
  const long long l0 = (long long)rt::launch_counter().load();
  RT_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  int rc = body();
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(st, &g);
  *launches = (long long)rt::launch_counter().load() - l0;
  rt::launch_counter() -= *launches;   // nothing ran: every replay adds them instead
  if (rc != RT_OK || e != cudaSuccess || !g) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    if (rc != RT_OK) return rc;
    return rt::fail(RT_ERR_CUDA, "stream capture of the update failed: %s", cudaGetErrorString(e));
  }
  e = cudaGraphInstantiate(out, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return rt::fail(RT_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  }
  return RT_OK;
}

// fp32-frame buffer of a batch (by its frame pointer); up to 3 replay slots get their own, everything else
// shares the default buffer
float* xf_for(rt_learner* h, const void* all_x, bool create) {
  auto it = h->xf_slots.find(all_x);
  if (it != h->xf_slots.end()) return it->second;
  if (!create || h->xf_slots.size() >= RT_BATCH_SLOTS || h->conv.empty()) return h->xf0;
  float* p = nullptr;
  if (cudaMalloc(&p, h->xf_floats * sizeof(float)) != cudaSuccess) {
    cudaGetLastError();
    return h->xf0;
  }
  h->allocs.push_back(p);
  h->xf_slots[all_x] = p;
  return p;
}
// the shared-frames fast path of the update: recurrent model, rnn_bootstrap, double-Q (config 3)
bool shared_frames_path(const rt_learner* h, const rt_batch* b, const rt_learner_io* io) {
  return h->U > 0 && h->td.rnn_bootstrap && h->R > 1 && h->td.double_q && !h->conv.empty() &&
         b->target_states[io->field_x] == nullptr;
}

int learner_step_impl(rt_learner* h, const rt_batch* b, const rt_learner_io* io,
                      const float* const* taus_host, void* stream, bool apply) {
  RT_REQUIRE(h && b && io, "null argument");
  RT_REQUIRE(b->B == h->B && b->S == h->S && b->n == h->n,
             "batch geometry (B=%d,S=%d,n=%d) does not match the learner (B=%d,S=%d,n=%d)", b->B,
             b->S, b->n, h->B, h->S, h->n);
  RT_CUDA(cudaSetDevice(h->device));
  cudaStream_t caller = (cudaStream_t)stream;
  // the update runs on the learner's own stream (graph capture is not possible on the legacy
  // default stream), forked from and joined back into the caller's stream
  const bool forked = h->graphs_enabled != 0;
  cudaStream_t st = forked ? h->own : caller;
  if (forked) {
    RT_CUDA(cudaEventRecord(h->ev_fork, caller));
    RT_CUDA(cudaStreamWaitEvent(st, h->ev_fork, 0));
  }
  const int B = h->B, P = h->P, n = h->n, U = h->U, M = h->M, Nq = h->Nq;
  const size_t frame = (size_t)h->md.in_c * h->md.in_h * h->md.in_w;
  const uint8_t* all_x = (const uint8_t*)b->all_states[io->field_x];
  // frames already converted by rt_learner_prefetch (on the replay stream, behind the gather)?
  const bool prefetched = h->prefetched_x == (const void*)all_x && shared_frames_path(h, b, io);
  h->xf = prefetched ? xf_for(h, all_x, false) : h->xf0;
  if (prefetched) RT_CUDA(cudaStreamWaitEvent(st, h->ev_prefetch, 0));
  h->prefetched_x = nullptr;
  float* all_hx = U ? (float*)b->all_states[io->field_hx] : nullptr;
  float* all_cx = U ? (float*)b->all_states[io->field_cx] : nullptr;
  const float* all_init = U ? (const float*)b->all_states[io->field_initials] : nullptr;
  const int X = h->X;
  RT_REQUIRE(!X || (io->field_extra >= 0 && io->field_extra < RT_MAX_FIELDS && b->all_states[io->field_extra]),
             "the model takes %d extra features: rt_learner_io.field_extra must name that batch leaf", X);
  const float* all_extra = X ? (const float*)b->all_states[io->field_extra] : nullptr;
  // bytes per observation row: uint8 frames, or float32 vectors without a CNN
  const size_t xrow = frame * (h->conv.empty() ? sizeof(float) : 1);
  // separately stacked target states (online history: per-row n-step): no row sharing between the passes
  const bool sep_targets = b->target_states[io->field_x] != nullptr;
  auto tview = [&](int row0) {
    StateView sv;
    sv.x = (const uint8_t*)b->target_states[io->field_x] + (size_t)row0 * B * xrow;
    sv.extra = X ? (const float*)b->target_states[io->field_extra] + (size_t)row0 * B * X : nullptr;
    sv.hx = U ? (float*)b->target_states[io->field_hx] + (size_t)row0 * B * U : nullptr;
    sv.cx = U ? (float*)b->target_states[io->field_cx] + (size_t)row0 * B * U : nullptr;
    sv.initials = U ? (const float*)b->target_states[io->field_initials] + (size_t)row0 * B : nullptr;
    return sv;
  };
  auto view = [&](int row0) {
    StateView sv;
    sv.x = all_x + (size_t)row0 * B * xrow;
    sv.extra = X ? all_extra + (size_t)row0 * B * X : nullptr;
    sv.hx = U ? all_hx + (size_t)row0 * B * U : nullptr;
    sv.cx = U ? all_cx + (size_t)row0 * B * U : nullptr;
    sv.initials = U ? all_init + (size_t)row0 * B : nullptr;
    return sv;
  };
  const float* feat = nullptr;
  // (M, h->feat) trunk features of the training pass: deterministic buffer addresses, so the value set
  // while the forward phase is issued / captured is also right when the phases replay from graphs
  const float* train_feat = nullptr;
  const int rnn_boot = h->td.rnn_bootstrap ? 1 : 0;

  // ---- quantile fractions: injected (parity) or drawn on the device (one launch for the three
  // segments: target / selection / training pass)
  const float* tau_seg[3];
  for (int s = 0; s < 3; ++s) tau_seg[s] = h->tau_stage + (size_t)s * h->MQ;
  if (!h->dqn) {
    bool any_host = false;
    for (int s = 0; s < 3; ++s) any_host = any_host || (taus_host && taus_host[s]);
    if (!any_host) {
      k_uniform<<<cdiv((size_t)3 * h->MQ, 256), 256, 0, st>>>(h->tau_stage, (size_t)3 * h->MQ, h->td.seed,
                                                             h->rng_counter);
      RT_LAUNCH_CHECK();
      h->rng_counter += (unsigned long long)3 * h->MQ;
    } else {
      for (int s = 0; s < 3; ++s) {
        float* dst = h->tau_stage + (size_t)s * h->MQ;
        if (taus_host && taus_host[s]) {
          RT_CUDA(cudaMemcpyAsync(dst, taus_host[s], (size_t)h->MQ * sizeof(float), cudaMemcpyHostToDevice, st));
        } else {
          k_uniform<<<cdiv(h->MQ, 256), 256, 0, st>>>(dst, (size_t)h->MQ, h->td.seed, h->rng_counter);
          RT_LAUNCH_CHECK();
          h->rng_counter += (unsigned long long)h->MQ;
        }
      }
    }
  }

  StateView svt = view(P);
  const long long* actions = (const long long*)b->policy_outputs[io->po_field_actions] + (size_t)P * B;
  const double* weights = b->importance_weights ? b->importance_weights + (size_t)P * B : nullptr;

  // ---- forward phase: burn-in, bootstrap targets, training forward, losses
  auto forward_phase = [&]() -> int {
    h->side_active = forked && !h->gx.profile;
    struct Off { rt_learner* h; ~Off() { h->side_active = false; } } off{h};
    bool train_heads_done = false;
    // ---- burn-in (multi_step_trainer.py:90-131): only the recurrent state is needed, so the
    // heads the reference also evaluates are skipped.  states / target_states alias one stack,
    // and the write-back order (online first) is the reference's.
    if (P > 0) {
      RT_REQUIRE(!sep_targets, "burn-in needs the overlapped state stack (replay history buffers)");
      for (int pass = 0; pass < 1 + rnn_boot; ++pass) {
        int row0 = pass == 0 ? 0 : n;
        StateView sv = view(row0);
        RT_TRY(trunk_forward(h, st, h->pr[pass], sv, P * B, P, &feat));
        StateView dst = view(row0 + P);
        size_t last = (size_t)(P - 1) * B * U;
        k_store_state<<<cdiv((size_t)B * U, 256), 256, 0, st>>>(h->h_all + last, h->c_all + last,
                                                               dst.initials, dst.hx, dst.cx, B, U);
        RT_LAUNCH_CHECK();
      }
    }

    // ---- bootstrap target (iqn.py:15-52): target net, then the action-selection net
    bool shared_cnn = false;
    const int R = h->R, BR = M / h->R;     // LSTM view of a pass: R steps of BR rows (R = T, BR = B by default)
    const bool merged = U > 0 && rnn_boot && R > 1 && !sep_targets;
    if (merged) {
      // recurrent model with rnn_bootstrap: the target, selection and training recurrences are
      // independent, so run both CNNs + input-gate GEMMs first and then ALL recurrences in one
      // launch (20 dependent steps instead of 60).  The online input gates are computed once
      // over the T+n distinct rows: the selection pass reads rows [n, T+n), training rows [0, T).
      StateView sv = view(P + n);
      // Two branches: the TARGET network's work (CNN + input gates before the recurrences, its heads
      // pass after them) runs on the side branch with the second activation set while the main
      // branch does the online network's; they meet at the one LSTM launch and at the target kernel.
      const bool fork_fwd = h->side_active && h->overlap_fwd && h->td.double_q && cnn_all_implicit(h);
      const bool has_cnn = !h->conv.empty();
      const float* f = nullptr;
      bool pair1 = false;
      if (h->td.double_q) {
        // the target pass (rows [n, T+n)) and the online pass (rows [0, T+n)) read the same frames:
        // convert them to fp32 NHWC once
        if (has_cnn && !prefetched)
          RT_TRY(launch_frames_to_nhwc(st, svt.x, h->xf, M + n * B, h->md.in_c, h->md.in_h, h->md.in_w,
                                       (float)(1.0 / 255.0), h->rn));
        const float* xf_t = has_cnn ? h->xf + (size_t)n * B * frame : nullptr;
        if (fork_fwd) {
          // conv1 of both networks as one product over the shared frames, then the branches split
          pair1 = has_cnn && conv1_pair_eligible(h, n * B);
          if (pair1) RT_TRY(conv1_pair_forward(h, st, h->pr[0], h->pr[1], h->xf, M + n * B, n * B));
          SideCtx sd;
          RT_TRY(side_begin(h, st, &sd));
          RT_TRY(feature_forward(h, sd.st, h->pr[1], sv.x, M, &f, xf_t, true, pair1 ? 1 : 0));
          RT_TRY(lstm_xgates(h, sd.st, h->pr[1], f, M, h->xg2, sv.extra, sd.gx));
        } else {
          RT_TRY(feature_forward(h, st, h->pr[1], sv.x, M, &f, xf_t));
          RT_TRY(lstm_xgates(h, st, h->pr[1], f, M, h->xg2, sv.extra));
        }
      } else {
        RT_TRY(feature_forward(h, st, h->pr[1], sv.x, M, &f));
        RT_TRY(lstm_xgates(h, st, h->pr[1], f, M, h->xg2, sv.extra));
      }
      SeqDesc seqs[3];
      int ns = 0;
      seqs[ns++] = SeqDesc{h->pr[1], h->xg2, sv.hx, sv.cx, sv.initials, h->h_all2, 1, false};
      if (h->td.double_q) {
        RT_TRY(feature_forward(h, st, h->pr[0], svt.x, M + n * B, &f, has_cnn ? h->xf : nullptr, false, pair1 ? 1 : 0));
        RT_TRY(lstm_xgates(h, st, h->pr[0], f, M + n * B, h->xg, svt.extra));
        seqs[ns++] = SeqDesc{h->pr[0], h->xg + (size_t)n * B * 4 * U, sv.hx, sv.cx, sv.initials, h->h_all3, 2, false};
      } else {
        RT_TRY(feature_forward(h, st, h->pr[0], svt.x, M, &f));
        RT_TRY(lstm_xgates(h, st, h->pr[0], f, M, h->xg, svt.extra));
      }
      train_feat = f;
      seqs[ns++] = SeqDesc{h->pr[0], h->xg, svt.hx, svt.cx, svt.initials, h->h_all, 0, true};
      if (fork_fwd) RT_TRY(side_join(h, st));
      // The quantile embeddings of the three heads passes do not depend on the trunk: they run on the side
      // branches NEXT TO the recurrence (whose 128 latency-bound CTAs leave the rest of the GPU idle)
      const bool phi_early = fork_fwd && !h->dqn && h->phi_early;
      const HeadSet hs2 = second_set(h, h->tq);
      const HeadSet hs_sel = third_set(h, h->sq);
      if (phi_early) {
        SideCtx sd;
        RT_TRY(side_begin(h, st, &sd));
        RT_TRY(heads_forward(h, sd.st, h->pr[1], nullptr, M, tau_seg[0], &hs2, 1));
        RT_CUDA(cudaEventRecord(h->ev_side_b[0], st));
        RT_CUDA(cudaStreamWaitEvent(h->side_b, h->ev_side_b[0], 0));
        RT_TRY(heads_forward(h, h->side_b, h->pr[0], nullptr, M, tau_seg[1], &hs_sel, 1));
        RT_TRY(heads_forward(h, h->side_b, h->pr[0], nullptr, M, tau_seg[2], nullptr, 1));
        RT_CUDA(cudaEventRecord(h->ev_phi, h->side_b));
      }
      RT_TRY(lstm_run(h, st, seqs, ns, R, BR));
      const int hp = phi_early ? 2 : 0;
      if (phi_early) RT_CUDA(cudaStreamWaitEvent(st, h->ev_phi, 0));
      if (fork_fwd) {
        // target heads on the side branch (second set, q straight into tq); selection and training
        // heads on the main branch; the bootstrap target needs all three
        SideCtx sd;
        RT_TRY(side_begin(h, st, &sd));
        RT_TRY(heads_forward(h, sd.st, h->pr[1], h->h_all2, M, tau_seg[0], &hs2, hp));
        // ... and the selection pass on a third branch with its own set
        RT_CUDA(cudaEventRecord(h->ev_side_b[0], st));
        RT_CUDA(cudaStreamWaitEvent(h->side_b, h->ev_side_b[0], 0));
        RT_TRY(heads_forward(h, h->side_b, h->pr[0], h->h_all3, M, tau_seg[1], &hs_sel, hp));
        RT_TRY(heads_forward(h, st, h->pr[0], h->h_all, M, tau_seg[2], nullptr, hp));
        RT_CUDA(cudaEventRecord(h->ev_side_b[1], h->side_b));
        RT_CUDA(cudaStreamWaitEvent(st, h->ev_side_b[1], 0));
        RT_TRY(side_join(h, st));
        train_heads_done = true;
      } else {
        RT_TRY(heads_forward(h, st, h->pr[1], h->h_all2, M, tau_seg[0]));
        RT_CUDA(cudaMemcpyAsync(h->tq, h->q, (size_t)h->MQ * h->A * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (h->td.double_q) RT_TRY(heads_forward(h, st, h->pr[0], h->h_all3, M, tau_seg[1]));
        else RT_TRY(heads_forward(h, st, h->pr[1], h->h_all2, M, tau_seg[1]));
        RT_CUDA(cudaMemcpyAsync(h->sq, h->q, (size_t)h->MQ * h->A * sizeof(float), cudaMemcpyDeviceToDevice, st));
      }
      size_t off = (size_t)P * B;
      rtk::k_iqn_target<<<cdiv(M, 4), 128, (size_t)4 * Nq * h->A * sizeof(float), st>>>(
          h->tq, h->sq, b->returns + off, b->target_masks + off, (const long long*)b->nsteps + off,
          h->targets, M, Nq, h->A, (float)h->td.gamma, h->td.vf_scale_epsilon);
      RT_LAUNCH_CHECK();
      feat = h->h_all;
    } else {
      StateView sv = sep_targets ? tview(P) : view(P + n);
      int ts = rnn_boot ? R : 1;
      RT_TRY(trunk_forward(h, st, h->pr[1], sv, M, ts, &feat));
      RT_TRY(heads_forward(h, st, h->pr[1], feat, M, tau_seg[0]));
      RT_CUDA(cudaMemcpyAsync(h->tq, h->q, (size_t)h->MQ * h->A * sizeof(float), cudaMemcpyDeviceToDevice, st));
      if (!h->td.double_q) {
        // same network, same states: only the quantile fractions differ -> reuse the whole trunk
        RT_TRY(heads_forward(h, st, h->pr[1], feat, M, tau_seg[1]));
      } else if (sep_targets) {
        // separately stacked target states: the online network's selection pass is a pass of its own
        const float* f = nullptr;
        RT_TRY(trunk_forward(h, st, h->pr[0], sv, M, ts, &f));
        RT_TRY(heads_forward(h, st, h->pr[0], f, M, tau_seg[1]));
      } else {
        // online net on target_states = rows [n, T+n) of the stack; the training forward below
        // needs rows [0, T): run the online CNN ONCE over the T+n distinct rows and let both
        // passes read their slice (saves (T-n)/(2T) of the online conv work)
        StateView s0 = view(P);
        RT_TRY(feature_forward(h, st, h->pr[0], s0.x, M + n * B, &train_feat));
        shared_cnn = true;
        const float* f = train_feat + (size_t)n * B * h->feat;
        if (U) {
          RT_TRY(lstm_forward(h, st, h->pr[0], f, M, ts, sv.hx, sv.cx, sv.initials, sv.extra));
          f = h->h_all;
        }
        RT_TRY(heads_forward(h, st, h->pr[0], f, M, tau_seg[1]));
      }
      RT_CUDA(cudaMemcpyAsync(h->sq, h->q, (size_t)h->MQ * h->A * sizeof(float), cudaMemcpyDeviceToDevice, st));
      size_t off = (size_t)P * B;
      rtk::k_iqn_target<<<cdiv(M, 4), 128, (size_t)4 * Nq * h->A * sizeof(float), st>>>(
          h->tq, h->sq, b->returns + off, b->target_masks + off, (const long long*)b->nsteps + off,
          h->targets, M, Nq, h->A, (float)h->td.gamma, h->td.vf_scale_epsilon);
      RT_LAUNCH_CHECK();

      // ---- training forward (iqn.py:54-129)
      if (shared_cnn) {
        feat = train_feat;
        if (U) {
          RT_TRY(lstm_forward(h, st, h->pr[0], feat, M, R, svt.hx, svt.cx, svt.initials, svt.extra));
          feat = h->h_all;
        }
      } else {
        RT_TRY(trunk_forward(h, st, h->pr[0], svt, M, R, &feat));
        // trunk features of the training pass (the LSTM's input): what the backward pass reads
        train_feat = h->pre.empty() ? (h->conv.empty() ? reinterpret_cast<const float*>(svt.x) : h->c_out.back())
                                    : h->pre_out.back();
      }
    }
    // ---- training heads + loss (iqn.py:54-129)
    if (!train_heads_done) RT_TRY(heads_forward(h, st, h->pr[0], feat, M, tau_seg[2]));
    RT_CUDA(cudaMemcpyAsync(h->tau, tau_seg[2], (size_t)h->MQ * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (h->dqn) {
      rtk::k_dqn_loss<<<cdiv(M, 128), 128, 0, st>>>(h->q, h->targets, actions, weights, h->dtheta, h->row_loss,
                                                   h->report, h->row_q, M, h->A, (float)h->td.huber_kappa,
                                                   h->td.loss_mse, h->loss_scale);
      RT_LAUNCH_CHECK();
      rtk::k_loss_stats<<<1, 256, 0, st>>>(h->row_loss, h->row_q, h->stats, M, h->loss_scale);
      RT_LAUNCH_CHECK();
    } else {
      int threads = ((Nq + 31) / 32) * 32;
      size_t smem = (3 * (size_t)Nq + 2 * threads) * sizeof(float);
      rtk::k_iqn_loss<<<M, threads, smem, st>>>(h->q, h->targets, h->tau, actions, weights, h->dtheta,
                                               h->row_loss, h->report, Nq, h->A,
                                               (float)h->td.huber_kappa, h->loss_scale);
      RT_LAUNCH_CHECK();
      rtk::k_loss_stats<<<1, 256, 0, st>>>(h->row_loss, h->report, h->stats, M, h->loss_scale);
      RT_LAUNCH_CHECK();
    }

    return RT_OK;
  };
  // ---- backward phase
  // part 0: the whole backward pass; 1: heads + LSTM (every non-conv gradient); 2: conv stack.
  // Data-parallel updates (apply == false) run parts 1 and 2 with an event in between, so the
  // all-reduce of the non-conv gradients (99 % of the bytes) overlaps the conv backward.
  // Data-parallel updates (apply == false) additionally publish the point where the gradients of every
  // non-conv parameter (99 % of the bytes) are final: an EXTERNAL event-record node on the weight-gradient
  // branch right behind its last such leaf, inside the one backward graph -- the caller's all-reduce of that
  // bucket then overlaps the conv backward without cutting the schedule in two (rt_learner_wait_late_grads).
  const bool dp_event = !apply && h->dp_split == 0;
  auto backward_part = [&](int part) -> int {
    h->side_active = h->overlap_bwd && forked && !h->gx.profile;
    struct Off { rt_learner* h; ~Off() { h->side_active = false; } } off{h};
    if (part != 2) {
      RT_CUDA(cudaMemsetAsync(h->grad, 0, h->nparams * sizeof(float), st));
      RT_TRY(heads_backward(h, st, h->pr[0], feat, M, actions));
      if (U) RT_TRY(lstm_backward(h, st, h->pr[0], train_feat, M, h->R, svt.initials, svt.extra));
      if (part == 0 && dp_event) {
        cudaStream_t es = h->side_active ? h->side : st;
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        RT_CUDA(cudaStreamIsCapturing(es, &cs));
        RT_CUDA(cudaEventRecordWithFlags(h->ev_late, es, cs == cudaStreamCaptureStatusActive
                                                              ? cudaEventRecordExternal : cudaEventRecordDefault));
      }
    }
    // with an LSTM the ReLU derivative of the trunk's last layer is already applied to dfeat (lstm_backward)
    if (part != 1) RT_TRY(features_backward(h, st, h->pr[0], svt.x, M, U ? h->dfeat : h->dfeatq, U > 0));
    return side_join(h, st);
  };
  auto backward_phase = [&]() -> int { return backward_part(0); };
  auto backward_late = [&]() -> int { return backward_part(1); };
  auto backward_conv = [&]() -> int { return backward_part(2); };
  const bool split_bwd = !apply && h->dp_split != 0;    // RT_DP_SPLIT=1: the older two-graph schedule

  // ---- run: replayed from CUDA graphs once the handle is warm (every lazy allocation / kernel
  // attribute of these shapes has happened), keyed by the batch's device pointers (the replay
  // buffer rotates three batch slots)
  const bool want_graph = h->graphs_enabled > 0 && h->steps_done >= 2 && !h->gx.profile && !h->lstm_dbg;
  rt_learner::StepGraphs* sg = nullptr;
  if (want_graph) {
    const void* key[12] = {all_x, all_hx, all_cx, all_init, b->returns, b->nsteps, b->target_masks,
                           b->policy_outputs[io->po_field_actions], b->importance_weights,
                           (const void*)(uintptr_t)((split_bwd ? 1 : 0) | (prefetched ? 2 : 0) | (dp_event ? 4 : 0)),
                           all_extra,
                           b->target_states[io->field_x]};
    for (auto& g : h->graphs)
      if (memcmp(g.key, key, sizeof(key)) == 0) sg = &g;
    if (!sg) {
      if (h->graphs.size() >= 8) {
        if (h->graphs[0].fwd) cudaGraphExecDestroy(h->graphs[0].fwd);
        if (h->graphs[0].bwd) cudaGraphExecDestroy(h->graphs[0].bwd);
        if (h->graphs[0].bwd2) cudaGraphExecDestroy(h->graphs[0].bwd2);
        h->graphs.erase(h->graphs.begin());
      }
      rt_learner::StepGraphs ng;
      memcpy(ng.key, key, sizeof(key));
      int rc = capture_graph(st, forward_phase, &ng.fwd, &ng.n_fwd);
      if (rc == RT_OK && !split_bwd) rc = capture_graph(st, backward_phase, &ng.bwd, &ng.n_bwd);
      if (rc == RT_OK && split_bwd) rc = capture_graph(st, backward_late, &ng.bwd, &ng.n_bwd);
      if (rc == RT_OK && split_bwd) rc = capture_graph(st, backward_conv, &ng.bwd2, &ng.n_bwd2);
      if (rc != RT_OK) {
        // capture is an optimisation: fall back to issuing the launches one by one
        if (ng.fwd) cudaGraphExecDestroy(ng.fwd);
        if (ng.bwd) cudaGraphExecDestroy(ng.bwd);
        if (ng.bwd2) cudaGraphExecDestroy(ng.bwd2);
        h->graphs_enabled = -1;   // stay on the own stream, never try again
        fprintf(stderr, "rltime_b200: CUDA-graph capture of the update disabled: %s\n", rt::last_error().c_str());
      } else {
        h->graphs.push_back(ng);
        sg = &h->graphs.back();
      }
    }
  }
  if (sg) {
    RT_CUDA(cudaGraphLaunch(sg->fwd, st));
    rt::launch_counter() += sg->n_fwd;
  } else {
    RT_TRY(forward_phase());
  }
  // losses, |td| and the loss statistics are final here: the priority write-back and the next
  // draw only depend on this point, not on the backward pass (rt_learner_wait_loss)
  RT_CUDA(cudaEventRecord(h->ev_loss, st));
  if (sg) {
    RT_CUDA(cudaGraphLaunch(sg->bwd, st));
    rt::launch_counter() += sg->n_bwd;
  } else {
    RT_TRY(split_bwd ? backward_late() : backward_phase());
  }
  if (split_bwd) {
    RT_CUDA(cudaEventRecord(h->ev_late, st));
    if (sg) {
      RT_CUDA(cudaGraphLaunch(sg->bwd2, st));
      rt::launch_counter() += sg->n_bwd2;
    } else {
      RT_TRY(backward_conv());
    }
  }
  h->steps_done++;
  h->xf = h->xf0;     // everything of this update is enqueued: acting / hand-built batches use the default buffer
  int rc_apply = apply ? apply_grads(h, st, 1.0f) : RT_OK;
  if (forked) {
    RT_CUDA(cudaEventRecord(h->ev_join, st));
    RT_CUDA(cudaStreamWaitEvent(caller, h->ev_join, 0));
  }
  return rc_apply;
}

}  // namespace

extern "C" {

int rt_learner_step(rt_learner* h, const rt_batch* b, const rt_learner_io* io,
                    const float* const* taus_host, void* stream) {
  return learner_step_impl(h, b, io, taus_host, stream, true);
}

int rt_learner_compute_grads(rt_learner* h, const rt_batch* b, const rt_learner_io* io,
                             const float* const* taus_host, void* stream) {
  return learner_step_impl(h, b, io, taus_host, stream, false);
}

int rt_learner_prefetch(rt_learner* h, const rt_batch* b, const rt_learner_io* io, void* stream) {
  RT_REQUIRE(h && b && io, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  if (!shared_frames_path(h, b, io) || b->B != h->B || b->S != h->S || b->n != h->n) return RT_OK;
  const uint8_t* all_x = (const uint8_t*)b->all_states[io->field_x];
  float* xf = xf_for(h, all_x, true);
  if (xf == h->xf0) return RT_OK;     // no private buffer for this batch: the update converts the frames itself
  const size_t frame = (size_t)h->md.in_c * h->md.in_h * h->md.in_w;
  cudaStream_t st = (cudaStream_t)stream;
  RT_TRY(launch_frames_to_nhwc(st, all_x + (size_t)h->P * h->B * frame, xf, h->M + h->n * h->B, h->md.in_c,
                               h->md.in_h, h->md.in_w, (float)(1.0 / 255.0), h->rn));
  RT_CUDA(cudaEventRecord(h->ev_prefetch, st));
  h->prefetched_x = all_x;
  return RT_OK;
}

int rt_learner_apply_grads(rt_learner* h, double grad_scale, void* stream) {
  RT_REQUIRE(h, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  return apply_grads(h, (cudaStream_t)stream, (float)grad_scale);
}

int rt_learner_flat_buffer(rt_learner* h, int32_t which, float** dev_ptr, int64_t* count) {
  RT_REQUIRE(h && dev_ptr && count, "null argument");
  float* p = which_buffer(h, which);
  RT_REQUIRE(p, "bad buffer selector");
  *dev_ptr = p;
  *count = (int64_t)h->nparams;
  return RT_OK;
}

int rt_learner_act(rt_learner* h, int32_t E, const void* x, const float* extra, const float* hx,
                   const float* cx, const float* initials, const float* taus_host, float* qvalues,
                   float* h_out, float* c_out, void* stream) {
  RT_REQUIRE(h && x && qvalues && E >= 1, "bad argument");
  RT_REQUIRE(E <= h->max_rows && (size_t)E * h->Nq <= (size_t)h->MQ,
             "acting batch of %d envs exceeds the learner's buffers (max %d)", E, h->MQ / h->Nq);
  RT_REQUIRE(!h->U || (hx && cx && initials && h_out && c_out), "recurrent model needs hx/cx/initials");
  RT_REQUIRE(!h->X || extra, "the model takes %d extra features per observation", h->X);
  RT_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  StateView sv;
  sv.x = static_cast<const uint8_t*>(x);
  sv.extra = extra;
  sv.hx = const_cast<float*>(hx);
  sv.cx = const_cast<float*>(cx);
  sv.initials = initials;
  // quantile fractions first: they are the only per-call input that is not behind a stable pointer
  const size_t nq = (size_t)E * h->Nq;
  if (!h->dqn) {
    if (taus_host) {
      RT_CUDA(cudaMemcpyAsync(h->tau_stage, taus_host, nq * sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
      k_uniform<<<cdiv(nq, 256), 256, 0, st>>>(h->tau_stage, nq, h->td.seed ^ 0xA5A5A5A5ULL, h->rng_counter);
      RT_LAUNCH_CHECK();
      h->rng_counter += nq;
    }
  }
  auto body = [&]() -> int {
    const float* feat = nullptr;
    RT_TRY(trunk_forward(h, st, h->pr[0], sv, E, 1, &feat));
    RT_TRY(heads_forward(h, st, h->pr[0], feat, E, h->tau_stage));
    rtk::k_quantile_mean<<<cdiv((size_t)E * h->A, 128), 128, 0, st>>>(h->q, qvalues, E, h->Nq, h->A);
    RT_LAUNCH_CHECK();
    if (h->U) {
      RT_CUDA(cudaMemcpyAsync(h_out, h->h_all, (size_t)E * h->U * sizeof(float), cudaMemcpyDeviceToDevice, st));
      RT_CUDA(cudaMemcpyAsync(c_out, h->c_all, (size_t)E * h->U * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return RT_OK;
  };
  // A host that acts from persistent buffers (DevicePolicy: two alternating state buffers) gets the step
  // replayed from a CUDA graph: the launches are ~5 us of GPU work each, issued one by one they are bound
  // by the launch path.  Graphs cannot be captured on the legacy default stream.
  rt_learner::ActGraph* ag = nullptr;
  if (h->act_graphs_enabled > 0 && st != nullptr && st != cudaStreamLegacy && !h->gx.profile && !h->lstm_dbg) {
    const void* key[10] = {(const void*)(intptr_t)E, x, extra, hx, cx, initials, qvalues, h_out, c_out, h->xf};
    for (auto& g : h->act_graphs)
      if (memcmp(g.key, key, sizeof(key)) == 0) ag = &g;
    if (!ag) {
      if (h->act_graphs.size() >= 4) {
        if (h->act_graphs[0].g) cudaGraphExecDestroy(h->act_graphs[0].g);
        h->act_graphs.erase(h->act_graphs.begin());
      }
      h->act_graphs.emplace_back();
      ag = &h->act_graphs.back();
      memcpy(ag->key, key, sizeof(key));
    }
    if (!ag->g && ++ag->seen >= 3) {
      // every lazy allocation / kernel attribute of these shapes happened in the two calls before
      if (capture_graph(st, body, &ag->g, &ag->n) != RT_OK) {
        ag->g = nullptr;
        h->act_graphs_enabled = -1;
        fprintf(stderr, "rltime_b200: CUDA-graph capture of the acting step disabled: %s\n", rt::last_error().c_str());
      }
    }
  }
  if (ag && ag->g) {
    RT_CUDA(cudaGraphLaunch(ag->g, st));
    rt::launch_counter() += ag->n;
    return RT_OK;
  }
  return body();
}

int rt_learner_wait_late_grads(rt_learner* h, void* stream, int64_t* first, int64_t* count) {
  RT_REQUIRE(h, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  if (stream != (void*)-1) RT_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_late, 0));
  if (first) *first = (int64_t)h->conv_param_end;
  if (count) *count = (int64_t)(h->nparams - h->conv_param_end);
  return RT_OK;
}

int rt_learner_wait_loss(rt_learner* h, void* stream) {
  RT_REQUIRE(h, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  RT_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->ev_loss, 0));
  return RT_OK;
}

int rt_learner_read_loss(rt_learner* h, float* loss, float* td_mean, void* stream) {
  RT_REQUIRE(h, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  RT_CUDA(cudaStreamWaitEvent(st, h->ev_loss, 0));
  RT_CUDA(cudaMemcpyAsync(h->h_stats, h->stats, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  RT_CUDA(cudaStreamSynchronize(st));
  if (loss) *loss = h->h_stats[0];
  if (td_mean) *td_mean = h->h_stats[1];
  return RT_OK;
}

int rt_learner_td_abs(rt_learner* h, float** out_device) {
  RT_REQUIRE(h && out_device, "null argument");
  *out_device = h->report;
  return RT_OK;
}

int rt_learner_read_stats(rt_learner* h, float* loss, float* td_mean, float* grad_norm, void* stream) {
  RT_REQUIRE(h, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  float s[4];
  RT_CUDA(cudaMemcpyAsync(s, h->stats, sizeof(s), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  RT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (loss) *loss = s[0];
  if (td_mean) *td_mean = s[1];
  if (grad_norm) *grad_norm = s[2];
  return RT_OK;
}

int rt_learner_debug_tensor(rt_learner* h, const char* name, void** dev_ptr, int64_t* count) {
  RT_REQUIRE(h && name && dev_ptr && count, "null argument");
  auto it = h->debug.find(name);
  RT_REQUIRE(it != h->debug.end(), "no debug tensor named '%s'", name);
  *dev_ptr = it->second.first;
  *count = it->second.second;
  return RT_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ data parallelism (NCCL)
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch
// already loaded), so the library itself has no link-time dependency on it.
namespace {
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
int nccl_api(Nccl** out) {
  static Nccl api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRankConfig = (decltype(api.CommInitRankConfig))dlsym(api.lib, "ncclCommInitRankConfig");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
      api.Broadcast = (decltype(api.Broadcast))dlsym(api.lib, "ncclBroadcast");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    }
  }
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Broadcast || !api.CommDestroy)
    return rt::fail(RT_ERR_NCCL, "libnccl.so.2 could not be loaded (%s)", api.lib ? "missing symbols" : dlerror());
  *out = &api;
  return RT_OK;
}
#define RT_NCCL(api, call)                                                                         \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess)                                                                        \
      return rt::fail(RT_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call,                   \
                      (api)->GetErrorString ? (api)->GetErrorString(r__) : "nccl error");          \
  } while (0)
}  // namespace

extern "C" int rt_comm_unique_id(uint8_t* id128) {
  RT_REQUIRE(id128, "null argument");
  Nccl* api = nullptr;
  RT_TRY(nccl_api(&api));
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == RT_COMM_ID_BYTES, "ncclUniqueId size");
  RT_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return RT_OK;
}

extern "C" int rt_comm_init(rt_learner* h, const uint8_t* id128, int32_t rank, int32_t world) {
  RT_REQUIRE(h && id128 && world >= 1 && rank >= 0 && rank < world, "bad argument");
  RT_REQUIRE(!h->comm, "communicator already initialised");
  RT_CUDA(cudaSetDevice(h->device));
  Nccl* api = nullptr;
  RT_TRY(nccl_api(&api));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  // few CTAs: the all-reduce of the late-gradient bucket runs next to the conv backward; the persistent conv
  // data-gradient kernels leave as many SMs free (reserve_sms) as NCCL may use (RT_NCCL_MAX_CTAS overrides)
  int max_ctas = 32;     // measured at 8 GPUs: 32 CTAs + 32 reserved SMs 1.396 ms, 16 + 16: 1.423 ms, 32 + 0: 1.430 ms
  if (const char* e = getenv("RT_NCCL_MAX_CTAS")) max_ctas = atoi(e);
  if (api->CommInitRankConfig && max_ctas > 0) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.maxCTAs = max_ctas;
    RT_NCCL(api, api->CommInitRankConfig(&h->comm, world, id, rank, &cfg));
  } else {
    RT_NCCL(api, api->CommInitRank(&h->comm, world, id, rank));
  }
  h->comm_rank = rank;
  h->comm_world = world;
  if (world > 1) {
    h->reserve_sms = max_ctas > 0 && max_ctas < 64 ? max_ctas : 16;
    if (const char* e = getenv("RT_DP_RESERVE_SMS")) h->reserve_sms = atoi(e);
  }
  // highest priority: NCCL's few CTAs must all become resident to make progress, and they compete with the
  // conv backward's CTAs for every slot that frees up (RT_DP_COMM_PRIO=0: lowest priority)
  int lo = 0, hi = 0;
  RT_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  int prio = hi;
  if (const char* e = getenv("RT_DP_COMM_PRIO")) prio = atoi(e) ? hi : lo;
  RT_CUDA(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, prio));
  RT_CUDA(cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
  return RT_OK;
}

extern "C" int rt_comm_destroy(rt_learner* h) {
  RT_REQUIRE(h, "null argument");
  if (!h->comm) return RT_OK;
  Nccl* api = nullptr;
  RT_TRY(nccl_api(&api));
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  api->CommDestroy(h->comm);
  h->comm = nullptr;
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->ev_comm) cudaEventDestroy(h->ev_comm);
  h->comm_stream = nullptr;
  h->ev_comm = nullptr;
  h->comm_world = 1;
  return RT_OK;
}

extern "C" int rt_comm_broadcast_params(rt_learner* h, int32_t root, void* stream) {
  RT_REQUIRE(h && h->comm, "rt_comm_init first");
  RT_CUDA(cudaSetDevice(h->device));
  Nccl* api = nullptr;
  RT_TRY(nccl_api(&api));
  cudaStream_t st = (cudaStream_t)stream;
  for (int w = 0; w < 2; ++w)
    RT_NCCL(api, api->Broadcast(h->p[w], h->p[w], h->nparams, ncclFloat, root, h->comm, st));
  return rt_learner_params_changed(h, stream);
}

extern "C" int rt_comm_allreduce_max_f64(rt_learner* h, double* dev_values, int32_t count, void* stream) {
  RT_REQUIRE(h && h->comm && dev_values && count > 0, "bad argument");
  RT_CUDA(cudaSetDevice(h->device));
  Nccl* api = nullptr;
  RT_TRY(nccl_api(&api));
  RT_NCCL(api, api->AllReduce(dev_values, dev_values, (size_t)count, ncclDouble, ncclMax, h->comm, (cudaStream_t)stream));
  return RT_OK;
}

// One data-parallel update entirely inside the library: local gradients -> NCCL sum over NVLink in two
// buckets (every non-conv gradient on the communication stream while the conv backward still runs, the
// small conv bucket afterwards) -> identical clip + Adam on every rank with grad_scale = 1 / world.
extern "C" int rt_learner_step_dp(rt_learner* h, const rt_batch* b, const rt_learner_io* io,
                                  const float* const* taus_host, void* stream) {
  RT_REQUIRE(h && h->comm, "rt_comm_init first");
  Nccl* api = nullptr;
  RT_TRY(nccl_api(&api));
  RT_TRY(learner_step_impl(h, b, io, taus_host, stream, false));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t first = h->conv_param_end, count = h->nparams - h->conv_param_end;
  if (h->comm_world > 1) {
    if (first > 0 && count > 0) {
      RT_CUDA(cudaStreamWaitEvent(h->comm_stream, h->ev_late, 0));
      RT_NCCL(api, api->AllReduce(h->grad + first, h->grad + first, count, ncclFloat, ncclSum, h->comm, h->comm_stream));
      RT_CUDA(cudaEventRecord(h->ev_comm, h->comm_stream));
      RT_NCCL(api, api->AllReduce(h->grad, h->grad, first, ncclFloat, ncclSum, h->comm, st));
      RT_CUDA(cudaStreamWaitEvent(st, h->ev_comm, 0));
    } else {
      RT_NCCL(api, api->AllReduce(h->grad, h->grad, h->nparams, ncclFloat, ncclSum, h->comm, st));
    }
  }
  return apply_grads(h, st, 1.0f / (float)h->comm_world);
}

extern "C" int rt_gemm_test(int32_t mode, int32_t M, int32_t N, int32_t K, int32_t transA,
                            int32_t transB, const float* A, const float* B, const float* bias,
                            int32_t relu, float* C, int32_t device) {
  RT_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "bad argument");
  RT_CUDA(cudaSetDevice(device));
  GemmCtx cx;
  cx.mode = mode;
  if (const char* e = getenv("RT_TC_PERSISTENT")) cx.persistent = atoi(e);
  if (const char* e = getenv("RT_TC_PAIR")) cx.pair = atoi(e);
  cx.ws_floats = (size_t)16 << 20;
  float *dA = nullptr, *dB = nullptr, *dC = nullptr, *dbias = nullptr;
  size_t nA = (size_t)M * K, nB = (size_t)N * K, nC = (size_t)M * N;
  RT_CUDA(cudaMalloc(&cx.ws, cx.ws_floats * sizeof(float)));
  RT_CUDA(cudaMalloc(&dA, nA * sizeof(float)));
  RT_CUDA(cudaMalloc(&dB, nB * sizeof(float)));
  RT_CUDA(cudaMalloc(&dC, nC * sizeof(float)));
  RT_CUDA(cudaMemcpy(dA, A, nA * sizeof(float), cudaMemcpyHostToDevice));
  RT_CUDA(cudaMemcpy(dB, B, nB * sizeof(float), cudaMemcpyHostToDevice));
  RT_CUDA(cudaMemset(dC, 0, nC * sizeof(float)));
  if (bias) {
    RT_CUDA(cudaMalloc(&dbias, (size_t)N * sizeof(float)));
    RT_CUDA(cudaMemcpy(dbias, bias, (size_t)N * sizeof(float), cudaMemcpyHostToDevice));
  }
  rtk::GemmArgs g = mk(dA, transA ? M : K, transA, dB, transB ? K : N, transB, dC, N, M, N, K);
  g.bias = dbias;
  g.relu = relu;
  int rc = gemm(cx, 0, g);
  if (rc == RT_OK) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = rt::fail(RT_ERR_CUDA, "gemm test kernel failed: %s", cudaGetErrorString(e));
  }
  if (rc == RT_OK && mode == 1 && cx.tc_launches == 0)
    rc = rt::fail(RT_ERR_INVALID, "shape not eligible for the tcgen05 path");
  if (rc == RT_OK) RT_CUDA(cudaMemcpy(C, dC, nC * sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(cx.ws);
  if (dbias) cudaFree(dbias);
  return rc;
}


extern "C" int rt_gemm_bench(int32_t mode, int32_t M, int32_t N, int32_t K, int32_t transA,
                             int32_t transB, int32_t force_bn, int32_t force_stages, int32_t iters,
                             double* avg_us, int32_t device) {
  RT_REQUIRE(avg_us && M > 0 && N > 0 && K > 0 && iters > 0, "bad argument");
  RT_CUDA(cudaSetDevice(device));
  GemmCtx cx;
  cx.mode = mode;
  if (const char* e = getenv("RT_TC_PERSISTENT")) cx.persistent = atoi(e);
  if (const char* e = getenv("RT_TC_PAIR")) cx.pair = atoi(e);
  cx.force_bn = force_bn;
  cx.force_stages = force_stages;
  cx.ws_floats = (size_t)64 << 20;
  float *dA = nullptr, *dB = nullptr, *dC = nullptr;
  size_t nA = (size_t)M * K, nB = (size_t)N * K, nC = (size_t)M * N;
  RT_CUDA(cudaMalloc(&cx.ws, cx.ws_floats * sizeof(float)));
  RT_CUDA(cudaMalloc(&dA, nA * sizeof(float)));
  RT_CUDA(cudaMalloc(&dB, nB * sizeof(float)));
  RT_CUDA(cudaMalloc(&dC, nC * sizeof(float)));
  RT_CUDA(cudaMemset(dA, 0, nA * sizeof(float)));
  RT_CUDA(cudaMemset(dB, 0, nB * sizeof(float)));
  rtk::GemmArgs g = mk(dA, transA ? M : K, transA, dB, transB ? K : N, transB, dC, N, M, N, K);
  cudaEvent_t e0, e1;
  RT_CUDA(cudaEventCreate(&e0));
  RT_CUDA(cudaEventCreate(&e1));
  int rc = RT_OK;
  for (int i = 0; i < 3 && rc == RT_OK; ++i) rc = gemm(cx, 0, g);
  RT_CUDA(cudaEventRecord(e0, 0));
  for (int i = 0; i < iters && rc == RT_OK; ++i) rc = gemm(cx, 0, g);
  RT_CUDA(cudaEventRecord(e1, 0));
  cudaError_t e = cudaDeviceSynchronize();
  if (rc == RT_OK && e != cudaSuccess) rc = rt::fail(RT_ERR_CUDA, "gemm bench failed: %s", cudaGetErrorString(e));
  float ms = 0;
  if (rc == RT_OK) {
    cudaEventElapsedTime(&ms, e0, e1);
    *avg_us = 1e3 * ms / iters;
  }
  if (rc == RT_OK && getenv("RT_DEBUG_TIMELINE") && mode == 1) {
    long long* d = nullptr;
    cudaMalloc(&d, 8 * sizeof(long long));
    cudaMemset(d, 0, 8 * sizeof(long long));
    cx.dbg = d;
    rc = gemm(cx, 0, g);
    long long hst[8] = {0};
    cudaMemcpy(hst, d, sizeof(hst), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[timeline M=%d N=%d K=%d tA=%d tB=%d bn=%d st=%d] cycles since setup: first_full=%lld "
                    "mma_issued=%lld acc_ready=%lld epilogue_done=%lld\n",
            M, N, K, transA, transB, force_bn, force_stages, hst[1] - hst[0], hst[2] - hst[0], hst[3] - hst[0],
            hst[4] - hst[0]);
    cudaFree(d);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(cx.ws);
  return rc;
}


extern "C" int rt_learner_profile(rt_learner* h, int32_t enable) {
  RT_REQUIRE(h, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  GemmCtx& cx = h->gx;
  if (enable && cx.prof_ev.empty()) {
    cx.prof_ev.resize(2 * 8192);
    for (auto& e : cx.prof_ev) RT_CUDA(cudaEventCreate(&e));
  }
  cx.profile = enable != 0;
  cx.prof_used = 0;
  cx.prof_flops = 0;
  cx.prof_fl.clear();
  cx.prof_shape.clear();
  return RT_OK;
}

extern "C" int rt_learner_gemm_time(rt_learner* h, double* total_ms, double* total_flops, int64_t* launches) {
  RT_REQUIRE(h && total_ms && total_flops && launches, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  RT_CUDA(cudaDeviceSynchronize());
  GemmCtx& cx = h->gx;
  double t = 0;
  for (size_t i = 0; i + 1 < cx.prof_used; i += 2) {
    float ms = 0;
    RT_CUDA(cudaEventElapsedTime(&ms, cx.prof_ev[i], cx.prof_ev[i + 1]));
    t += ms;
  }
  *total_ms = t;
  *total_flops = cx.prof_flops;
  *launches = (int64_t)(cx.prof_used / 2);
  cx.prof_used = 0;
  cx.prof_flops = 0;
  cx.prof_fl.clear();
  cx.prof_shape.clear();
  return RT_OK;
}

extern "C" int rt_learner_gemm_shapes(rt_learner* h, int64_t cap, int32_t* shapes6, int64_t* count) {
  RT_REQUIRE(h && shapes6 && count, "null argument");
  GemmCtx& cx = h->gx;
  int64_t n = (int64_t)(cx.prof_shape.size() / 6);
  if (n > cap) n = cap;
  if (n > (int64_t)(cx.prof_used / 2)) n = (int64_t)(cx.prof_used / 2);
  for (int64_t i = 0; i < n * 6; ++i) shapes6[i] = cx.prof_shape[i];
  *count = n;
  return RT_OK;
}

extern "C" int rt_learner_gemm_launches(rt_learner* h, int64_t cap, double* flops, double* ms, int64_t* count) {
  RT_REQUIRE(h && flops && ms && count, "null argument");
  RT_CUDA(cudaSetDevice(h->device));
  RT_CUDA(cudaDeviceSynchronize());
  GemmCtx& cx = h->gx;
  int64_t n = 0;
  for (size_t i = 0; i + 1 < cx.prof_used && n < cap; i += 2, ++n) {
    float t = 0;
    RT_CUDA(cudaEventElapsedTime(&t, cx.prof_ev[i], cx.prof_ev[i + 1]));
    flops[n] = cx.prof_fl[i / 2];
    ms[n] = t;
  }
  *count = n;
  return RT_OK;
}
