// Device-resident multi-env / multi-step (prioritized) sequence replay for B200 (sm_100a).
//
// Replaces rltime/history/{history,replay_history,prioritized_replay_history}.py and
// data_structures/segment_tree.py of opherlieber/rltime behind the C ABI in
// include/rltime_b200.h.  Layout in HBM (struct of arrays over storage *slots*):
//
//   state field f : uint8 [NS][bytes_f]     next_state leaves (frames 28,224 B, hx, cx, ...)
//   po field f    : uint8 [NS][bytes_f]     policy_output leaves (actions, qvalues)
//   reward f64[NS], done u8[NS]
//   pos2slot      : int32 [max_envs][N+1]   (env, env offset mod (N+1)) -> slot
//   sum/min tree  : f64 [2*cap]             root at 1, leaves at cap+idx (segment_tree.py)
//   seq_env/base  : prioritization idx -> (dense env, base env offset)
//
// NS = N + max_envs: a transition keeps its slot until its *successor* in the same env is
// evicted, because the state of offset k is the next_state of offset k-1
// (history.py:159-167) and the reference keeps that array alive by reference after k-1
// itself has left the buffer.
//
// What runs where:
//   host (this file, C++): O(1)-per-transition integer bookkeeping — global FIFO eviction
//     (replay_history.py:77-91), sequence activation / free-list (prioritized_replay_
//     history.py:136-172, 210-230), and the sequence priority (eta*max+(1-eta)*mean)^alpha
//     (:174-208).  The priority stays on the host on purpose: sampled indices must be
//     bit-identical to the reference, whose `**` is glibc pow(); CUDA's pow() is not
//     bit-equal to it.  numpy's pairwise summation order for np.mean is restated below.
//   device (kernels below): fp64 sum/min tree maintenance, stratified prefix-sum descent,
//     sequence -> slot resolution, n-step return/mask assembly, importance weights, loss
//     indices, and the byte gather of every state / policy-output leaf into the time-major
//     (S+n, B) batch.  No host synchronisation on the sample path.
#include <cmath>
#include <cstring>
#include <deque>
#include <unordered_map>
#include <vector>

#include "rt_common.h"

namespace {

// ------------------------------------------------------------------------------ kernels

// Batched leaf writes followed by bottom-up re-summation of the touched ancestors.
// Node values are a pure function of the leaves (parent = fl(left + right),
// segment_tree.py:87-97), so any schedule reproduces the reference's incremental updates
// bit for bit.  Single CTA; threads own updates; one __syncthreads per level.
__global__ void k_tree_set(double* __restrict__ sum_tree, double* __restrict__ min_tree,
                           int cap, int depth, const int* __restrict__ idx,
                           const double* __restrict__ sum_val,
                           const double* __restrict__ min_val, int m) {
  for (int base = 0; base < m; base += blockDim.x) {
    int j = base + threadIdx.x;
    bool act = j < m;
    int leaf = act ? cap + idx[j] : 0;
    if (act) {
      sum_tree[leaf] = sum_val[j];
      if (min_tree) min_tree[leaf] = min_val[j];
    }
    __syncthreads();
    for (int d = 1; d <= depth; ++d) {
      if (act) {
        int node = leaf >> d;
        sum_tree[node] = __dadd_rn(sum_tree[2 * node], sum_tree[2 * node + 1]);
        if (min_tree) min_tree[node] = fmin(min_tree[2 * node], min_tree[2 * node + 1]);
      }
      __syncthreads();
    }
  }
}

// find_prefixsum_idx (segment_tree.py:116-142): left iff node[2i] > mass else subtract.
__device__ __forceinline__ int tree_descend(const double* __restrict__ tree, int cap,
                                            double mass) {
  int i = 1;
  while (i < cap) {
    double left = tree[2 * i];
    if (left > mass) {
      i = 2 * i;
    } else {
      mass = __dsub_rn(mass, left);
      i = 2 * i + 1;
    }
  }
  return i - cap;
}

__global__ void k_tree_find(const double* __restrict__ tree, int cap,
                            const double* __restrict__ mass, int* __restrict__ out, int m) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) out[j] = tree_descend(tree, cap, mass[j]);
}

// Applies the host's bookkeeping deltas of one append() call.
__global__ void k_apply_updates(double* __restrict__ reward, uint8_t* __restrict__ done,
                                const int* __restrict__ u_slot,
                                const double* __restrict__ u_reward,
                                const uint8_t* __restrict__ u_done, int n_slot,
                                int* __restrict__ pos2slot, const long long* __restrict__ u_p2s_at,
                                const int* __restrict__ u_p2s_slot, int n_p2s,
                                int* __restrict__ seq_env, long long* __restrict__ seq_base,
                                const int* __restrict__ u_seq_idx, const int* __restrict__ u_seq_env,
                                const long long* __restrict__ u_seq_base, int n_seq,
                                long long* __restrict__ env_ids, const int* __restrict__ u_env,
                                const long long* __restrict__ u_env_id, int n_env) {
  int stride = gridDim.x * blockDim.x;
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = tid; j < n_slot; j += stride) {
    reward[u_slot[j]] = u_reward[j];
    done[u_slot[j]] = u_done[j];
  }
  for (int j = tid; j < n_p2s; j += stride) pos2slot[u_p2s_at[j]] = u_p2s_slot[j];
  for (int j = tid; j < n_seq; j += stride) {
    seq_env[u_seq_idx[j]] = u_seq_env[j];
    seq_base[u_seq_idx[j]] = u_seq_base[j];
  }
  for (int j = tid; j < n_env; j += stride) env_ids[u_env[j]] = u_env_id[j];
}

struct DrawParams {
  const double* sum_tree;
  const double* min_tree;
  int cap;
  int B;
  int P;
  double beta;
  double total_items;
  int global_scaling;
  const double* uniforms;
  const int* seq_env;
  const long long* seq_base;
  int* idxes;
  int* col_env;
  long long* col_start;   // base - P : env offset of row t = 0
  long long* col_base;
  double* col_weight;     // normalised importance weight per column
  double* weight_max;     // [1] the weight the batch was normalised by (sharded replay: rescale to the global max)
};

// _sample_proportional + per-sequence importance weight (prioritized_replay_history.py:
// 232-241, 327, 347-354).  One CTA; thread i owns stratum i.
__global__ void k_per_draw(DrawParams p) {
  extern __shared__ double s_w[];
  int i = threadIdx.x;
  double p_total = p.sum_tree[1];
  double w = 0.0;
  if (i < p.B) {
    double every = __ddiv_rn(p_total, (double)p.B);
    double mass = __dadd_rn(__dmul_rn(p.uniforms[i], every), __dmul_rn((double)i, every));
    int idx = tree_descend(p.sum_tree, p.cap, mass);
    p.idxes[i] = idx;
    int e = p.seq_env[idx];
    long long base = p.seq_base[idx];
    p.col_env[i] = e;
    p.col_base[i] = base;
    p.col_start[i] = base - p.P;
    double prob = __ddiv_rn(p.sum_tree[p.cap + idx], p_total);
    w = pow(__dmul_rn(prob, p.total_items), -p.beta);
  }
  s_w[i] = w;
  __syncthreads();
  double max_w;
  if (p.global_scaling) {
    double p_min = __ddiv_rn(p.min_tree[1], p_total);
    max_w = pow(__dmul_rn(p_min, p.total_items), -p.beta);
  } else {
    // np.max over the batch: exact, order independent
    max_w = 0.0;
    for (int j = 0; j < p.B; ++j) max_w = fmax(max_w, s_w[j]);
  }
  if (i < p.B) p.col_weight[i] = __ddiv_rn(w, max_w);
  if (i == 0) p.weight_max[0] = max_w;
}

struct AssembleParams {
  int B, S, n, P, T;
  long long N;
  int prioritized;
  const int* pos2slot;
  const int* col_env;
  const long long* col_start;
  const long long* col_base;
  const long long* col_count;  // transitions the env held at the draw (null: every n-step target exists)
  const double* col_weight;
  const long long* env_ids;
  const double* reward;
  const uint8_t* done;
  const double* gpow;       // gamma ** k, k < n (host libm pow == Python float pow)
  int* slots;               // (S+n) * B
  double* returns;
  long long* nsteps;
  double* target_masks;
  double* weights;
  long long* loss_indices;
};

__device__ __forceinline__ int slot_of(const int* __restrict__ pos2slot, long long N, int e,
                                       long long pos) {
  // The state of an env's very first transition is its own next_state (history.py:159-163).
  if (pos < 0) pos = 0;
  return pos2slot[(long long)e * N + (pos % N)];
}

// Rows j < S+n: slot of the *state* at env offset start + j, i.e. the next_state stored by
// offset start + j - 1.  Rows t < S additionally get the n-step return / mask
// (history.py:71-108), importance weight and loss index.
__global__ void k_assemble(AssembleParams p) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  int rows = (p.S + p.n) * p.B;
  if (r >= rows) return;
  int j = r / p.B, b = r - j * p.B;
  int e = p.col_env[b];
  long long start = p.col_start[b];
  // A window shifted to the very end of an env (avoid_episode_crossing, replay_history.py:167-168)
  // has fewer than n successors: _update_nstep stops at the end of the buffer (history.py:82),
  // so nstep < n and the target state is the last next_state that exists.
  const long long count = p.col_count ? p.col_count[b] : (1LL << 62);
  long long spos = start + j - 1;
  if (spos > count - 1) spos = count - 1;
  p.slots[r] = slot_of(p.pos2slot, p.N, e, spos);
  if (j >= p.S) return;
  long long k = start + j;  // env offset of this transition
  int own = slot_of(p.pos2slot, p.N, e, k);
  double ret = p.reward[own];
  bool mask = !p.done[own];
  int ns = 1;
  for (int q = 1; q < p.n && k + q < count; ++q, ++ns) {
    int s = slot_of(p.pos2slot, p.N, e, k + q);
    if (mask) ret = __dadd_rn(ret, __dmul_rn(p.gpow[q], p.reward[s]));
    if (p.done[s]) mask = false;
  }
  p.returns[r] = ret;
  p.nsteps[r] = ns;
  p.target_masks[r] = mask ? 1.0 : 0.0;
  if (p.prioritized) {
    p.weights[r] = p.col_weight[b];
    if (j < p.P) {
      p.loss_indices[2 * r] = -1;
      p.loss_indices[2 * r + 1] = -1;
    } else {
      p.loss_indices[2 * r] = p.env_ids[e];
      p.loss_indices[2 * r + 1] = p.col_base[b] + (j - p.P);
    }
  }
}

// Byte gather of all leaves into the time-major batch.  Work item = (field, row, chunk of
// CHUNK bytes); a CTA of 256 threads moves one chunk with 16-byte vector accesses when the
// leaf size allows (frames: 28,224 B = 1,764 x 16 B), streaming past L1.
#define RT_GATHER_THREADS 256
struct GatherField {
  const uint8_t* src;    // [NS][nb]
  uint8_t* dst;          // [rows][nb]
  const int* slots;      // row -> slot
  long long nb;          // bytes per item
  int rows;
  int row_lo, row_hi;    // rows in [row_lo, row_hi) are skipped (n >= S case)
  int chunks_per_row;
  long long first_item;  // prefix sum of work items
};
struct GatherParams {
  GatherField f[2 * RT_MAX_FIELDS];
  int num_fields;
  long long total_items;
  int chunk;
};

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w));
}

template <int VPT>  // 16-byte vectors per thread per chunk
__global__ void __launch_bounds__(RT_GATHER_THREADS) k_gather(const __grid_constant__ GatherParams p) {
  for (long long item = blockIdx.x; item < p.total_items; item += gridDim.x) {
    int fi = 0;
    while (fi + 1 < p.num_fields && item >= p.f[fi + 1].first_item) ++fi;
    const GatherField& f = p.f[fi];
    long long local = item - f.first_item;
    int row = (int)(local / f.chunks_per_row);
    int chunk = (int)(local - (long long)row * f.chunks_per_row);
    if (row >= f.row_lo && row < f.row_hi) continue;
    long long off = (long long)chunk * p.chunk;
    long long len = f.nb - off;
    if (len > p.chunk) len = p.chunk;
    const uint8_t* src = f.src + (long long)f.slots[row] * f.nb + off;
    uint8_t* dst = f.dst + (long long)row * f.nb + off;
    if ((f.nb & 15) == 0) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      int nv = (int)(len >> 4);
      uint4 v[VPT];
#pragma unroll
      for (int u = 0; u < VPT; ++u) {
        int i = threadIdx.x + u * RT_GATHER_THREADS;
        if (i < nv) v[u] = ld_stream(s4 + i);
      }
#pragma unroll
      for (int u = 0; u < VPT; ++u) {
        int i = threadIdx.x + u * RT_GATHER_THREADS;
        if (i < nv) st_stream(d4 + i, v[u]);
      }
    } else {
      for (int i = threadIdx.x; i < len; i += RT_GATHER_THREADS) dst[i] = src[i];
    }
  }
}


// ---- bulk-copy gather (TMA engine, no tensor map): global -> shared -> global
// Every field whose item size is a multiple of 16 bytes (frames 28,224 B, hx / cx 2,048 B) moves
// in chunks of <= GB_CHUNK bytes through a ring of GB_STAGES shared-memory stages: one
// cp.async.bulk per chunk into a stage (mbarrier complete_tx), one cp.async.bulk out of it
// (bulk_group).  Warp 0 drives the ring: its lanes resolve 32 chunk descriptors at a time in
// parallel (row -> slot is a dependent global load), lane 0 issues the copies; GB_LOOK loads stay in
// flight per CTA.  The other warps move the fields that are not 16-byte multiples (initials,
// actions, q-values: a few KB in total) with plain loads and stores.  Chunks are dealt round-robin
// to the CTAs, so with 4 chunks per frame the per-CTA load differs by one chunk at most.
#define GB_THREADS 128

__device__ __forceinline__ uint32_t gb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct BulkParams {
  GatherField f[2 * RT_MAX_FIELDS];  // bulk-eligible fields; chunks_per_row in units of the kernel's chunk size
  int num_fields;
  long long total_items;
  GatherField s[2 * RT_MAX_FIELDS];  // the other fields: one plain-copy item per (field, row)
  int num_small;
  long long small_items;
};

template <int GB_CHUNK, int GB_STAGES, int GB_LOOK>
__global__ void __launch_bounds__(GB_THREADS) k_gather_bulk(const __grid_constant__ BulkParams p) {
  extern __shared__ __align__(128) uint8_t gb_smem[];
  __shared__ __align__(8) uint64_t bars[GB_STAGES];
  __shared__ uint8_t* st_dst[GB_STAGES];
  __shared__ uint32_t st_bytes[GB_STAGES];
  __shared__ uint32_t st_armed[GB_STAGES];   // loads armed on this stage so far (phase parity of the next wait)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < GB_STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gb_smem_u32(&bars[s])), "r"(1));
      st_armed[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    // local items of this CTA: global item = blockIdx.x + local * gridDim.x
    const long long n = p.total_items > (long long)blockIdx.x
                            ? (p.total_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint8_t* d_src = nullptr;   // descriptor of local item (32-block base + lane)
    uint8_t* d_dst = nullptr;
    uint32_t d_bytes = 0;
    for (long long it = 0; it < n + GB_LOOK; ++it) {
      if (it < n) {
        if ((it & 31) == 0) {
          const long long li = it + lane;
          d_bytes = 0;
          if (li < n) {
            const long long item = blockIdx.x + li * (long long)gridDim.x;
            int fi = 0;
            while (fi + 1 < p.num_fields && item >= p.f[fi + 1].first_item) ++fi;
            const GatherField& f = p.f[fi];
            const long long local = item - f.first_item;
            const int row = (int)(local / f.chunks_per_row);
            const int chunk = (int)(local - (long long)row * f.chunks_per_row);
            if (!(row >= f.row_lo && row < f.row_hi)) {
              const long long off = (long long)chunk * GB_CHUNK;
              long long len = f.nb - off;
              if (len > GB_CHUNK) len = GB_CHUNK;
              d_src = f.src + (long long)f.slots[row] * f.nb + off;
              d_dst = f.dst + (long long)row * f.nb + off;
              d_bytes = (uint32_t)len;
            }
          }
        }
        const int sl = (int)(it & 31);
        const uint8_t* src = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)d_src, sl);
        uint8_t* dst = (uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)d_dst, sl);
        const uint32_t bytes = __shfl_sync(0xffffffffu, d_bytes, sl);
        if (lane == 0) {
          const int s = (int)(it % GB_STAGES);
          // the store that last read this stage was committed GB_STAGES - GB_LOOK iterations ago:
          // all but the newest GB_STAGES - GB_LOOK - 1 groups must have finished reading
          if (it >= GB_STAGES)
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(GB_STAGES - GB_LOOK - 1) : "memory");
          st_dst[s] = dst;
          st_bytes[s] = bytes;
          if (bytes) {
            st_armed[s]++;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gb_smem_u32(&bars[s])),
                         "r"(bytes)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    gb_smem_u32(gb_smem + (size_t)s * GB_CHUNK)),
                "l"(src), "r"(bytes), "r"(gb_smem_u32(&bars[s]))
                : "memory");
          }
        }
      }
      const long long j = it - GB_LOOK;
      if (lane == 0 && j >= 0 && j < n) {
        const int s = (int)(j % GB_STAGES);
        const uint32_t bytes = st_bytes[s];
        if (bytes) {
          // wait for the landed bytes: phase parity = number of earlier loads armed on this stage
          // (skipped rows never arm it)
          const uint32_t parity = (st_armed[s] - 1) & 1;
          uint32_t ok = 0;
          while (!ok) {
            asm volatile(
                "{\n\t.reg .pred q;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, q;\n\t}"
                : "=r"(ok)
                : "r"(gb_smem_u32(&bars[s])), "r"(parity)
                : "memory");
          }
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(st_dst[s]),
                       "r"(gb_smem_u32(gb_smem + (size_t)s * GB_CHUNK)), "r"(bytes)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // plain path for the small fields: one warp per (field, row) item, byte loop
    const long long w = (long long)blockIdx.x * (GB_THREADS / 32 - 1) + (warp - 1);
    const long long nw = (long long)gridDim.x * (GB_THREADS / 32 - 1);
    for (long long item = w; item < p.small_items; item += nw) {
      int fi = 0;
      while (fi + 1 < p.num_small && item >= p.s[fi + 1].first_item) ++fi;
      const GatherField& f = p.s[fi];
      const int row = (int)(item - f.first_item);
      if (row >= f.row_lo && row < f.row_hi) continue;
      const uint8_t* src = f.src + (long long)f.slots[row] * f.nb;
      uint8_t* dst = f.dst + (long long)row * f.nb;
      if ((f.nb & 3) == 0) {
        for (int i = lane; i < (int)(f.nb >> 2); i += 32)
          reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
      } else {
        for (int i = lane; i < (int)f.nb; i += 32) dst[i] = src[i];
      }
    }
  }
}

// ------------------------------------------------------------------- host-side numerics

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE),
// the order np.mean uses for a contiguous 1-D float64 array
// (prioritized_replay_history.py:199).
double np_pairwise_sum(const double* a, int64_t n) {
  if (n < 8) {
    double res = 0.;
    for (int64_t i = 0; i < n; i++) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int64_t i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
  } else {
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
  }
}

struct EnvState {
  bool seen = false;
  int64_t first = 0;   // env offset of the oldest live transition (_env_sample_offsets)
  int64_t count = 0;   // transitions ever appended == next env offset
  int64_t env_id = 0;
  int32_t zombie = -1; // slot of offset first-1, kept alive as the state of `first`
  std::vector<int32_t> ring;  // env offset -> slot for offsets in [first-1, count)
  std::vector<uint8_t> dring; // env offset -> done flag (host mirror for _refine_sample_range)
  int32_t get(int64_t pos) const { return ring[pos & (int64_t)(ring.size() - 1)]; }
  bool done_at(int64_t pos) const { return dring[pos & (int64_t)(dring.size() - 1)] != 0; }
  void put(int64_t pos, int32_t slot) {
    int64_t live = count - first + 2;
    if ((int64_t)ring.size() < live + 1) {
      size_t ncap = ring.empty() ? 64 : ring.size() * 2;
      while ((int64_t)ncap < live + 1) ncap *= 2;
      std::vector<int32_t> nr(ncap, -1);
      std::vector<uint8_t> nd(ncap, 0);
      if (!ring.empty())
        for (int64_t q = (first > 0 ? first - 1 : 0); q < pos; ++q) {
          nr[q & (int64_t)(ncap - 1)] = ring[q & (int64_t)(ring.size() - 1)];
          nd[q & (int64_t)(ncap - 1)] = dring[q & (int64_t)(dring.size() - 1)];
        }
      ring.swap(nr);
      dring.swap(nd);
    }
    ring[pos & (int64_t)(ring.size() - 1)] = slot;
  }
};

struct BatchSlot {
  int B = 0;
  void* all_states[RT_MAX_FIELDS] = {};
  void* po[RT_MAX_FIELDS] = {};
  double* returns = nullptr;
  long long* nsteps = nullptr;
  double* masks = nullptr;
  double* weights = nullptr;
  long long* loss_indices = nullptr;
  int* idxes = nullptr;
  int* slots = nullptr;
  int* col_env = nullptr;
  long long* col_start = nullptr;
  long long* col_count = nullptr;
  long long* col_base = nullptr;
  double* col_weight = nullptr;
  double* weight_max = nullptr;
  double* uniforms = nullptr;
  double* h_uniforms = nullptr;   // pinned staging (pageable H2D would serialise the stream)
  cudaEvent_t h2d_done = nullptr;
};

template <typename T>
struct DevVec {  // grow-only device scratch
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    cap = n + n / 2 + 64;
    return cudaMalloc(reinterpret_cast<void**>(&p), cap * sizeof(T));
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = reserve(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    // pageable source: the runtime stages it before returning, so `v` may be reused
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// One host->device upload per flush: every small delta array of a flush is packed into ONE pinned
// staging block and moved with ONE cudaMemcpyAsync (pageable sources make the runtime stage each
// copy and stall the issuing thread ~25 us per call).  Stages rotate; a stage is reused only after
// the event recorded behind its consumer kernels has completed.
struct Stage {
  uint8_t* host = nullptr;
  uint8_t* dev = nullptr;
  size_t cap = 0, used = 0;
  cudaEvent_t done = nullptr;
  cudaError_t begin(size_t bytes) {
    cudaError_t e;
    if (!done && (e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventSynchronize(done)) != cudaSuccess) return e;
    if (bytes > cap) {
      if (host) cudaFreeHost(host);
      if (dev) cudaFree(dev);
      host = dev = nullptr;
      cap = bytes + bytes / 2 + 4096;
      if ((e = cudaMallocHost(reinterpret_cast<void**>(&host), cap)) != cudaSuccess) return e;
      if ((e = cudaMalloc(reinterpret_cast<void**>(&dev), cap)) != cudaSuccess) return e;
    }
    used = 0;
    return cudaSuccess;
  }
  static size_t padded(size_t bytes) { return (bytes + 15) & ~(size_t)15; }
  template <typename T>
  const T* put(const std::vector<T>& v) {   // returns the DEVICE address the block will have
    size_t at = used;
    if (!v.empty()) memcpy(host + at, v.data(), v.size() * sizeof(T));
    used += padded(v.size() * sizeof(T));
    return reinterpret_cast<const T*>(dev + at);
  }
  cudaError_t send(cudaStream_t s) {
    if (!used) return cudaSuccess;
    return cudaMemcpyAsync(dev, host, used, cudaMemcpyHostToDevice, s);
  }
  cudaError_t fence(cudaStream_t s) { return cudaEventRecord(done, s); }
  void release() {
    if (host) cudaFreeHost(host);
    if (dev) cudaFree(dev);
    if (done) cudaEventDestroy(done);
    host = dev = nullptr;
    done = nullptr;
    cap = used = 0;
  }
};
#define RT_STAGES 4

}  // namespace

struct rt_tree {
  int cap = 0, depth = 0, device = 0;
  double* v = nullptr;
  DevVec<int> d_idx;
  DevVec<double> d_val;
  DevVec<double> d_mass;
  DevVec<int> d_out;
};

struct rt_replay {
  rt_replay_config cfg;
  double train_frequency = 0.0, train_quota = 0.0;   // replay_history.py:62-75,173-184 (0 = no quota)
  int64_t N = 0, NS = 0;
  int T = 0, P = 0, n = 0, S = 0, gap = 1;
  bool per = false;
  int target_capacity = 0, cap = 1, depth = 0;

  // host bookkeeping
  std::vector<EnvState> envs;
  std::vector<int32_t> env_order;       // first-appearance order (History.buffer dict order)
  std::vector<int32_t> fifo;            // linear_history: env of each live transition
  int64_t fifo_head = 0, fifo_len = 0;
  std::vector<int32_t> free_slots;      // FIFO ring of free storage slots
  int64_t fs_head = 0, fs_len = 0;
  std::vector<double> loss;             // per slot (sample['loss'])
  std::vector<int32_t> prio;            // per slot prioritization index or -1
  std::deque<int32_t> free_idx;         // _free_indexes
  std::vector<int32_t> seq_env;         // _index_data
  std::vector<int64_t> seq_base;
  int64_t active = 0;
  std::vector<double> gpow;

  // device storage
  uint8_t* d_state[RT_MAX_FIELDS] = {};
  uint8_t* d_po[RT_MAX_FIELDS] = {};
  double* d_reward = nullptr;
  uint8_t* d_done = nullptr;
  int* d_pos2slot = nullptr;
  double* d_sum = nullptr;
  double* d_min = nullptr;
  int* d_seq_env = nullptr;
  long long* d_seq_base = nullptr;
  long long* d_env_ids = nullptr;
  double* d_gpow = nullptr;

  BatchSlot batch[RT_BATCH_SLOTS];
  int cur = -1;
  int last_B = 0;                       // mbatch size of the last draw
  cudaEvent_t ev = nullptr;

  // pending device deltas (filled by bookkeeping, flushed once per append)
  std::vector<int> u_slot;
  std::vector<double> u_reward;
  std::vector<uint8_t> u_done;
  std::vector<long long> u_p2s_at;
  std::vector<int> u_p2s_slot;
  std::unordered_map<int, std::pair<int, long long>> u_seq;
  std::unordered_map<int, std::pair<double, double>> u_tree;
  std::vector<int> u_env;
  std::vector<long long> u_env_id;
  // packed pinned upload stages (one H2D copy per flush)
  Stage stage[RT_STAGES];
  int stage_cur = 0;
  int gather_bulk = 1;                // bulk-copy (TMA) gather; 0 = register-staged k_gather (RT_GATHER_BULK)
  int gather_cfg = 0, gather_cps = 3;  // ring geometry / CTAs per SM of the bulk gather (RT_GB_CFG, RT_GB_CPS)
  int num_sms = 148;
  // optional live timing of the gather kernel (bench.py roofline)
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev;   // pairs (start, stop)
  size_t prof_used = 0;
  // pinned read-back for update_losses_last
  float* h_td = nullptr;
  int* h_idx = nullptr;
  size_t h_td_cap = 0;
};

namespace {

int32_t pop_free_slot(rt_replay* h) {
  int32_t s = h->free_slots[h->fs_head];
  h->fs_head = (h->fs_head + 1) % h->NS;
  h->fs_len--;
  return s;
}
void push_free_slot(rt_replay* h, int32_t s) {
  h->free_slots[(h->fs_head + h->fs_len) % h->NS] = s;
  h->fs_len++;
}

void queue_tree(rt_replay* h, int idx, double sum_v, double min_v) {
  h->u_tree[idx] = std::make_pair(sum_v, min_v);
}

// _recalc_weighted_priority (prioritized_replay_history.py:174-208)
void recalc_priority(rt_replay* h, int idx) {
  const EnvState& es = h->envs[h->seq_env[idx]];
  int64_t base = h->seq_base[idx];
  double w;
  if (h->T == 1) {
    w = h->loss[es.get(base)];
  } else {
    double buf[1024];
    std::vector<double> big;
    double* a = buf;
    if (h->T > 1024) {
      big.resize(h->T);
      a = big.data();
    }
    double mx = -INFINITY;
    for (int t = 0; t < h->T; ++t) {
      a[t] = h->loss[es.get(base + t)];
      if (a[t] > mx) mx = a[t];
    }
    double mean = np_pairwise_sum(a, h->T) / (double)h->T;
    double eta = h->cfg.max_weight_factor;
    w = eta * mx + (1 - eta) * mean;
  }
  double pr = std::pow(w, h->cfg.alpha);  // glibc pow == Python float ** float
  queue_tree(h, idx, pr, pr);
}

// _sample_removed (prioritized_replay_history.py:210-230) for the oldest transition of e.
int evict_oldest(rt_replay* h, int e) {
  EnvState& es = h->envs[e];
  if (h->per) {
    int64_t k = es.first + h->P;
    if (k >= es.count)
      return rt::fail(RT_ERR_STATE,
                      "eviction from env %d with only %lld live transitions <= prefix_steps "
                      "(the reference indexes buffer[env][prefix_steps] here)",
                      e, (long long)(es.count - es.first));
    int32_t s = es.get(k);
    if (k % h->gap == 0 && h->prio[s] >= 0) {
      int idx = h->prio[s];
      h->prio[s] = -1;
      queue_tree(h, idx, 0.0, INFINITY);
      h->free_idx.push_back(idx);
      h->seq_env[idx] = -1;
      h->active--;
    }
  }
  // offset `first` leaves; its slot becomes the zombie that still backs the state of
  // first+1, and the previous zombie is finally recycled.
  if (es.zombie >= 0) push_free_slot(h, es.zombie);
  es.zombie = es.get(es.first);
  es.first++;
  return RT_OK;
}

int flush_updates(rt_replay* h, cudaStream_t st) {
  std::vector<int> seq_idx, seq_env;
  std::vector<long long> seq_base;
  for (auto& kv : h->u_seq) {
    seq_idx.push_back(kv.first);
    seq_env.push_back(kv.second.first);
    seq_base.push_back(kv.second.second);
  }
  std::vector<int> tidx;
  std::vector<double> tsum, tmin;
  for (auto& kv : h->u_tree) {
    tidx.push_back(kv.first);
    tsum.push_back(kv.second.first);
    tmin.push_back(kv.second.second);
  }
  int n_slot = (int)h->u_slot.size(), n_p2s = (int)h->u_p2s_at.size(),
      n_seq = (int)seq_idx.size(), n_env = (int)h->u_env.size(), n_tree = (int)tidx.size();
  const bool deltas = n_slot + n_p2s + n_seq + n_env > 0;
  if (deltas || n_tree > 0) {
    auto pb = [](size_t count, size_t elem) { return Stage::padded(count * elem); };
    size_t bytes = 0;
    if (deltas)
      bytes += pb(n_slot, 4) + pb(n_slot, 8) + pb(n_slot, 1) + pb(n_p2s, 8) + pb(n_p2s, 4) +
               pb(n_seq, 4) + pb(n_seq, 4) + pb(n_seq, 8) + pb(n_env, 4) + pb(n_env, 8);
    if (n_tree > 0) bytes += pb(n_tree, 4) + 2 * pb(n_tree, 8);
    Stage& sg = h->stage[h->stage_cur];
    h->stage_cur = (h->stage_cur + 1) % RT_STAGES;
    RT_CUDA(sg.begin(bytes));
    const int *d_slot = nullptr, *d_p2s_slot = nullptr, *d_seq_idx = nullptr, *d_seq_env = nullptr,
              *d_env = nullptr, *d_tidx = nullptr;
    const double *d_reward = nullptr, *d_tsum = nullptr, *d_tmin = nullptr;
    const uint8_t* d_done = nullptr;
    const long long *d_p2s_at = nullptr, *d_seq_base = nullptr, *d_env_id = nullptr;
    if (deltas) {
      d_slot = sg.put(h->u_slot);
      d_reward = sg.put(h->u_reward);
      d_done = sg.put(h->u_done);
      d_p2s_at = sg.put(h->u_p2s_at);
      d_p2s_slot = sg.put(h->u_p2s_slot);
      d_seq_idx = sg.put(seq_idx);
      d_seq_env = sg.put(seq_env);
      d_seq_base = sg.put(seq_base);
      d_env = sg.put(h->u_env);
      d_env_id = sg.put(h->u_env_id);
    }
    if (n_tree > 0) {
      d_tidx = sg.put(tidx);
      d_tsum = sg.put(tsum);
      d_tmin = sg.put(tmin);
    }
    RT_CUDA(sg.send(st));
    if (deltas) {
      int work = n_slot > n_p2s ? n_slot : n_p2s;
      int blocks = (work + 255) / 256;
      if (blocks < 1) blocks = 1;
      if (blocks > 296) blocks = 296;
      k_apply_updates<<<blocks, 256, 0, st>>>(
          h->d_reward, h->d_done, d_slot, d_reward, d_done, n_slot, h->d_pos2slot, d_p2s_at,
          d_p2s_slot, n_p2s, h->d_seq_env, h->d_seq_base, d_seq_idx, d_seq_env, d_seq_base, n_seq,
          h->d_env_ids, d_env, d_env_id, n_env);
      RT_LAUNCH_CHECK();
    }
    if (n_tree > 0) {
      int threads = n_tree < 1024 ? ((n_tree + 31) / 32) * 32 : 1024;
      k_tree_set<<<1, threads, 0, st>>>(h->d_sum, h->d_min, h->cap, h->depth, d_tidx, d_tsum, d_tmin,
                                        n_tree);
      RT_LAUNCH_CHECK();
    }
    RT_CUDA(sg.fence(st));
  }
  h->u_slot.clear();
  h->u_reward.clear();
  h->u_done.clear();
  h->u_p2s_at.clear();
  h->u_p2s_slot.clear();
  h->u_seq.clear();
  h->u_tree.clear();
  h->u_env.clear();
  h->u_env_id.clear();
  return RT_OK;
}

int ensure_batch(rt_replay* h, BatchSlot& bs, int B) {
  if (bs.B >= B) return RT_OK;
  auto fr = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  for (int f = 0; f < RT_MAX_FIELDS; ++f) {
    fr(bs.all_states[f]);
    fr(bs.po[f]);
  }
  fr(bs.returns); fr(bs.nsteps); fr(bs.masks); fr(bs.weights); fr(bs.loss_indices);
  fr(bs.idxes); fr(bs.slots); fr(bs.col_env); fr(bs.col_start); fr(bs.col_count); fr(bs.col_base);
  fr(bs.col_weight); fr(bs.weight_max); fr(bs.uniforms);
  if (bs.h_uniforms) cudaFreeHost(bs.h_uniforms);
  bs.h_uniforms = nullptr;
  RT_CUDA(cudaMallocHost(&bs.h_uniforms, (size_t)B * sizeof(double)));
  if (!bs.h2d_done) RT_CUDA(cudaEventCreateWithFlags(&bs.h2d_done, cudaEventDisableTiming));
  size_t rows_all = (size_t)(h->S + h->n) * B, rows = (size_t)h->S * B;
  for (int f = 0; f < h->cfg.num_state_fields; ++f)
    RT_CUDA(cudaMalloc(&bs.all_states[f], rows_all * h->cfg.state_field_bytes[f]));
  for (int f = 0; f < h->cfg.num_po_fields; ++f)
    RT_CUDA(cudaMalloc(&bs.po[f], rows * h->cfg.po_field_bytes[f]));
  RT_CUDA(rt::dmalloc(&bs.returns, rows));
  RT_CUDA(rt::dmalloc(&bs.nsteps, rows));
  RT_CUDA(rt::dmalloc(&bs.masks, rows));
  RT_CUDA(rt::dmalloc(&bs.weights, rows));
  RT_CUDA(rt::dmalloc(&bs.loss_indices, rows * 2));
  RT_CUDA(rt::dmalloc(&bs.idxes, (size_t)B));
  RT_CUDA(rt::dmalloc(&bs.slots, rows_all));
  RT_CUDA(rt::dmalloc(&bs.col_env, (size_t)B));
  RT_CUDA(rt::dmalloc(&bs.col_start, (size_t)B));
  RT_CUDA(rt::dmalloc(&bs.col_count, (size_t)B));
  RT_CUDA(rt::dmalloc(&bs.col_base, (size_t)B));
  RT_CUDA(rt::dmalloc(&bs.col_weight, (size_t)B));
  RT_CUDA(rt::dmalloc(&bs.weight_max, (size_t)1));
  RT_CUDA(rt::dmalloc(&bs.uniforms, (size_t)B));
  bs.B = B;
  return RT_OK;
}

// Assemble + gather for the columns already written to bs.col_*.
int assemble_and_gather(rt_replay* h, BatchSlot& bs, int B, cudaStream_t st) {
  AssembleParams ap;
  ap.B = B; ap.S = h->S; ap.n = h->n; ap.P = h->P; ap.T = h->T; ap.N = h->N + 1;   // pos2slot ring length
  ap.prioritized = h->per ? 1 : 0;
  ap.pos2slot = h->d_pos2slot; ap.col_env = bs.col_env; ap.col_start = bs.col_start;
  ap.col_base = bs.col_base; ap.col_weight = bs.col_weight; ap.env_ids = h->d_env_ids;
  ap.col_count = h->per ? nullptr : bs.col_count;
  ap.reward = h->d_reward; ap.done = h->d_done; ap.gpow = h->d_gpow; ap.slots = bs.slots;
  ap.returns = bs.returns; ap.nsteps = bs.nsteps; ap.target_masks = bs.masks;
  ap.weights = bs.weights; ap.loss_indices = bs.loss_indices;
  int rows_all = (h->S + h->n) * B;
  k_assemble<<<(rows_all + 127) / 128, 128, 0, st>>>(ap);
  RT_LAUNCH_CHECK();

  GatherParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.chunk = 4 * RT_GATHER_THREADS * 16;  // 16 KiB per work item
  long long items = 0;
  int nf = 0;
  for (int f = 0; f < h->cfg.num_state_fields; ++f) {
    GatherField& g = gp.f[nf++];
    g.src = h->d_state[f]; g.dst = (uint8_t*)bs.all_states[f]; g.slots = bs.slots;
    g.nb = h->cfg.state_field_bytes[f]; g.rows = rows_all;
    // rows in [S, n) belong to neither view when n > S (history.py:266-270 path)
    g.row_lo = h->S * B; g.row_hi = (h->n > h->S ? h->n : h->S) * B;
    g.chunks_per_row = (int)((g.nb + gp.chunk - 1) / gp.chunk);
    g.first_item = items;
    items += (long long)g.rows * g.chunks_per_row;
  }
  for (int f = 0; f < h->cfg.num_po_fields; ++f) {
    GatherField& g = gp.f[nf++];
    g.src = h->d_po[f]; g.dst = (uint8_t*)bs.po[f];
    g.slots = bs.slots + B;  // own slot of row t == state slot of row t+1
    g.nb = h->cfg.po_field_bytes[f]; g.rows = h->S * B; g.row_lo = g.row_hi = 0;
    g.chunks_per_row = (int)((g.nb + gp.chunk - 1) / gp.chunk);
    g.first_item = items;
    items += (long long)g.rows * g.chunks_per_row;
  }
  gp.num_fields = nf;
  gp.total_items = items;
  if (items > 0) {
    bool timed = h->profile && h->prof_used + 2 <= h->prof_ev.size();
    if (timed) RT_CUDA(cudaEventRecord(h->prof_ev[h->prof_used], st));
    if (h->gather_bulk) {
      // split the fields: 16-byte multiples go through the bulk-copy ring, the rest the plain path
      static const int cfgs[4][3] = {{8192, 8, 6}, {16384, 6, 4}, {4096, 16, 12}, {32768, 4, 2}};
      const int* cf = cfgs[h->gather_cfg & 3];
      const int chunk_bytes = cf[0];
      BulkParams bp;
      memset(&bp, 0, sizeof(bp));
      long long bi = 0, si = 0;
      for (int f = 0; f < nf; ++f) {
        const GatherField& g = gp.f[f];
        if ((g.nb & 15) == 0 && g.nb >= 256) {
          GatherField& d = bp.f[bp.num_fields++];
          d = g;
          d.chunks_per_row = (int)((g.nb + chunk_bytes - 1) / chunk_bytes);
          d.first_item = bi;
          bi += (long long)d.rows * d.chunks_per_row;
        } else {
          GatherField& d = bp.s[bp.num_small++];
          d = g;
          d.chunks_per_row = 1;
          d.first_item = si;
          si += d.rows;
        }
      }
      bp.total_items = bi;
      bp.small_items = si;
      long long grid = (long long)h->num_sms * h->gather_cps;
      long long want = bi > si ? bi : si;
      if (want < grid) grid = want < 1 ? 1 : want;
      const int smem = cf[0] * cf[1];
#define RT_GB_LAUNCH(C, S, L)                                                                         \
  {                                                                                                   \
    static bool configured = false;                                                                   \
    if (!configured) {                                                                                \
      RT_CUDA(cudaFuncSetAttribute(k_gather_bulk<C, S, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   smem));                                                            \
      configured = true;                                                                              \
    }                                                                                                 \
    k_gather_bulk<C, S, L><<<(int)grid, GB_THREADS, smem, st>>>(bp);                                  \
  }
      switch (h->gather_cfg & 3) {
        case 0: RT_GB_LAUNCH(8192, 8, 6) break;
        case 1: RT_GB_LAUNCH(16384, 6, 4) break;
        case 2: RT_GB_LAUNCH(4096, 16, 12) break;
        default: RT_GB_LAUNCH(32768, 4, 2) break;
      }
#undef RT_GB_LAUNCH
    } else {
      long long grid = items < 148LL * 16 ? items : 148LL * 16;
      k_gather<4><<<(int)grid, RT_GATHER_THREADS, 0, st>>>(gp);
    }
    RT_LAUNCH_CHECK();
    if (timed) {
      RT_CUDA(cudaEventRecord(h->prof_ev[h->prof_used + 1], st));
      h->prof_used += 2;
    }
  }
  return RT_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

const char* rt_last_error(void) { return rt::last_error().c_str(); }
int rt_version(void) { return 1; }
int64_t rt_launch_count(void) { return rt::launch_counter().load(); }

int rt_replay_create(const rt_replay_config* c, rt_replay** out) {
  RT_REQUIRE(c && out, "null argument");
  RT_REQUIRE(c->size > 0 && c->nstep_train >= 1 && c->nstep_target >= 1 && c->prefix_steps >= 0,
             "bad size/nstep arguments");
  RT_REQUIRE(c->max_envs >= 1, "max_envs must be >= 1");
  RT_REQUIRE(c->num_state_fields >= 1 && c->num_state_fields <= RT_MAX_FIELDS &&
                 c->num_po_fields >= 0 && c->num_po_fields <= RT_MAX_FIELDS,
             "field counts out of range");
  RT_REQUIRE(c->kind == RT_KIND_UNIFORM || c->kind == RT_KIND_PRIORITIZED, "bad kind");
  RT_CUDA(cudaSetDevice(c->device));
  rt_replay* h = new rt_replay();
  h->cfg = *c;
  {
    cudaDeviceProp prop;
    RT_CUDA(cudaGetDeviceProperties(&prop, c->device));
    h->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("RT_GATHER_BULK")) h->gather_bulk = atoi(e);
    if (const char* e = getenv("RT_GB_CFG")) h->gather_cfg = atoi(e);
    if (const char* e = getenv("RT_GB_CPS")) h->gather_cps = atoi(e) > 0 ? atoi(e) : 1;
  }
  h->N = c->size;
  h->NS = c->size + c->max_envs;
  h->T = c->nstep_train; h->P = c->prefix_steps; h->n = c->nstep_target;
  h->S = h->T + h->P;
  h->per = c->kind == RT_KIND_PRIORITIZED;
  RT_REQUIRE(h->NS < (1LL << 31), "size too large for int32 slots");
  if (h->per) {
    RT_REQUIRE(c->overlap >= 0 && c->overlap < h->T, "overlap must be in [0, nstep_train)");
    h->gap = h->T - c->overlap;
    h->target_capacity = (int)(h->N / h->gap);
    RT_REQUIRE(h->target_capacity >= 1, "size smaller than one sequence gap");
    while (h->cap < h->target_capacity) h->cap *= 2;
    while ((1 << h->depth) < h->cap) h->depth++;
    for (int i = 0; i < h->target_capacity; ++i) h->free_idx.push_back(i);
    h->seq_env.assign(h->target_capacity, -1);
    h->seq_base.assign(h->target_capacity, 0);
  }
  h->envs.resize(c->max_envs);
  h->fifo.assign(h->N, -1);
  h->free_slots.resize(h->NS);
  for (int64_t i = 0; i < h->NS; ++i) h->free_slots[i] = (int32_t)i;
  h->fs_len = h->NS;
  h->loss.assign(h->NS, 0.0);
  h->prio.assign(h->NS, -1);
  for (int k = 0; k < h->n; ++k) h->gpow.push_back(std::pow(c->gamma, (double)k));

  for (int f = 0; f < c->num_state_fields; ++f) {
    RT_REQUIRE(c->state_field_bytes[f] > 0, "state field %d has no bytes", f);
    RT_CUDA(cudaMalloc(&h->d_state[f], (size_t)h->NS * c->state_field_bytes[f]));
  }
  for (int f = 0; f < c->num_po_fields; ++f) {
    RT_REQUIRE(c->po_field_bytes[f] > 0, "policy_output field %d has no bytes", f);
    RT_CUDA(cudaMalloc(&h->d_po[f], (size_t)h->NS * c->po_field_bytes[f]));
  }
  RT_CUDA(rt::dmalloc(&h->d_reward, (size_t)h->NS));
  RT_CUDA(rt::dmalloc(&h->d_done, (size_t)h->NS));
  // N + 1 entries per env: an env that alone holds all N transitions still needs the entry of the evicted
  // "zombie" at offset first-1 (it backs the state of `first`), i.e. N + 1 live positions
  RT_CUDA(rt::dmalloc(&h->d_pos2slot, (size_t)c->max_envs * (h->N + 1)));
  RT_CUDA(rt::dmalloc(&h->d_env_ids, (size_t)c->max_envs));
  RT_CUDA(rt::dmalloc(&h->d_gpow, (size_t)h->n));
  RT_CUDA(cudaMemcpy(h->d_gpow, h->gpow.data(), h->n * sizeof(double), cudaMemcpyHostToDevice));
  if (h->per) {
    RT_CUDA(rt::dmalloc(&h->d_sum, (size_t)2 * h->cap));
    RT_CUDA(cudaMemset(h->d_sum, 0, (size_t)2 * h->cap * sizeof(double)));
    if (c->global_importance_scaling) {
      RT_CUDA(rt::dmalloc(&h->d_min, (size_t)2 * h->cap));
      std::vector<double> inf((size_t)2 * h->cap, INFINITY);
      RT_CUDA(cudaMemcpy(h->d_min, inf.data(), inf.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    RT_CUDA(rt::dmalloc(&h->d_seq_env, (size_t)h->target_capacity));
    RT_CUDA(rt::dmalloc(&h->d_seq_base, (size_t)h->target_capacity));
  }
  RT_CUDA(cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming));
  *out = h;
  return RT_OK;
}

void rt_replay_destroy(rt_replay* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (int f = 0; f < RT_MAX_FIELDS; ++f) {
    if (h->d_state[f]) cudaFree(h->d_state[f]);
    if (h->d_po[f]) cudaFree(h->d_po[f]);
  }
  void* ptrs[] = {h->d_reward, h->d_done, h->d_pos2slot, h->d_sum, h->d_min, h->d_seq_env,
                  h->d_seq_base, h->d_env_ids, h->d_gpow};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto& bs : h->batch) {
    for (int f = 0; f < RT_MAX_FIELDS; ++f) {
      if (bs.all_states[f]) cudaFree(bs.all_states[f]);
      if (bs.po[f]) cudaFree(bs.po[f]);
    }
    void* bp[] = {bs.returns, bs.nsteps, bs.masks, bs.weights, bs.loss_indices, bs.idxes,
                  bs.slots, bs.col_env, bs.col_start, bs.col_count, bs.col_base, bs.col_weight, bs.uniforms};
    for (void* p : bp)
      if (p) cudaFree(p);
    if (bs.h_uniforms) cudaFreeHost(bs.h_uniforms);
    if (bs.h2d_done) cudaEventDestroy(bs.h2d_done);
  }
  for (auto& sg : h->stage) sg.release();
  if (h->h_td) cudaFreeHost(h->h_td);
  if (h->h_idx) cudaFreeHost(h->h_idx);
  if (h->ev) cudaEventDestroy(h->ev);
  for (auto& e : h->prof_ev) cudaEventDestroy(e);
  delete h;
}

int rt_replay_append(rt_replay* h, int64_t m, const int32_t* env, const int64_t* env_ids,
                     const double* reward, const uint8_t* done,
                     const void* const* state_fields, const void* const* po_fields,
                     int32_t fields_on_device, void* stream) {
  RT_REQUIRE(h && env && reward && done && state_fields, "null argument");
  RT_REQUIRE(m >= 0 && m <= h->N, "append of %lld transitions exceeds capacity %lld per call",
             (long long)m, (long long)h->N);
  RT_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->train_frequency > 0) h->train_quota += h->train_frequency * (double)m;   // replay_history.py:91
  std::vector<int32_t> slots((size_t)m);
  for (int64_t i = 0; i < m; ++i) {
    int e = env[i];
    RT_REQUIRE(e >= 0 && e < h->cfg.max_envs, "env index %d outside [0, max_envs=%d)", e,
               h->cfg.max_envs);
    EnvState& es = h->envs[e];
    if (!es.seen) {
      es.seen = true;
      es.env_id = env_ids ? env_ids[i] : e;
      h->env_order.push_back(e);
      h->u_env.push_back(e);
      h->u_env_id.push_back(es.env_id);
    }
    int64_t pos = es.count;
    // History.update appends to the env buffer first (history.py:169-171) ...
    es.count++;
    // ... then _sample_added evicts the globally oldest transition when full
    // (replay_history.py:79-87)
    if (h->fifo_len >= h->N) {
      int victim = h->fifo[h->fifo_head];
      h->fifo_head = (h->fifo_head + 1) % h->N;
      h->fifo_len--;
      int rc = evict_oldest(h, victim);
      if (rc != RT_OK) return rc;
    }
    h->fifo[(h->fifo_head + h->fifo_len) % h->N] = e;
    h->fifo_len++;
    if (h->fs_len <= 0) return rt::fail(RT_ERR_STATE, "slot pool exhausted");
    int32_t s = pop_free_slot(h);
    slots[i] = s;
    es.put(pos, s);
    es.dring[pos & (int64_t)(es.dring.size() - 1)] = done[i] ? 1 : 0;
    h->loss[s] = 1.0;  // _max_loss, never updated by the reference (:120,141)
    h->prio[s] = -1;
    h->u_slot.push_back(s);
    h->u_reward.push_back(reward[i]);
    h->u_done.push_back(done[i] ? 1 : 0);
    h->u_p2s_at.push_back((long long)e * (h->N + 1) + (pos % (h->N + 1)));
    h->u_p2s_slot.push_back(s);
    if (h->per) {
      // activation of a new overlapped sequence (:152-172)
      int64_t base = pos - h->T + 1 - h->n + 1;
      if (base >= 0 && base % h->gap == 0 && base >= es.first + h->P) {
        if (h->free_idx.empty()) return rt::fail(RT_ERR_STATE, "no free prioritization index");
        int idx = h->free_idx.front();
        h->free_idx.pop_front();
        h->prio[es.get(base)] = idx;
        h->seq_env[idx] = e;
        h->seq_base[idx] = base;
        h->u_seq[idx] = std::make_pair(e, (long long)base);
        h->active++;
        recalc_priority(h, idx);
      }
    }
  }
  // payload: runs of consecutive slots move with one copy each
  cudaMemcpyKind kind = fields_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (int64_t i = 0; i < m;) {
    int64_t j = i + 1;
    while (j < m && slots[j] == slots[j - 1] + 1) ++j;
    for (int f = 0; f < h->cfg.num_state_fields; ++f) {
      int64_t nb = h->cfg.state_field_bytes[f];
      RT_CUDA(cudaMemcpyAsync(h->d_state[f] + (size_t)slots[i] * nb,
                              (const uint8_t*)state_fields[f] + (size_t)i * nb,
                              (size_t)(j - i) * nb, kind, st));
    }
    for (int f = 0; f < h->cfg.num_po_fields; ++f) {
      int64_t nb = h->cfg.po_field_bytes[f];
      RT_CUDA(cudaMemcpyAsync(h->d_po[f] + (size_t)slots[i] * nb,
                              (const uint8_t*)po_fields[f] + (size_t)i * nb,
                              (size_t)(j - i) * nb, kind, st));
    }
    i = j;
  }
  return flush_updates(h, st);
}

int rt_replay_set_train_frequency(rt_replay* h, double train_frequency) {
  RT_REQUIRE(h, "null argument");
  h->train_frequency = train_frequency > 0 ? train_frequency : 0.0;
  return RT_OK;
}

int64_t rt_replay_needed_feed(const rt_replay* h, int32_t mbatch, int32_t num_envs) {
  (void)mbatch;
  if (!h || h->train_frequency <= 0) return 0;            // take whatever is ready
  if (h->train_quota > 0) return -1;                       // None: do not act, train first
  int64_t want = (int64_t)(-h->train_quota / h->train_frequency);
  return want > num_envs ? want : num_envs;
}

int rt_replay_consume_quota(rt_replay* h, int32_t mbatch) {
  RT_REQUIRE(h && mbatch >= 1, "bad argument");
  if (h->train_frequency <= 0) return RT_OK;
  const double step = (double)mbatch * h->T;
  h->train_quota -= step;
  if (!(h->train_quota < 100.0 * step && h->train_quota > -100.0 * step))
    return rt::fail(RT_ERR_STATE, "train quota %.1f outside +-100 x mbatch x nstep_train (replay_history.py:179-181)",
                    h->train_quota);
  return RT_OK;
}

double rt_replay_train_quota(const rt_replay* h) { return h ? h->train_quota : 0.0; }

int64_t rt_replay_len(const rt_replay* h) { return h ? h->fifo_len : 0; }
int64_t rt_replay_active_sequences(const rt_replay* h) { return h ? h->active : 0; }

int64_t rt_replay_uniform_available(rt_replay* h) {
  if (!h) return 0;
  int64_t total = 0;
  for (int e : h->env_order) {
    const EnvState& es = h->envs[e];
    int64_t a = (es.count - es.first) - (h->S + h->n - 1);
    if (a > 0) total += a;
  }
  return total;
}

int rt_replay_sample_prioritized(rt_replay* h, int32_t B, double beta, const double* uniforms,
                                 void* stream) {
  RT_REQUIRE(h && uniforms, "null argument");
  RT_REQUIRE(h->per, "not a prioritized buffer");
  RT_REQUIRE(B >= 1 && B <= 1024, "mbatch_size must be in [1, 1024]");
  RT_CUDA(cudaSetDevice(h->cfg.device));
  // The reference draws first and only then checks availability (:284 vs :295-299); the
  // caller has already consumed its B uniforms either way.
  if (h->active < B) {
    if (h->fifo_len >= h->N)
      return rt::fail(RT_ERR_STATE, "buffer full but fewer than mbatch_size sequences active");
    return RT_NEED_MORE_DATA;
  }
  cudaStream_t st = (cudaStream_t)stream;
  h->cur = (h->cur + 1) % RT_BATCH_SLOTS;
  BatchSlot& bs = h->batch[h->cur];
  int rc = ensure_batch(h, bs, B);
  if (rc != RT_OK) return rc;
  RT_CUDA(cudaEventSynchronize(bs.h2d_done));  // previous use of the pinned staging
  memcpy(bs.h_uniforms, uniforms, B * sizeof(double));
  RT_CUDA(cudaMemcpyAsync(bs.uniforms, bs.h_uniforms, B * sizeof(double), cudaMemcpyHostToDevice, st));
  RT_CUDA(cudaEventRecord(bs.h2d_done, st));
  DrawParams dp;
  dp.sum_tree = h->d_sum; dp.min_tree = h->d_min; dp.cap = h->cap; dp.B = B; dp.P = h->P;
  dp.beta = beta; dp.total_items = (double)h->active;
  dp.global_scaling = h->cfg.global_importance_scaling ? 1 : 0;
  dp.uniforms = bs.uniforms; dp.seq_env = h->d_seq_env; dp.seq_base = h->d_seq_base;
  dp.idxes = bs.idxes; dp.col_env = bs.col_env; dp.col_start = bs.col_start;
  dp.col_base = bs.col_base; dp.col_weight = bs.col_weight; dp.weight_max = bs.weight_max;
  int threads = ((B + 31) / 32) * 32;
  k_per_draw<<<1, threads, threads * sizeof(double), st>>>(dp);
  RT_LAUNCH_CHECK();
  h->last_B = B;
  return assemble_and_gather(h, bs, B, st);
}

int rt_replay_sample_uniform(rt_replay* h, int32_t B, const int64_t* choices, void* stream) {
  RT_REQUIRE(h && choices, "null argument");
  RT_REQUIRE(B >= 1 && B <= 1024, "mbatch_size must be in [1, 1024]");
  RT_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int> col_env(B);
  std::vector<long long> col_start(B), col_count(B);
  for (int b = 0; b < B; ++b) {
    int64_t c = choices[b];
    bool found = false;
    for (int e : h->env_order) {
      const EnvState& es = h->envs[e];
      int64_t a = (es.count - es.first) - (h->S + h->n - 1);
      if (a <= 0) continue;
      if (c < a) {
        if (h->cfg.avoid_episode_crossing) {
          // _refine_sample_range (replay_history.py:142-171): a `done` inside the window (its last
          // step excepted) shifts the window to the end of that episode or the start of the next
          const int64_t amount = h->S, len = es.count - es.first;
          for (int64_t i = 0; i < amount - 1; ++i) {
            if (es.done_at(es.first + c + i)) {
              if ((double)i < (double)amount / 2.0) c = std::max<int64_t>(c - (amount - i - 1), 0);
              else c = std::min<int64_t>(c + i + 1, len - amount);
              break;
            }
          }
        }
        col_env[b] = e;
        col_start[b] = es.first + c;
        col_count[b] = es.count;
        found = true;
        break;
      }
      c -= a;
    }
    RT_REQUIRE(found, "choice %lld outside the available range", (long long)choices[b]);
  }
  h->cur = (h->cur + 1) % RT_BATCH_SLOTS;
  BatchSlot& bs = h->batch[h->cur];
  int rc = ensure_batch(h, bs, B);
  if (rc != RT_OK) return rc;
  RT_CUDA(cudaMemcpyAsync(bs.col_env, col_env.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
  RT_CUDA(cudaMemcpyAsync(bs.col_start, col_start.data(), B * sizeof(long long),
                          cudaMemcpyHostToDevice, st));
  RT_CUDA(cudaMemcpyAsync(bs.col_count, col_count.data(), B * sizeof(long long),
                          cudaMemcpyHostToDevice, st));
  h->last_B = B;
  return assemble_and_gather(h, bs, B, st);
}

int rt_replay_batch(rt_replay* h, rt_batch* out) {
  RT_REQUIRE(h && out, "null argument");
  RT_REQUIRE(h->cur >= 0 && h->last_B > 0, "no batch drawn yet");
  BatchSlot& bs = h->batch[h->cur];
  memset(out, 0, sizeof(*out));
  out->B = h->last_B;
  out->S = h->S;
  out->n = h->n;
  out->num_state_fields = h->cfg.num_state_fields;
  out->num_po_fields = h->cfg.num_po_fields;
  for (int f = 0; f < RT_MAX_FIELDS; ++f) {
    out->all_states[f] = bs.all_states[f];
    out->policy_outputs[f] = bs.po[f];
  }
  out->returns = bs.returns;
  out->nsteps = (int64_t*)bs.nsteps;
  out->target_masks = bs.masks;
  out->importance_weights = h->per ? bs.weights : nullptr;
  out->loss_indices = h->per ? (int64_t*)bs.loss_indices : nullptr;
  out->idxes = h->per ? bs.idxes : nullptr;
  out->slots = bs.slots;
  out->weight_max = h->per ? bs.weight_max : nullptr;
  return RT_OK;
}

int rt_replay_update_losses(rt_replay* h, int64_t m, const int64_t* pairs, const double* losses,
                            void* stream) {
  RT_REQUIRE(h && pairs && losses, "null argument");
  RT_REQUIRE(h->per, "not a prioritized buffer");
  RT_CUDA(cudaSetDevice(h->cfg.device));
  std::vector<int> affected;
  std::vector<uint8_t> mark;  // lazily sized
  mark.assign(h->target_capacity, 0);
  for (int64_t i = 0; i < m; ++i) {
    int64_t e = pairs[2 * i], offset = pairs[2 * i + 1];
    RT_REQUIRE(e >= 0 && e < h->cfg.max_envs && h->envs[e].seen,
               "update_losses: unknown env index %lld (prefix rows (-1,-1) must not be sent back)",
               (long long)e);
    EnvState& es = h->envs[e];
    if (offset < es.first) continue;  // evicted since the batch was drawn (:254-257)
    RT_REQUIRE(offset < es.count, "update_losses: offset %lld beyond env head", (long long)offset);
    h->loss[es.get(offset)] = std::fabs(losses[i]) + h->cfg.eps;
    int64_t base = offset - (offset % h->gap);
    while (base + h->T > offset && base >= es.first) {
      int idx = h->prio[es.get(base)];
      if (idx >= 0 && !mark[idx]) {
        mark[idx] = 1;
        affected.push_back(idx);
      }
      base -= h->gap;
    }
  }
  for (int idx : affected) recalc_priority(h, idx);
  return flush_updates(h, (cudaStream_t)stream);
}

int rt_replay_update_losses_last(rt_replay* h, const float* td_abs_device, void* stream) {
  RT_REQUIRE(h && td_abs_device, "null argument");
  RT_REQUIRE(h->per && h->cur >= 0 && h->last_B > 0, "no prioritized batch drawn");
  RT_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  int B = h->last_B;
  size_t rows = (size_t)h->T * B;
  if (h->h_td_cap < rows) {
    if (h->h_td) cudaFreeHost(h->h_td);
    if (h->h_idx) cudaFreeHost(h->h_idx);
    RT_CUDA(cudaMallocHost(&h->h_td, rows * sizeof(float)));
    RT_CUDA(cudaMallocHost(&h->h_idx, 3 * 1024 * sizeof(long long)));
    h->h_td_cap = rows;
  }
  RT_REQUIRE(B <= 1024, "batch of %d sequences exceeds the write-back staging", B);
  BatchSlot& bs = h->batch[h->cur];
  // (env, base offset) of every drawn sequence AS RECORDED BY THE DRAW (k_per_draw writes col_env /
  // col_base into the batch slot): the reference writes back through the loss_indices captured at draw
  // time (prioritized_replay_history.py:335-338) and skips rows evicted since (:254-257).  Resolving
  // the drawn prioritization indices through seq_env / seq_base now instead would go wrong when an
  // append between the draw and this call freed or re-assigned one of them.
  long long* h_base = reinterpret_cast<long long*>(h->h_idx);
  int* h_env = reinterpret_cast<int*>(h_base + 1024);
  RT_CUDA(cudaMemcpyAsync(h->h_td, td_abs_device, rows * sizeof(float), cudaMemcpyDeviceToHost, st));
  RT_CUDA(cudaMemcpyAsync(h_base, bs.col_base, B * sizeof(long long), cudaMemcpyDeviceToHost, st));
  RT_CUDA(cudaMemcpyAsync(h_env, bs.col_env, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  RT_CUDA(cudaEventRecord(h->ev, st));
  RT_CUDA(cudaEventSynchronize(h->ev));
  std::vector<int64_t> pairs(rows * 2);
  std::vector<double> losses(rows);
  for (int t = 0; t < h->T; ++t)
    for (int b = 0; b < B; ++b) {
      size_t r = (size_t)t * B + b;
      pairs[2 * r] = h_env[b];
      pairs[2 * r + 1] = h_base[b] + t;
      losses[r] = (double)h->h_td[r];
    }
  return rt_replay_update_losses(h, (int64_t)rows, pairs.data(), losses.data(), stream);
}

int rt_replay_profile(rt_replay* h, int32_t enable) {
  RT_REQUIRE(h, "null argument");
  RT_CUDA(cudaSetDevice(h->cfg.device));
  if (enable && h->prof_ev.empty()) {
    h->prof_ev.resize(2 * 512);
    for (auto& e : h->prof_ev) RT_CUDA(cudaEventCreate(&e));
  }
  h->profile = enable != 0;
  h->prof_used = 0;
  return RT_OK;
}

int rt_replay_gather_time(rt_replay* h, double* total_ms, int64_t* launches) {
  RT_REQUIRE(h && total_ms && launches, "null argument");
  RT_CUDA(cudaSetDevice(h->cfg.device));
  RT_CUDA(cudaDeviceSynchronize());
  double t = 0;
  for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
    float ms = 0;
    RT_CUDA(cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]));
    t += ms;
  }
  *total_ms = t;
  *launches = (int64_t)(h->prof_used / 2);
  h->prof_used = 0;
  return RT_OK;
}

static int read_scalar(int device, const double* dptr, double* out, void* stream) {
  RT_CUDA(cudaSetDevice(device));
  RT_CUDA(cudaMemcpyAsync(out, dptr, sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  RT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return RT_OK;
}
int rt_replay_tree_sum(rt_replay* h, double* out, void* stream) {
  RT_REQUIRE(h && out && h->per, "no sum tree");
  return read_scalar(h->cfg.device, h->d_sum + 1, out, stream);
}
int rt_replay_tree_min(rt_replay* h, double* out, void* stream) {
  RT_REQUIRE(h && out && h->d_min, "no min tree");
  return read_scalar(h->cfg.device, h->d_min + 1, out, stream);
}
int rt_replay_tree_leaf(rt_replay* h, int32_t idx, double* out, void* stream) {
  RT_REQUIRE(h && out && h->per && idx >= 0 && idx < h->cap, "bad leaf");
  return read_scalar(h->cfg.device, h->d_sum + h->cap + idx, out, stream);
}

int rt_tree_create(int32_t capacity, int32_t device, rt_tree** out) {
  RT_REQUIRE(out && capacity > 0 && (capacity & (capacity - 1)) == 0,
             "capacity must be a positive power of two");
  RT_CUDA(cudaSetDevice(device));
  rt_tree* t = new rt_tree();
  t->cap = capacity;
  t->device = device;
  while ((1 << t->depth) < capacity) t->depth++;
  RT_CUDA(rt::dmalloc(&t->v, (size_t)2 * capacity));
  RT_CUDA(cudaMemset(t->v, 0, (size_t)2 * capacity * sizeof(double)));
  *out = t;
  return RT_OK;
}
void rt_tree_destroy(rt_tree* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  cudaDeviceSynchronize();
  if (t->v) cudaFree(t->v);
  t->d_idx.release(); t->d_val.release(); t->d_mass.release(); t->d_out.release();
  delete t;
}
int rt_tree_set(rt_tree* t, int32_t m, const int32_t* idx, const double* val, void* stream) {
  RT_REQUIRE(t && idx && val && m >= 0, "bad argument");
  RT_CUDA(cudaSetDevice(t->device));
  cudaStream_t st = (cudaStream_t)stream;
  // duplicates: last write wins, like sequential __setitem__ calls
  std::unordered_map<int, double> last;
  for (int i = 0; i < m; ++i) {
    RT_REQUIRE(idx[i] >= 0 && idx[i] < t->cap, "leaf index out of range");
    last[idx[i]] = val[i];
  }
  std::vector<int> vi;
  std::vector<double> vv;
  for (auto& kv : last) {
    vi.push_back(kv.first);
    vv.push_back(kv.second);
  }
  if (vi.empty()) return RT_OK;
  RT_CUDA(t->d_idx.upload(vi, st));
  RT_CUDA(t->d_val.upload(vv, st));
  int mm = (int)vi.size();
  int threads = mm < 1024 ? ((mm + 31) / 32) * 32 : 1024;
  k_tree_set<<<1, threads, 0, st>>>(t->v, nullptr, t->cap, t->depth, t->d_idx.p, t->d_val.p,
                                    nullptr, mm);
  RT_LAUNCH_CHECK();
  return RT_OK;
}
int rt_tree_sum(rt_tree* t, double* out, void* stream) {
  RT_REQUIRE(t && out, "null argument");
  return read_scalar(t->device, t->v + 1, out, stream);
}
int rt_tree_find(rt_tree* t, int32_t m, const double* mass, int32_t* out_idx, void* stream) {
  RT_REQUIRE(t && mass && out_idx && m >= 0, "bad argument");
  RT_CUDA(cudaSetDevice(t->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (m == 0) return RT_OK;
  std::vector<double> vm(mass, mass + m);
  RT_CUDA(t->d_mass.upload(vm, st));
  RT_CUDA(t->d_out.reserve(m));
  k_tree_find<<<(m + 127) / 128, 128, 0, st>>>(t->v, t->cap, t->d_mass.p, t->d_out.p, m);
  RT_LAUNCH_CHECK();
  RT_CUDA(cudaMemcpyAsync(out_idx, t->d_out.p, m * sizeof(int), cudaMemcpyDeviceToHost, st));
  RT_CUDA(cudaStreamSynchronize(st));
  return RT_OK;
}

}  // extern "C"
