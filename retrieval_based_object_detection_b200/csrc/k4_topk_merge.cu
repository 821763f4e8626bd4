// K4 family -- everything between the tensor-core pass and the answer:
//
//   merge_partials   per query, the per-slice candidate lists of K3 -> the kc best approximate
//                    candidates and tau = the kc-th best approximate score (every row K3 dropped
//                    scores <= tau).
//   rescore          exact cosine of every (query, candidate) pair in fp64 on the stored values:
//                    dot / (|q| * |g|), the formula of cosine_similarity
//                    (33_run_all_experiments.py:76-77).
//   select           per query: order candidates by (score desc, row asc), emit the top k, and
//                    certify them: if tau + eps < (k-th exact score) no dropped row can belong to
//                    the top k.  Uncertified queries are appended to a flag list.
//   exact_collect /  exact fp64 sweep over the whole gallery for flagged queries only: collect
//   select_collected every row whose exact score >= the query's provisional k-th score, then
//                    select among those.  This is the guarantee behind "identical ids".
//   merge_topk       (K4 proper) G sorted per-GPU lists -> global top k, after the NCCL
//                    all-gather in the row-sharded multi-GPU search.
//
// All of it is small, latency/HBM-bound integer and fp64 work on CUDA cores.
#include "rbod_common.cuh"
#include "rbod_internal.h"

namespace rbod {

namespace {

__device__ __forceinline__ bool beats(double sa, uint32_t ia, double sb, uint32_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// Exact score of one (query, stored row) pair, one warp per pair, fp64 on the stored values; every lane returns it.
//   COSINE : dot / (|q| |g|)  -- cosine_similarity, 33_run_all_experiments.py:76-77
//   DOT    : dot
//   EUCLID : the ordering key -sum (q_i - g_i)^2, accumulated directly (no |q|^2 - 2 q.g + |g|^2 cancellation)
__device__ __forceinline__ double exact_pair_score(const float* __restrict__ qv, double qq,
                                                   const float* __restrict__ g32, const uint16_t* __restrict__ g16,
                                                   int kind16, int dim, int metric, int lane) {
  double a = 0.0, b = 0.0;
  for (int c = lane; c < dim; c += 32) {
    const double x = g32 ? (double)g32[c] : (double)h16_to_f32(g16[c], kind16);
    const double qc = (double)qv[c];
    if (metric == RBOD_EUCLID) {
      const double d = qc - x;
      a = fma(d, d, a);
    } else {
      a = fma(qc, x, a);
      b = fma(x, x, b);
    }
  }
  a = warp_sum_f64(a);
  if (metric == RBOD_EUCLID) return -a;
  if (metric == RBOD_DOT) return a;
  b = warp_sum_f64(b);
  const double den = sqrt(qq) * sqrt(b);
  return den > 0.0 ? a / den : 0.0;
}

// What the caller sees for an internal score: EUCLID keys (-d^2) come out as the distance d (fp32) and keep the
// key in the fp64 output (the order the multi-GPU merge uses); "no result" is an infinite distance.
__device__ __forceinline__ float user_score(double s, int metric) {
  return metric == RBOD_EUCLID ? (float)sqrt(-s) : (float)s;
}
__device__ __forceinline__ float user_no_result(int metric) { return metric == RBOD_EUCLID ? INFINITY : -INFINITY; }

// ---------------------------------------------------------------------------------------------
// merge_partials: one CTA per query, bitonic sort (descending) of packed keys in shared memory.
// key = ordered(score) << 32 | ~idx   (ties: smaller row index first); 0 = padding.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
merge_partials_kernel(const float* __restrict__ part_score, const uint32_t* __restrict__ part_idx, int slices,
                      int64_t q_pad, int kc, int m_pow2, const float* __restrict__ tau_init,
                      uint32_t* __restrict__ cand_idx, float* __restrict__ cand_tau) {
  extern __shared__ unsigned long long keys[];
  const int64_t q = blockIdx.x;
  const int m = slices * kc;
  for (int i = threadIdx.x; i < m_pow2; i += blockDim.x) {
    unsigned long long key = 0ull;
    if (i < m) {
      const int s = i / kc, j = i - s * kc;
      const size_t off = ((size_t)s * q_pad + q) * kc + j;
      const uint32_t idx = part_idx[off];
      if (idx != 0xffffffffu)
        key = (static_cast<unsigned long long>(f32_to_ordered(part_score[off])) << 32) |
              static_cast<unsigned long long>(~idx);
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= m_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (m_pow2 >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = keys[lo], b = keys[hi];
        if (desc ? (a < b) : (a > b)) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < kc; j += blockDim.x) {
    const unsigned long long key = j < m_pow2 ? keys[j] : 0ull;
    cand_idx[q * kc + j] = key ? ~static_cast<uint32_t>(key & 0xffffffffull) : 0xffffffffu;
  }
  if (threadIdx.x == 0) {
    const unsigned long long key = (kc - 1) < m_pow2 ? keys[kc - 1] : 0ull;
    float tau = key ? ordered_to_f32(static_cast<uint32_t>(key >> 32)) : -INFINITY;
    // rows below the pre-sampled starting threshold were dropped without ever entering a list
    if (tau_init != nullptr) tau = fmaxf(tau, tau_init[q]);
    cand_tau[q] = tau;
  }
}

// ---------------------------------------------------------------------------------------------
// tau_init: starting threshold of a query = the smallest of its `groups` sample-group maxima.  Every
// group holds one row scoring at least that much, and with N / (rows per group) = 16 * kc the expected
// number of gallery rows above it is ~ 16 * kc * H(groups): far fewer than N, so the candidate heaps
// skip their cold start, yet fewer than k rows beat it only with probability ~ (k / (16 kc))^groups
// per query (8 groups: < 1e-9 even at k = 100, kc = 128); the host then redoes the call without it.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tau_init_kernel(const float* __restrict__ groupmax, int groups, int splits, int64_t q_pad,
                uint32_t* __restrict__ tau_shared, float* __restrict__ tau_init) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= q_pad) return;
  // a group's maximum = the largest of its splits' maxima (each split visited part of the group's comb)
  float t = INFINITY;
  for (int g = 0; g < groups; ++g) {
    float m = -INFINITY;
    for (int s = 0; s < splits; ++s) m = fmaxf(m, groupmax[((size_t)g * splits + s) * q_pad + q]);
    t = fminf(t, m);
  }
  if (!(t > -INFINITY)) t = -INFINITY;   // an empty (or fully masked) group: no starting threshold
  tau_init[q] = t;
  tau_shared[q] = f32_to_ordered(t);
}

// ---------------------------------------------------------------------------------------------
// rescore: one warp per (query, candidate)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rescore_kernel(const float* __restrict__ q, const double* __restrict__ q_qq, const float* __restrict__ master32,
               const uint16_t* __restrict__ rows16, int kind16, int dim, int64_t ld32, int64_t ld16, int metric,
               const uint32_t* __restrict__ cand_idx, int64_t n_pairs, int kc, double* __restrict__ cand_score) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * 8;
  for (int64_t p = w0; p < n_pairs; p += nw) {
    const uint32_t idx = cand_idx[p];
    if (idx == 0xffffffffu) {
      if (lane == 0) cand_score[p] = -INFINITY;
      continue;
    }
    const int64_t qi = p / kc;
    const double sc = exact_pair_score(q + qi * dim, q_qq[qi], master32 ? master32 + (int64_t)idx * ld32 : nullptr,
                                       master32 ? nullptr : rows16 + (int64_t)idx * ld16, kind16, dim, metric, lane);
    if (lane == 0) cand_score[p] = sc;
  }
}

// ---------------------------------------------------------------------------------------------
// select: one warp per query, rank by counting over kc <= 128 candidates
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
select_kernel(const double* __restrict__ cand_score, const uint32_t* __restrict__ cand_idx,
              const float* __restrict__ cand_tau, const float* __restrict__ q_dq, const float* __restrict__ stats,
              const double* __restrict__ q_qq, int metric, int master16, int shadow, int dp, int64_t Q, int kc, int k, float* __restrict__ out_scores,
              int64_t* __restrict__ out_rows,
              double* __restrict__ out_scores64, int* __restrict__ n_flag, int* __restrict__ flag_q,
              double* __restrict__ flag_thr, float* __restrict__ flag_lo, float* __restrict__ max_eps) {
  __shared__ double s_sc[4][K3_MAX_KC];
  __shared__ uint32_t s_ix[4][K3_MAX_KC];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * 4 + w;
  if (q >= Q) return;
  for (int j = lane; j < kc; j += 32) {
    s_sc[w][j] = cand_score[q * kc + j];
    s_ix[w][j] = cand_idx[q * kc + j];
  }
  for (int j = lane; j < k; j += 32) {
    out_scores[q * k + j] = user_no_result(metric);
    out_rows[q * k + j] = -1;
    if (out_scores64) out_scores64[q * k + j] = -INFINITY;
  }
  __syncwarp();
  double kth = -INFINITY;  // exact score of rank k-1, if it exists
  int have_kth = 0;
  for (int j = lane; j < kc; j += 32) {
    const double s = s_sc[w][j];
    const uint32_t ix = s_ix[w][j];
    if (ix == 0xffffffffu) continue;
    int rank = 0;
    for (int i = 0; i < kc; ++i) {
      const uint32_t oi = s_ix[w][i];
      if (oi != 0xffffffffu && beats(s_sc[w][i], oi, s, ix)) ++rank;
    }
    if (rank < k) {
      out_scores[q * k + rank] = user_score(s, metric);
      out_rows[q * k + rank] = (int64_t)ix;
      if (out_scores64) out_scores64[q * k + rank] = s;
      if (rank == k - 1) { kth = s; have_kth = 1; }
    }
  }
  // broadcast the k-th score
  const uint32_t who = __ballot_sync(FULL_MASK, have_kth);
  const float tau = cand_tau[q];
  if (tau == -INFINITY) return;  // nothing was dropped for this query: exact by construction
  // Certification margin: every row the tensor-core pass dropped has approximate score <= tau, and
  //   |approx - exact cosine| <= ||q16 - unit(q)|| * max||g16||      (query rounding, Cauchy-Schwarz)
  //                            + dp * 2^-23 * max||g16||              (fp32 accumulation in the tensor core)
  //                            + row term
  // The row term is ||g16 - unit(G)|| for fp32 masters (their 16-bit shadow is rounded independently);
  // for 16-bit masters g16 IS the stored row and only its norm deviates from 1 by d = stats[1], which
  // scales the score instead of adding to it: (|tau| + e) * d / (1 - d).
  // With an fp16 shadow as the search operand, the operand norm is stats[2] and its distance to the stored
  // row, stats[3], adds to the query term.
  const float gmax = (shadow ? stats[2] : stats[0]) * 1.000001f, gdev = stats[1] * 1.000001f;
  // DOT collections: nothing is normalised, so every term scales with the query norm |q| and the row term is
  // |q| * ||g16 - g|| (zero for 16-bit masters, whose operand is the stored row).
  // EUCLID collections: the tensor-core pass scores a = q16 . g16 + bias32 with bias32 = fp32(-|g|^2 / 2), an
  // approximation of (key + |q|^2) / 2 for the exact key = -|q - g|^2.  Nothing is normalised, so the dot-product
  // terms are those of DOT, plus the rounding of the bias and of its addition: 2^-23 (|q| G + G^2), G = max |g|.
  const bool unnorm = metric != RBOD_COSINE;
  const float qn = unnorm ? (float)sqrt(q_qq[q]) * 1.000001f + q_dq[q] : 1.0f;
  const float e = q_dq[q] * gmax + (float)dp * 1.2e-7f * gmax * qn + (shadow ? stats[3] * 1.000001f * qn : 0.0f);
  const float row_term = unnorm ? (master16 ? 0.0f : qn * gdev)
                                : (master16 ? (fabsf(tau) + e) * gdev / (1.0f - gdev) : gdev);
  const float gbig = gmax + (master16 ? 0.0f : gdev);
  const float bias_term = metric == RBOD_EUCLID ? 1.2e-7f * (qn * gbig + gbig * gbig) : 0.0f;
  const float eps = e + row_term + bias_term + fabsf(tau) * 1e-6f + 1e-7f;
  const double qq_half = metric == RBOD_EUCLID ? 0.5 * q_qq[q] : 0.0;
  bool flagged = true;
  if (who) {
    const int src = __ffs(who) - 1;
    const double kth_b = __shfl_sync(FULL_MASK, kth, src);
    // the k-th exact score in the domain the tensor-core pass works in
    const double kth_a = metric == RBOD_EUCLID ? 0.5 * kth_b + qq_half : kth_b;
    flagged = !((double)tau + (double)eps < kth_a);
    kth = kth_b;
  } else {
    // Fewer than k candidates although rows were dropped: the pre-sampled starting threshold sat above this
    // query's k-th best score (possible only when the sample misrepresents the gallery).  The host reruns the
    // search without the pre-pass.
    kth = -INFINITY;
    if (lane == 0) atomicAdd(n_flag + 4, 1);
  }
  if (lane == 0) {
    atomic_max_nonneg(max_eps, eps);
    if (flagged) {
      const int slot = atomicAdd(n_flag, 1);
      flag_q[slot] = (int)q;
      flag_thr[slot] = kth;
      // Threshold for the collecting second pass: every row whose exact score reaches kth has an
      // approximate score above lo (same error model, applied from the exact side).
      const float kf = (float)(metric == RBOD_EUCLID ? 0.5 * kth + qq_half : kth);
      const float lo = unnorm ? kf - e - row_term - bias_term
                              : (master16 ? kf - fabsf(kf) * gdev - e : kf - e - gdev);
      flag_lo[slot] = kth == -INFINITY ? -INFINITY : lo - fabsf(kf) * 2e-6f - 2e-7f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// exact fallback
// ---------------------------------------------------------------------------------------------
constexpr int EX_NMAX = 32;  // dim <= 1024

__global__ void __launch_bounds__(256)
exact_collect_kernel(const float* __restrict__ q, const double* __restrict__ q_qq,
                     const float* __restrict__ master32, const uint16_t* __restrict__ rows16, int kind16, int dim,
                     int64_t ld32, int64_t ld16, int metric, int64_t n_rows, const uint32_t* __restrict__ row_mask,
                     const int* __restrict__ flag_q, const double* __restrict__ flag_thr, int f0, int nf, int cap,
                     double* __restrict__ coll_score, uint32_t* __restrict__ coll_idx, int* __restrict__ coll_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * 8;
  for (int64_t r = w0; r < n_rows; r += nw) {
    if (row_mask && !((row_mask[r >> 5] >> (r & 31)) & 1u)) continue;
    double g[EX_NMAX];
    double gg = 0.0;
#pragma unroll
    for (int i = 0; i < EX_NMAX; ++i) {
      const int c = lane + 32 * i;
      double x = 0.0;
      if (c < dim) x = master32 ? (double)master32[r * ld32 + c] : (double)h16_to_f32(rows16[r * ld16 + c], kind16);
      g[i] = x;
      gg = fma(x, x, gg);
    }
    gg = warp_sum_f64(gg);
    const double gn = sqrt(gg);
    for (int f = 0; f < nf; ++f) {
      const int qi = flag_q[f0 + f];
      const float* qv = q + (int64_t)qi * dim;
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < EX_NMAX; ++i) {
        const int c = lane + 32 * i;
        if (c < dim) {
          if (metric == RBOD_EUCLID) {
            const double d = (double)qv[c] - g[i];
            acc = fma(d, d, acc);
          } else {
            acc = fma((double)qv[c], g[i], acc);
          }
        }
      }
      acc = warp_sum_f64(acc);
      if (lane == 0) {
        const double den = sqrt(q_qq[qi]) * gn;
        const double s = metric == RBOD_EUCLID ? -acc : (metric == RBOD_DOT ? acc : (den > 0.0 ? acc / den : 0.0));
        if (s >= flag_thr[f0 + f]) {
          const int slot = atomicAdd(coll_cnt + f, 1);
          if (slot < cap) {
            coll_score[(size_t)f * cap + slot] = s;
            coll_idx[(size_t)f * cap + slot] = (uint32_t)r;
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
select_collected_kernel(const double* __restrict__ coll_score, const uint32_t* __restrict__ coll_idx,
                        const int* __restrict__ coll_cnt, const int* __restrict__ flag_q, int f0, int cap, int k,
                        int metric, float* __restrict__ out_scores, int64_t* __restrict__ out_rows,
                        double* __restrict__ out_scores64, int* __restrict__ overflow) {
  const int f = blockIdx.x;
  const int q = flag_q[f0 + f];
  int cnt = coll_cnt[f];
  if (cnt > cap) {
    if (overflow == nullptr) return;   // caller re-runs this query through the exact sweep
    if (threadIdx.x == 0) atomicExch(overflow, 1);
    cnt = cap;
  }
  const double* sc = coll_score + (size_t)f * cap;
  const uint32_t* ix = coll_idx + (size_t)f * cap;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    out_scores[(int64_t)q * k + j] = user_no_result(metric);
    out_rows[(int64_t)q * k + j] = -1;
    if (out_scores64) out_scores64[(int64_t)q * k + j] = -INFINITY;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
    const double s = sc[j];
    const uint32_t id = ix[j];
    int rank = 0;
    for (int i = 0; i < cnt; ++i)
      if (beats(sc[i], ix[i], s, id)) ++rank;
    if (rank < k) {
      out_scores[(int64_t)q * k + rank] = user_score(s, metric);
      out_rows[(int64_t)q * k + rank] = (int64_t)id;
      if (out_scores64) out_scores64[(int64_t)q * k + rank] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// second pass for uncertified queries (tensor-core collect): pack their 16-bit rows into a dense
// query matrix, then rescore whatever the collecting K3 launch recorded.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_flagged_kernel(const uint16_t* __restrict__ q16, int dp, const int* __restrict__ flag_q, int f0, int nf,
                      int64_t nf_pad, uint16_t* __restrict__ fq16, int* __restrict__ coll_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * 8;
  for (int64_t f = w0; f < nf_pad; f += nw) {
    uint4* dst = reinterpret_cast<uint4*>(fq16 + f * dp);
    if (f < nf) {
      const uint4* src = reinterpret_cast<const uint4*>(q16 + (int64_t)flag_q[f0 + f] * dp);
      for (int c = lane; c < dp / 8; c += 32) dst[c] = src[c];
    } else {
      for (int c = lane; c < dp / 8; c += 32) dst[c] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (lane == 0) coll_cnt[f] = 0;
  }
}

// one warp per (flagged query, recorded row); grid.y = flagged query
__global__ void __launch_bounds__(256)
rescore_collected_kernel(const float* __restrict__ q, const double* __restrict__ q_qq,
                         const float* __restrict__ master32, const uint16_t* __restrict__ rows16, int kind16,
                         int dim, int64_t ld32, int64_t ld16, int metric, const int* __restrict__ flag_q, int f0, int cap,
                         const uint32_t* __restrict__ coll_idx, const int* __restrict__ coll_cnt,
                         double* __restrict__ coll_score) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.y;
  const int cnt = min(coll_cnt[f], cap);
  const int qi = flag_q[f0 + f];
  const float* qv = q + (int64_t)qi * dim;
  for (int j = blockIdx.x * 8 + (threadIdx.x >> 5); j < cnt; j += gridDim.x * 8) {
    const uint32_t idx = coll_idx[(size_t)f * cap + j];
    const double sc = exact_pair_score(qv, q_qq[qi], master32 ? master32 + (int64_t)idx * ld32 : nullptr,
                                       master32 ? nullptr : rows16 + (int64_t)idx * ld16, kind16, dim, metric, lane);
    if (lane == 0) coll_score[(size_t)f * cap + j] = sc;
  }
}

// ---------------------------------------------------------------------------------------------
// merge_topk: one warp per query, lane g walks the (sorted) list of shard g.  G <= 32.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
merge_topk_kernel(const double* __restrict__ scores64, const int64_t* __restrict__ ids, int G, int64_t Q, int k,
                  float* __restrict__ out_scores, int64_t* __restrict__ out_ids, double* __restrict__ out_scores64) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= Q) return;
  int pos = 0;
  const double* sc = scores64 + ((size_t)lane * Q + q) * k;
  const int64_t* id = ids + ((size_t)lane * Q + q) * k;
  const long long kEmpty = 0x7fffffffffffffffll;
  double hs = -INFINITY;
  long long hi = kEmpty;
  if (lane < G && k > 0 && id[0] >= 0) { hs = sc[0]; hi = id[0]; }
  for (int j = 0; j < k; ++j) {
    // warp arg-best over (score desc, id asc); empty heads lose
    double bs = hs;
    long long bi = hi;
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double os = __shfl_xor_sync(FULL_MASK, bs, o);
      const long long oi = __shfl_xor_sync(FULL_MASK, bi, o);
      const int ol = __shfl_xor_sync(FULL_MASK, bl, o);
      const bool take = (oi != kEmpty) && (bi == kEmpty || os > bs || (os == bs && (oi < bi || (oi == bi && ol < bl))));
      if (take) { bs = os; bi = oi; bl = ol; }
    }
    if (lane == 0) {
      const bool valid = bi != kEmpty;
      out_scores[q * k + j] = valid ? (float)bs : -INFINITY;
      out_ids[q * k + j] = valid ? (int64_t)bi : -1;
      if (out_scores64) out_scores64[q * k + j] = valid ? bs : -INFINITY;
    }
    if (bi == kEmpty) continue;  // uniform: all lanes agree on the winner
    if (lane == bl) {
      ++pos;
      if (pos < k && id[pos] >= 0) { hs = sc[pos]; hi = id[pos]; }
      else { hs = -INFINITY; hi = kEmpty; }
    }
  }
}

}  // namespace

int launch_merge_partials(const float* part_score, const uint32_t* part_idx, int slices, int64_t q_pad,
                          int64_t Q, int kc, const float* tau_init, uint32_t* cand_idx, float* cand_tau,
                          cudaStream_t st) {
  if (Q <= 0) return RBOD_OK;
  int m = slices * kc, p2 = 32;
  while (p2 < m) p2 <<= 1;
  if (p2 > 8192) return set_error(RBOD_E_INVAL, "merge_partials: %d candidates per query exceed 8192", m);
  const size_t smem = (size_t)p2 * 8;
  RBOD_CUDA(cudaFuncSetAttribute(merge_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  merge_partials_kernel<<<(unsigned)Q, 256, smem, st>>>(part_score, part_idx, slices, q_pad, kc, p2, tau_init,
                                                        cand_idx, cand_tau);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_tau_init(const float* groupmax, int groups, int splits, int64_t q_pad, uint32_t* tau_shared,
                    float* tau_init, cudaStream_t st) {
  if (q_pad <= 0) return RBOD_OK;
  tau_init_kernel<<<(unsigned)((q_pad + 255) / 256), 256, 0, st>>>(groupmax, groups, splits, q_pad, tau_shared,
                                                                   tau_init);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_rescore(const float* q, const double* q_qq, const float* master32, const uint16_t* rows16, int kind16,
                   int dim, int64_t ld32, int64_t ld16, int metric, const uint32_t* cand_idx, int64_t Q, int kc,
                   double* cand_score, cudaStream_t st) {
  const int64_t n_pairs = Q * kc;
  if (n_pairs <= 0) return RBOD_OK;
  const int64_t want = (n_pairs + 7) / 8;
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  rescore_kernel<<<grid, 256, 0, st>>>(q, q_qq, master32, rows16, kind16, dim, ld32, ld16, metric, cand_idx, n_pairs,
                                       kc, cand_score);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_select(const double* cand_score, const uint32_t* cand_idx, const float* cand_tau, const float* q_dq,
                  const float* stats, const double* q_qq, int metric, int master16, int shadow, int dp, int64_t Q, int kc,
                  int k, float* out_scores, int64_t* out_rows, double* out_scores64, int* n_flag, int* flag_q,
                  double* flag_thr, float* flag_lo, float* max_eps, cudaStream_t st) {
  if (Q <= 0) return RBOD_OK;
  if (kc > K3_MAX_KC) return set_error(RBOD_E_INVAL, "select: kc %d > %d", kc, K3_MAX_KC);
  select_kernel<<<(unsigned)((Q + 3) / 4), 128, 0, st>>>(cand_score, cand_idx, cand_tau, q_dq, stats, q_qq, metric, master16, shadow, dp, Q, kc, k,
                                                         out_scores, out_rows, out_scores64, n_flag, flag_q,
                                                         flag_thr, flag_lo, max_eps);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_exact_collect(const float* q, const double* q_qq, const float* master32, const uint16_t* rows16,
                         int kind16, int dim, int64_t ld32, int64_t ld16, int metric, int64_t n_rows,
                         const uint32_t* row_mask, const int* flag_q, const double* flag_thr, int f0, int nf,
                         int cap, double* coll_score, uint32_t* coll_idx, int* coll_cnt, int num_sms,
                         cudaStream_t st) {
  if (nf <= 0 || n_rows <= 0) return RBOD_OK;
  if (dim > 32 * EX_NMAX) return set_error(RBOD_E_UNSUPPORTED, "exact fallback supports dim <= %d", 32 * EX_NMAX);
  const int64_t want = (n_rows + 7) / 8;
  const int grid = (int)(want < (int64_t)num_sms * 6 ? want : (int64_t)num_sms * 6);
  exact_collect_kernel<<<grid, 256, 0, st>>>(q, q_qq, master32, rows16, kind16, dim, ld32, ld16, metric, n_rows, row_mask,
                                             flag_q, flag_thr, f0, nf, cap, coll_score, coll_idx, coll_cnt);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_select_collected(const double* coll_score, const uint32_t* coll_idx, const int* coll_cnt,
                            const int* flag_q, int f0, int nf, int cap, int k, int metric, float* out_scores,
                            int64_t* out_rows, double* out_scores64, int* overflow, cudaStream_t st) {
  if (nf <= 0) return RBOD_OK;
  select_collected_kernel<<<nf, 256, 0, st>>>(coll_score, coll_idx, coll_cnt, flag_q, f0, cap, k, metric, out_scores,
                                              out_rows, out_scores64, overflow);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_gather_flagged(const uint16_t* q16, int dp, const int* flag_q, int f0, int nf, int64_t nf_pad,
                          uint16_t* fq16, int* coll_cnt, cudaStream_t st) {
  if (nf_pad <= 0) return RBOD_OK;
  const int64_t want = (nf_pad + 7) / 8;
  gather_flagged_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, st>>>(q16, dp, flag_q, f0, nf, nf_pad,
                                                                                    fq16, coll_cnt);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_rescore_collected(const float* q, const double* q_qq, const float* master32, const uint16_t* rows16,
                             int kind16, int dim, int64_t ld32, int64_t ld16, int metric, const int* flag_q, int f0,
                             int nf, int cap, const uint32_t* coll_idx, const int* coll_cnt, double* coll_score,
                             cudaStream_t st) {
  if (nf <= 0) return RBOD_OK;
  for (int y0 = 0; y0 < nf; y0 += 32768) {   // grid.y limit
    const int ny = nf - y0 < 32768 ? nf - y0 : 32768;
    rescore_collected_kernel<<<dim3(4, (unsigned)ny), 256, 0, st>>>(
        q, q_qq, master32, rows16, kind16, dim, ld32, ld16, metric, flag_q, f0 + y0, cap, coll_idx + (size_t)y0 * cap,
        coll_cnt + y0, coll_score + (size_t)y0 * cap);
  }
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_merge_topk(const double* scores64, const int64_t* ids, int G, int64_t Q, int k, float* out_scores,
                      int64_t* out_ids, double* out_scores64, cudaStream_t st) {
  if (Q <= 0 || k <= 0) return RBOD_OK;
  if (G < 1 || G > 32) return set_error(RBOD_E_UNSUPPORTED, "merge_topk: G=%d outside [1, 32]", G);
  merge_topk_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, st>>>(scores64, ids, G, Q, k, out_scores, out_ids,
                                                            out_scores64);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

}  // namespace rbod
