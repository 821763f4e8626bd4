// K2  segmented_mean_renorm -- delegate ("average") vectors for every class in one launch.
//
// Replaces: compute_average (32_create_delegate_vector.py:9-10) applied per class to the rows
// scrolled at :123-137, followed by the renormalisation the COSINE collection applies when the
// mean is upserted (:41-42).
//
// Work decomposition: a work item is (class, chunk of <= K2_SEG_CHUNK rows) and belongs to ONE WARP.
// A one-block planning kernel turns the CSR offsets into an exclusive prefix of item counts; the
// main kernel launches an upper bound of items and each warp finds its (class, chunk) by binary
// search, so skewed class sizes spread over many warps with no host round trip.  Lanes cover the
// row with 128-bit loads, two rows in flight per warp, and every lane accumulates its columns in
// fp64 registers -- no shared memory, no block barrier, so a warp that is normalising and writing
// its class never stalls the warps that are still streaming.  Single-chunk classes are finished
// in place; multi-chunk classes park an fp64 partial per chunk and the last warp to arrive adds
// the partials in chunk order (deterministic) and finishes.
//
// Numerics (mirrored by oracle/oracle_np.py::segment_mean_renorm): m = fp32(sum_fp64 / len),
// then exactly K1's normalisation of m.  HBM-bound: bytes = n*dim*sizeof(row) + n*8 + C*dim*4.
#include "rbod_common.cuh"
#include "rbod_internal.h"

#include <cstdlib>
#include <cstring>

namespace rbod {

namespace {

constexpr int SEG_CHUNK = K2_SEG_CHUNK;
constexpr int SEG_WARPS = 4;

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
constexpr int SEG_THREADS = SEG_WARPS * 32;

// prefix[c]  = number of items before class c            (c = 0..C)
// prefix[C+1+c] = number of parked partial slots before class c (multi-chunk classes only)
__global__ void __launch_bounds__(1024) seg_plan_kernel(const int64_t* __restrict__ offsets, int64_t C,
                                                        int* __restrict__ prefix) {
  // one block, 1024 classes per step: warp-shuffle inclusive scans, warp totals scanned by warp 0 (2 barriers per
  // step instead of the 20 of a shared-memory Hillis-Steele scan: 40 000 classes plan in ~20 us)
  __shared__ int w_items[32], w_parts[32];
  __shared__ int carry_items, carry_parts;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_items = 0; carry_parts = 0; }
  __syncthreads();
  for (int64_t base = 0; base < C; base += 1024) {
    const int64_t c = base + threadIdx.x;
    int items = 0, parts = 0;
    if (c < C) {
      const int64_t len = offsets[c + 1] - offsets[c];
      items = len > 0 ? (int)((len + SEG_CHUNK - 1) / SEG_CHUNK) : 1;  // empty class: 1 item writes zeros
      parts = items > 1 ? items : 0;
    }
    int si = items, sp = parts;       // inclusive scans inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ti = __shfl_up_sync(FULL_MASK, si, o), tp = __shfl_up_sync(FULL_MASK, sp, o);
      if (lane >= o) { si += ti; sp += tp; }
    }
    if (lane == 31) { w_items[warp] = si; w_parts[warp] = sp; }
    __syncthreads();
    if (warp == 0) {                  // exclusive scan of the 32 warp totals
      const int ti0 = w_items[lane], tp0 = w_parts[lane];
      int ti = ti0, tp = tp0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ui = __shfl_up_sync(FULL_MASK, ti, o), up = __shfl_up_sync(FULL_MASK, tp, o);
        if (lane >= o) { ti += ui; tp += up; }
      }
      w_items[lane] = ti - ti0;
      w_parts[lane] = tp - tp0;
    }
    __syncthreads();
    const int ex_items = carry_items + w_items[warp] + si - items;
    const int ex_parts = carry_parts + w_parts[warp] + sp - parts;
    if (c < C) {
      prefix[c] = ex_items;
      prefix[C + 1 + c] = ex_parts;
    }
    __syncthreads();                  // everyone has read the carries
    if (threadIdx.x == 1023) { carry_items = ex_items + items; carry_parts = ex_parts + parts; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { prefix[C] = carry_items; prefix[2 * C + 1] = carry_parts; }
}

// One WARP per work item (class, chunk).  VEC = 4: dim == 128 * NG, 128-bit loads; VEC = 1: dim <= 32 * NG,
// scalar loads.  Lane l owns columns {(l + 32 g) * VEC + e}; rows are streamed UN at a time (2 fp32 rows or
// 4 16-bit rows = 6 KB in flight per warp at dim 768) with the next UN rows prefetched into L2.  No shared
// memory and no block barrier: while one warp normalises and writes its class, its neighbours keep streaming.
//
// Classes longer than one chunk are reduced over a fixed binary tree whose nodes live in the chunks' own
// partial slots (node (level, i) is stored in the slot of its leftmost leaf): a warp stores its sum, sets the
// parent's bit in that slot's flag word, and the second of two siblings to arrive adds left + right and climbs.
// The shape of the tree depends only on the chunk count, so the sum is deterministic, and no warp ever waits.
// Column accumulators of one lane.  FF = 0: fp64 registers (one F2F + DADD per element).  FF = 1: unevaluated
// fp32 pairs (hi, lo) updated with Knuth's TwoSum -- the element work stays on the fp32 pipe and the pair
// carries ~48 bits, so fp32(sum / len) is the same value the fp64 accumulator gives (both are far below
// half an fp32 ulp from the exact sum).
template <int N, int FF>
struct K2Acc {
  double a[N];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = 0.0;
  }
  __device__ __forceinline__ void add(int i, float x) { a[i] += (double)x; }
  __device__ __forceinline__ double get(int i) const { return a[i]; }
};
template <int N>
struct K2Acc<N, 1> {
  float h[N], l[N];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < N; ++i) { h[i] = 0.f; l[i] = 0.f; }
  }
  __device__ __forceinline__ void add(int i, float x) {
    const float t = __fadd_rn(h[i], x);
    const float bp = __fsub_rn(t, h[i]);
    const float err = __fadd_rn(__fsub_rn(h[i], __fsub_rn(t, bp)), __fsub_rn(x, bp));
    h[i] = t;
    l[i] = __fadd_rn(l[i], err);
  }
  __device__ __forceinline__ double get(int i) const { return (double)h[i] + (double)l[i]; }
};

template <int IS16, int NG>
struct K2Row {
  uint2 v[NG];
  __device__ __forceinline__ void load(const uint16_t* rows16, int64_t r, int64_t ld, int lane) {
#pragma unroll
    for (int g = 0; g < NG; ++g) v[g] = __ldcs(reinterpret_cast<const uint2*>(rows16 + r * ld) + lane + 32 * g);
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int g = 0; g < NG; ++g) v[g] = make_uint2(0u, 0u);
  }
  template <class ACC>
  __device__ __forceinline__ void add_to(ACC& acc, int kind16) const {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      acc.add(4 * g + 0, h16_to_f32((uint16_t)(v[g].x & 0xffffu), kind16));
      acc.add(4 * g + 1, h16_to_f32((uint16_t)(v[g].x >> 16), kind16));
      acc.add(4 * g + 2, h16_to_f32((uint16_t)(v[g].y & 0xffffu), kind16));
      acc.add(4 * g + 3, h16_to_f32((uint16_t)(v[g].y >> 16), kind16));
    }
  }
};
template <int NG>
struct K2Row<0, NG> {
  float4 v[NG];
  __device__ __forceinline__ void load(const float* master32, int64_t r, int64_t ld, int lane) {
#pragma unroll
    for (int g = 0; g < NG; ++g) v[g] = __ldcs(reinterpret_cast<const float4*>(master32 + r * ld) + lane + 32 * g);
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int g = 0; g < NG; ++g) v[g] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  template <class ACC>
  __device__ __forceinline__ void add_to(ACC& acc, int) const {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      acc.add(4 * g + 0, v[g].x);
      acc.add(4 * g + 1, v[g].y);
      acc.add(4 * g + 2, v[g].z);
      acc.add(4 * g + 3, v[g].w);
    }
  }
};

template <int VEC, int NG, int IS16, int FF>
__global__ void __launch_bounds__(SEG_THREADS)
seg_mean_kernel(const float* __restrict__ master32, const uint16_t* __restrict__ rows16, int kind16, int dim,
                int64_t ld32, int64_t ld16, int64_t n_valid, const int64_t* __restrict__ row_idx,
                const int64_t* __restrict__ offsets, int64_t C, const int* __restrict__ prefix,
                double* __restrict__ partials, unsigned int* __restrict__ node_bits, float* __restrict__ out,
                double* __restrict__ sums_out, int normalize, int* __restrict__ err_flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * SEG_WARPS + warp;
  if (item >= prefix[C]) return;
  // largest c with prefix[c] <= item
  int64_t lo = 0, hi = C - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (prefix[mid] <= item) lo = mid; else hi = mid - 1;
  }
  const int64_t c = lo;
  const int chunk = item - prefix[c];
  const int nchunks = prefix[c + 1] - prefix[c];
  const int64_t seg0 = offsets[c], seg1 = offsets[c + 1];
  const int64_t r0 = seg0 + (int64_t)chunk * SEG_CHUNK;
  const int64_t r1 = (r0 + SEG_CHUNK < seg1) ? r0 + SEG_CHUNK : seg1;

  constexpr int NACC = VEC * NG;
  K2Acc<NACC, FF> run;
  run.clear();

  auto row_of = [&](int64_t i) -> int64_t {
    const int64_t r = row_idx ? row_idx[i] : i;
    if (r < 0 || r >= n_valid) {
      if (lane == 0) atomicExch(err_flag, 1);
      return -1;
    }
    return r;
  };

  if constexpr (VEC == 4) {
    constexpr int UN = IS16 ? 4 : 2;
    constexpr int LINES = IS16 ? NG * 2 : NG * 4;   // 128-byte lines per row
    int64_t i = r0;
    for (; i + UN <= r1; i += UN) {
      int64_t rr[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) rr[u] = row_of(i + u);
      if (i + 2 * UN <= r1 && lane < LINES) {
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int64_t pr = row_idx ? row_idx[i + UN + u] : i + UN + u;
          if (pr >= 0 && pr < n_valid) {
            if (IS16) prefetch_l2(rows16 + pr * ld16 + lane * 64);
            else prefetch_l2(master32 + pr * ld32 + lane * 32);
          }
        }
      }
      K2Row<IS16, NG> row[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (rr[u] >= 0) {
          if constexpr (IS16) row[u].load(rows16, rr[u], ld16, lane);
          else row[u].load(master32, rr[u], ld32, lane);
        } else {
          row[u].zero();
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) row[u].add_to(run, kind16);
    }
    for (; i < r1; ++i) {
      const int64_t r = row_of(i);
      if (r < 0) continue;
      K2Row<IS16, NG> row;
      if constexpr (IS16) row.load(rows16, r, ld16, lane);
      else row.load(master32, r, ld32, lane);
      row.add_to(run, kind16);
    }
  } else {
    for (int64_t i = r0; i < r1; ++i) {
      const int64_t r = row_of(i);
      if (r < 0) continue;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int col = lane + 32 * g;
        if (col < dim) {
          const float x = IS16 ? h16_to_f32(rows16[r * ld16 + col], kind16) : master32[r * ld32 + col];
          run.add(g, x);
        }
      }
    }
  }

  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = run.get(i);

  // column index of accumulator a of this lane
  auto col_of = [&](int a) -> int { return VEC == 4 ? (lane + 32 * (a >> 2)) * 4 + (a & 3) : lane + 32 * a; };

  if (nchunks > 1) {
    const int part0 = prefix[C + 1 + c];
    double* slots = partials + (size_t)part0 * dim;
    unsigned int* bits = node_bits + part0;
    int idx = chunk;      // node index on its level
    int level = 0;
    int count = nchunks;  // nodes on this level
    while (count > 1) {
      if ((idx ^ 1) >= count) {   // no sibling on this level: the node is its own parent
        idx >>= 1;
        ++level;
        count = (count + 1) >> 1;
        continue;
      }
      // publish this node's sum in the slot of its leftmost leaf, then claim the parent
      double* mine = slots + (size_t)((int64_t)idx << level) * dim;
#pragma unroll
      for (int a = 0; a < NACC; ++a)
        if (col_of(a) < dim) mine[col_of(a)] = acc[a];
      __threadfence();
      __syncwarp();
      const int parent_leaf = (idx >> 1) << (level + 1);
      unsigned int prev = 0;
      if (lane == 0) prev = atomicOr(bits + parent_leaf, 1u << level);
      prev = __shfl_sync(FULL_MASK, prev, 0);
      if (!(prev & (1u << level))) return;   // first of the two siblings: the other one carries on
      __threadfence();
      const double* other = slots + (size_t)((int64_t)(idx ^ 1) << level) * dim;
      if (idx & 1) {   // left + right, whoever arrives last
#pragma unroll
        for (int a = 0; a < NACC; ++a)
          if (col_of(a) < dim) acc[a] = __ldcg(other + col_of(a)) + acc[a];
      } else {
#pragma unroll
        for (int a = 0; a < NACC; ++a)
          if (col_of(a) < dim) acc[a] = acc[a] + __ldcg(other + col_of(a));
      }
      idx >>= 1;
      ++level;
      count = (count + 1) >> 1;
    }
  }

  // shard mode: hand the raw fp64 column sums to the caller (they are all-reduced over the ranks that hold
  // the other rows of this class and finished by seg_finish_kernel)
  if (sums_out != nullptr) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
      if (col_of(a) < dim) sums_out[c * dim + col_of(a)] = acc[a];
    return;
  }

  // finish: m = fp32(sum / len); out = K1 normalisation of m
  const int64_t len = seg1 - seg0;
  double ss = 0.0;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    const float m = (len > 0 && col_of(a) < dim) ? (float)(acc[a] / (double)len) : 0.0f;
    acc[a] = (double)m;
    ss = fma((double)m, (double)m, ss);
  }
  ss = warp_sum_f64(ss);
  // COSINE collections store the mean renormalised (what the upsert at 32_…py:41-42 leaves); the others keep it
  const double r = normalize ? (ss > 0.0 ? 1.0 / sqrt(ss) : 0.0) : 1.0;
  float* o = out + c * dim;
  if constexpr (VEC == 4) {
#pragma unroll
    for (int g = 0; g < NG; ++g)
      reinterpret_cast<float4*>(o)[lane + 32 * g] =
          make_float4((float)(acc[4 * g + 0] * r), (float)(acc[4 * g + 1] * r), (float)(acc[4 * g + 2] * r),
                      (float)(acc[4 * g + 3] * r));
  } else {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
      if (col_of(a) < dim) o[col_of(a)] = (float)(acc[a] * r);
  }
}

// Second half of the sharded delegate build: sums [C, dim] fp64 (already reduced over the shards) and counts
// [C] -> out[c] = K1 normalisation of fp32(sum / count); one warp per class, same arithmetic as the tail above.
__global__ void __launch_bounds__(256)
seg_finish_kernel(const double* __restrict__ sums, const int64_t* __restrict__ counts, int64_t C, int dim,
                  int normalize, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;
  const int64_t len = counts[c];
  const double* s = sums + c * dim;
  double ss = 0.0;
  for (int i = lane; i < dim; i += 32) {
    const float m = len > 0 ? (float)(s[i] / (double)len) : 0.0f;
    ss = fma((double)m, (double)m, ss);
  }
  ss = warp_sum_f64(ss);
  const double r = normalize ? (ss > 0.0 ? 1.0 / sqrt(ss) : 0.0) : 1.0;
  for (int i = lane; i < dim; i += 32) {
    const float m = len > 0 ? (float)(s[i] / (double)len) : 0.0f;
    out[c * dim + i] = (float)((double)m * r);
  }
}

}  // namespace

int launch_segment_finish(const double* sums, const int64_t* counts, int64_t n_classes, int dim, int normalize,
                          float* out, cudaStream_t st) {
  if (n_classes <= 0) return RBOD_OK;
  seg_finish_kernel<<<(unsigned)((n_classes + 7) / 8), 256, 0, st>>>(sums, counts, n_classes, dim, normalize, out);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_segment_mean(const float* master32, const uint16_t* rows16, int kind16, int dim, int64_t ld32,
                        int64_t ld16, int64_t n_valid, const int64_t* row_idx, const int64_t* offsets,
                        int64_t n_classes, int64_t n_items_upper, double* partials, int* chunk_prefix,
                        unsigned int* arrive_cnt, float* out, double* sums_out, int normalize, int* err_flag,
                        cudaStream_t st) {
  if (n_classes <= 0) return RBOD_OK;
  seg_plan_kernel<<<1, 1024, 0, st>>>(offsets, n_classes, chunk_prefix);
  RBOD_CUDA(cudaGetLastError());
  const int grid = (int)((n_items_upper + SEG_WARPS - 1) / SEG_WARPS);
  // Accumulators: fp32 rows add up as TwoSum pairs on the fp32 pipe, 16-bit rows in fp64 registers (measured on
  // B200: 1M x 768 fp32 Zipf classes 75% -> 83% of the HBM peak with pairs; 4M x 768 bf16 79% with fp64 vs 59%
  // with pairs, whose 7 fp32 ops per element outweigh the element's 2 bytes).  RBOD_K2_ACC=f64|ff overrides.
  static const int acc_env = [] {
    const char* e = getenv("RBOD_K2_ACC");
    return !e ? -1 : (!strcmp(e, "f64") ? 0 : 1);
  }();
  const int acc_ff = acc_env >= 0 ? acc_env : (master32 != nullptr ? 1 : 0);
  const bool vec_ok = dim % 128 == 0 && dim / 128 <= 8 &&
                      (master32 ? (reinterpret_cast<uintptr_t>(master32) % 16 == 0 && ld32 % 4 == 0)
                                : (reinterpret_cast<uintptr_t>(rows16) % 8 == 0 && ld16 % 4 == 0));
#define RBOD_K2_ARGS                                                                                         \
  master32, rows16, kind16, dim, ld32, ld16, n_valid, row_idx, offsets, n_classes, chunk_prefix, partials,   \
      arrive_cnt, out, sums_out, normalize, err_flag
#define RBOD_K2_LAUNCH(VEC, NG)                                                                              \
  do {                                                                                                       \
    if (master32) {                                                                                          \
      if (acc_ff) seg_mean_kernel<VEC, NG, 0, 1><<<grid, SEG_THREADS, 0, st>>>(RBOD_K2_ARGS);                \
      else seg_mean_kernel<VEC, NG, 0, 0><<<grid, SEG_THREADS, 0, st>>>(RBOD_K2_ARGS);                       \
    } else {                                                                                                 \
      if (acc_ff) seg_mean_kernel<VEC, NG, 1, 1><<<grid, SEG_THREADS, 0, st>>>(RBOD_K2_ARGS);                \
      else seg_mean_kernel<VEC, NG, 1, 0><<<grid, SEG_THREADS, 0, st>>>(RBOD_K2_ARGS);                       \
    }                                                                                                        \
  } while (0)
  if (vec_ok) {
    switch (dim / 128) {
      case 1: RBOD_K2_LAUNCH(4, 1); break;
      case 2: RBOD_K2_LAUNCH(4, 2); break;
      case 3: RBOD_K2_LAUNCH(4, 3); break;
      case 4: RBOD_K2_LAUNCH(4, 4); break;
      case 5: RBOD_K2_LAUNCH(4, 5); break;
      case 6: RBOD_K2_LAUNCH(4, 6); break;
      case 7: RBOD_K2_LAUNCH(4, 7); break;
      default: RBOD_K2_LAUNCH(4, 8); break;
    }
  } else if (dim <= 256) {
    RBOD_K2_LAUNCH(1, 8);
  } else if (dim <= 1024) {
    RBOD_K2_LAUNCH(1, 32);
  } else {
    return set_error(RBOD_E_UNSUPPORTED, "segment_mean: dim %d > 1024 must be a multiple of 128", dim);
  }
#undef RBOD_K2_LAUNCH
#undef RBOD_K2_ARGS
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

}  // namespace rbod
