// K5  distance_topk -- exact top-k for the two distances of the collection menu that are not inner
// products: Distance.EUCLID and Distance.MANHATTAN (util/qdrant_manager.py:61-66, chosen at :68-79 and
// passed to recreate_collection at :82-85).  Neither is served by the tcgen05 pass: L1 distance is not a
// contraction at all, and the reference only ever creates such collections at script scale, so both take
// one exact route on the CUDA cores in fp64 (B200 issues 64 DFMA per clock per SM):
//
//   key(q, g) = -sum_i (q_i - g_i)^2   (EUCLID)      key(q, g) = -sum_i |q_i - g_i|   (MANHATTAN)
//
// computed in fp64 from the fp32 query and the stored row, so larger key = closer and the ordering
// (key desc, row asc) needs no certification.  Selection without a Q x N matrix:
//   1. sample pass: keys of a strided sample of <= cap rows; the k-th best sample key is a threshold no
//      member of the true top-k can fall below;
//   2. sweep: one pass over the gallery records every row whose key >= threshold (expected k * stride
//      rows); a list that overflows `cap` tightens its threshold to the k-th best key it did record --
//      still a lower bound of the true k-th key -- and only those queries are swept again;
//   3. select: rank by counting, emit (distance, row) ascending by distance.
// One warp owns one gallery row (kept in fp64 registers) and scores it against a batch of queries whose
// fp64 copies sit in L1/L2, so the gallery is read once per batch of 32 queries.
#include "rbod_common.cuh"
#include "rbod_internal.h"

#include <algorithm>

namespace rbod {

namespace {

__device__ __forceinline__ bool key_beats(double sa, uint32_t ia, double sb, uint32_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// fp32 queries of one batch -> fp64 (done once, so the sweep's inner loop has no conversions)
__global__ void __launch_bounds__(256)
dist_widen_queries_kernel(const float* __restrict__ q, const int* __restrict__ qsel, int nf, int dim,
                          double* __restrict__ q64) {
  const int64_t n = (int64_t)nf * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i / dim), c = (int)(i - (int64_t)f * dim);
    q64[i] = (double)q[(int64_t)qsel[f] * dim + c];
  }
}

// Rows r = row0, row0 + stride, ... < n_rows; every (row, query f) with key >= thr[f] is appended to list f.
template <int METRIC, int NMAX>
__global__ void __launch_bounds__(256)
dist_collect_kernel(const double* __restrict__ q64, const float* __restrict__ master32,
                    const uint16_t* __restrict__ rows16, int kind16, int dim, int64_t ld32, int64_t ld16,
                    int64_t n_rows, int64_t row0, int64_t stride, const uint32_t* __restrict__ row_mask,
                    const double* __restrict__ thr, const int* __restrict__ active, int nf, int cap,
                    double* __restrict__ coll_key, uint32_t* __restrict__ coll_idx, int* __restrict__ coll_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * 8;
  for (int64_t r = row0 + w0 * stride; r < n_rows; r += nw * stride) {
    if (row_mask && !((row_mask[r >> 5] >> (r & 31)) & 1u)) continue;
    double g[NMAX];
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
      const int c = lane + 32 * i;
      double x = 0.0;
      if (c < dim) x = master32 ? (double)master32[r * ld32 + c] : (double)h16_to_f32(rows16[r * ld16 + c], kind16);
      g[i] = x;
    }
    for (int f = 0; f < nf; ++f) {
      if (!active[f]) continue;
      const double* qv = q64 + (int64_t)f * dim;
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
        const int c = lane + 32 * i;
        if (c < dim) {
          const double d = qv[c] - g[i];
          if (METRIC == RBOD_EUCLID) acc = fma(d, d, acc);
          else acc += fabs(d);
        }
      }
      acc = warp_sum_f64(acc);
      if (lane == 0) {
        const double key = -acc;
        if (key >= thr[f]) {
          const int slot = atomicAdd(coll_cnt + f, 1);
          if (slot < cap) {
            coll_key[(size_t)f * cap + slot] = key;
            coll_idx[(size_t)f * cap + slot] = (uint32_t)r;
          }
        }
      }
    }
  }
}

// One CTA per query of the batch.
//   sample != 0         : thr[f] = k-th best recorded key (-inf if fewer than k were recorded); no output.
//   list fits (<= cap)  : rank by counting, write the top k, active[f] = 0.
//   list overflowed     : thr[f] = k-th best key among the `cap` rows that were recorded, active[f] stays 1.
__global__ void __launch_bounds__(256)
dist_select_kernel(const double* __restrict__ coll_key, const uint32_t* __restrict__ coll_idx,
                   int* __restrict__ coll_cnt, const int* __restrict__ qsel, int cap, int k, int sample, int metric,
                   double* __restrict__ thr, int* __restrict__ active, int* __restrict__ n_active,
                   float* __restrict__ out_scores, int64_t* __restrict__ out_rows, double* __restrict__ out_keys) {
  const int f = blockIdx.x;
  if (!active[f]) return;
  const int total = coll_cnt[f];
  const int cnt = total < cap ? total : cap;
  const bool overflow = total > cap;
  const double* sc = coll_key + (size_t)f * cap;
  const uint32_t* ix = coll_idx + (size_t)f * cap;
  const int64_t q = qsel[f];
  const bool emit = !sample && !overflow;
  if (emit) {
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      out_scores[q * k + j] = INFINITY;       // "no result": infinitely far
      out_rows[q * k + j] = -1;
      if (out_keys) out_keys[q * k + j] = -INFINITY;
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
    const double s = sc[j];
    const uint32_t id = ix[j];
    int rank = 0;
    for (int i = 0; i < cnt; ++i)
      if (key_beats(sc[i], ix[i], s, id)) ++rank;
    if (emit) {
      if (rank < k) {
        out_scores[q * k + rank] = (float)(metric == RBOD_EUCLID ? sqrt(-s) : -s);
        out_rows[q * k + rank] = (int64_t)id;
        if (out_keys) out_keys[q * k + rank] = s;
      }
    } else if (rank == k - 1) {
      thr[f] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (emit) active[f] = 0;
    else atomicAdd(n_active, 1);
    coll_cnt[f] = 0;
  }
}

template <int METRIC>
void launch_collect_t(int nmax, int grid, cudaStream_t st, const double* q64, const float* master32,
                      const uint16_t* rows16, int kind16, int dim, int64_t ld32, int64_t ld16, int64_t n_rows,
                      int64_t row0, int64_t stride, const uint32_t* row_mask, const double* thr, const int* active,
                      int nf, int cap, double* coll_key, uint32_t* coll_idx, int* coll_cnt) {
#define RBOD_K5_GO(NM)                                                                                          \
  dist_collect_kernel<METRIC, NM><<<grid, 256, 0, st>>>(q64, master32, rows16, kind16, dim, ld32, ld16, n_rows,   \
                                                        row0, stride, row_mask, thr, active, nf, cap, coll_key,  \
                                                        coll_idx, coll_cnt)
  if (nmax <= 4) RBOD_K5_GO(4);
  else if (nmax <= 8) RBOD_K5_GO(8);
  else if (nmax <= 16) RBOD_K5_GO(16);
  else if (nmax <= 24) RBOD_K5_GO(24);
  else RBOD_K5_GO(32);
#undef RBOD_K5_GO
}

}  // namespace

int launch_dist_widen_queries(const float* q, const int* qsel, int nf, int dim, double* q64, cudaStream_t st) {
  if (nf <= 0) return RBOD_OK;
  const int64_t n = (int64_t)nf * dim;
  dist_widen_queries_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, qsel, nf, dim, q64);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_dist_collect(int metric, const double* q64, const float* master32, const uint16_t* rows16, int kind16,
                        int dim, int64_t ld32, int64_t ld16, int64_t n_rows, int64_t row0, int64_t stride,
                        const uint32_t* row_mask, const double* thr, const int* active, int nf, int cap,
                        double* coll_key, uint32_t* coll_idx, int* coll_cnt, int num_sms, cudaStream_t st) {
  if (nf <= 0 || n_rows <= 0) return RBOD_OK;
  if (dim > 1024) return set_error(RBOD_E_UNSUPPORTED, "EUCLID / MANHATTAN search supports dim <= 1024");
  const int64_t rows_visited = (n_rows - row0 + stride - 1) / stride;
  const int64_t want = (rows_visited + 7) / 8;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms * 4));
  const int nmax = (dim + 31) / 32;
  if (metric == RBOD_EUCLID)
    launch_collect_t<RBOD_EUCLID>(nmax, grid, st, q64, master32, rows16, kind16, dim, ld32, ld16, n_rows, row0,
                                  stride, row_mask, thr, active, nf, cap, coll_key, coll_idx, coll_cnt);
  else
    launch_collect_t<RBOD_MANHATTAN>(nmax, grid, st, q64, master32, rows16, kind16, dim, ld32, ld16, n_rows, row0,
                                     stride, row_mask, thr, active, nf, cap, coll_key, coll_idx, coll_cnt);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_dist_select(const double* coll_key, const uint32_t* coll_idx, int* coll_cnt, const int* qsel, int nf,
                       int cap, int k, int sample, int metric, double* thr, int* active, int* n_active,
                       float* out_scores, int64_t* out_rows, double* out_keys, cudaStream_t st) {
  if (nf <= 0) return RBOD_OK;
  dist_select_kernel<<<nf, 256, 0, st>>>(coll_key, coll_idx, coll_cnt, qsel, cap, k, sample, metric, thr, active,
                                         n_active, out_scores, out_rows, out_keys);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

}  // namespace rbod
