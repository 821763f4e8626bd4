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
// A warp scores two gallery rows at a time against a batch of 32 queries whose fp64 copies sit in L1/L2 (64
// independent DFMA chains per lane), so the gallery is read once per batch of 32 queries.
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

// Lane l ends up with the sum over all lanes of acc[l] (a 32 x 32 transpose-reduce: 31 shuffle-adds instead of
// 32 five-step butterflies).
__device__ __forceinline__ double warp_transpose_sum(double (&acc)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const double send = up ? acc[i] : acc[i + o];
      const double keep = up ? acc[i + o] : acc[i];
      acc[i] = keep + __shfl_xor_sync(FULL_MASK, send, o);
    }
  }
  return acc[0];
}

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Rows r = row0, row0 + stride, ... < n_rows.  A warp takes R rows at a time: it copies them into its own
// shared-memory buffer with cp.async (all of a row group's loads in flight at once, the next group's copy
// overlapping this group's arithmetic), then lane l walks columns l, l + 32, ... keeping one fp64 accumulator
// per (row, query) -- 32 queries x R rows of independent DFMA chains, each query element loaded once per R
// rows -- and the transpose-reduce leaves query f's key in lane f, which owns that query's threshold and
// appends to its list.  q64 is [32, dim], zero rows beyond the batch.  `words` = 4-byte words per stored row.
template <int METRIC, int R>
__global__ void __launch_bounds__(256, 1)
dist_collect_kernel(const double* __restrict__ q64, const float* __restrict__ master32,
                    const uint16_t* __restrict__ rows16, int kind16, int dim, int64_t ld32, int64_t ld16, int words,
                    int64_t n_rows, int64_t row0, int64_t stride, const uint32_t* __restrict__ row_mask,
                    const double* __restrict__ thr, const int* __restrict__ active, int nf, int cap,
                    double* __restrict__ coll_key, uint32_t* __restrict__ coll_idx, int* __restrict__ coll_cnt) {
  extern __shared__ uint32_t k5_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + warp;
  const int64_t nw = (int64_t)gridDim.x * 8;
  const bool mine = lane < nf && active[lane] != 0;
  const double my_thr = mine ? thr[lane] : INFINITY;
  const int64_t n_visit = n_rows > row0 ? (n_rows - row0 + stride - 1) / stride : 0;   // rows of this pass
  uint32_t* wbuf = k5_smem + (size_t)warp * 2 * R * words;

  auto row_ok = [&](int64_t v) -> bool {
    if (v >= n_visit) return false;
    const int64_t r = row0 + v * stride;
    return row_mask == nullptr || ((row_mask[r >> 5] >> (r & 31)) & 1u);
  };
  auto issue = [&](int64_t v0, int b) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (!row_ok(v0 + j)) continue;
      const int64_t r = row0 + (v0 + j) * stride;
      const uint32_t* src = master32 ? reinterpret_cast<const uint32_t*>(master32 + r * ld32)
                                     : reinterpret_cast<const uint32_t*>(rows16 + r * ld16);
      uint32_t* dst = wbuf + (size_t)(b * R + j) * words;
      for (int w = lane; w < words; w += 32) cp_async_4(dst + w, src + w);
    }
    cp_async_commit();
  };

  int b = 0;
  int64_t v0 = w0 * R;
  if (v0 < n_visit) issue(v0, 0);
  for (; v0 < n_visit; v0 += nw * R, b ^= 1) {
    const int64_t vn = v0 + nw * R;
    if (vn < n_visit) issue(vn, b ^ 1); else cp_async_commit();
    cp_async_wait_1();
    __syncwarp();
    bool ok[R];
#pragma unroll
    for (int j = 0; j < R; ++j) ok[j] = row_ok(v0 + j);
    double acc[R][32];
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int f = 0; f < 32; ++f) acc[j][f] = 0.0;
    const uint32_t* rows_s = wbuf + (size_t)b * R * words;
    for (int c = lane; c < dim; c += 32) {
      double x[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        x[j] = 0.0;
        if (ok[j]) {
          if (master32) {
            x[j] = (double)__uint_as_float(rows_s[(size_t)j * words + c]);
          } else {
            const uint32_t w = rows_s[(size_t)j * words + (c >> 1)];
            x[j] = (double)h16_to_f32((uint16_t)((c & 1) ? (w >> 16) : (w & 0xffffu)), kind16);
          }
        }
      }
#pragma unroll
      for (int f = 0; f < 32; ++f) {
        const double qf = q64[(int64_t)f * dim + c];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const double d = qf - x[j];
          if (METRIC == RBOD_EUCLID) acc[j][f] = fma(d, d, acc[j][f]);
          else acc[j][f] += fabs(d);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const double key = -warp_transpose_sum(acc[j], lane);
      if (ok[j] && mine && key >= my_thr) {
        const int slot = atomicAdd(coll_cnt + lane, 1);
        if (slot < cap) {
          coll_key[(size_t)lane * cap + slot] = key;
          coll_idx[(size_t)lane * cap + slot] = (uint32_t)(row0 + (v0 + j) * stride);
        }
      }
    }
    __syncwarp();   // every lane is done with buffer b before the next iteration's copy lands in it
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
  if (nf > 32) return set_error(RBOD_E_INVAL, "dist_collect: at most 32 queries per pass");
  constexpr int R = 2;
  // a row is staged as 4-byte words (16-bit rows are stored padded to a multiple of 64 elements)
  const int words = master32 ? dim : (int)(ld16 / 2);
  const size_t smem = (size_t)8 * 2 * R * words * 4;
  if (smem > 200 * 1024)
    return set_error(RBOD_E_UNSUPPORTED, "EUCLID / MANHATTAN search: dim %d needs %zu bytes of staging per SM", dim, smem);
  const int64_t rows_visited = (n_rows - row0 + stride - 1) / stride;
  const int64_t want = (rows_visited + 8 * R - 1) / (8 * R);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms));
  if (metric == RBOD_EUCLID) {
    RBOD_CUDA(cudaFuncSetAttribute(dist_collect_kernel<RBOD_EUCLID, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    dist_collect_kernel<RBOD_EUCLID, R><<<grid, 256, smem, st>>>(q64, master32, rows16, kind16, dim, ld32, ld16,
                                                                 words, n_rows, row0, stride, row_mask, thr, active,
                                                                 nf, cap, coll_key, coll_idx, coll_cnt);
  } else {
    RBOD_CUDA(cudaFuncSetAttribute(dist_collect_kernel<RBOD_MANHATTAN, R>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dist_collect_kernel<RBOD_MANHATTAN, R><<<grid, 256, smem, st>>>(q64, master32, rows16, kind16, dim, ld32, ld16,
                                                                    words, n_rows, row0, stride, row_mask, thr,
                                                                    active, nf, cap, coll_key, coll_idx, coll_cnt);
  }
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
