// K2b  segment_delegates -- the other three delegate ("representative") vectors of a class, for every
// class in one launch: centroid, weighted average and medoid.
//
// Replaces, per class, what 32_create_delegate_vector.py computes in float64 numpy on the class's
// scrolled vectors (:137) and then upserts (:41-42, :147-156):
//   compute_centroid          :12-15   member closest (L2) to the mean
//   compute_weighted_average  :17-21   sum_i v_i * softmax_i(-alpha * ||v_i - mean||), alpha = 2.0
//   compute_medoid            :23-26   member with the smallest sum of L2 distances to all members
// (compute_average :9-10 is K2, k2_segment_mean.cu).  Like the reference the arithmetic is float64 on the
// stored (float32 / 16-bit) rows; the output is the STORED form of the delegate, i.e. what the COSINE
// collection keeps when the script upserts it: fp32(v) L2-normalised exactly as K1 does.
//
// One CTA (8 warps) per class.  Lane l of a warp owns columns l, l+32, ...; a warp streams rows with
// coalesced loads and keeps its partial sums in fp64 registers; warps combine through shared memory in a
// fixed order (deterministic).  HBM traffic: the class's rows are read 2x (centroid), 3x (weighted) or
// (n/8 + 1)x (medoid, O(n^2 d) flops -- the one super-linear step of the reference); classes are small
// (hundreds of rows) so the re-reads are L2 hits.  Ties (duplicate members) resolve to the first member,
// as numpy's argmin does.
#include "rbod_common.cuh"
#include "rbod_internal.h"

#include <algorithm>

namespace rbod {

namespace {

constexpr int DG_WARPS = 8;
constexpr int DG_THREADS = DG_WARPS * 32;

struct RowSrc {
  const float* master32;
  const uint16_t* rows16;
  int kind16;
  int64_t ld32, ld16;
  __device__ __forceinline__ double at(int64_t r, int c) const {
    return master32 ? (double)master32[r * ld32 + c] : (double)h16_to_f32(rows16[r * ld16 + c], kind16);
  }
};

// (value, index) argmin with first-index tie break, over the block
__device__ __forceinline__ void block_argmin(double& v, int& i, double* s_v, int* s_i) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(FULL_MASK, v, o);
    const int oi = __shfl_xor_sync(FULL_MASK, i, o);
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  if (lane == 0) { s_v[warp] = v; s_i[warp] = i; }
  __syncthreads();
  v = s_v[0];
  i = s_i[0];
  for (int w = 1; w < DG_WARPS; ++w)
    if (s_v[w] < v || (s_v[w] == v && s_i[w] < i)) { v = s_v[w]; i = s_i[w]; }
  __syncthreads();
}

template <int NJ>
__global__ void __launch_bounds__(DG_THREADS)
segment_delegates_kernel(RowSrc src, int dim, int64_t n_valid, const int64_t* __restrict__ row_idx,
                         const int64_t* __restrict__ offsets, int kind, double alpha, int cosine,
                         double* __restrict__ scratch, float* __restrict__ out, int64_t* __restrict__ out_member,
                         int* __restrict__ err_flag, int skip_upto) {
  extern __shared__ double s_vec[];   // [dim] mean, later the delegate itself
  __shared__ double s_red[DG_WARPS];
  __shared__ int s_redi[DG_WARPS];
  __shared__ int s_bad;
  const int64_t c = blockIdx.x;
  const int64_t seg0 = offsets[c], seg1 = offsets[c + 1];
  const int n = (int)(seg1 - seg0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* o = out + c * dim;
  if (n > 0 && n <= skip_upto) return;   // the tiled medoid kernel answered this class
  if (n <= 0) {   // empty class: zero vector, no member
    for (int col = threadIdx.x; col < dim; col += DG_THREADS) o[col] = 0.0f;
    if (threadIdx.x == 0 && out_member) out_member[c] = -1;
    return;
  }
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  auto row_of = [&](int i) -> int64_t {
    const int64_t r = row_idx ? row_idx[seg0 + i] : seg0 + i;
    if (r < 0 || r >= n_valid) { s_bad = 1; return 0; }
    return r;
  };
  double* dist = scratch + seg0;   // one double per member

  // ---- mean (float64, warps combined in warp order) -- not needed by the medoid
  if (kind != RBOD_DELEGATE_MEDOID) {
    double acc[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[j] = 0.0;
    for (int i = warp; i < n; i += DG_WARPS) {
      const int64_t r = row_of(i);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int col = lane + 32 * j;
        if (col < dim) acc[j] += src.at(r, col);
      }
    }
    for (int w = 0; w < DG_WARPS; ++w) {
      if (warp == w) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int col = lane + 32 * j;
          if (col < dim) s_vec[col] = (w == 0 ? 0.0 : s_vec[col]) + acc[j];
        }
      }
      __syncthreads();
    }
    for (int col = threadIdx.x; col < dim; col += DG_THREADS) s_vec[col] = s_vec[col] / (double)n;
    __syncthreads();
  }

  int member = -1;
  if (kind == RBOD_DELEGATE_CENTROID || kind == RBOD_DELEGATE_WEIGHTED) {
    // ---- distance of every member to the mean
    double best = INFINITY;
    int best_i = 0x7fffffff;
    double wsum = 0.0;
    for (int i = warp; i < n; i += DG_WARPS) {
      const int64_t r = row_of(i);
      double d2 = 0.0;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int col = lane + 32 * j;
        if (col < dim) {
          const double d = src.at(r, col) - s_vec[col];
          d2 = fma(d, d, d2);
        }
      }
      d2 = warp_sum_f64(d2);
      const double d = sqrt(d2);
      if (kind == RBOD_DELEGATE_CENTROID) {
        if (d < best || (d == best && i < best_i)) { best = d; best_i = i; }
      } else {
        const double w = exp(-alpha * d);
        if (lane == 0) dist[i] = w;
        wsum += w;   // same value in every lane
      }
    }
    if (kind == RBOD_DELEGATE_CENTROID) {
      block_argmin(best, best_i, s_red, s_redi);
      member = best_i;
    } else {
      if (lane == 0) s_red[warp] = wsum;
      __threadfence_block();
      __syncthreads();
      double total = 0.0;
      for (int w = 0; w < DG_WARPS; ++w) total += s_red[w];
      __syncthreads();
      // ---- weighted sum with weights / total
      double acc[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[j] = 0.0;
      for (int i = warp; i < n; i += DG_WARPS) {
        const int64_t r = row_of(i);
        const double w = dist[i] / total;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int col = lane + 32 * j;
          if (col < dim) acc[j] = fma(src.at(r, col), w, acc[j]);
        }
      }
      for (int w = 0; w < DG_WARPS; ++w) {
        if (warp == w) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int col = lane + 32 * j;
            if (col < dim) s_vec[col] = (w == 0 ? 0.0 : s_vec[col]) + acc[j];
          }
        }
        __syncthreads();
      }
    }
  } else if (kind == RBOD_DELEGATE_MEDOID) {
    // ---- sum of distances from member i to every member; warp per i, row i held in registers
    double best = INFINITY;
    int best_i = 0x7fffffff;
    for (int i = warp; i < n; i += DG_WARPS) {
      const int64_t ri = row_of(i);
      double vi[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int col = lane + 32 * j;
        vi[j] = col < dim ? src.at(ri, col) : 0.0;
      }
      double total = 0.0;
      for (int k = 0; k < n; ++k) {
        const int64_t rk = row_of(k);
        double d2 = 0.0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int col = lane + 32 * j;
          if (col < dim) {
            const double d = vi[j] - src.at(rk, col);
            d2 = fma(d, d, d2);
          }
        }
        d2 = warp_sum_f64(d2);
        total += sqrt(d2);
      }
      if (total < best || (total == best && i < best_i)) { best = total; best_i = i; }
    }
    block_argmin(best, best_i, s_red, s_redi);
    member = best_i;
  }

  if (member >= 0) {   // centroid / medoid: the delegate is the member's stored row
    const int64_t r = row_of(member);
    for (int col = threadIdx.x; col < dim; col += DG_THREADS) s_vec[col] = src.at(r, col);
    if (threadIdx.x == 0 && out_member) out_member[c] = s_bad ? -1 : r;
    __syncthreads();
  } else if (threadIdx.x == 0 && out_member) {
    out_member[c] = -1;
  }

  // ---- stored form: m = fp32(delegate); cosine collections keep K1's normalisation of m
  double ss = 0.0;
  for (int col = threadIdx.x; col < dim; col += DG_THREADS) {
    const float m = (float)s_vec[col];
    s_vec[col] = (double)m;
    ss = fma((double)m, (double)m, ss);
  }
  ss = warp_sum_f64(ss);
  if (lane == 0) s_red[warp] = ss;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < DG_WARPS; ++w) tot += s_red[w];
  const double rn = cosine ? (tot > 0.0 ? 1.0 / sqrt(tot) : 0.0) : 1.0;
  for (int col = threadIdx.x; col < dim; col += DG_THREADS) o[col] = (float)(s_vec[col] * rn);
  if (threadIdx.x == 0 && s_bad) atomicExch(err_flag, 1);
}

// ---------------------------------------------------------------------------------------------
// Medoid of classes of up to MD_NMAX members: all pairwise distances with register tiling.
// The generic kernel above gives one warp a member i and walks every other member with a warp reduction per pair:
// correct for any class size, but < 1 fp64 pair-column per clock per SM (10^4 classes x 100 rows x 768: 314 ms).
// Here a thread owns 4 x 4 blocks of the UPPER TRIANGLE of the n x n distance matrix (16 fp64 accumulators each, two
// blocks per pass; a block off the diagonal feeds the row sums of its i-rows and, transposed, of its j-rows);
// the class is staged through shared memory 32 columns at a time, TRANSPOSED (column-major, padded pitch), so a thread
// reads the four i-values and the four j-values of a column with 16-byte loads -- 4 loads per 32 fp64 operations -- and
// the sum over columns runs in column order.  Every (row, block column) cell of the partial-sum table is written by
// exactly one thread and the cells are added in a fixed order, so the result is deterministic, duplicate members get
// bit-identical distance sums and the argmin takes the first of them, as numpy does.
// ---------------------------------------------------------------------------------------------
constexpr int MD_THREADS = 256;
constexpr int MD_CH = 32;        // columns per staged chunk
constexpr int MD_NB = 2;         // 4x4 blocks per thread and pass (32 fp64 accumulators: two CTAs per SM)
constexpr int MD_NMAX = 256;

__host__ __device__ inline int md_pitch(int n4) { return n4 + ((18 - n4 % 16) % 16); }   // pitch % 16 == 2: even, spreads banks
__host__ inline size_t md_smem_bytes(int n_max, int dim) {
  const int n4 = (n_max + 3) / 4 * 4;
  const size_t tile = (size_t)MD_CH * md_pitch(n4) * 8;
  return std::max(tile, (size_t)dim * 8) + (size_t)n4 * (n4 / 4) * 8 + (size_t)n4 * 8;
}

__global__ void __launch_bounds__(MD_THREADS, 2)
medoid_tiled_kernel(RowSrc src, int dim, int64_t n_valid, const int64_t* __restrict__ row_idx,
                    const int64_t* __restrict__ offsets, int cosine, int n_max, float* __restrict__ out,
                    int64_t* __restrict__ out_member, int* __restrict__ err_flag) {
  extern __shared__ double md_smem[];
  __shared__ double s_red[MD_THREADS / 32];
  __shared__ int s_redi[MD_THREADS / 32];
  __shared__ int s_bad;
  const int64_t c = blockIdx.x;
  const int64_t seg0 = offsets[c], seg1 = offsets[c + 1];
  const int n = (int)(seg1 - seg0);
  if (n <= 0 || n > n_max) return;          // empty and over-sized classes belong to the generic kernel
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n4m = (n_max + 3) / 4 * 4;
  const int n4 = (n + 3) / 4 * 4, nbk = n4 / 4, pitch = md_pitch(n4);
  const size_t tile_a = (size_t)MD_CH * md_pitch(n4m);
  const size_t tile_doubles = tile_a > (size_t)dim ? tile_a : (size_t)dim;
  double* T = md_smem;                                   // [MD_CH][pitch], later the delegate [dim]
  double* SP = md_smem + tile_doubles;                   // [n4][nbk] partial row sums
  int64_t* rows = reinterpret_cast<int64_t*>(SP + (size_t)n4m * (n4m / 4));   // [n4] member row slots
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n4; i += MD_THREADS) {
    int64_t r = 0;
    if (i < n) {
      r = row_idx ? row_idx[seg0 + i] : seg0 + i;
      if (r < 0 || r >= n_valid) { s_bad = 1; r = 0; }
    }
    rows[i] = r;
  }
  __syncthreads();

  const int n_blocks = nbk * (nbk + 1) / 2;              // blocks (bi, bj) with bi <= bj, numbered row by row
  for (int b0 = 0; b0 < n_blocks; b0 += MD_THREADS * MD_NB) {
    double acc[MD_NB][16];
#pragma unroll
    for (int s = 0; s < MD_NB; ++s)
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[s][e] = 0.0;
    int bi[MD_NB], bj[MD_NB];
#pragma unroll
    for (int s = 0; s < MD_NB; ++s) {
      int b = b0 + s * MD_THREADS + (int)threadIdx.x;
      bi[s] = -1;
      bj[s] = 0;
      if (b < n_blocks) {
        int r = 0;
        while (b >= nbk - r) {
          b -= nbk - r;
          ++r;
        }
        bi[s] = r;
        bj[s] = r + b;
      }
    }
    for (int c0 = 0; c0 < dim; c0 += MD_CH) {
      __syncthreads();   // the previous chunk has been consumed
      for (int i = warp; i < n4; i += MD_THREADS / 32) {
        const int col = c0 + lane;
        T[lane * pitch + i] = (i < n && col < dim) ? src.at(rows[i], col) : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int s = 0; s < MD_NB; ++s) {
        if (bi[s] < 0) continue;
        const double* pa = T + 4 * bi[s];
        const double* pb = T + 4 * bj[s];
#pragma unroll 4
        for (int cc = 0; cc < MD_CH; ++cc) {
          const double2 a01 = *reinterpret_cast<const double2*>(pa + cc * pitch);
          const double2 a23 = *reinterpret_cast<const double2*>(pa + cc * pitch + 2);
          const double2 b01 = *reinterpret_cast<const double2*>(pb + cc * pitch);
          const double2 b23 = *reinterpret_cast<const double2*>(pb + cc * pitch + 2);
          const double a[4] = {a01.x, a01.y, a23.x, a23.y};
          const double bb[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const double d = a[r] - bb[t];
              acc[s][4 * r + t] = fma(d, d, acc[s][4 * r + t]);
            }
        }
      }
    }
    // distances of this pass's blocks -> partial row sums: SP[i][bj] = sum over the block's j, and for blocks off the
    // diagonal SP[j][bi] = sum over the block's i (padding members >= n do not count)
#pragma unroll
    for (int s = 0; s < MD_NB; ++s) {
      if (bi[s] < 0) continue;
      double dd[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) dd[e] = sqrt(acc[s][e]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double sum = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (4 * bj[s] + t < n) sum += dd[4 * r + t];
        SP[(size_t)(4 * bi[s] + r) * nbk + bj[s]] = sum;
      }
      if (bi[s] != bj[s]) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          double sum = 0.0;
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (4 * bi[s] + r < n) sum += dd[4 * r + t];
          SP[(size_t)(4 * bj[s] + t) * nbk + bi[s]] = sum;
        }
      }
    }
  }
  __syncthreads();
  double best = INFINITY;
  int best_i = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += MD_THREADS) {
    double tot = 0.0;
    for (int b = 0; b < nbk; ++b) tot += SP[(size_t)i * nbk + b];
    if (tot < best || (tot == best && i < best_i)) { best = tot; best_i = i; }
  }
  block_argmin(best, best_i, s_red, s_redi);
  // the delegate is the member's stored row, in stored form (fp32, K1's normalisation for cosine collections)
  const int64_t rm = rows[best_i];
  for (int col = threadIdx.x; col < dim; col += MD_THREADS) T[col] = src.at(rm, col);
  if (threadIdx.x == 0 && out_member) out_member[c] = s_bad ? -1 : rm;
  __syncthreads();
  double ss = 0.0;
  for (int col = threadIdx.x; col < dim; col += MD_THREADS) {
    const float m = (float)T[col];
    T[col] = (double)m;
    ss = fma((double)m, (double)m, ss);
  }
  ss = warp_sum_f64(ss);
  if (lane == 0) s_red[warp] = ss;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < MD_THREADS / 32; ++w) tot += s_red[w];
  const double rn = cosine ? (tot > 0.0 ? 1.0 / sqrt(tot) : 0.0) : 1.0;
  float* o = out + c * dim;
  for (int col = threadIdx.x; col < dim; col += MD_THREADS) o[col] = (float)(T[col] * rn);
  if (threadIdx.x == 0 && s_bad) atomicExch(err_flag, 1);
}

}  // namespace

int launch_segment_delegates(const float* master32, const uint16_t* rows16, int kind16, int dim, int64_t ld32,
                             int64_t ld16, int64_t n_valid, const int64_t* row_idx, const int64_t* offsets,
                             int64_t n_classes, int kind, double alpha, int cosine, double* scratch, float* out,
                             int64_t* out_member, int* err_flag, int64_t max_class_rows, cudaStream_t st) {
  if (n_classes <= 0) return RBOD_OK;
  if (dim > 1024) return set_error(RBOD_E_UNSUPPORTED, "segment_delegates: dim %d > 1024", dim);
  RowSrc src{master32, rows16, kind16, ld32, ld16};
  // medoid: classes of up to MD_NMAX members take the register-tiled pairwise-distance kernel, the rest (and every
  // other delegate kind) the generic one, which then skips what is already answered
  int skip_upto = 0;
  if (kind == RBOD_DELEGATE_MEDOID && max_class_rows > 0) {
    const int n_max = (int)std::min<int64_t>(max_class_rows, MD_NMAX);
    const size_t msmem = md_smem_bytes(n_max, dim);
    static bool configured[64] = {false};
    int dev = 0;
    RBOD_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      RBOD_CUDA(cudaFuncSetAttribute(medoid_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)md_smem_bytes(MD_NMAX, 1024)));
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    medoid_tiled_kernel<<<(unsigned)n_classes, MD_THREADS, msmem, st>>>(src, dim, n_valid, row_idx, offsets, cosine, n_max,
                                                                      out, out_member, err_flag);
    RBOD_CUDA(cudaGetLastError());
    skip_upto = n_max;
    if (max_class_rows <= MD_NMAX) {
      // every non-empty class is done; empty ones still need their zero vector from the generic kernel below
    }
  }
  const size_t smem = (size_t)dim * sizeof(double);
  const unsigned grid = (unsigned)n_classes;
#define RBOD_DG_LAUNCH(NJ)                                                                                       \
  segment_delegates_kernel<NJ><<<grid, DG_THREADS, smem, st>>>(src, dim, n_valid, row_idx, offsets, kind, alpha, \
                                                               cosine, scratch, out, out_member, err_flag, skip_upto)
  const int nj = (dim + 31) / 32;
  if (nj <= 8) RBOD_DG_LAUNCH(8);
  else if (nj <= 16) RBOD_DG_LAUNCH(16);
  else if (nj <= 24) RBOD_DG_LAUNCH(24);
  else RBOD_DG_LAUNCH(32);
#undef RBOD_DG_LAUNCH
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

}  // namespace rbod
