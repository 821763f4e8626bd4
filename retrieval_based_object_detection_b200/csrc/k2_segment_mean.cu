// K2  segmented_mean_renorm -- delegate ("average") vectors for every class in one launch.
//
// Replaces: compute_average (32_create_delegate_vector.py:9-10) applied per class to the rows
// scrolled at :123-137, followed by the renormalisation the COSINE collection applies when the
// mean is upserted (:41-42).
//
// Work decomposition: a work item is (class, chunk of <= SEG_CHUNK rows).  A one-block planning
// kernel turns the CSR offsets into an exclusive prefix of item counts; the main kernel launches
// an upper bound of items and each CTA finds its (class, chunk) by binary search, so skewed class
// sizes spread over many CTAs with no host round trip.  Inside a CTA the 4 warps take rows
// round-robin, lanes cover the row with 128-bit loads, and every lane accumulates its columns in
// fp64 registers; warps are combined through shared memory.  Single-chunk classes are finished
// in place; multi-chunk classes park an fp64 partial per chunk and the last CTA to arrive adds
// the partials in chunk order (deterministic) and finishes.
//
// Numerics (mirrored by oracle/oracle_np.py::segment_mean_renorm): m = fp32(sum_fp64 / len),
// then exactly K1's normalisation of m.  HBM-bound: bytes = n*dim*sizeof(row) + n*8 + C*dim*4.
#include "rbod_common.cuh"
#include "rbod_internal.h"

namespace rbod {

namespace {

constexpr int SEG_CHUNK = 1024;
constexpr int SEG_WARPS = 4;
constexpr int SEG_THREADS = SEG_WARPS * 32;

// prefix[c]  = number of items before class c            (c = 0..C)
// prefix[C+1+c] = number of parked partial slots before class c (multi-chunk classes only)
__global__ void seg_plan_kernel(const int64_t* __restrict__ offsets, int64_t C, int* __restrict__ prefix,
                                unsigned int* __restrict__ arrive) {
  __shared__ int s_items[1024];
  __shared__ int s_parts[1024];
  __shared__ int carry_items, carry_parts;
  if (threadIdx.x == 0) { carry_items = 0; carry_parts = 0; }
  __syncthreads();
  for (int64_t base = 0; base < C; base += 1024) {
    const int64_t c = base + threadIdx.x;
    int items = 0, parts = 0;
    if (c < C) {
      const int64_t len = offsets[c + 1] - offsets[c];
      items = len > 0 ? (int)((len + SEG_CHUNK - 1) / SEG_CHUNK) : 1;  // empty class: 1 item writes zeros
      parts = items > 1 ? items : 0;
      arrive[c] = 0u;
    }
    s_items[threadIdx.x] = items;
    s_parts[threadIdx.x] = parts;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 entries
    for (int o = 1; o < 1024; o <<= 1) {
      int a = 0, b = 0;
      if ((int)threadIdx.x >= o) { a = s_items[threadIdx.x - o]; b = s_parts[threadIdx.x - o]; }
      __syncthreads();
      s_items[threadIdx.x] += a;
      s_parts[threadIdx.x] += b;
      __syncthreads();
    }
    if (c < C) {
      prefix[c] = carry_items + s_items[threadIdx.x] - items;
      prefix[C + 1 + c] = carry_parts + s_parts[threadIdx.x] - parts;
    }
    __syncthreads();
    if (threadIdx.x == 1023) { carry_items += s_items[1023]; carry_parts += s_parts[1023]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { prefix[C] = carry_items; prefix[2 * C + 1] = carry_parts; }
}

// VEC = 4: dim == 128 * NG, 128-bit loads.  VEC = 1: dim <= 32 * NG, scalar loads.
template <int VEC, int NG>
__global__ void __launch_bounds__(SEG_THREADS)
seg_mean_kernel(const float* __restrict__ master32, const uint16_t* __restrict__ rows16, int kind16, int dim,
                int64_t ld32, int64_t ld16, int64_t n_valid, const int64_t* __restrict__ row_idx,
                const int64_t* __restrict__ offsets, int64_t C, const int* __restrict__ prefix,
                double* __restrict__ partials, unsigned int* __restrict__ arrive, float* __restrict__ out,
                int* __restrict__ err_flag) {
  extern __shared__ double s_acc[];  // [SEG_WARPS][dim]
  __shared__ double s_red[SEG_WARPS];
  __shared__ int s_last;

  const int item = blockIdx.x;
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NACC = VEC * NG;
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

  for (int64_t i = r0 + warp; i < r1; i += SEG_WARPS) {
    const int64_t r = row_idx ? row_idx[i] : i;
    if (r < 0 || r >= n_valid) {
      if (lane == 0) atomicExch(err_flag, 1);
      continue;
    }
    if constexpr (VEC == 4) {
      if (master32) {
        const float4* src = reinterpret_cast<const float4*>(master32 + r * ld32);
        float4 v[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) v[g] = __ldcs(src + lane + 32 * g);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          acc[4 * g + 0] += (double)v[g].x;
          acc[4 * g + 1] += (double)v[g].y;
          acc[4 * g + 2] += (double)v[g].z;
          acc[4 * g + 3] += (double)v[g].w;
        }
      } else {
        const uint2* src = reinterpret_cast<const uint2*>(rows16 + r * ld16);
        uint2 v[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) v[g] = __ldcs(src + lane + 32 * g);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          acc[4 * g + 0] += (double)h16_to_f32((uint16_t)(v[g].x & 0xffffu), kind16);
          acc[4 * g + 1] += (double)h16_to_f32((uint16_t)(v[g].x >> 16), kind16);
          acc[4 * g + 2] += (double)h16_to_f32((uint16_t)(v[g].y & 0xffffu), kind16);
          acc[4 * g + 3] += (double)h16_to_f32((uint16_t)(v[g].y >> 16), kind16);
        }
      }
    } else {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int col = lane + 32 * g;
        if (col < dim) {
          const float x = master32 ? master32[r * ld32 + col] : h16_to_f32(rows16[r * ld16 + col], kind16);
          acc[g] += (double)x;
        }
      }
    }
  }

  // park per-warp sums: column of acc[...] for this lane
  double* mine = s_acc + (size_t)warp * dim;
  if constexpr (VEC == 4) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
#pragma unroll
      for (int e = 0; e < 4; ++e) mine[(lane + 32 * g) * 4 + e] = acc[4 * g + e];
    }
  } else {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int col = lane + 32 * g;
      if (col < dim) mine[col] = acc[g];
    }
  }
  __syncthreads();
  // combine warps in fixed order into s_acc[0][*]
  for (int col = threadIdx.x; col < dim; col += SEG_THREADS) {
    double s = s_acc[col];
#pragma unroll
    for (int w = 1; w < SEG_WARPS; ++w) s += s_acc[(size_t)w * dim + col];
    s_acc[col] = s;
  }
  __syncthreads();

  if (nchunks > 1) {
    const int part0 = prefix[C + 1 + c];
    double* dst = partials + (size_t)(part0 + chunk) * dim;
    for (int col = threadIdx.x; col < dim; col += SEG_THREADS) dst[col] = s_acc[col];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(arrive + c, 1u);
      s_last = (prev == (unsigned int)(nchunks - 1));
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double* src = partials + (size_t)part0 * dim;
    for (int col = threadIdx.x; col < dim; col += SEG_THREADS) {
      double s = 0.0;
      for (int j = 0; j < nchunks; ++j) s += __ldcg(src + (size_t)j * dim + col);
      s_acc[col] = s;
    }
    __syncthreads();
  }

  // finish: m = fp32(sum / len); out = K1 normalisation of m
  const int64_t len = seg1 - seg0;
  const double inv_len = len > 0 ? 1.0 / (double)len : 0.0;
  double ss = 0.0;
  for (int col = threadIdx.x; col < dim; col += SEG_THREADS) {
    const float m = len > 0 ? (float)(s_acc[col] / (double)len) : 0.0f;
    (void)inv_len;
    s_acc[col] = (double)m;
    ss = fma((double)m, (double)m, ss);
  }
  ss = warp_sum_f64(ss);
  if (lane == 0) s_red[warp] = ss;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < SEG_WARPS; ++w) tot += s_red[w];
  const double r = tot > 0.0 ? 1.0 / sqrt(tot) : 0.0;
  for (int col = threadIdx.x; col < dim; col += SEG_THREADS) out[c * dim + col] = (float)(s_acc[col] * r);
}

}  // namespace

int launch_segment_mean(const float* master32, const uint16_t* rows16, int kind16, int dim, int64_t ld32,
                        int64_t ld16, int64_t n_valid, const int64_t* row_idx, const int64_t* offsets,
                        int64_t n_classes, int64_t n_items_upper, double* partials, int* chunk_prefix,
                        unsigned int* arrive_cnt, float* out, int* err_flag, cudaStream_t st) {
  if (n_classes <= 0) return RBOD_OK;
  seg_plan_kernel<<<1, 1024, 0, st>>>(offsets, n_classes, chunk_prefix, arrive_cnt);
  RBOD_CUDA(cudaGetLastError());
  const size_t smem = (size_t)SEG_WARPS * dim * sizeof(double);
  const int grid = (int)n_items_upper;
  const bool vec_ok = dim % 128 == 0 && dim / 128 <= 8 &&
                      (master32 ? (reinterpret_cast<uintptr_t>(master32) % 16 == 0 && ld32 % 4 == 0)
                                : (reinterpret_cast<uintptr_t>(rows16) % 8 == 0 && ld16 % 4 == 0));
#define RBOD_K2_LAUNCH(VEC, NG)                                                                              \
  do {                                                                                                       \
    RBOD_CUDA(cudaFuncSetAttribute(seg_mean_kernel<VEC, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                   (int)smem));                                                              \
    seg_mean_kernel<VEC, NG><<<grid, SEG_THREADS, smem, st>>>(master32, rows16, kind16, dim, ld32, ld16,      \
                                                              n_valid, row_idx, offsets, n_classes,          \
                                                              chunk_prefix, partials, arrive_cnt, out,       \
                                                              err_flag);                                     \
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
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

}  // namespace rbod
