// K3  cosine_topk -- query x gallery cosine scores on the tcgen05 tensor cores with the top-k
// selection fused into the epilogue, so the Q x N score matrix never leaves the SM.
//
// Replaces: cosine_similarity (33_run_all_experiments.py:76-77, used at :151), generalised from
// one (test vector, delegate) pair to Q queries against N stored rows, plus the top-k a Qdrant
// `search` would return.
//
// Shape of the computation (D[M=queries, N=gallery rows] = A[M,K] * B[N,K]^T, K = dim):
//   * One CTA owns 128 queries = the 128 TMEM lanes.  In variant 0 the whole 16-bit query tile
//     (128 x dp) is parked in TMEM columns [0, dp/2) for the life of a work unit and used as the
//     A operand straight from tensor memory (tcgen05.mma with A in TMEM), so the only operand
//     that streams is the gallery: HBM -> L2 -> shared memory by TMA (128-byte swizzle, boxes of
//     64 rows x 64 elements), 4 k-blocks per pipeline stage.
//   * Accumulators: two 128x64 fp32 buffers in TMEM columns [384,448) and [448,512); the MMA of
//     gallery tile j+1 overlaps the epilogue of tile j.
//   * Epilogue (4 warps, one thread per query row): tcgen05.ld the 64 scores of the row, compare
//     against the row's running threshold (one FMNMX per score on the fast path); survivors are
//     inserted by the whole warp into the row's candidate list in shared memory (replace-min).
//   * Work unit = (gallery slice, query tile).  Units are ordered slice-major so that CTAs
//     running at the same time stream the same gallery slice and share it through L2.
// Output: per (slice, query) the `kc` best approximate scores and their row indices.  The K4
// kernels merge slices, rescore exactly in fp64 and certify the result.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp % 4).
// Variant 1 streams the A tile through shared memory as well (plain SS MMA); it exists to
// cross-check the TMEM-resident path.
#include "rbod_common.cuh"
#include "rbod_internal.h"

#include <cstring>

namespace rbod {

namespace {

constexpr int TMEM_COLS = 512;
constexpr int ACC_COL0 = 384;
constexpr int B_KBLOCK_BYTES = K3_TILE_N * 128;   // 8 KB : 64 rows x 128 B
constexpr int A_KBLOCK_BYTES = K3_TILE_M * 128;   // 16 KB: 128 rows x 128 B
constexpr int MAX_STAGES = 8;

struct alignas(64) K3Params {
  CUtensorMap tmap_b;
  CUtensorMap tmap_a;
  const uint16_t* q16;
  float* part_score;
  uint32_t* part_idx;
  const uint32_t* row_mask;
  float* dump;
  int64_t dump_ld;
  int64_t n_rows;
  int64_t q_valid;
  int64_t q_pad;
  int dp;
  int num_kb;
  int tiles_total;
  int num_qt;
  int slices;
  int kc;
  int num_stages;
  int variant;
  uint32_t idesc;
};

struct K3Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint64_t a_ready;
  uint32_t tmem_base;
  uint32_t pad;
};

// Whole-warp insertion of the lanes flagged in `ball` (each with its own val / col) into the
// candidate lists of their rows.  Returns the calling lane's updated (threshold, min slot).
__device__ __noinline__ float2 k3_insert(float val, uint32_t col, uint32_t ball, float tau, int minpos, float* sc,
                                         uint32_t* ix, int wrow0, int kc, int lane) {
  while (ball) {
    const int l = __ffs(ball) - 1;
    ball &= ball - 1;
    float* rs = sc + (wrow0 + l) * kc;
    uint32_t* ri = ix + (wrow0 + l) * kc;
    if (lane == l) {
      rs[minpos] = val;
      ri[minpos] = col;
    }
    __syncwarp();
    float lm = INFINITY;
    int lp = 0;
    for (int j = lane; j < kc; j += 32) {
      const float x = rs[j];
      if (x < lm) { lm = x; lp = j; }
    }
    const uint32_t key = f32_to_ordered(lm);
    const uint32_t mn = __reduce_min_sync(FULL_MASK, key);
    const uint32_t who = __ballot_sync(FULL_MASK, key == mn);
    const int src = __ffs(who) - 1;
    const int p = __shfl_sync(FULL_MASK, lp, src);
    if (lane == l) {
      tau = ordered_to_f32(mn);
      minpos = p;
    }
    __syncwarp();
  }
  return make_float2(tau, __int_as_float(minpos));
}

__global__ void __launch_bounds__(K3_THREADS, 1) k3_cosine_topk_kernel(const __grid_constant__ K3Params P) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled operand tiles
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stage_bytes = K3_KB_PER_STAGE * (B_KBLOCK_BYTES + (P.variant == 1 ? A_KBLOCK_BYTES : 0));
  uint8_t* stage_base = smem;
  float* sc = reinterpret_cast<float*>(smem + (size_t)P.num_stages * stage_bytes);
  uint32_t* ix = reinterpret_cast<uint32_t*>(sc + K3_TILE_M * P.kc);
  K3Barriers* bars = reinterpret_cast<K3Barriers*>(ix + K3_TILE_M * P.kc);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmap_b);
    if (P.variant == 1) tma_prefetch_desc(&P.tmap_a);
    for (int s = 0; s < P.num_stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->tfull[b], 1);
      mbar_init(&bars->tempty[b], 128);
    }
    mbar_init(&bars->a_ready, 128);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int num_units = P.slices * P.num_qt;
  const int num_chunks = (P.num_kb + K3_KB_PER_STAGE - 1) / K3_KB_PER_STAGE;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // The whole warp walks the schedule (warp-uniform control flow); one elected lane issues.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t kb_bytes = B_KBLOCK_BYTES + (P.variant == 1 ? A_KBLOCK_BYTES : 0);
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int slice = u / P.num_qt, qt = u - slice * P.num_qt;
      const int t0 = (int)(((int64_t)slice * P.tiles_total) / P.slices);
      const int t1 = (int)(((int64_t)(slice + 1) * P.tiles_total) / P.slices);
      for (int t = t0; t < t1; ++t) {
        for (int ch = 0; ch < num_chunks; ++ch) {
          const int kb0 = ch * K3_KB_PER_STAGE;
          const int nkb = min(K3_KB_PER_STAGE, P.num_kb - kb0);
          mbar_wait(&bars->empty[stage], phase ^ 1u, 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&bars->full[stage], nkb * kb_bytes);
            uint8_t* sb = stage_base + (size_t)stage * stage_bytes;
            for (int j = 0; j < nkb; ++j)
              tma_load_2d(sb + j * B_KBLOCK_BYTES, &P.tmap_b, &bars->full[stage], (kb0 + j) * K3_KBLOCK,
                          t * K3_TILE_N);
            if (P.variant == 1) {
              uint8_t* sa = sb + K3_KB_PER_STAGE * B_KBLOCK_BYTES;
              for (int j = 0; j < nkb; ++j)
                tma_load_2d(sa + j * A_KBLOCK_BYTES, &P.tmap_a, &bars->full[stage], (kb0 + j) * K3_KBLOCK,
                            qt * K3_TILE_M);
            }
          }
          __syncwarp();
          if (++stage == P.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    // Warp-converged loop; elect.sync picks the issuing lane so descriptors stay in uniform registers.
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0, unit_par = 0;
    const uint32_t tmem_b = __shfl_sync(FULL_MASK, tmem_base, 0);
    const uint32_t smem_stage0 = __shfl_sync(FULL_MASK, smem_u32(stage_base), 0);
    const uint64_t desc_hi = make_smem_desc_sw128(0);  // everything except the start address
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int slice = u / P.num_qt;
      const int t0 = (int)(((int64_t)slice * P.tiles_total) / P.slices);
      const int t1 = (int)(((int64_t)(slice + 1) * P.tiles_total) / P.slices);
      if (P.variant == 0) {
        mbar_wait(&bars->a_ready, unit_par, 2);
        unit_par ^= 1u;
        tc_fence_after();
      }
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&bars->tempty[acc], acc_phase ^ 1u, 3);
        tc_fence_after();
        const uint32_t d_tmem = tmem_b + ACC_COL0 + acc * K3_TILE_N;
        for (int ch = 0; ch < num_chunks; ++ch) {
          const int kb0 = ch * K3_KB_PER_STAGE;
          const int nkb = min(K3_KB_PER_STAGE, P.num_kb - kb0);
          mbar_wait(&bars->full[stage], phase, 4);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sb = smem_stage0 + (uint32_t)stage * (uint32_t)stage_bytes;
            const uint32_t sa = sb + K3_KB_PER_STAGE * B_KBLOCK_BYTES;
            for (int j = 0; j < nkb; ++j) {
              const uint64_t bdesc0 = desc_hi | (uint64_t)(((sb + j * B_KBLOCK_BYTES) >> 4) & 0x3fffu);
              const uint64_t adesc0 = desc_hi | (uint64_t)(((sa + j * A_KBLOCK_BYTES) >> 4) & 0x3fffu);
              const int k16 = (kb0 + j) * 4;  // index of the first K=16 step of this k-block
              if (P.variant == 0) {
                const uint32_t a_tmem = tmem_b + (uint32_t)k16 * 8u;
                mma_f16_ts(d_tmem, a_tmem, bdesc0, P.idesc, k16 > 0 ? 1u : 0u);
                mma_f16_ts(d_tmem, a_tmem + 8u, bdesc0 + 2u, P.idesc, 1u);
                mma_f16_ts(d_tmem, a_tmem + 16u, bdesc0 + 4u, P.idesc, 1u);
                mma_f16_ts(d_tmem, a_tmem + 24u, bdesc0 + 6u, P.idesc, 1u);
              } else {
                mma_f16_ss(d_tmem, adesc0, bdesc0, P.idesc, k16 > 0 ? 1u : 0u);
                mma_f16_ss(d_tmem, adesc0 + 2u, bdesc0 + 2u, P.idesc, 1u);
                mma_f16_ss(d_tmem, adesc0 + 4u, bdesc0 + 4u, P.idesc, 1u);
                mma_f16_ss(d_tmem, adesc0 + 6u, bdesc0 + 6u, P.idesc, 1u);
              }
            }
            mma_commit(&bars->empty[stage]);
            if (ch == num_chunks - 1) mma_commit(&bars->tfull[acc]);
          }
          __syncwarp();
          if (++stage == P.num_stages) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ================================ epilogue ====================================
    const int quad = warp & 3;
    const int wrow0 = quad * 32;
    const int row = wrow0 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(wrow0) << 16;
    const int kc = P.kc;
    int acc = 0;
    uint32_t acc_phase = 0;

    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const int slice = u / P.num_qt, qt = u - slice * P.num_qt;
      const int t0 = (int)(((int64_t)slice * P.tiles_total) / P.slices);
      const int t1 = (int)(((int64_t)(slice + 1) * P.tiles_total) / P.slices);
      const int64_t qg = (int64_t)qt * K3_TILE_M + row;

      if (P.variant == 0) {
        // park this thread's query row in TMEM: lane = row, column c holds elements 2c, 2c+1
        const uint4* src = reinterpret_cast<const uint4*>(P.q16 + qg * P.dp);
        const int n16 = P.dp / 16;
        for (int c = 0; c < n16; ++c) {
          const uint4 x0 = __ldg(src + 2 * c), x1 = __ldg(src + 2 * c + 1);
          const uint32_t r[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          tmem_st_32x32b_x8(tmem_base + lane_addr + (uint32_t)c * 8u, r);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars->a_ready);
      }
      // reset the candidate lists of this warp's 32 rows
      for (int r = 0; r < 32; ++r)
        for (int j = lane; j < kc; j += 32) {
          sc[(wrow0 + r) * kc + j] = -INFINITY;
          ix[(wrow0 + r) * kc + j] = 0xffffffffu;
        }
      __syncwarp();
      float tau = -INFINITY;
      int minpos = 0;

      for (int t = t0; t < t1; ++t) {
        mbar_wait(&bars->tfull[acc], acc_phase, 5);
        tc_fence_after();
        uint32_t raw0[32], raw1[32];
        const uint32_t taddr = tmem_base + lane_addr + ACC_COL0 + acc * K3_TILE_N;
        tmem_ld_32x32b_x32(taddr, raw0);
        tmem_ld_32x32b_x32(taddr + 32, raw1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bars->tempty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;

        float v[64];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          v[c] = __uint_as_float(raw0[c]);
          v[32 + c] = __uint_as_float(raw1[c]);
        }
        const int64_t col0 = (int64_t)t * K3_TILE_N;

        if (P.dump != nullptr && qg < P.q_valid) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (col0 + c < P.n_rows) P.dump[qg * P.dump_ld + col0 + c] = v[c];
        }

        if (col0 + K3_TILE_N > P.n_rows || P.row_mask != nullptr) {
          uint64_t allow = ~0ull;
          if (col0 + K3_TILE_N > P.n_rows) {
            const int valid = (int)(P.n_rows - col0);
            allow = valid <= 0 ? 0ull : (valid >= 64 ? ~0ull : ((1ull << valid) - 1ull));
          }
          if (P.row_mask != nullptr) {
            const int64_t w0 = col0 >> 5;
            const int64_t nwords = (P.n_rows + 31) >> 5;
            const uint64_t lo = w0 < nwords ? P.row_mask[w0] : 0u;
            const uint64_t hi = (w0 + 1) < nwords ? P.row_mask[w0 + 1] : 0u;
            allow &= (lo | (hi << 32));
          }
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (!((allow >> c) & 1ull)) v[c] = -INFINITY;
        }

        float m = v[0];
#pragma unroll
        for (int c = 1; c < 64; ++c) m = fmaxf(m, v[c]);
        if (__any_sync(FULL_MASK, m > tau)) {
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            const uint32_t ball = __ballot_sync(FULL_MASK, v[c] > tau);
            if (ball) {
              const float2 r = k3_insert(v[c], (uint32_t)(col0 + c), ball, tau, minpos, sc, ix, wrow0, kc, lane);
              tau = r.x;
              minpos = __float_as_int(r.y);
            }
          }
        }
      }

      // unit done: publish this warp's 32 candidate lists
      __syncwarp();
      for (int r = 0; r < 32; ++r) {
        const int64_t q = (int64_t)qt * K3_TILE_M + wrow0 + r;
        if (q < P.q_valid) {
          const size_t base = ((size_t)slice * P.q_pad + q) * kc;
          for (int j = lane; j < kc; j += 32) {
            P.part_score[base + j] = sc[(wrow0 + r) * kc + j];
            P.part_idx[base + j] = ix[(wrow0 + r) * kc + j];
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// One warp per query: |q|^2 in fp64, unit query rounded to the 16-bit operand type, and
// dq = || fp(q16) - q/|q| ||_2 (the query's share of the certification margin).
__global__ void __launch_bounds__(256)
prep_queries_kernel(const float* __restrict__ q, int64_t Q, int64_t q_pad, int dim, int dp, int kind16,
                    uint16_t* __restrict__ q16, float* __restrict__ q_dq, double* __restrict__ q_qq) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * 8;
  for (int64_t i = w0; i < q_pad; i += nw) {
    uint16_t* dst = q16 + i * dp;
    if (i >= Q) {
      for (int c = lane; c < dp; c += 32) dst[c] = 0;
      continue;
    }
    const float* src = q + i * dim;
    double ss = 0.0;
    for (int c = lane; c < dim; c += 32) {
      const double x = (double)src[c];
      ss = fma(x, x, ss);
    }
    ss = warp_sum_f64(ss);
    const double r = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
    double dd = 0.0;
    for (int c = lane; c < dp; c += 32) {
      uint16_t h = 0;
      if (c < dim) {
        const double u = (double)src[c] * r;
        h = f32_to_h16((float)u, kind16);
        const double e = (double)h16_to_f32(h, kind16) - u;
        dd = fma(e, e, dd);
      }
      dst[c] = h;
    }
    dd = warp_sum_f64(dd);
    if (lane == 0) {
      q_dq[i] = (float)(sqrt(dd) * 1.000001 + 1e-12);
      q_qq[i] = ss;
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

}  // namespace

// [rows, dp] 16-bit row-major -> 2D map with box {64 elements, box_rows}, 128-byte swizzle.
int make_tmap_2d_sw128(CUtensorMap* out, const void* base, int64_t rows, int dp, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(RBOD_E_IO, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)dp, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)dp * 2};
  cuuint32_t box[2] = {(cuuint32_t)K3_KBLOCK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return set_error(RBOD_E_IO, "cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
  return RBOD_OK;
}

size_t k3_smem_bytes(int variant, int kc, int num_stages) {
  const size_t stage = (size_t)K3_KB_PER_STAGE * (B_KBLOCK_BYTES + (variant == 1 ? A_KBLOCK_BYTES : 0));
  return 1024 + (size_t)num_stages * stage + (size_t)K3_TILE_M * kc * 8 + sizeof(K3Barriers);
}

int k3_configure(int device) {
  int optin = 0;
  RBOD_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  RBOD_CUDA(cudaFuncSetAttribute(k3_cosine_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  return optin;
}

int launch_k3(const K3Launch& L, cudaStream_t st) {
  K3Params P;
  memset(&P, 0, sizeof(P));
  P.tmap_b = L.tmap_b;
  P.tmap_a = L.tmap_a;
  P.q16 = L.q16;
  P.part_score = L.part_score;
  P.part_idx = L.part_idx;
  P.row_mask = L.row_mask;
  P.dump = L.dump;
  P.dump_ld = L.dump_ld;
  P.n_rows = L.n_rows;
  P.q_valid = L.q_valid;
  P.q_pad = L.q_pad;
  P.dp = L.dp;
  P.num_kb = L.dp / K3_KBLOCK;
  P.tiles_total = L.tiles_total;
  P.num_qt = L.num_qt;
  P.slices = L.slices;
  P.kc = L.kc;
  P.num_stages = L.num_stages;
  P.variant = L.variant;
  P.idesc = make_idesc_f16(L.a_fmt, L.b_fmt, K3_TILE_M, K3_TILE_N);
  if (L.num_stages < 1 || L.num_stages > MAX_STAGES)
    return set_error(RBOD_E_INVAL, "k3: bad stage count %d", L.num_stages);
  k3_cosine_topk_kernel<<<L.grid, K3_THREADS, L.smem_bytes, st>>>(P);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_prep_queries(const float* q, int64_t Q, int64_t q_pad, int dim, int dp, int kind16, uint16_t* q16,
                        float* q_dq, double* q_qq, cudaStream_t st) {
  const int64_t want = (q_pad + 7) / 8;
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  prep_queries_kernel<<<grid, 256, 0, st>>>(q, Q, q_pad, dim, dp, kind16, q16, q_dq, q_qq);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

}  // namespace rbod
