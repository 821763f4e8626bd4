// K1  l2norm_pack -- the normalise-on-upsert step of a COSINE collection.
//
// Replaces: the per-point `client.upsert` of the reference
// (31_clip_embedding_and_save_vector.py:178-179, 32_create_delegate_vector.py:41-42) together
// with the L2 normalisation Qdrant applies to vectors stored in a Distance.COSINE collection
// (util/qdrant_manager.py:74,82-85).
//
// Layout: one warp owns one row.  For dim == 128*NV the row lives in registers as NV float4 per
// lane (coalesced 512-byte warp loads), is reduced with an fp64 warp shuffle tree, scaled and
// written once: fp32 master row (optional) + 16-bit search operand (bf16 or fp16, row stride
// padded to a multiple of 64 elements so the tcgen05 pass can TMA it with a 128-byte swizzle).
// HBM-bound: algorithmic bytes per row = dim * (4 + sizeof(outputs)).
//
// Numerics (mirrored by oracle/oracle_np.py::l2_normalize_store): s = sum of x_i^2 in fp64,
// r = 1/sqrt(s) in fp64 (0 if s == 0), y_i = fp32(fp64(x_i) * r); 16-bit copy = RNE(y_i).
//
// It also maintains two per-gallery maxima used by the search certification margin:
//   stats[0] = max ||row16||            stats[1] = max ||row16 - target||
// where target = unit(master row) for cosine collections and the master row for dot.
//
// bf16 collections can keep an fp16 SHADOW of the stored rows as the search operand (option
// "shadow16"): fp16 carries 3 more mantissa bits, so queries rounded to fp16 sit 8x closer to the
// exact unit query and the certification margin of the search shrinks accordingly.  A bf16 value
// converts to fp16 exactly unless it is smaller than 2^-17 in magnitude (then off by < 2^-25);
// stats[2] = max ||shadow row||, stats[3] = max ||shadow row - stored row|| account for that.
#include "rbod_common.cuh"
#include "rbod_internal.h"

#include <cstdlib>

namespace rbod {

namespace {

constexpr int K1_WARPS = 8;
constexpr int K1_PREFETCH_ROWS = 2;

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ void finish_row_stats(float s16, float sy, float sd, bool master_is_f32, bool cosine,
                                                 float& wmax_norm, float& wmax_dev) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s16 += __shfl_xor_sync(FULL_MASK, s16, o);
    sy += __shfl_xor_sync(FULL_MASK, sy, o);
    sd += __shfl_xor_sync(FULL_MASK, sd, o);
  }
  const float n16 = sqrtf(s16);
  float dev;
  if (master_is_f32) {
    dev = sqrtf(sd) + (cosine ? fabsf(sqrtf(sy) - 1.0f) : 0.0f);
    if (cosine && sy == 0.0f) dev = sqrtf(sd);
  } else {
    dev = cosine ? fabsf(n16 - 1.0f) : 0.0f;
    if (cosine && s16 == 0.0f) dev = 0.0f;
  }
  wmax_norm = fmaxf(wmax_norm, n16);
  wmax_dev = fmaxf(wmax_dev, dev);
}

// HAS32: the collection keeps an fp32 master row (then the 16-bit row is a shadow and its distance to the
// master is tracked); otherwise the 16-bit row IS the stored vector and only its norm matters.
template <int NV, int HAS32>
__global__ void __launch_bounds__(256)
l2norm_pack_vec_kernel(const float* __restrict__ in, int64_t n, int dim, const int64_t* __restrict__ slots,
                       int64_t slot0, int normalize, int cosine, float* __restrict__ master32, int64_t ld32,
                       uint16_t* __restrict__ out16, int64_t ld16, int kind16, uint16_t* __restrict__ shadow16,
                       float* __restrict__ out_norms, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;   // warps per block
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5);
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * wpb;
  float wmax_norm = 0.f, wmax_dev = 0.f, wmax_snorm = 0.f, wmax_sdev = 0.f;

  for (int64_t row = warp0; row < n; row += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(in + row * dim);
    // pull the row this warp handles two iterations from now into L2 (one 128-byte line per lane): the
    // demand loads below then see L2 latency instead of HBM latency
    {
      const int64_t ahead = row + K1_PREFETCH_ROWS * nwarps;
      if (ahead < n && lane < NV * 4) prefetch_l2(in + ahead * dim + lane * 32);
    }
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldcs(src + lane + 32 * i);

    double ss = 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      ss = fma((double)v[i].x, (double)v[i].x, ss);
      ss = fma((double)v[i].y, (double)v[i].y, ss);
      ss = fma((double)v[i].z, (double)v[i].z, ss);
      ss = fma((double)v[i].w, (double)v[i].w, ss);
    }
    ss = warp_sum_f64(ss);
    if (out_norms != nullptr && lane == 0) out_norms[row] = (float)sqrt(ss);

    if (normalize) {
      const double r = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        v[i].x = (float)((double)v[i].x * r);
        v[i].y = (float)((double)v[i].y * r);
        v[i].z = (float)((double)v[i].z * r);
        v[i].w = (float)((double)v[i].w * r);
      }
    }

    const int64_t slot = slots ? slots[row] : slot0 + row;
    float s16 = 0.f, sy = 0.f, sd = 0.f;
    if (HAS32) {
      float4* dst = reinterpret_cast<float4*>(master32 + slot * ld32);
#pragma unroll
      for (int i = 0; i < NV; ++i) dst[lane + 32 * i] = v[i];
    }
    uint2* dst16 = reinterpret_cast<uint2*>(out16 + slot * ld16);
    uint2* dsts = shadow16 ? reinterpret_cast<uint2*>(shadow16 + slot * ld16) : nullptr;
    float ssn = 0.f, ssd = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint16_t h0 = f32_to_h16(v[i].x, kind16), h1 = f32_to_h16(v[i].y, kind16);
      const uint16_t h2 = f32_to_h16(v[i].z, kind16), h3 = f32_to_h16(v[i].w, kind16);
      const float f0 = h16_to_f32(h0, kind16), f1 = h16_to_f32(h1, kind16);
      const float f2 = h16_to_f32(h2, kind16), f3 = h16_to_f32(h3, kind16);
      if (dsts) {
        const uint16_t t0 = f32_to_h16(f0, 2), t1 = f32_to_h16(f1, 2), t2 = f32_to_h16(f2, 2), t3 = f32_to_h16(f3, 2);
        const float g0 = h16_to_f32(t0, 2), g1 = h16_to_f32(t1, 2), g2 = h16_to_f32(t2, 2), g3 = h16_to_f32(t3, 2);
        ssn += g0 * g0 + g1 * g1 + g2 * g2 + g3 * g3;
        ssd += (g0 - f0) * (g0 - f0) + (g1 - f1) * (g1 - f1) + (g2 - f2) * (g2 - f2) + (g3 - f3) * (g3 - f3);
        uint2 ps;
        ps.x = static_cast<uint32_t>(t0) | (static_cast<uint32_t>(t1) << 16);
        ps.y = static_cast<uint32_t>(t2) | (static_cast<uint32_t>(t3) << 16);
        dsts[lane + 32 * i] = ps;
      }
      s16 += f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3;
      if (HAS32) {
        sy += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        const float d0 = f0 - v[i].x, d1 = f1 - v[i].y, d2 = f2 - v[i].z, d3 = f3 - v[i].w;
        sd += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      }
      uint2 p;
      p.x = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
      p.y = static_cast<uint32_t>(h2) | (static_cast<uint32_t>(h3) << 16);
      if (out16) dst16[lane + 32 * i] = p;
    }
    finish_row_stats(s16, sy, sd, HAS32 != 0, cosine != 0, wmax_norm, wmax_dev);
    if (dsts) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ssn += __shfl_xor_sync(FULL_MASK, ssn, o);
        ssd += __shfl_xor_sync(FULL_MASK, ssd, o);
      }
      wmax_snorm = fmaxf(wmax_snorm, sqrtf(ssn));
      wmax_sdev = fmaxf(wmax_sdev, sqrtf(ssd));
    }
  }
  if (lane == 0 && stats) {
    atomic_max_nonneg(stats + 0, wmax_norm);
    atomic_max_nonneg(stats + 1, wmax_dev);
    if (shadow16) {
      atomic_max_nonneg(stats + 2, wmax_snorm);
      atomic_max_nonneg(stats + 3, wmax_sdev);
    }
  }
}

// Any dim: two passes over the row (the second hits L1/L2).
__global__ void __launch_bounds__(K1_WARPS * 32)
l2norm_pack_generic_kernel(const float* __restrict__ in, int64_t n, int dim, const int64_t* __restrict__ slots,
                           int64_t slot0, int normalize, int cosine, float* __restrict__ master32, int64_t ld32,
                           uint16_t* __restrict__ out16, int64_t ld16, int kind16, uint16_t* __restrict__ shadow16,
                           float* __restrict__ out_norms, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * K1_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * K1_WARPS;
  float wmax_norm = 0.f, wmax_dev = 0.f, wmax_snorm = 0.f, wmax_sdev = 0.f;

  for (int64_t row = warp0; row < n; row += nwarps) {
    const float* src = in + row * dim;
    double ss = 0.0;
    for (int i = lane; i < dim; i += 32) {
      const double x = (double)src[i];
      ss = fma(x, x, ss);
    }
    ss = warp_sum_f64(ss);
    if (lane == 0 && out_norms) out_norms[row] = (float)sqrt(ss);
    const double r = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
    const int64_t slot = slots ? slots[row] : slot0 + row;
    float s16 = 0.f, sy = 0.f, sd = 0.f, ssn = 0.f, ssd = 0.f;
    for (int i = lane; i < dim; i += 32) {
      float y = src[i];
      if (normalize) y = (float)((double)y * r);
      if (master32) master32[slot * ld32 + i] = y;
      const uint16_t h = f32_to_h16(y, kind16);
      const float f = h16_to_f32(h, kind16);
      if (out16) out16[slot * ld16 + i] = h;
      if (shadow16) {
        const uint16_t sh = f32_to_h16(f, 2);
        const float gsh = h16_to_f32(sh, 2);
        shadow16[slot * ld16 + i] = sh;
        ssn += gsh * gsh;
        ssd += (gsh - f) * (gsh - f);
      }
      s16 += f * f;
      sy += y * y;
      sd += (f - y) * (f - y);
    }
    finish_row_stats(s16, sy, sd, master32 != nullptr, cosine != 0, wmax_norm, wmax_dev);
    if (shadow16) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ssn += __shfl_xor_sync(FULL_MASK, ssn, o);
        ssd += __shfl_xor_sync(FULL_MASK, ssd, o);
      }
      wmax_snorm = fmaxf(wmax_snorm, sqrtf(ssn));
      wmax_sdev = fmaxf(wmax_sdev, sqrtf(ssd));
    }
  }
  if (lane == 0 && stats) {
    atomic_max_nonneg(stats + 0, wmax_norm);
    atomic_max_nonneg(stats + 1, wmax_dev);
    if (shadow16) {
      atomic_max_nonneg(stats + 2, wmax_snorm);
      atomic_max_nonneg(stats + 3, wmax_sdev);
    }
  }
}

// fp16 shadow of a bf16 collection built after the fact (a search with k > 40 asks for it, see rbod_api.cu): one warp
// per stored row, 16-byte loads and stores over the padded row, each bf16 value rounded to fp16 exactly as K1 does
// when the shadow exists from the start; stats[2] / stats[3] = max ||shadow row|| / max ||shadow row - stored row||.
// HBM-bound: 4 bytes per element.
__global__ void __launch_bounds__(256)
build_shadow_kernel(const uint16_t* __restrict__ rows16, int64_t n, int dp, uint16_t* __restrict__ shadow16,
                    float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t nw = static_cast<int64_t>(gridDim.x) * 8;
  float wmax_n = 0.f, wmax_d = 0.f;
  for (int64_t r = w0; r < n; r += nw) {
    const uint4* src = reinterpret_cast<const uint4*>(rows16 + r * dp);
    uint4* dst = reinterpret_cast<uint4*>(shadow16 + r * dp);
    float ssn = 0.f, ssd = 0.f;
    for (int c = lane; c < dp / 8; c += 32) {
      const uint4 h = __ldcs(src + c);
      const uint32_t w[4] = {h.x, h.y, h.z, h.w};
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float f0 = h16_to_f32(static_cast<uint16_t>(w[i] & 0xffffu), 1);
        const float f1 = h16_to_f32(static_cast<uint16_t>(w[i] >> 16), 1);
        const uint16_t t0 = f32_to_h16(f0, 2), t1 = f32_to_h16(f1, 2);
        const float g0 = h16_to_f32(t0, 2), g1 = h16_to_f32(t1, 2);
        ssn += g0 * g0 + g1 * g1;
        ssd += (g0 - f0) * (g0 - f0) + (g1 - f1) * (g1 - f1);
        o[i] = static_cast<uint32_t>(t0) | (static_cast<uint32_t>(t1) << 16);
      }
      dst[c] = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ssn += __shfl_xor_sync(FULL_MASK, ssn, o);
      ssd += __shfl_xor_sync(FULL_MASK, ssd, o);
    }
    wmax_n = fmaxf(wmax_n, sqrtf(ssn));
    wmax_d = fmaxf(wmax_d, sqrtf(ssd));
  }
  if (lane == 0) {
    atomic_max_nonneg(stats + 2, wmax_n);
    atomic_max_nonneg(stats + 3, wmax_d);
  }
}

// EUCLID collections: row_bias[slot] = fp32(-|stored row|^2 / 2), the per-row term the K3 epilogue adds so that the
// tensor-core score q . g - |g|^2 / 2 orders rows by Euclidean distance.  Sum in fp64 over the row AS STORED (after
// K1), one warp per row; slots = slots_dev[i] or slot0 + i.
__global__ void __launch_bounds__(256)
row_bias_kernel(const float* __restrict__ master32, const uint16_t* __restrict__ rows16, int kind16, int dim,
                int64_t ld32, int64_t ld16, const int64_t* __restrict__ slots, int64_t slot0, int64_t n,
                float* __restrict__ row_bias) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t nw = static_cast<int64_t>(gridDim.x) * 8;
  for (int64_t i = w0; i < n; i += nw) {
    const int64_t slot = slots ? slots[i] : slot0 + i;
    double ss = 0.0;
    for (int c = lane; c < dim; c += 32) {
      const double x = master32 ? (double)master32[slot * ld32 + c] : (double)h16_to_f32(rows16[slot * ld16 + c], kind16);
      ss = fma(x, x, ss);
    }
    ss = warp_sum_f64(ss);
    if (lane == 0) row_bias[slot] = (float)(-0.5 * ss);
  }
}

// out[i, :] = widen(stored row rows[i])
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ master32, const uint16_t* __restrict__ rows16, int kind16, int dim,
                   int64_t ld32, int64_t ld16, const int64_t* __restrict__ rows, int64_t n, int64_t n_valid,
                   float* __restrict__ out, int* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * 8;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int64_t r = rows[i];
    if (r < 0 || r >= n_valid) {
      if (lane == 0) atomicExch(err_flag, 1);
      continue;
    }
    if (master32) {
      for (int c = lane; c < dim; c += 32) out[i * dim + c] = master32[r * ld32 + c];
    } else {
      for (int c = lane; c < dim; c += 32) out[i * dim + c] = h16_to_f32(rows16[r * ld16 + c], kind16);
    }
  }
}

}  // namespace

int launch_l2norm_pack(const float* in, int64_t n, int dim, const int64_t* slots_dev, int64_t slot0,
                          int normalize, int cosine, float* master32, int64_t ld32, uint16_t* out16,
                          int64_t ld16, int kind16, uint16_t* shadow16, float* out_norms, float* stats, int num_sms,
                          cudaStream_t st) {
  if (n <= 0) return RBOD_OK;
  const int64_t want = (n + K1_WARPS - 1) / K1_WARPS;
  int grid = (int)(want < (int64_t)num_sms * 8 ? want : (int64_t)num_sms * 8);
  const bool aligned = (reinterpret_cast<uintptr_t>(in) % 16 == 0) &&
                       (master32 == nullptr || (reinterpret_cast<uintptr_t>(master32) % 16 == 0 && ld32 % 4 == 0)) &&
                       (out16 == nullptr || (reinterpret_cast<uintptr_t>(out16) % 8 == 0 && ld16 % 4 == 0));
  // Launch shape of the vector kernel: blocks of 4 warps, 128 blocks per SM in the grid (many short waves of
  // row-strided warps).  Same-box A/B on B200s whose b.copy_(a) bandwidth measured 6594-6602 GB/s, 8M x 768 -> bf16 /
  // 4M x 512 -> fp32, % of that copy bandwidth: grid = exactly the resident set 86 / 93; 8 warps x 64 blocks per
  // SM 88-92 / 93-101; 4 warps x 128 per SM 94 / 101; one row per warp (no grid-stride) 67 / 85.  Also tried and
  // dropped (slower or equal): error-free fp32 arithmetic instead of fp64 (77 / 90), four fp64 chains + Newton rsqrt
  // (89 / 92), streaming stores, L2 prefetch distances 0 / 1 / 3 / 4 rows instead of 2.
  // RBOD_K1_CTAS / RBOD_K1_WARPS override for probing.
  static const int tune_ctas = getenv("RBOD_K1_CTAS") ? atoi(getenv("RBOD_K1_CTAS")) : 128;
  static const int tune_warps = getenv("RBOD_K1_WARPS") ? atoi(getenv("RBOD_K1_WARPS")) : 4;
  const int64_t vwant = (n + tune_warps - 1) / tune_warps;
  const int64_t vcap = (int64_t)num_sms * tune_ctas;
  const int vgrid = (int)(vwant < vcap ? vwant : vcap);
#define RBOD_K1_ARGS in, n, dim, slots_dev, slot0, normalize, cosine, master32, ld32, out16, ld16, kind16, shadow16, out_norms, stats
#define RBOD_K1_CASE(NV)                                                                                     \
  case NV:                                                                                                   \
    if (master32) l2norm_pack_vec_kernel<NV, 1><<<vgrid, tune_warps * 32, 0, st>>>(RBOD_K1_ARGS);            \
    else l2norm_pack_vec_kernel<NV, 0><<<vgrid, tune_warps * 32, 0, st>>>(RBOD_K1_ARGS);                     \
    break;
  if (aligned && dim % 128 == 0 && dim / 128 >= 1 && dim / 128 <= 8) {
    switch (dim / 128) {
      RBOD_K1_CASE(1)
      RBOD_K1_CASE(2)
      RBOD_K1_CASE(3)
      RBOD_K1_CASE(4)
      RBOD_K1_CASE(5)
      RBOD_K1_CASE(6)
      RBOD_K1_CASE(7)
      RBOD_K1_CASE(8)
    }
  } else {
    l2norm_pack_generic_kernel<<<grid, K1_WARPS * 32, 0, st>>>(in, n, dim, slots_dev, slot0, normalize, cosine,
                                                               master32, ld32, out16, ld16, kind16, shadow16,
                                                               out_norms, stats);
  }
#undef RBOD_K1_CASE
#undef RBOD_K1_ARGS
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_build_shadow(const uint16_t* rows16, int64_t n, int dp, uint16_t* shadow16, float* stats, cudaStream_t st) {
  if (n <= 0) return RBOD_OK;
  const int64_t want = (n + 7) / 8;
  const int grid = (int)(want < 148 * 32 ? want : 148 * 32);
  build_shadow_kernel<<<grid, 256, 0, st>>>(rows16, n, dp, shadow16, stats);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_row_bias(const float* master32, const uint16_t* rows16, int kind16, int dim, int64_t ld32, int64_t ld16,
                    const int64_t* slots_dev, int64_t slot0, int64_t n, float* row_bias, cudaStream_t st) {
  if (n <= 0) return RBOD_OK;
  const int64_t want = (n + 7) / 8;
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  row_bias_kernel<<<grid, 256, 0, st>>>(master32, rows16, kind16, dim, ld32, ld16, slots_dev, slot0, n, row_bias);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

int launch_gather_rows(const float* master32, const uint16_t* rows16, int kind16, int dim, int64_t ld32,
                       int64_t ld16, const int64_t* rows, int64_t n, int64_t n_valid, float* out, int* err_flag,
                       cudaStream_t st) {
  if (n <= 0) return RBOD_OK;
  const int64_t want = (n + 7) / 8;
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  gather_rows_kernel<<<grid, 256, 0, st>>>(master32, rows16, kind16, dim, ld32, ld16, rows, n, n_valid, out,
                                           err_flag);
  RBOD_CUDA(cudaGetLastError());
  return RBOD_OK;
}

}  // namespace rbod
