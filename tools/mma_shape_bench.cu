// Microbenchmark (bring-up tool, not product): sustained tcgen05.mma throughput under the power cap
// as a function of instruction shape.  One CTA per SM; one elected lane issues back-to-back MMAs
// (no TMA, operands are whatever is in smem/TMEM), alternating between two accumulators; commits
// every 16 MMAs and waits so the pipe never holds more than 2 batches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_shape_bench tools/mma_shape_bench.cu
#include <cstdio>
#include <cstdlib>
#include "../retrieval_based_object_detection_b200/csrc/rbod_common.cuh"

using namespace rbod;

template <int N, int TS>
__global__ void __launch_bounds__(128, 1) bench_kernel(long long iters, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(1, 1, 128, N);
    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem));
    const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem) + 65536);
    uint32_t phase[2] = {0, 0};
    const unsigned long long t0 = clock64();
    for (long long it = 0; it < iters; ++it) {
      const int b = (int)(it & 1);
      if (it >= 2) { mbar_wait(&bar[b], phase[b]); phase[b] ^= 1u; }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem + 256 + (N <= 128 ? b * N : 0);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (TS) mma_f16_ts(d, tmem + (uint32_t)(i * 8), bdesc + (uint64_t)((i & 3) * 2), idesc, i > 0);
          else mma_f16_ss(d, adesc + (uint64_t)((i & 3) * 2), bdesc + (uint64_t)((i & 3) * 2), idesc, i > 0);
        }
        mma_commit(&bar[b]);
      }
      __syncwarp();
    }
    mbar_wait(&bar[0], phase[0]);
    mbar_wait(&bar[1], phase[1]);
    if (threadIdx.x == 32 && blockIdx.x == 0) *cycles = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int TS>
void run(const char* name, double seconds) {
  unsigned long long* d_cyc;
  cudaMalloc(&d_cyc, 8);
  const size_t smem = 1024 + 65536 + 32768;
  cudaFuncSetAttribute(bench_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  // calibrate
  long long iters = 20000;
  bench_kernel<N, TS><<<148, 128, smem>>>(iters, d_cyc);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench_kernel<N, TS><<<148, 128, smem>>>(iters, d_cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  iters = (long long)(iters * (seconds * 1e3 / ms));
  cudaEventRecord(e0);
  bench_kernel<N, TS><<<148, 128, smem>>>(iters, d_cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long cyc = 0;
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  const double flops = 2.0 * 128 * N * 16 * 16.0 * (double)iters * 148;
  fflush(stdout);
  printf("%-22s N=%3d %s : %8.1f TFLOP/s  (%.2f s, %.0f MHz avg, %.1f cyc/MMA, err=%s)\n", name, N, TS ? "A=tmem" : "A=smem",
         flops / (ms * 1e-3) / 1e12, ms * 1e-3, cyc / (ms * 1e3), (double)cyc / (iters * 16.0),
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_cyc);
}

int main(int argc, char** argv) {
  const double s = argc > 1 ? atof(argv[1]) : 2.0;
  run<64, 1>("ts", s);
  run<128, 1>("ts", s);
  run<256, 1>("ts", s);
  run<64, 0>("ss", s);
  run<128, 0>("ss", s);
  run<256, 0>("ss", s);
  run<64, 1>("ts (again)", s);
  return 0;
}
