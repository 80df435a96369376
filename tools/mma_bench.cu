// Microbenchmark: issue cost of tcgen05.mma.cta_group::1.kind::f16 (bf16, K = 16) on sm_100a as a function of
// the instruction shape, with A from shared memory (SS) and A from tensor memory (TS).  One CTA per SM, one elected
// thread issues a chain of `iters` MMAs into one accumulator, commit, wait; cycles = clock64 delta / iters.
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hairfastgan_b200/csrc \
//                    tools/mma_bench.cu -o gpurun_out/mma_bench -lcuda && gpurun_out/mma_bench
#include <cstdio>
#include <vector>

#include "hf_common.cuh"

using namespace hf;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int m, int n, int ts_mode, int iters, int a_stride_k,
                                                           long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_f16(HF_BF16, m, n);
    const uint32_t hi = kmajor_desc_hi(1024, UMMA_LAYOUT_SW128);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t koff = (uint32_t)((i & 3) * 32 * a_stride_k);     // walk the 4 K16 slices of the 64-wide chunk
      if (ts_mode) umma_ts(tmem_base, tmem_base + 256 + (i & 3) * 8, kmajor_desc(hi, b_addr + koff), idesc, i > 0);
      else umma_f16(tmem_base, kmajor_desc(hi, a_addr + koff), kmajor_desc(hi, b_addr + koff), idesc, i > 0);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * sms);
  const int smem_bytes = 16384 + 32768 + 2048;
  cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int iters = 4096;
  std::vector<long long> h(sms);
  printf("mode  M    N    cycles/MMA   MAC/cycle/SM   (grid = %d CTAs, %d MMAs each, bf16 K=16)\n", sms, iters);
  for (int ts = 0; ts < 2; ++ts)
    for (int m : {64, 128})
      for (int n : {16, 32, 64, 96, 128, 192, 256}) {
        if (m == 128 && n % 16) continue;
        for (int rep = 0; rep < 2; ++rep) {
          mma_bench_kernel<<<sms, 128, smem_bytes>>>(m, n, ts, iters, 1, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s M=%d N=%d: %s\n", ts ? "TS" : "SS", m, n, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (long long v : h) avg += (double)v;
        avg /= sms * (double)iters;
        printf("%s   %4d %4d   %9.1f   %10.0f\n", ts ? "TS" : "SS", m, n, avg, (double)m * n * 16 / avg);
      }
  cudaFree(d_out);
  return 0;
}
