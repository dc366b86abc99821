// Micro-benchmark: issue-to-completion cost of tcgen05.mma (SS form) per kind / N on one CTA per SM.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}

// kind: 0 = tf32 (K=8), 1 = bf16 (K=16).  ndist = number of distinct accumulators cycled through.
template <int KIND>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int ndist, int ts_mode, long long* out, int nissuers, int rot) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~static_cast<uintptr_t>(1023));
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"((unsigned)nissuers));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const int wq = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t tm = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (wq < nissuers) {
    const uint64_t adesc = desc_sw128(smem_u32(base)), bdesc = desc_sw128(smem_u32(base + 16384));
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tm + static_cast<uint32_t>(((i % ndist) + wq * ndist) * N);
      const uint64_t ro = rot ? static_cast<uint64_t>((i & 3) * 2) : 0ull;  // rotate k-offset inside the atom
      uint32_t el;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(el));
      if (el) {
      if (ts_mode) {
        const uint32_t a = tm + 448u;   // some columns used as the A operand
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bdesc + ro), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bdesc + ro), "r"(idesc), "r"(1u) : "memory");
      } else {
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(adesc + ro), "l"(bdesc + ro), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(adesc + ro), "l"(bdesc + ro), "r"(idesc), "r"(1u) : "memory");
      }
      }
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* d_out; long long h[2];
  cudaMalloc(&d_out, 16);
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  struct Cfg { int kind, ts, N, ndist, niss, rot; };
  Cfg cfgs[] = {{1,0,64,1,1,0},{1,0,64,4,1,0},{1,0,64,1,2,0},{1,0,64,1,4,0},{1,0,128,1,4,0},{1,0,128,1,2,1},{1,0,256,1,1,1},{1,0,256,1,2,0},
                {0,0,128,1,2,0},{0,0,128,1,4,0},{0,0,256,1,2,0},{0,1,64,1,4,0},{0,0,32,1,1,0},{1,0,16,1,1,0}};
  for (auto c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      if (c.kind == 0) mma_rate_kernel<0><<<148, 128, smem>>>(c.N, iters, c.ndist, c.ts, d_out, c.niss, c.rot);
      else mma_rate_kernel<1><<<148, 128, smem>>>(c.N, iters, c.ndist, c.ts, d_out, c.niss, c.rot);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    const double k = c.kind == 0 ? 8 : 16;
    printf("%s %s N%3d acc/issuer=%d issuers=%d rot=%d : issue %.1f, complete %.1f clk per mma per issuer -> %.0f FLOP/clk/SM\n",
           c.kind == 0 ? "tf32" : "bf16", c.ts ? "TS" : "SS", c.N, c.ndist, c.niss, c.rot, double(h[0]) / iters, double(h[1]) / iters,
           2.0 * 128 * c.N * k * iters * c.niss / double(h[1]));
  }
  return 0;
}
