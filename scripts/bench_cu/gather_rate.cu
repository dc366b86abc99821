// Micro-benchmark: how fast can one SM gather 256 rows x 128 bytes (one K chunk of the conv A operand) into shared
// memory?  Models upconv(4,1)'s skip input: rows of 1024 floats, 3x3 taps over a 64-wide image, 32 channel chunks.
// Variants: 0 = 12 warps LDGSTS.16 + barrier per chunk (the producers of conv_tc.cu), 1 = W warps LDGSTS.16 with
// two chunks in flight, 2 = W warps LDG.128 -> STS.128, 3 = one 128-byte cp.async.bulk per row.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_rate gather_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROWS = 256, LDF = 1024, IMG_W = 64;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(512, 1) gather_kernel(const float* __restrict__ x, int total_rows, int variant, int nwarps,
                                                        int tiles_per_cta, long long* out, float* sink, int order) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ int tab[9 * ROWS];
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x, nthr = nwarps * 32;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar[s])), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  long long t_total = 0, chunks = 0;
  uint32_t round = 0;
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int m0 = ((blockIdx.x * tiles_per_cta + t) * ROWS) % (total_rows - ROWS);
    for (int e = tid; e < 9 * ROWS; e += blockDim.x) {
      const int tap = e / ROWS, r = e % ROWS;
      int q = m0 + r + (tap / 3 - 1) * IMG_W + (tap % 3 - 1);
      q = min(max(q, 0), total_rows - 1);
      tab[e] = q;
    }
    __syncthreads();
    const long long t0 = clock64();
    if (tid < nthr) {
      const int nch = 9 * (LDF / 32);
      if (variant == 3) {
        for (int c = 0; c < nch; ++c, ++round) {
          const int tap = order ? c % 9 : c / 32, ci0 = (order ? c / 9 : c % 32) * 32;   // order 1: taps innermost
          unsigned char* st = smem + (round & 1) * ROWS * 128;
          const uint32_t b = smem_u32(&bar[round & 1]);
          if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(ROWS * 128u) : "memory");
          asm volatile("bar.sync 1, %0;\n" ::"r"(nthr) : "memory");
          for (int r = tid; r < ROWS; r += nthr) {
            const float* src = x + static_cast<long long>(tab[tap * ROWS + r]) * LDF + ci0;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(st + r * 128)),
                         "l"(src), "r"(128u), "r"(b) : "memory");
          }
          mbar_wait(b, (round >> 1) & 1);
          acc += reinterpret_cast<const float*>(st)[tid];
        }
      } else if (variant == 2) {
        for (int c = 0; c < nch; ++c, ++round) {
          const int tap = order ? c % 9 : c / 32, ci0 = (order ? c / 9 : c % 32) * 32;   // order 1: taps innermost
          unsigned char* st = smem + (round & 1) * ROWS * 128;
          float4 v[16];
          const int per = ROWS * 8 / nthr;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < per) {
              const int p = tid + i * nthr, r = p >> 3, j = p & 7;
              v[i] = __ldg(reinterpret_cast<const float4*>(x + static_cast<long long>(tab[tap * ROWS + r]) * LDF + ci0) + j);
            }
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < per) {
              const int p = tid + i * nthr, r = p >> 3, j = p & 7;
              *reinterpret_cast<float4*>(st + r * 128 + ((j ^ (r & 7)) << 4)) = v[i];
            }
          asm volatile("bar.sync 1, %0;\n" ::"r"(nthr) : "memory");
          acc += reinterpret_cast<const float*>(st)[tid];
        }
      } else {
        const int depth = variant == 0 ? 0 : 1;
        for (int c = 0; c < nch; ++c, ++round) {
          const int tap = order ? c % 9 : c / 32, ci0 = (order ? c / 9 : c % 32) * 32;   // order 1: taps innermost
          unsigned char* st = smem + (round & 1) * ROWS * 128;
          for (int p = tid; p < ROWS * 8; p += nthr) {
            const int r = p >> 3, j = p & 7;
            cp16(st + r * 128 + ((j ^ (r & 7)) << 4), reinterpret_cast<const float4*>(x + static_cast<long long>(tab[tap * ROWS + r]) * LDF + ci0) + j);
          }
          asm volatile("cp.async.commit_group;\n" ::: "memory");
          if (depth == 0) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
          else asm volatile("cp.async.wait_group 1;\n" ::: "memory");
          asm volatile("bar.sync 1, %0;\n" ::"r"(nthr) : "memory");
          acc += reinterpret_cast<const float*>(smem + ((round - depth) & 1) * ROWS * 128)[tid];
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      }
      t_total += clock64() - t0;
      chunks += 9 * (LDF / 32);
    }
    __syncthreads();
  }
  if (acc == 123.456f) sink[0] = acc;
  if (blockIdx.x == 0 && tid == 0) { out[0] = t_total; out[1] = chunks; }
}

int main() {
  const int total_rows = 40960;
  float* x; long long* d_out; float* sink;
  cudaMalloc(&x, sizeof(float) * static_cast<size_t>(total_rows) * LDF);
  cudaMemset(x, 0, sizeof(float) * static_cast<size_t>(total_rows) * LDF);
  cudaMalloc(&d_out, 16); cudaMalloc(&sink, 16);
  const int smem = 2 * ROWS * 128 + 1024;
  cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Cfg { int variant, nwarps, order, grid; };
  Cfg cfgs[] = {{0, 12, 0, 148}, {1, 12, 0, 148}, {1, 4, 0, 148}, {0, 12, 1, 148}, {1, 12, 1, 148}, {1, 4, 1, 148}, {1, 8, 1, 148}, {1, 2, 1, 148},
                {2, 8, 1, 148}, {3, 8, 1, 148}, {1, 4, 1, 37}, {1, 12, 1, 37}, {1, 4, 0, 37}, {1, 4, 1, 8}};
  for (auto c : cfgs) {
    long long h[2] = {0, 0};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      gather_kernel<<<c.grid, 512, smem>>>(x, total_rows, c.variant, c.nwarps, 2, d_out, sink, c.order);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    const double bytes = double(c.grid) * 2 * 288 * ROWS * 128;
    printf("variant %d warps %2d order %d grid %3d : %7.0f clk per chunk (CTA 0), %.3f ms, %.0f GB/s aggregate gather\n", c.variant, c.nwarps, c.order, c.grid,
           double(h[0]) / double(h[1]), ms, bytes / ms / 1e6);
  }
  return 0;
}
