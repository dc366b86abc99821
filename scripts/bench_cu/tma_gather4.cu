// Semantics + rate check of the TMA gather4 load (cp.async.bulk.tensor.2d...tile::gather4) for the conv A operand:
// tensor = rows x C fp32 (row pitch LD floats), box = {32 channels, 1 row}, SWIZZLE_128B.  Each instruction fetches the
// 128-byte slices of 4 arbitrary rows into 512 contiguous bytes of shared memory.  Checks: swizzle pattern == the
// (piece ^ (row & 7)) rule the kernel reads with, rows -1 / >= extent and channels >= C are zero-filled; then the
// clocks per 256-row chunk with W warps issuing.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_gather4 tma_gather4.cu   (no -lcuda needed)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 26) && !done; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  if (!done) __trap();
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* tm, int col, int4 rows, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(col), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(bar)
      : "memory");
}

constexpr int ROWS = 256;
// idx: [nchunks][ROWS] row indices; cols: [nchunks] channel offsets.  mode 0: copy the first chunk's smem image out.
__global__ void __launch_bounds__(512, 1) gather_kernel(const __grid_constant__ CUtensorMap tm, const int* __restrict__ idx,
                                                        const int* __restrict__ cols, int nchunks, int nwarps, float* out,
                                                        long long* clk) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(16) int tab[2][ROWS];
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar[s])), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int c = 0; c < nchunks; ++c) {
    const int st = c & 1;
    if (tid < ROWS) tab[st][tid] = idx[(static_cast<long long>(blockIdx.x) * nchunks + c) * ROWS + tid];
    __syncthreads();
    if (warp < nwarps) {
      if (warp == 0 && elect_one())
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar[st])), "r"(ROWS * 128u) : "memory");
      const int col = cols[c];
      for (int g = warp; g < ROWS / 4; g += nwarps) {
        const int4 r4 = *reinterpret_cast<const int4*>(&tab[st][4 * g]);
        if (elect_one()) gather4(smem_u32(smem + st * ROWS * 128 + g * 512), &tm, col, r4, smem_u32(&bar[st]));
      }
    }
    // consume the previous chunk while this one is in flight
    if (c > 0) {
      mbar_wait(smem_u32(&bar[(c - 1) & 1]), ((c - 1) >> 1) & 1);
      acc += reinterpret_cast<const float*>(smem + ((c - 1) & 1) * ROWS * 128)[tid];
      if (c == 1 && out && blockIdx.x == 0)
        for (int i = tid; i < ROWS * 32; i += blockDim.x) out[i] = reinterpret_cast<const float*>(smem)[i];
    }
    __syncthreads();
  }
  mbar_wait(smem_u32(&bar[(nchunks - 1) & 1]), ((nchunks - 1) >> 1) & 1);
  if (nchunks == 1 && out && blockIdx.x == 0) {
    __syncthreads();
    for (int i = tid; i < ROWS * 32; i += blockDim.x) out[i] = reinterpret_cast<const float*>(smem)[i];
  }
  if (acc == 123.456f && out) out[0] = acc;
  if (blockIdx.x == 0 && tid == 0 && clk) clk[0] = clock64() - t0;
}

// rate kernel closer to the conv pipeline: row indices computed in registers (no global loads, no CTA barrier per
// chunk), 3-stage ring: gather warps run up to two chunks ahead of a consumer warp group that reads the tile.
__global__ void __launch_bounds__(512, 1) rate_kernel(const __grid_constant__ CUtensorMap tm, int nchunks, int nwarps, int R,
                                                      float* sink, long long* clk) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[3], empty[3];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 3; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&full[s])), "r"(1u));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&empty[s])), "r"(8u));
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const int m0 = (blockIdx.x * ROWS) % (R - ROWS);
  float acc = 0.f;
  const long long t0 = clock64();
  if (warp >= 8 && warp < 8 + nwarps) {                 // gather warps
    const int gw = warp - 8;
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % 3, tap = c % 9, col = (c / 9) * 32 % 992;
      if (c >= 3) mbar_wait(smem_u32(&empty[st]), ((c - 3) / 3) & 1);
      if (gw == 0 && elect_one())
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&full[st])), "r"(ROWS * 128u) : "memory");
      const int off = (tap / 3 - 1) * 64 + (tap % 3 - 1);
      for (int g = gw; g < ROWS / 4; g += nwarps) {
        int q0 = m0 + 4 * g + off;
        int4 r4 = make_int4(q0, q0 + 1, q0 + 2, q0 + 3);
        if (q0 < 0 || q0 + 3 >= R) r4 = make_int4(-1, -1, -1, -1);
        if (elect_one()) gather4(smem_u32(smem + st * ROWS * 128 + g * 512), &tm, col, r4, smem_u32(&full[st]));
      }
    }
  } else if (warp < 8) {                                // consumers: read the tile like the split warps do
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % 3;
      mbar_wait(smem_u32(&full[st]), (c / 3) & 1);
      const unsigned char* rowp = smem + st * ROWS * 128 + tid * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(rowp + ((j ^ (tid & 7)) << 4));
        acc += v.x + v.y + v.z + v.w;
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&empty[st])) : "memory");
    }
  }
  __syncthreads();
  if (acc == 123.456f && sink) sink[0] = acc;
  if (blockIdx.x == 0 && tid == 0 && clk) clk[0] = clock64() - t0;
}

int main() {
  const int R = 40960, C = 1000, LD = 1024;          // C < LD and not a multiple of 32: channel tail must zero-fill
  std::vector<float> hx(static_cast<size_t>(R) * LD);
  for (size_t i = 0; i < hx.size(); ++i) hx[i] = static_cast<float>((i / LD) * 0.001 + (i % LD));   // row*0.001 + channel
  float* x; cudaMalloc(&x, hx.size() * 4); cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode), cudaEnableDefault, &q) != cudaSuccess || !encode) {
    printf("no cuTensorMapEncodeTiled\n"); return 1;
  }
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(R)};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(LD) * 4};
  const cuuint32_t box[2] = {32, 1}, estr[2] = {1, 1};
  CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", int(rc)); return 1; }

  const int smem = 2 * ROWS * 128 + 1024;
  cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // ---- semantics
  {
    std::vector<int> hidx(ROWS);
    for (int i = 0; i < ROWS; ++i) hidx[i] = (i * 977 + 13) % R;
    hidx[5] = -1; hidx[6] = R; hidx[7] = R + 12345; hidx[100] = -7; hidx[255] = 0;
    const int hcol = 992;                              // channels 992..1023: only 992..999 exist
    int *didx, *dcol; float* dout;
    cudaMalloc(&didx, ROWS * 4); cudaMalloc(&dcol, 4); cudaMalloc(&dout, ROWS * 32 * 4);
    cudaMemcpy(didx, hidx.data(), ROWS * 4, cudaMemcpyHostToDevice); cudaMemcpy(dcol, &hcol, 4, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0xff, ROWS * 32 * 4);
    gather_kernel<<<1, 512, smem>>>(tm, didx, dcol, 1, 4, dout, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("semantics kernel error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> ho(ROWS * 32);
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < ROWS; ++r)
      for (int k = 0; k < 32; ++k) {
        const int piece = k >> 2, within = k & 3;
        const float got = ho[r * 32 + ((piece ^ (r & 7)) << 2) + within];
        const int src = hidx[r], ch = hcol + k;
        const float want = (src >= 0 && src < R && ch < C) ? hx[static_cast<size_t>(src) * LD + ch] : 0.f;
        if (got != want && bad++ < 8) printf("  mismatch row %d k %d: got %g want %g (src %d)\n", r, k, got, want, src);
      }
    printf("semantics: %d mismatches of %d (swizzle = piece ^ (row & 7); OOB rows/channels zero)\n", bad, ROWS * 32);
  }
  // ---- rate: 148 CTAs x 288 chunks (9 taps x 32 channel chunks, taps innermost), like upconv(4,1)'s skip input
  {
    const int nchunks = 288, grid = 148;
    std::vector<int> hidx(static_cast<size_t>(grid) * nchunks * ROWS), hcols(nchunks);
    for (int b = 0; b < grid; ++b)
      for (int c = 0; c < nchunks; ++c) {
        const int tap = c % 9, m0 = (b * ROWS) % (R - ROWS);
        for (int r = 0; r < ROWS; ++r) {
          int qq = m0 + r + (tap / 3 - 1) * 64 + (tap % 3 - 1);
          hidx[(static_cast<size_t>(b) * nchunks + c) * ROWS + r] = qq < 0 ? -1 : (qq >= R ? -1 : qq);
        }
      }
    for (int c = 0; c < nchunks; ++c) hcols[c] = (c / 9) * 32 % 992;
    int *didx, *dcol; long long* dclk;
    cudaMalloc(&didx, hidx.size() * 4); cudaMalloc(&dcol, nchunks * 4); cudaMalloc(&dclk, 8);
    cudaMemcpy(didx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dcol, hcols.data(), nchunks * 4, cudaMemcpyHostToDevice);
    for (int nw : {1, 2, 4, 6, 8, 16}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        gather_kernel<<<grid, 512, smem>>>(tm, didx, dcol, nchunks, nw, nullptr, dclk);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rate kernel error %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      long long h = 0; cudaMemcpy(&h, dclk, 8, cudaMemcpyDeviceToHost);
      printf("gather4 issued by %2d warps: %6.0f clk per 256-row chunk, %.3f ms, %.0f GB/s aggregate\n", nw, double(h) / nchunks, ms,
             double(grid) * nchunks * ROWS * 128 / ms / 1e6);
    }
  }
  {
    const int smem3 = 3 * ROWS * 128 + 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3);
    long long* dclk; cudaMalloc(&dclk, 8);
    const int nchunks = 288 * 4;
    for (int nw : {1, 2, 4, 6, 8}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        rate_kernel<<<148, 512, smem3>>>(tm, nchunks, nw, R, nullptr, dclk);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rate kernel error %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      long long h = 0; cudaMemcpy(&h, dclk, 8, cudaMemcpyDeviceToHost);
      printf("pipelined (3 stages): gather4 issued by %d warps: %6.0f clk per 256-row chunk, %.3f ms, %.0f GB/s aggregate\n", nw,
             double(h) / nchunks, ms, 148.0 * nchunks * ROWS * 128 / ms / 1e6);
    }
  }
  return 0;
}
