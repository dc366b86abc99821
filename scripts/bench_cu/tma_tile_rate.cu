// Next-round question for the conv A operand: is the TMA's cost per LOAD or per ROW?
// The gather4 im2col sustains one 4-row load per ~24 clk per SM, which bounds the N <= 64 layers.  Where the source rows of
// a tile are consecutive (dense levels, long runs of active pixels) a plain 2-D tiled load moves BR consecutive rows x 128 B
// per instruction.  This measures, in the conv kernel's 3-stage producer/consumer setting, clocks per 256-row x 32-channel
// chunk for  (a) 64 gather4 loads,  (b) 256/BR tiled loads, BR = 4..256,  (c) one 258-row tiled load serving the three
// dx taps of a row (consumers read rows r-1, r, r+1 of the same stage: one third of the loads and of the L2->SM bytes).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_tile_rate tma_tile_rate.cu   (no -lcuda needed)
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
__device__ __forceinline__ void tile2d(uint32_t dst, const CUtensorMap* tm, int col, int row, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(col), "r"(row), "r"(bar)
      : "memory");
}

constexpr int ROWS = 256;
constexpr int STAGE = (ROWS + 8) * 128;            // room for the 258-row halo stage, 1024-byte multiple

// mode 0: gather4 (tm4, box {32,1});  mode 1: tiled loads of BR rows (tmb, box {32,BR});  mode 2: halo - one (ROWS+2)-row
// tiled load per THREE chunks (tmh, box {32, ROWS+2}), the consumers read rows r + dx of the same stage.
__global__ void __launch_bounds__(512, 1) rate_kernel(const __grid_constant__ CUtensorMap tm4, const __grid_constant__ CUtensorMap tmb,
                                                      const __grid_constant__ CUtensorMap tmh, int mode, int BR, int nchunks,
                                                      int nwarps, int R, float* sink, long long* clk) {
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
  const int m0 = 64 + (blockIdx.x * ROWS) % (R - ROWS - 192);
  const int per = mode == 2 ? 3 : 1;                   // chunks served by one stage fill
  const int nfills = (nchunks + per - 1) / per;
  float acc = 0.f;
  const long long t0 = clock64();
  if (warp >= 8 && warp < 8 + nwarps) {                 // producer warps
    const int gw = warp - 8;
    for (int f = 0; f < nfills; ++f) {
      const int st = f % 3;
      const int c = f * per, tap = c % 9, col = (c / 9) * 32 % 992;
      if (f >= 3) mbar_wait(smem_u32(&empty[st]), ((f - 3) / 3) & 1);
      const int off = (tap / 3 - 1) * 64 + (mode == 2 ? -1 : (tap % 3 - 1));
      const uint32_t bytes = mode == 2 ? (ROWS + 2) * 128u : ROWS * 128u;
      if (gw == 0 && elect_one())
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&full[st])), "r"(bytes) : "memory");
      if (mode == 0) {
        for (int g = gw; g < ROWS / 4; g += nwarps) {
          const int q0 = m0 + 4 * g + off;
          if (elect_one()) gather4(smem_u32(smem + st * STAGE + g * 512), &tm4, col, make_int4(q0, q0 + 1, q0 + 2, q0 + 3), smem_u32(&full[st]));
        }
      } else if (mode == 1) {
        for (int g = gw; g < ROWS / BR; g += nwarps)
          if (elect_one()) tile2d(smem_u32(smem + st * STAGE + g * BR * 128), &tmb, col, m0 + g * BR + off, smem_u32(&full[st]));
      } else if (gw == 0) {
        if (elect_one()) tile2d(smem_u32(smem + st * STAGE), &tmh, col, m0 + off, smem_u32(&full[st]));
      }
    }
  } else if (warp < 8) {                                // consumers: read the tile like the split warps do
    for (int c = 0; c < nchunks; ++c) {
      const int f = c / per, st = f % 3;
      mbar_wait(smem_u32(&full[st]), (f / 3) & 1);
      const int row = tid + (mode == 2 ? c % 3 : 0);      // halo: tap dx reads row r + dx (+1 for the stage's leading row)
      const unsigned char* rowp = smem + st * STAGE + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(rowp + ((j ^ (row & 7)) << 4));
        acc += v.x + v.y + v.z + v.w;
      }
      __syncwarp();
      if (lane == 0 && (c % per == per - 1 || c == nchunks - 1))
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&empty[st])) : "memory");
    }
  }
  __syncthreads();
  if (acc == 123.456f && sink) sink[0] = acc;
  if (blockIdx.x == 0 && tid == 0 && clk) clk[0] = clock64() - t0;
}

static bool make_map(EncodeFn encode, CUtensorMap* tm, void* x, int C, int R, int LD, int box_rows) {
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(R)};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(LD) * 4};
  const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(box_rows)}, estr[2] = {1, 1};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main() {
  const int R = 163840, C = 1024, LD = 1024;            // 671 MB of rows: larger than L2
  float* x;
  if (cudaMalloc(&x, static_cast<size_t>(R) * LD * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(x, 0, static_cast<size_t>(R) * LD * 4);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode), cudaEnableDefault, &q) != cudaSuccess || !encode) {
    printf("no cuTensorMapEncodeTiled\n"); return 1;
  }
  CUtensorMap tm4, tmh;
  if (!make_map(encode, &tm4, x, C, R, LD, 1) || !make_map(encode, &tmh, x, C, R, LD, 256)) { printf("encode failed\n"); return 1; }
  // the halo box is 258 rows > the 256-row box limit: use 2 loads of 129 rows instead if the encode of 258 fails
  CUtensorMap tmh258;
  const bool halo_ok = make_map(encode, &tmh258, x, C, R, LD, ROWS + 2);
  const int smem3 = 3 * STAGE + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3);
  long long* dclk; cudaMalloc(&dclk, 8);
  const int nchunks = 288 * 4, nw = 6;
  auto run = [&](const char* name, int mode, int BR, const CUtensorMap& tmb, const CUtensorMap& th) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      rate_kernel<<<148, 512, smem3>>>(tm4, tmb, th, mode, BR, nchunks, nw, R, nullptr, dclk);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: kernel error %s\n", name, cudaGetErrorString(e)); exit(1); }
      cudaEventElapsedTime(&ms, e0, e1);
    }
    long long h = 0; cudaMemcpy(&h, dclk, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %6.0f clk per 256-row chunk, %.3f ms\n", name, double(h) / nchunks, ms);
  };
  run("gather4 x64", 0, 4, tm4, tmh);
  for (int BR : {4, 8, 16, 32, 64, 128, 256}) {
    CUtensorMap tmb;
    if (!make_map(encode, &tmb, x, C, R, LD, BR)) { printf("encode BR=%d failed\n", BR); continue; }
    char name[64]; snprintf(name, sizeof name, "tiled, %d rows per load", BR);
    run(name, 1, BR, tmb, tmh);
  }
  if (halo_ok) run("halo: 258 rows / 3 taps", 2, 0, tm4, tmh258);
  else printf("halo: a 258-row box does not encode (box limit 256): use 2 x 129-row loads\n");
  return 0;
}
