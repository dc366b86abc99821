// K3 / K4: per-sample range threshold, the six per-level pixel sets, and batch-wide mask compaction.
// All HBM-bound byte/int work: one pass over the inputs, coalesced, counts never leave the device.
#include "common.cuh"

namespace wmd {

// ------------------------------------------------------------------------------------ range -> threshold
constexpr int kRangeThreads = 256;
constexpr int kRangeMaxBlocks = 64;
// The ticket counters live in a FIXED-size prefix of the workspace (independent of N), so the "left zeroed"
// invariant survives calls with different batch sizes sharing one scratch buffer.
constexpr int kRangeMaxN = 16384;
constexpr size_t kRangeCounterBytes = static_cast<size_t>(kRangeMaxN) * sizeof(unsigned);

static inline int range_blocks(long long per_sample) {
  long long b = (per_sample + 4095) / 4096;
  if (b < 1) b = 1;
  if (b > kRangeMaxBlocks) b = kRangeMaxBlocks;
  return static_cast<int>(b);
}

// torch.max / torch.min PROPAGATE NaN (fminf / fmaxf drop it): a NaN anywhere in yl makes the reference's threshold
// NaN and every `> thresh` test false (depth_decoder.py:308-309).  Same here.
__device__ __forceinline__ float nan_min(float a, float b) { return (a != a || b != b) ? NAN : fminf(a, b); }
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a || b != b) ? NAN : fmaxf(a, b); }

__device__ __forceinline__ void block_minmax(float& mn, float& mx, float* smn, float* smx) {
  for (int o = 16; o > 0; o >>= 1) {
    mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    mn = lane < nw ? smn[lane] : INFINITY;
    mx = lane < nw ? smx[lane] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) {
      mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
  }
  __syncthreads();
}

// grid (B, N).  Each block reduces a strided slice of sample n; the last block to finish for that
// sample (ticket counter, self-resetting) folds the B partials and writes thresh[n].
__global__ void __launch_bounds__(kRangeThreads) range_thresh_kernel(const float* __restrict__ x, long long per_sample,
                                                                     float ratio, float* __restrict__ thresh,
                                                                     float* __restrict__ minmax,
                                                                     unsigned* __restrict__ counters,
                                                                     float* __restrict__ partial) {
  __shared__ float smn[8], smx[8];
  __shared__ bool is_last;
  const int n = blockIdx.y, b = blockIdx.x, B = gridDim.x;
  const float* xs = x + static_cast<long long>(n) * per_sample;
  float mn = INFINITY, mx = -INFINITY;
  const bool vec = (per_sample % 4 == 0) && ((reinterpret_cast<uintptr_t>(xs) & 15) == 0);
  if (vec) {
    const long long nv = per_sample >> 2;
    const float4* xv = reinterpret_cast<const float4*>(xs);
    for (long long i = static_cast<long long>(b) * blockDim.x + threadIdx.x; i < nv;
         i += static_cast<long long>(B) * blockDim.x) {
      const float4 v = __ldg(xv + i);
      mn = nan_min(nan_min(mn, v.x), nan_min(v.y, nan_min(v.z, v.w)));
      mx = nan_max(nan_max(mx, v.x), nan_max(v.y, nan_max(v.z, v.w)));
    }
  } else {
    for (long long i = static_cast<long long>(b) * blockDim.x + threadIdx.x; i < per_sample;
         i += static_cast<long long>(B) * blockDim.x) {
      const float v = __ldg(xs + i);
      mn = nan_min(mn, v);
      mx = nan_max(mx, v);
    }
  }
  block_minmax(mn, mx, smn, smx);
  if (threadIdx.x == 0) {
    volatile float* pp = partial + (static_cast<long long>(n) * B + b) * 2;
    pp[0] = mn; pp[1] = mx;
    __threadfence();
    const unsigned ticket = atomicAdd(&counters[n], 1u);
    is_last = (ticket == static_cast<unsigned>(B - 1));
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  mn = INFINITY; mx = -INFINITY;
  const volatile float* pr = partial + static_cast<long long>(n) * B * 2;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    mn = nan_min(mn, pr[2 * i]);
    mx = nan_max(mx, pr[2 * i + 1]);
  }
  block_minmax(mn, mx, smn, smx);
  if (threadIdx.x == 0) {
    thresh[n] = __fmul_rn(__fsub_rn(mx, mn), ratio);
    if (minmax) { minmax[2 * n] = mn; minmax[2 * n + 1] = mx; }
    counters[n] = 0u;   // leave the workspace ready for the next call
  }
}

// ------------------------------------------------------------------------------------ level masks
constexpr int kMT = 32;          // low-res tile edge
constexpr int kMHalo = 2;
constexpr int kMS = kMT + 2 * kMHalo;

__global__ void __launch_bounds__(256) level_masks_kernel(const float* __restrict__ yh, const float* __restrict__ thresh,
                                                          uint8_t* __restrict__ s0, uint8_t* __restrict__ s1,
                                                          uint8_t* __restrict__ s2, uint8_t* __restrict__ s3,
                                                          uint8_t* __restrict__ s4, uint8_t* __restrict__ s5, int H,
                                                          int W) {
  __shared__ uint8_t t0[kMS][kMS + 4];
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kMT, x0 = blockIdx.x * kMT;
  const long long HW = static_cast<long long>(H) * W;
  const float* ph = yh ? yh + static_cast<long long>(n) * 3 * HW : nullptr;
  const float th = thresh ? thresh[n] : 0.f;
  for (int e = threadIdx.x; e < kMS * kMS; e += blockDim.x) {
    const int ly = e / kMS, lx = e % kMS;
    const int y = y0 + ly - kMHalo, x = x0 + lx - kMHalo;
    uint8_t v = 0;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      if (!thresh) {
        v = 1;
      } else {
        const long long o = static_cast<long long>(y) * W + x;
        // torch.abs(yh).max(2)[0] > thresh: the band maximum propagates NaN, and NaN > thresh is false
        const float m = nan_max(nan_max(fabsf(__ldg(ph + o)), fabsf(__ldg(ph + HW + o))), fabsf(__ldg(ph + 2 * HW + o)));
        v = m > th ? 1 : 0;
      }
    }
    t0[ly][lx] = v;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + tx;
  if (x >= W) return;
  const long long lo_base = static_cast<long long>(n) * HW;
  const long long hi_base = static_cast<long long>(n) * 4 * HW;
  const int W2 = 2 * W;
  for (int r = ty; r < kMT; r += 8) {
    const int y = y0 + r;
    if (y >= H) break;
    const int cy = r + kMHalo, cx = tx + kMHalo;
    unsigned rows3[5];   // per-row OR over the 3-wide and 5-wide windows
    unsigned rows5[5];
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
      const uint8_t* row = t0[cy + dy];
      const unsigned c3 = row[cx - 1] | row[cx] | row[cx + 1];
      rows3[dy + 2] = c3;
      rows5[dy + 2] = c3 | row[cx - 2] | row[cx + 2];
    }
    const uint8_t v0 = t0[cy][cx];
    const uint8_t v1 = static_cast<uint8_t>(rows3[1] | rows3[2] | rows3[3]);
    const uint8_t v2 = static_cast<uint8_t>(rows5[0] | rows5[1] | rows5[2] | rows5[3] | rows5[4]);
    const long long lo = lo_base + static_cast<long long>(y) * W + x;
    if (s0) s0[lo] = v0;
    if (s1) s1[lo] = v1;
    if (s2) s2[lo] = v2;
    // high-resolution sets: pixel (2y+a, 2x+b).  S5 = S0; S3 = dilate5(up2(S0)) = up2(S1);
    // S4 = dilate3(up2(S0)) = OR of S0 over rows {y+a-1, y+a} x cols {x+b-1, x+b}.
    const uint8_t l = t0[cy][cx - 1], rgt = t0[cy][cx + 1];
    const uint8_t up_l = t0[cy - 1][cx - 1], up_c = t0[cy - 1][cx], up_r = t0[cy - 1][cx + 1];
    const uint8_t dn_l = t0[cy + 1][cx - 1], dn_c = t0[cy + 1][cx], dn_r = t0[cy + 1][cx + 1];
    const long long hi = hi_base + static_cast<long long>(2 * y) * W2 + 2 * x;
    if (s5) {
      *reinterpret_cast<uchar2*>(s5 + hi) = make_uchar2(v0, v0);
      *reinterpret_cast<uchar2*>(s5 + hi + W2) = make_uchar2(v0, v0);
    }
    if (s3) {
      *reinterpret_cast<uchar2*>(s3 + hi) = make_uchar2(v1, v1);
      *reinterpret_cast<uchar2*>(s3 + hi + W2) = make_uchar2(v1, v1);
    }
    if (s4) {
      const uint8_t a00 = v0 | l | up_c | up_l, a01 = v0 | rgt | up_c | up_r;
      const uint8_t a10 = v0 | l | dn_c | dn_l, a11 = v0 | rgt | dn_c | dn_r;
      *reinterpret_cast<uchar2*>(s4 + hi) = make_uchar2(a00, a01);
      *reinterpret_cast<uchar2*>(s4 + hi + W2) = make_uchar2(a10, a11);
    }
  }
}

// ------------------------------------------------------------------------------------ compaction
constexpr int kCThreads = 256;
constexpr int kCPer = 8;                       // pixels per thread
constexpr int kCTile = kCThreads * kCPer;      // 2048 pixels per block

__device__ __forceinline__ unsigned load_mask8(const uint8_t* __restrict__ mask, long long base, long long total,
                                               uint8_t (&m)[kCPer]) {
  unsigned c = 0;
  if (base + kCPer <= total) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(mask + base));
    const unsigned w[2] = {v.x, v.y};
#pragma unroll
    for (int k = 0; k < kCPer; ++k) {
      m[k] = ((w[k >> 2] >> (8 * (k & 3))) & 0xffu) ? 1 : 0;
      c += m[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < kCPer; ++k) {
      m[k] = (base + k < total && mask[base + k]) ? 1 : 0;
      c += m[k];
    }
  }
  return c;
}

__device__ __forceinline__ unsigned block_sum(unsigned v, unsigned* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  unsigned t = 0;
  for (int w = 0; w < (kCThreads >> 5); ++w) t += sh[w];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kCThreads) compact_count_kernel(const uint8_t* __restrict__ mask, long long total,
                                                                  unsigned* __restrict__ block_counts) {
  __shared__ unsigned sh[kCThreads / 32];
  uint8_t m[kCPer];
  const long long base = static_cast<long long>(blockIdx.x) * kCTile + static_cast<long long>(threadIdx.x) * kCPer;
  const unsigned c = base < total ? load_mask8(mask, base, total, m) : 0u;
  const unsigned t = block_sum(c, sh);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kCThreads) compact_fill_kernel(const uint8_t* __restrict__ mask, long long total,
                                                                 long long HW, const unsigned* __restrict__ block_counts,
                                                                 int32_t* __restrict__ idxmap,
                                                                 int32_t* __restrict__ pixels,
                                                                 int32_t* __restrict__ offsets, int N) {
  __shared__ unsigned sh[kCThreads / 32];
  __shared__ unsigned warp_excl[kCThreads / 32];
  // rows contributed by all earlier blocks (redundant per-block prefix: <= a few thousand ints from L2)
  unsigned before = 0;
  for (int b = threadIdx.x; b < static_cast<int>(blockIdx.x); b += kCThreads) before += block_counts[b];
  before = block_sum(before, sh);

  uint8_t m[kCPer];
  const long long base = static_cast<long long>(blockIdx.x) * kCTile + static_cast<long long>(threadIdx.x) * kCPer;
  const unsigned c = base < total ? load_mask8(mask, base, total, m) : 0u;
  // exclusive scan of c over the block
  unsigned incl = c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) sh[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int w = 0; w < kCThreads / 32; ++w) { warp_excl[w] = run; run += sh[w]; }
  }
  __syncthreads();
  unsigned row = before + warp_excl[warp] + incl - c;
  if (base < total) {
    // per-sample offsets: one 64-bit division per thread (not two per pixel); `next` = first pixel of the next sample
    long long smp = 0, next = 0;
    if (offsets) {
      smp = base / HW;
      next = smp * HW;
      if (next < base) { ++smp; next += HW; }        // the next sample boundary at or after base
    }
    int32_t idx[kCPer];
#pragma unroll
    for (int k = 0; k < kCPer; ++k) {
      const long long p = base + k;
      idx[k] = -1;
      if (p < total) {
        if (offsets && p == next) { offsets[smp] = static_cast<int32_t>(row); ++smp; next += HW; }
        if (m[k]) {
          idx[k] = static_cast<int32_t>(row);
          if (pixels) pixels[row] = static_cast<int32_t>(p);
          ++row;
        }
        if (offsets && p == total - 1) offsets[N] = static_cast<int32_t>(row);
      }
    }
    if (idxmap) {
      if (base + kCPer <= total && (reinterpret_cast<uintptr_t>(idxmap) & 15) == 0) {   // base is a multiple of 8: 32-byte aligned
        int4* o = reinterpret_cast<int4*>(idxmap + base);
        o[0] = make_int4(idx[0], idx[1], idx[2], idx[3]);
        o[1] = make_int4(idx[4], idx[5], idx[6], idx[7]);
      } else {
#pragma unroll
        for (int k = 0; k < kCPer; ++k)
          if (base + k < total) idxmap[base + k] = idx[k];
      }
    }
  }
}

__global__ void gate_map_kernel(const uint8_t* __restrict__ gate, const int32_t* __restrict__ idxmap,
                                int32_t* __restrict__ out, long long count) {
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < count; p += step)
    out[p] = gate[p] ? (idxmap ? idxmap[p] : static_cast<int32_t>(p)) : -1;
}

}  // namespace wmd

// ---------------------------------------------------------------------------------------- C ABI
extern "C" size_t wmd_range_ws_bytes(int N, long long per_sample) {
  if (N <= 0) return wmd::kRangeCounterBytes;
  return wmd::kRangeCounterBytes + static_cast<size_t>(N) * wmd::range_blocks(per_sample) * 2 * sizeof(float);
}

extern "C" int wmd_range_thresh_f32(const float* x, int N, long long per_sample, float ratio, float* thresh,
                                    float* minmax, void* ws, size_t ws_bytes, wmd_stream_t stream) {
  using namespace wmd;
  WMD_REQUIRE(x && thresh && ws, WMD_ERR_ARG);
  WMD_REQUIRE(N >= 0 && N <= kRangeMaxN && per_sample > 0, WMD_ERR_SHAPE);
  if (N == 0) return WMD_OK;
  WMD_REQUIRE(ws_bytes >= wmd_range_ws_bytes(N, per_sample), WMD_ERR_WORKSPACE);
  unsigned* cnt = static_cast<unsigned*>(ws);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(ws) + kRangeCounterBytes);
  dim3 grid(range_blocks(per_sample), N);
  range_thresh_kernel<<<grid, kRangeThreads, 0, as_stream(stream)>>>(x, per_sample, ratio, thresh, minmax, cnt, partial);
  return launched();
}

extern "C" int wmd_level_masks(const float* yh, const float* thresh, uint8_t* s0, uint8_t* s1, uint8_t* s2,
                               uint8_t* s3, uint8_t* s4, uint8_t* s5, int N, int H, int W, wmd_stream_t stream) {
  using namespace wmd;
  WMD_REQUIRE(thresh == nullptr || yh != nullptr, WMD_ERR_ARG);
  WMD_REQUIRE(N >= 0 && H > 0 && W > 0, WMD_ERR_SHAPE);
  WMD_REQUIRE(N <= 65535, WMD_ERR_SHAPE);
  if (N == 0) return WMD_OK;
  dim3 grid(ceil_div(W, kMT), ceil_div(H, kMT), N);
  level_masks_kernel<<<grid, 256, 0, as_stream(stream)>>>(yh, thresh, s0, s1, s2, s3, s4, s5, H, W);
  return launched();
}

extern "C" size_t wmd_compact_ws_bytes(int N, int H, int W) {
  const long long total = static_cast<long long>(N) * H * W;
  const long long blocks = (total + wmd::kCTile - 1) / wmd::kCTile;
  return static_cast<size_t>(blocks < 1 ? 1 : blocks) * sizeof(unsigned);
}

extern "C" int wmd_compact_mask(const uint8_t* mask, int32_t* idxmap, int32_t* pixels, int32_t* offsets, int N, int H,
                                int W, void* ws, size_t ws_bytes, wmd_stream_t stream) {
  using namespace wmd;
  WMD_REQUIRE(N >= 0 && H > 0 && W > 0, WMD_ERR_SHAPE);
  if (N == 0) {                                   // an empty shard (world size > batch): no rows, offsets = [0]
    if (offsets) return record(cudaMemsetAsync(offsets, 0, sizeof(int32_t), as_stream(stream)));
    return WMD_OK;
  }
  WMD_REQUIRE(mask && ws, WMD_ERR_ARG);
  const long long total = static_cast<long long>(N) * H * W;
  WMD_REQUIRE(total < (1ll << 31), WMD_ERR_SHAPE);
  WMD_REQUIRE(ws_bytes >= wmd_compact_ws_bytes(N, H, W), WMD_ERR_WORKSPACE);
  WMD_REQUIRE((reinterpret_cast<uintptr_t>(mask) & 7) == 0, WMD_ERR_SHAPE);
  const int blocks = ceil_div(total, kCTile);
  unsigned* bc = static_cast<unsigned*>(ws);
  compact_count_kernel<<<blocks, kCThreads, 0, as_stream(stream)>>>(mask, total, bc);
  int rc = launched();
  if (rc != WMD_OK) return rc;
  compact_fill_kernel<<<blocks, kCThreads, 0, as_stream(stream)>>>(mask, total, static_cast<long long>(H) * W, bc,
                                                                  idxmap, pixels, offsets, N);
  return launched();
}

extern "C" int wmd_gate_map(const uint8_t* gate, const int32_t* idxmap, int32_t* out, long long count,
                            wmd_stream_t stream) {
  using namespace wmd;
  WMD_REQUIRE(gate && out, WMD_ERR_ARG);
  if (count <= 0) return WMD_OK;
  gate_map_kernel<<<stride_grid(count, 256), 256, 0, as_stream(stream)>>>(gate, idxmap, out, count);
  return launched();
}
