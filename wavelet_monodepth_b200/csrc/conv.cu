// K5: gather-GEMM convolution on pixel-major feature rows (fp32 SIMT, cp.async pipeline).
//
//   y[m, co] = act(bias[co] + sum_{tap, c} in(p_m + tap)[c] * w[tap][c][co])
//
// One CTA owns a BM x BN output tile (BM active pixels x BN output channels) and walks the reduction
// K = taps x (c0 + c1) in BK-wide chunks.  For every chunk the A operand is an *implicit im2col tile*:
// BM gathered row segments of BK contiguous channels, fetched straight into shared memory with 16-byte
// cp.async (zero-filled for inactive / padded taps and channel tails), so the gather costs no registers
// and overlaps the FMAs of the previous chunks (3-stage ring).  The per-tap source rows of the tile are
// resolved once per tile into a shared table (index map lookup, border rule, gate), which is the only
// integer work on the path - the reference spends 64 % of its time materialising these indices
// (BASELINE.md profile: gather 41 %, int64 add 23 %).  Persistent grid sized from the SM count; the
// active-row count is read on the device, so no host sync is needed between levels.
#include "common.cuh"

namespace wmd {

template <int BM_, int BN_, int BK_, int TM_, int TN_>
struct ConvCfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
  static constexpr int STAGES = 3;
  static constexpr int TY = BM / TM, TX = BN / TN;
  static constexpr int THREADS = TY * TX;
  static constexpr int A_LD = BK + 4;              // +4 floats: rows of one warp land in distinct banks
  static constexpr int A_STAGE = BM * A_LD;        // floats
  static constexpr int B_STAGE = BK * BN;
  static constexpr int SEG_A = BK / 4;             // 16-byte segments per A row
  static constexpr int SEG_B = BN / 4;
  static constexpr size_t SMEM = static_cast<size_t>(STAGES) * (A_STAGE + B_STAGE) * sizeof(float) +
                                 static_cast<size_t>(2) * 9 * BM * sizeof(int32_t);
  static_assert(THREADS == 256, "tile configs are written for 256 threads");
  static_assert(TN % 4 == 0 && BK % 4 == 0, "vector widths");
  static_assert((THREADS % SEG_A) == 0 && (THREADS % SEG_B) == 0, "loader mapping");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 2) conv_rows_kernel(const wmd_conv_desc d) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, TM = Cfg::TM, TN = Cfg::TN;
  constexpr int STAGES = Cfg::STAGES, A_LD = Cfg::A_LD, THREADS = Cfg::THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);
  float* Bs = As + STAGES * Cfg::A_STAGE;
  int32_t* tab0 = reinterpret_cast<int32_t*>(Bs + STAGES * Cfg::B_STAGE);
  int32_t* tab1 = tab0 + 9 * BM;

  const int tid = threadIdx.x;
  const int tx = tid % Cfg::TX, ty = tid / Cfg::TX;
  const long long HW = static_cast<long long>(d.H) * d.W;
  const int total_px = static_cast<int>(static_cast<long long>(d.N) * HW);
  int rows = d.pixels ? *d.count : total_px;
  rows = min(rows, d.max_rows);
  const int ctot = d.c0 + d.c1;
  const int nch0 = (d.c0 + BK - 1) / BK, nch1 = (d.c1 + BK - 1) / BK;
  const int per_tap = nch0 + nch1;
  const int nchunks = d.taps * per_tap;
  const int n_tiles = (d.cout + BN - 1) / BN;
  const long long tiles = static_cast<long long>((rows + BM - 1) / BM) * n_tiles;
  const int Hs = d.H >> d.shift0, Ws = d.W >> d.shift0;
  const bool aligned_rows = (d.taps == 1 && d.map0 == nullptr);

  // loader coordinates (fixed per thread)
  const int a_seg = tid % Cfg::SEG_A, a_row0 = tid / Cfg::SEG_A;
  constexpr int A_ROWS_PER_PASS = THREADS / Cfg::SEG_A;
  const int b_seg = tid % Cfg::SEG_B, b_row0 = tid / Cfg::SEG_B;
  constexpr int B_ROWS_PER_PASS = THREADS / Cfg::SEG_B;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int m0 = static_cast<int>(tile / n_tiles) * BM;
    const int n0 = static_cast<int>(tile % n_tiles) * BN;

    // ---- resolve the source rows of every (tap, tile row) once
    for (int e = tid; e < d.taps * BM; e += THREADS) {
      const int tap = e / BM, r = e - tap * BM;
      const int m = m0 + r;
      int32_t r0 = -1, r1 = -1;
      if (m < rows) {
        const int p = d.pixels ? d.pixels[m] : m;
        const int n = static_cast<int>(p / HW);
        const int rem = static_cast<int>(p - n * HW);
        const int y = rem / d.W, x = rem - y * d.W;
        int qy = y, qx = x;
        if (d.taps == 9) { qy += tap / 3 - 1; qx += tap % 3 - 1; }
        bool ok = pad_coord(qy, d.H, d.pad_mode);
        ok = pad_coord(qx, d.W, d.pad_mode) && ok;
        if (ok) {
          const int q = (n * d.H + qy) * d.W + qx;
          if (d.gate && !d.gate[q]) ok = false;
          if (ok) {
            r1 = d.map1 ? d.map1[q] : q;
            if (aligned_rows) {
              r0 = m;
            } else {
              const int qs = (n * Hs + (qy >> d.shift0)) * Ws + (qx >> d.shift0);
              r0 = d.map0 ? d.map0[qs] : qs;
            }
          }
        }
      }
      tab0[e] = r0;
      tab1[e] = r1;
    }
    __syncthreads();

    auto load_chunk = [&](int c, int stage) {
      const int tap = c / per_tap;
      const int rr = c - tap * per_tap;
      const bool src1 = rr >= nch0;
      const int ci0 = (src1 ? rr - nch0 : rr) * BK;
      const int csrc = src1 ? d.c1 : d.c0;
      const float* xb = src1 ? d.x1 : d.x0;
      const int ld = src1 ? d.ld1 : d.ld0;
      const int32_t* tab = (src1 ? tab1 : tab0) + tap * BM;
      float* as = As + stage * Cfg::A_STAGE;
      const int ci = ci0 + a_seg * 4;
      const int a_bytes = max(0, min(16, (csrc - ci) * 4));
#pragma unroll
      for (int r = a_row0; r < BM; r += A_ROWS_PER_PASS) {
        const int32_t row = tab[r];
        const bool live = row >= 0 && a_bytes > 0;
        const float* src = live ? xb + static_cast<long long>(row) * ld + ci : xb;
        cp_async16(as + r * A_LD + a_seg * 4, src, live ? a_bytes : 0);
      }
      float* bs = Bs + stage * Cfg::B_STAGE;
      const int kw = tap * ctot + (src1 ? d.c0 : 0) + ci0;
      const int kvalid = min(BK, csrc - ci0);
      const int co = n0 + b_seg * 4;
      const int b_bytes = max(0, min(16, (d.ldw - co) * 4));
#pragma unroll
      for (int k = b_row0; k < BK; k += B_ROWS_PER_PASS) {
        const bool live = k < kvalid && b_bytes > 0;
        const float* src = live ? d.w + static_cast<long long>(kw + k) * d.ldw + co : d.w;
        cp_async16(bs + k * BN + b_seg * 4, src, live ? b_bytes : 0);
      }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // ---- prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nchunks) load_chunk(s, s);
      cp_async_commit();
    }
    // ---- main loop
    for (int c = 0; c < nchunks; ++c) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      const int nxt = c + STAGES - 1;
      if (nxt < nchunks) load_chunk(nxt, nxt % STAGES);
      cp_async_commit();
      const float* as = As + (c % STAGES) * Cfg::A_STAGE;
      const float* bs = Bs + (c % STAGES) * Cfg::B_STAGE;
#pragma unroll
      for (int kk = 0; kk < BK; kk += 4) {
        float4 a4[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i)
          a4[i] = *reinterpret_cast<const float4*>(as + (ty + Cfg::TY * i) * A_LD + kk);
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
          float b[TN];
#pragma unroll
          for (int j = 0; j < TN; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(bs + (kk + kq) * BN + tx * TN + j);
            b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
          }
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            const float a = kq == 0 ? a4[i].x : (kq == 1 ? a4[i].y : (kq == 2 ? a4[i].z : a4[i].w));
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a, b[j], acc[i][j]);
          }
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();   // all stages + tables free for the next tile

    // ---- epilogue: bias, activation, row store
    const int co0 = n0 + tx * TN;
    float bv[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = (d.bias && co0 + j < d.cout) ? __ldg(d.bias + co0 + j) : 0.f;
    const bool vec_ok = (d.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15) == 0);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty + Cfg::TY * i;
      if (m >= rows) continue;
      float* yr = d.y + static_cast<long long>(m) * d.ldy;
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const int co = co0 + j;
        float4 v;
        v.x = activate(acc[i][j] + bv[j], d.act, d.act_param);
        v.y = activate(acc[i][j + 1] + bv[j + 1], d.act, d.act_param);
        v.z = activate(acc[i][j + 2] + bv[j + 2], d.act, d.act_param);
        v.w = activate(acc[i][j + 3] + bv[j + 3], d.act, d.act_param);
        if (vec_ok && co + 3 < d.cout) {
          *reinterpret_cast<float4*>(yr + co) = v;
        } else {
          if (co < d.cout) yr[co] = v.x;
          if (co + 1 < d.cout) yr[co + 1] = v.y;
          if (co + 2 < d.cout) yr[co + 2] = v.z;
          if (co + 3 < d.cout) yr[co + 3] = v.w;
        }
      }
    }
  }
}

using CfgWide = ConvCfg<128, 128, 32, 8, 8>;   // cout >= 96
using CfgMid = ConvCfg<128, 64, 32, 8, 4>;     // 48 <= cout < 96
using CfgThin = ConvCfg<256, 32, 16, 8, 4>;    // cout < 48

template <class Cfg>
static int launch_conv(const wmd_conv_desc& d, cudaStream_t stream) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {   // outside the cache: set it on every launch
    int rc = record(cudaFuncSetAttribute(conv_rows_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(Cfg::SMEM)));
    if (rc != WMD_OK) return rc;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const long long tiles = static_cast<long long>(ceil_div(d.max_rows, Cfg::BM)) * ceil_div(d.cout, Cfg::BN);
  const long long cap = static_cast<long long>(sm_count()) * 2;
  const int grid = static_cast<int>(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
  conv_rows_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(d);
  return launched();
}

}  // namespace wmd

extern "C" int wmd_conv_rows_f32(const wmd_conv_desc* dp, wmd_stream_t stream) {
  using namespace wmd;
  WMD_REQUIRE(dp, WMD_ERR_ARG);
  wmd_conv_desc d = *dp;
  WMD_REQUIRE(d.x0 && d.w && d.y, WMD_ERR_ARG);
  WMD_REQUIRE(d.taps == 1 || d.taps == 9, WMD_ERR_ARG);
  WMD_REQUIRE(d.pad_mode >= WMD_PAD_ZERO && d.pad_mode <= WMD_PAD_REPLICATE, WMD_ERR_ARG);
  WMD_REQUIRE(d.act >= WMD_ACT_NONE && d.act <= WMD_ACT_SIGMOID, WMD_ERR_ARG);
  WMD_REQUIRE(d.shift0 == 0 || d.shift0 == 1, WMD_ERR_ARG);
  WMD_REQUIRE((d.pixels == nullptr) == (d.count == nullptr), WMD_ERR_ARG);
  WMD_REQUIRE(d.N > 0 && d.H > 0 && d.W > 0 && d.c0 > 0 && d.cout > 0 && d.max_rows >= 0, WMD_ERR_SHAPE);
  WMD_REQUIRE(static_cast<long long>(d.N) * d.H * d.W < (1ll << 31), WMD_ERR_SHAPE);
  if (d.x1 == nullptr) { d.c1 = 0; d.ld1 = 0; }
  WMD_REQUIRE(d.c1 >= 0 && (d.c1 == 0 || d.x1), WMD_ERR_ARG);
  // 16-byte cp.async granularity
  WMD_REQUIRE(d.ld0 >= d.c0 && d.ld0 % 4 == 0 && (reinterpret_cast<uintptr_t>(d.x0) & 15) == 0, WMD_ERR_SHAPE);
  WMD_REQUIRE(d.c1 == 0 || (d.ld1 >= d.c1 && d.ld1 % 4 == 0 && (reinterpret_cast<uintptr_t>(d.x1) & 15) == 0),
              WMD_ERR_SHAPE);
  WMD_REQUIRE(d.ldw >= d.cout && d.ldw % 4 == 0 && (reinterpret_cast<uintptr_t>(d.w) & 15) == 0, WMD_ERR_SHAPE);
  WMD_REQUIRE(d.ldy >= d.cout, WMD_ERR_SHAPE);
  if (d.shift0 == 1) WMD_REQUIRE(d.H % 2 == 0 && d.W % 2 == 0, WMD_ERR_SHAPE);
  if (d.pad_mode == WMD_PAD_REFLECT && d.taps == 9) WMD_REQUIRE(d.H >= 2 && d.W >= 2, WMD_ERR_SHAPE);
  if (d.max_rows == 0) return WMD_OK;
  if (d.cout >= 96) return launch_conv<CfgWide>(d, as_stream(stream));
  if (d.cout >= 48) return launch_conv<CfgMid>(d, as_stream(stream));
  return launch_conv<CfgThin>(d, as_stream(stream));
}
