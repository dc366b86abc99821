// K5-TC: the gather-GEMM convolution on 5th-gen tensor cores (tcgen05 + TMEM), fp32-faithful via 3xTF32.
//
// Same contract as conv_rows_kernel (conv.cu), different engine.  One CTA per SM owns a 256 x 128 output tile
// (two UMMA M=128 halves sharing one B tile) and walks K = taps x (c0 + c1) in 32-channel chunks:
//   * A (implicit im2col: 256 gathered rows x 32 channels) is fetched with 16-byte zero-filling cp.async straight
//     into 128B-swizzled K-major shared-memory tiles - the layout a TMA tile load would have produced, which TMA
//     cannot do here because the rows come from an index-map gather.  The 128-byte line of each row piece is
//     prefetched into L2 several chunks ahead (prefetch.global.L2), so the cp.async itself is an L2 hit;
//   * B (weights) is pre-split and pre-swizzled by wmd_pack_conv_weight_tc_f32 into one [hi | lo] image per
//     (n-tile, chunk): a single 32 KB cp.async.bulk by one thread, completing on an mbarrier (async proxy);
//   * each of the 16 warps' threads then splits exactly the 16-byte A pieces it copied (visible to it after
//     cp.async.wait_group) into tf32 "hi" (in place) and "lo": x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi);
//   * one elected thread issues per chunk 2 halves x 4 k-steps x 3 tcgen05.mma.kind::tf32 (lo*hi + hi*lo + hi*hi)
//     into fp32 accumulators in TMEM and commits to an mbarrier that frees the stage (2-stage ring);
//   * the tensor core's fp32 accumulation rounds toward zero, a bias that grows linearly with K (measured
//     ~6.5e-9 * K relative).  So accumulation runs in EPOCHS of kFlushChunks chunks that alternate between two
//     TMEM column sets; a finished epoch is drained (tcgen05.ld) into per-thread fp32 registers with ordinary
//     round-to-nearest adds while the next epoch's MMAs run.  Each thread ends up owning one output row x 64
//     channels: bias + activation + one contiguous 256-byte row store.
#include "common.cuh"

namespace wmd {

constexpr int TC_BM = 256;                      // rows per CTA tile = 2 UMMA halves of 128
constexpr int TC_BN = 128;
constexpr int TC_BK = 32;                       // floats per chunk = one 128-byte swizzle-atom row
constexpr int TC_STAGES = 2;
constexpr int TC_THREADS = 512;                 // 16 warps: all gather/split; warp w drains TMEM lane quarter w&3
constexpr int TC_A_HALF = 128 * TC_BK * 4;      // 16 KB: one M=128 half of A (hi or lo)
constexpr int TC_B_TILE = TC_BN * TC_BK * 4;    // 16 KB (hi or lo)
constexpr int TC_STAGE = 4 * TC_A_HALF + 2 * TC_B_TILE;            // [A0hi A1hi A0lo A1lo Bhi Blo] = 96 KB
constexpr int TC_TABLES = 2 * 9 * TC_BM * 4;
constexpr size_t TC_SMEM = static_cast<size_t>(TC_STAGES) * TC_STAGE + TC_TABLES + 1024;
constexpr int TC_TMEM_COLS = 512;               // 2 epoch sets x 2 halves x 128 columns
constexpr int kFlushChunks = 32;                // epoch length: K = 1024 per TMEM accumulation run
constexpr int kPrefetchAhead = 4;               // chunks of L2 prefetch distance
constexpr uint32_t kNoRow = 0xFFFFFFFFu;        // tap-table entry of an inactive / padded source

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"); rows are 128 bytes, 8-row groups
// are 1024 bytes apart (SBO), LBO is the canonical 1 for swizzled K-major operands.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// in place: *ph <- hi, *pl <- lo
__device__ __forceinline__ void split_piece(float4* ph, float4* pl) {
  const float4 v = *ph;
  float4 h, l;
  h.x = tf32_rna(v.x); l.x = tf32_rna(v.x - h.x);
  h.y = tf32_rna(v.y); l.y = tf32_rna(v.y - h.y);
  h.z = tf32_rna(v.z); l.z = tf32_rna(v.z - h.z);
  h.w = tf32_rna(v.w); l.w = tf32_rna(v.w - h.w);
  *ph = h;
  *pl = l;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1) conv_rows_tc_kernel(const wmd_conv_desc d, const float* __restrict__ wtc) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bar_mma[TC_STAGES];     // stage consumed by the tensor pipe
  __shared__ __align__(8) uint64_t bar_b[TC_STAGES];       // weight image of the stage has landed (bulk copy)
  __shared__ __align__(8) uint64_t bar_epoch[2];           // accumulation epoch complete (per TMEM set)
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  uint32_t* tab0 = reinterpret_cast<uint32_t*>(base + TC_STAGES * TC_STAGE);   // 16-byte-unit offsets into x0
  uint32_t* tab1 = tab0 + 9 * TC_BM;                                            // ... into x1

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(smem_u32(&bar_mma[s]), 1);
      mbar_init(smem_u32(&bar_b[s]), 1);
    }
    mbar_init(smem_u32(&bar_epoch[0]), 1);
    mbar_init(smem_u32(&bar_epoch[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(static_cast<uint32_t>(TC_TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;

  const long long HW = static_cast<long long>(d.H) * d.W;
  const int total_px = static_cast<int>(static_cast<long long>(d.N) * HW);
  int rows = d.pixels ? *d.count : total_px;
  rows = min(rows, d.max_rows);
  const int nch0 = (d.c0 + TC_BK - 1) / TC_BK, nch1 = (d.c1 + TC_BK - 1) / TC_BK;
  const int per_tap = nch0 + nch1;
  const int nchunks = d.taps * per_tap;
  const int n_tiles = (d.cout + TC_BN - 1) / TC_BN;
  const long long tiles = static_cast<long long>((rows + TC_BM - 1) / TC_BM) * n_tiles;
  const int Hs = d.H >> d.shift0, Ws = d.W >> d.shift0;
  const bool aligned_rows = (d.taps == 1 && d.map0 == nullptr);
  const uint32_t ld0q = static_cast<uint32_t>(d.ld0 >> 2), ld1q = static_cast<uint32_t>(d.ld1 >> 2);
  // instruction descriptor: D=f32, A=B=tf32, both K-major, N = 128, M = 128
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(TC_BN >> 3) << 17) |
                              (static_cast<uint32_t>(128 >> 4) << 24);

  // UMMA descriptors are loop invariant up to a +2 (32 bytes >> 4) per k-step in the start-address field: build
  // them once so the single issuing thread spends ~3 instructions per tcgen05.mma instead of ~30
  uint64_t dsc[TC_STAGES][6];   // [A0hi A1hi A0lo A1lo Bhi Blo]
#pragma unroll
  for (int s = 0; s < TC_STAGES; ++s) {
    const uint32_t sA = smem_u32(base + s * TC_STAGE);
    dsc[s][0] = umma_desc_sw128(sA);
    dsc[s][1] = umma_desc_sw128(sA + TC_A_HALF);
    dsc[s][2] = umma_desc_sw128(sA + 2 * TC_A_HALF);
    dsc[s][3] = umma_desc_sw128(sA + 3 * TC_A_HALF);
    dsc[s][4] = umma_desc_sw128(sA + 4 * TC_A_HALF);
    dsc[s][5] = umma_desc_sw128(sA + 4 * TC_A_HALF + TC_B_TILE);
  }

  // producer mapping: thread -> 16-byte piece a_j of rows a_r0 + 64*i (i < 4); row r lives in half r>>7
  const int a_j = tid & 7, a_r0 = tid >> 3;
  uint32_t use0 = 0, use1 = 0;                 // MMA commits issued so far per stage (phase tracking)
  uint32_t ep_use0 = 0, ep_use1 = 0;           // epoch commits per TMEM set
  // accumulator ownership: TMEM lane quarter, M half, 64-column half
  const int my_q = warp & 3, my_half = (warp >> 2) & 1, my_ch = warp >> 3;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int m0 = static_cast<int>(tile / n_tiles) * TC_BM;
    const int nt = static_cast<int>(tile % n_tiles);
    const int n0 = nt * TC_BN;

    for (int e = tid; e < d.taps * TC_BM; e += TC_THREADS) {
      const int tap = e / TC_BM, r = e - tap * TC_BM;
      const int m = m0 + r;
      uint32_t o0 = kNoRow, o1 = kNoRow;
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
            o1 = static_cast<uint32_t>(q) * ld1q;
            int r0;
            if (aligned_rows) {
              r0 = m;
            } else {
              const int qs = (n * Hs + (qy >> d.shift0)) * Ws + (qx >> d.shift0);
              r0 = d.map0 ? d.map0[qs] : qs;
            }
            if (r0 >= 0) o0 = static_cast<uint32_t>(r0) * ld0q;
          }
        }
      }
      tab0[e] = o0;
      tab1[e] = o1;
    }
    __syncthreads();

    const unsigned char* wtile = reinterpret_cast<const unsigned char*>(wtc) +
                                 static_cast<long long>(nt) * nchunks * (2 * TC_B_TILE);

    // decode chunk c -> gather source (as float4 pointer), channels left, offset table
    auto chunk_src = [&](int c, const float4*& xq, int& cleft) -> const uint32_t* {
      const int tap = c / per_tap;
      const int rr = c - tap * per_tap;
      const bool src1 = rr >= nch0;
      const int ci0 = (src1 ? rr - nch0 : rr) * TC_BK;
      cleft = (src1 ? d.c1 : d.c0) - ci0;                                   // channels available from ci0 on
      xq = reinterpret_cast<const float4*>(src1 ? d.x1 : d.x0) + (ci0 >> 2);
      return (src1 ? tab1 : tab0) + tap * TC_BM;
    };

    auto prefetch_chunk = [&](int c) {
      if (a_j != 0) return;                      // one 128-byte line per row
      const float4* xq; int cleft;
      const uint32_t* tab = chunk_src(c, xq, cleft);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = tab[a_r0 + 64 * i];
        if (off != kNoRow) prefetch_l2(xq + off);
      }
    };

    auto load_chunk = [&](int c, int stage) {
      const float4* xq; int cleft;
      const uint32_t* tab = chunk_src(c, xq, cleft);
      unsigned char* sA = base + stage * TC_STAGE;
      const int a_bytes = max(0, min(16, (cleft - a_j * 4) * 4));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = a_r0 + 64 * i;             // 0..255; half = r >> 7
        const uint32_t off = tab[r];
        const bool live = off != kNoRow && a_bytes > 0;
        const float4* src = live ? xq + off + a_j : xq;
        cp_async16(sA + (r >> 7) * TC_A_HALF + (r & 127) * 128 + ((a_j ^ (r & 7)) << 4), src, live ? a_bytes : 0);
      }
      // weights: pre-split, pre-swizzled [Bhi | Blo] image of (n-tile, chunk): one 32 KB bulk copy
      if (tid == 0)
        bulk_g2s(smem_u32(sA + 4 * TC_A_HALF), wtile + static_cast<long long>(c) * (2 * TC_B_TILE), 2 * TC_B_TILE,
                 smem_u32(&bar_b[stage]));
    };

    auto split_chunk = [&](int stage) {
      unsigned char* sA = base + stage * TC_STAGE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = a_r0 + 64 * i;
        unsigned char* p = sA + (r >> 7) * TC_A_HALF + (r & 127) * 128 + ((a_j ^ (r & 7)) << 4);
        split_piece(reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(p + 2 * TC_A_HALF));
      }
    };

    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;

    // drains this thread's slice (its row, 64 columns) of TMEM set `set` into acc with round-to-nearest adds
    auto drain = [&](int set) {
      const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(my_q * 32) << 16) +
                             static_cast<uint32_t>(set * 256 + my_half * 128 + my_ch * 64);
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + cc, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[cc + j] += __uint_as_float(v[j]);
      }
    };

    // ---- prologue: first chunk in flight, first prefetches
    for (int c = 1; c <= kPrefetchAhead && c < nchunks; ++c) prefetch_chunk(c);
    if (use0 > 0) mbar_wait(smem_u32(&bar_mma[0]), (use0 - 1) & 1);     // previous tile's MMAs left stage 0
    load_chunk(0, 0);
    cp_async_commit();

    for (int c = 0; c < nchunks; ++c) {
      const int stage = c & 1;
      const int epoch = c / kFlushChunks, set = epoch & 1;
      const uint32_t fills = (stage == 0) ? use0 : use1;     // loads into this stage before this one == MMA rounds so far
      cp_async_wait<0>();
      split_chunk(stage);
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        mbar_wait(smem_u32(&bar_b[stage]), fills & 1);        // this stage's weight image has landed
        tc_fence_after();
        const bool first = (c % kFlushChunks) == 0;
        const uint64_t* ds = stage == 0 ? dsc[0] : dsc[1];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t dcol = tmem_acc + static_cast<uint32_t>(set * 256 + half * 128);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t ko = static_cast<uint64_t>(2 * k);   // 8 tf32 = 32 bytes along K inside the swizzle atom
            umma_tf32(dcol, ds[2 + half] + ko, ds[4] + ko, kIdesc, (first && k == 0) ? 0u : 1u);   // lo*hi first
            umma_tf32(dcol, ds[half] + ko, ds[5] + ko, kIdesc, 1u);                                   // hi*lo
            umma_tf32(dcol, ds[half] + ko, ds[4] + ko, kIdesc, 1u);                                   // hi*hi
          }
        }
        umma_commit(smem_u32(&bar_mma[stage]));
        if ((c + 1) % kFlushChunks == 0 || c == nchunks - 1) umma_commit(smem_u32(&bar_epoch[set]));
      }
      if (stage == 0) use0 += 1; else use1 += 1;
      const bool epoch_end = ((c + 1) % kFlushChunks == 0) || (c == nchunks - 1);
      if (epoch_end) { if (set == 0) ep_use0 += 1; else ep_use1 += 1; }

      // refill the other stage (chunk c+1) once chunk c-1's MMAs have drained it; keep L2 warm further ahead
      if (c + 1 < nchunks) {
        const uint32_t u = (stage == 0) ? use1 : use0;
        if (u > 0) mbar_wait(smem_u32(&bar_mma[stage ^ 1]), (u - 1) & 1);
        load_chunk(c + 1, stage ^ 1);
        if (c + 1 + kPrefetchAhead < nchunks) prefetch_chunk(c + 1 + kPrefetchAhead);
      }
      cp_async_commit();

      // a finished epoch (other than the last, handled below) is drained while the next epoch's MMAs run
      if (epoch_end && c != nchunks - 1) {
        const uint32_t eu = (set == 0) ? ep_use0 : ep_use1;
        mbar_wait(smem_u32(&bar_epoch[set]), (eu - 1) & 1);
        tc_fence_after();
        drain(set);
        tc_fence_before();
      }
    }
    cp_async_wait<0>();

    // ---- last epoch + epilogue: bias, activation, one contiguous 256-byte store per thread
    {
      const int set = ((nchunks - 1) / kFlushChunks) & 1;
      const uint32_t eu = (set == 0) ? ep_use0 : ep_use1;
      mbar_wait(smem_u32(&bar_epoch[set]), (eu - 1) & 1);
      tc_fence_after();
      drain(set);
      const int m = m0 + my_half * 128 + my_q * 32 + lane;
      if (m < rows) {
        float* yr = d.y + static_cast<long long>(m) * d.ldy;
        const bool vec_ok = (d.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15) == 0);
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const int co = n0 + my_ch * 64 + j;
          if (co < d.cout) {
            float4 o;
            o.x = activate(acc[j] + (d.bias ? __ldg(d.bias + min(co, d.cout - 1)) : 0.f), d.act, d.act_param);
            o.y = activate(acc[j + 1] + (d.bias ? __ldg(d.bias + min(co + 1, d.cout - 1)) : 0.f), d.act, d.act_param);
            o.z = activate(acc[j + 2] + (d.bias ? __ldg(d.bias + min(co + 2, d.cout - 1)) : 0.f), d.act, d.act_param);
            o.w = activate(acc[j + 3] + (d.bias ? __ldg(d.bias + min(co + 3, d.cout - 1)) : 0.f), d.act, d.act_param);
            if (vec_ok && co + 3 < d.cout) {
              *reinterpret_cast<float4*>(yr + co) = o;
            } else {
              yr[co] = o.x;
              if (co + 1 < d.cout) yr[co + 1] = o.y;
              if (co + 2 < d.cout) yr[co + 2] = o.z;
              if (co + 3 < d.cout) yr[co + 3] = o.w;
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // accumulators drained by every warp, tap tables free
    tc_fence_after();
  }

  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_acc),
                 "r"(static_cast<uint32_t>(TC_TMEM_COLS))
                 : "memory");
  }
}

// w (Cout, Cin, taps) fp32 -> per (n-tile, chunk) smem image [tf32 hi: 128 x 32 | tf32 lo: 128 x 32], K-major,
// 128B-swizzled.  Chunk order = the kernel's: tap-major, then source-0 channel chunks, then source-1 chunks.
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int c0, int c1,
                                      int taps, long long total) {
  const int nch0 = (c0 + TC_BK - 1) / TC_BK, nch1 = (c1 + TC_BK - 1) / TC_BK;
  const int per_tap = nch0 + nch1;
  const int nchunks = taps * per_tap;
  const int Cin = c0 + c1;
  const long long tile_floats = static_cast<long long>(TC_BN) * TC_BK;
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += step) {
    // i indexes LOGICAL (nt, chunk, hilo, n, kk); the store address applies the swizzle
    long long t = i;
    const int kk = static_cast<int>(t % TC_BK); t /= TC_BK;
    const int n = static_cast<int>(t % TC_BN); t /= TC_BN;
    const int hilo = static_cast<int>(t % 2); t /= 2;
    const int c = static_cast<int>(t % nchunks);
    const int nt = static_cast<int>(t / nchunks);
    const int tap = c / per_tap;
    const int rr = c - tap * per_tap;
    const bool src1 = rr >= nch0;
    const int ci_local = (src1 ? rr - nch0 : rr) * TC_BK + kk;
    const int csrc = src1 ? c1 : c0;
    const int co = nt * TC_BN + n;
    float v = 0.f;
    if (ci_local < csrc && co < Cout) {
      const int ci = (src1 ? c0 : 0) + ci_local;
      v = __ldg(w + (static_cast<long long>(co) * Cin + ci) * taps + tap);
    }
    const float hi = tf32_rna(v);
    const float val = hilo == 0 ? hi : tf32_rna(v - hi);
    const long long tile_base = ((static_cast<long long>(nt) * nchunks + c) * 2 + hilo) * tile_floats;
    const int piece = kk >> 2, within = kk & 3;
    out[tile_base + static_cast<long long>(n) * TC_BK + ((piece ^ (n & 7)) << 2) + within] = val;
  }
}

}  // namespace wmd

extern "C" int wmd_conv_tc_tile_n(int cout) { (void)cout; return wmd::TC_BN; }

extern "C" size_t wmd_conv_tc_weight_floats(int cout, int c0, int c1, int taps) {
  using namespace wmd;
  const int nchunks = taps * ((c0 + TC_BK - 1) / TC_BK + (c1 + TC_BK - 1) / TC_BK);
  return static_cast<size_t>(ceil_div(cout, TC_BN)) * nchunks * 2 * TC_BN * TC_BK;
}

extern "C" int wmd_pack_conv_weight_tc_f32(const float* w, float* packed, int Cout, int c0, int c1, int taps,
                                           wmd_stream_t stream) {
  using namespace wmd;
  WMD_REQUIRE(w && packed, WMD_ERR_ARG);
  WMD_REQUIRE(Cout > 0 && c0 > 0 && c1 >= 0 && (taps == 1 || taps == 9), WMD_ERR_SHAPE);
  const long long total = static_cast<long long>(wmd_conv_tc_weight_floats(Cout, c0, c1, taps));
  pack_weight_tc_kernel<<<stride_grid(total, 256), 256, 0, as_stream(stream)>>>(w, packed, Cout, c0, c1, taps, total);
  return launched();
}

extern "C" int wmd_conv_rows_tc_f32(const wmd_conv_desc* dp, wmd_stream_t stream) {
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
  WMD_REQUIRE(d.ld0 >= d.c0 && d.ld0 % 4 == 0 && (reinterpret_cast<uintptr_t>(d.x0) & 15) == 0, WMD_ERR_SHAPE);
  WMD_REQUIRE(d.c1 == 0 || (d.ld1 >= d.c1 && d.ld1 % 4 == 0 && (reinterpret_cast<uintptr_t>(d.x1) & 15) == 0),
              WMD_ERR_SHAPE);
  WMD_REQUIRE((reinterpret_cast<uintptr_t>(d.w) & 15) == 0, WMD_ERR_SHAPE);
  WMD_REQUIRE(d.ldy >= d.cout, WMD_ERR_SHAPE);
  if (d.shift0 == 1) WMD_REQUIRE(d.H % 2 == 0 && d.W % 2 == 0, WMD_ERR_SHAPE);
  if (d.pad_mode == WMD_PAD_REFLECT && d.taps == 9) WMD_REQUIRE(d.H >= 2 && d.W >= 2, WMD_ERR_SHAPE);
  // tap tables hold 32-bit offsets in 16-byte units
  WMD_REQUIRE(static_cast<long long>(d.N) * d.H * d.W * (d.ld0 / 4) < (1ll << 32) &&
                  static_cast<long long>(d.N) * d.H * d.W * (d.ld1 / 4) < (1ll << 32),
              WMD_ERR_UNSUPPORTED);
  if (d.max_rows == 0) return WMD_OK;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    int rc = record(cudaFuncSetAttribute(conv_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(TC_SMEM)));
    if (rc != WMD_OK) return rc;
    attr_done[dev] = true;
  }
  const long long tiles = static_cast<long long>(ceil_div(d.max_rows, TC_BM)) * ceil_div(d.cout, TC_BN);
  const long long cap = sm_count();
  const int grid = static_cast<int>(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
  conv_rows_tc_kernel<<<grid, TC_THREADS, TC_SMEM, as_stream(stream)>>>(d, d.w);
  return launched();
}
