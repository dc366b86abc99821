// K5-TC: the gather-GEMM convolution on 5th-gen tensor cores (tcgen05 + TMEM), fp32-faithful via 3xTF32.
//
// Same contract as conv_rows_kernel (conv.cu), different engine.  One CTA per SM owns a 256 x 128 output tile
// (two UMMA M=128 halves sharing one B tile) and walks K = taps x (c0 + c1) in 32-channel chunks.
// Warp-specialised, mbarrier-pipelined (no CTA-wide barrier inside the K loop):
//
//   producers (warps 0-11)
//     * gather A (implicit im2col: 256 rows x 32 channels) with 16-byte zero-filling cp.async into a raw fp32,
//       128B-swizzled shared tile (lanes = adjacent pieces of a row: 4 lines per warp request; rows come from an
//       index-map gather, which a TMA tile load cannot express), L2-prefetched a few chunks ahead;
//     * then each thread owns (row, 8 channels) slots: reads them back (the swizzle makes this transposed read
//       conflict-free), splits x = hi + lo (hi = x with the 13 low mantissa bits cleared, lo = x - hi: exact) and
//       writes hi / lo into TENSOR MEMORY with tcgen05.st - the MMAs take A from TMEM (".ts" form), because an
//       SS-mode M=128 x N=128 MMA would need the full 128 B/clk of shared-memory bandwidth three times per k-step;
//   issuers (lane 0 of warps 12-15, one k-step each)
//     * a single thread can only issue one tcgen05.mma per ~200 clk (measured, scripts/bench_cu/mma_rate.cu) while
//       a 128x128x8 tf32 MMA occupies the tensor pipe for 64: four issuers keep it fed (83 % of peak in isolation);
//     * per chunk each issues 2 halves x 3 terms (lo*hi + hi*lo + hi*hi) and commits to the mbarriers that recycle
//       the TMEM A stage / shared B stage; B (weights) is pre-split, pre-swizzled by wmd_pack_conv_weight_tc_f32
//       into one [hi | lo] image per (n-tile, chunk): a single 32 KB cp.async.bulk into a 4-deep ring;
//   all 16 warps
//     * the tensor core's fp32 accumulation rounds toward zero, a bias that grows linearly with K (measured
//       ~6.5e-9 * K relative).  So accumulation runs in EPOCHS of kFlushChunks chunks: a finished epoch is drained
//       (tcgen05.ld) into per-thread fp32 registers with round-to-nearest adds and the TMEM accumulators are
//       re-zeroed.  Each thread ends up owning one output row x 64 channels: bias + activation + one 256-byte store.
// TMEM map (512 columns): [0,256) accumulators (half h at h*128), [256,512) A operand: stage s, half h at
// 256 + s*128 + h*64, hi in the first 32 columns, lo in the next 32.
#include "common.cuh"

namespace wmd {

constexpr int TC_BM = 256;                      // rows per CTA tile = 2 UMMA halves of 128
constexpr int TC_BK = 32;                       // floats per chunk = one 128-byte swizzle-atom row
constexpr int TC_THREADS = 512;                 // 16 warps: 12 producers + 4 issuers; all drain (lane quarter w&3, half (w>>2)&1, cols w>>3)
constexpr int TC_PROD_WARPS = 12;
constexpr int TC_PROD_THREADS = TC_PROD_WARPS * 32;
constexpr int TC_A_STAGES = 2;                  // raw A tiles in shared memory (and split A stages in TMEM)
constexpr int TC_B_STAGES = 4;                  // [Bhi | Blo] images in shared memory
constexpr int TC_A_TILE = TC_BM * TC_BK * 4;    // 32 KB raw fp32
constexpr int TC_TABLES = 2 * 9 * TC_BM * 4;
constexpr int TC_TMEM_COLS = 512;

// Per N-tile configuration.  A single thread issues one tcgen05.mma per ~200 clk whatever its size, so narrow
// tiles need more issuers to keep the tensor pipe fed; every issuer gets a PRIVATE accumulator copy (one writer per
// accumulator => the order of the round-toward-zero accumulations, and with it every output bit, is fixed) and the
// copies are summed in registers at drain time.  Accumulators: (half h, copy j) at column (h*COPIES + j)*BN <= 256.
template <int BN>
struct TcCfg {
  static constexpr int COPIES = BN >= 128 ? 1 : 2;
  static constexpr int ISSUERS = 2 * COPIES;               // issuer i: half i / COPIES, k-steps (i % COPIES) + n*COPIES
  static constexpr int B_TILE = BN * TC_BK * 4;            // bytes of one of hi / lo
  static constexpr int ACC = BN / 2;                       // accumulator registers per thread: its row x BN/2 columns
  static constexpr size_t SMEM = static_cast<size_t>(TC_A_STAGES) * TC_A_TILE +
                                 static_cast<size_t>(TC_B_STAGES) * 2 * B_TILE + TC_TABLES + 1024;
  static_assert(2 * COPIES * BN <= 256, "accumulators must leave 256 TMEM columns for the A operand");
};
constexpr int kFlushChunks = 32;                // epoch length: K = 1024 per TMEM accumulation run
constexpr int kPrefetchAhead = 4;               // chunks of L2 prefetch distance
constexpr uint32_t kNoRow = 0xFFFFFFFFu;        // tap-table entry of an inactive / padded source


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
#ifdef WMD_TC_DEBUG
__device__ unsigned int g_dbg[4];
#define MBAR_FAIL(id) do { if (atomicCAS(&g_dbg[0], 0u, (id)) == 0u) { g_dbg[1] = blockIdx.x; g_dbg[2] = threadIdx.x; } return; } while (0)
#define MBAR_SPINS (1u << 21)
#else
#define MBAR_FAIL(id) __trap()
#define MBAR_SPINS (1u << 28)
#endif
// -DWMD_TC_TRACE (scripts/tc_trace.py builds a separate library): CTA 0 records clock64() at fixed points of the
// first kTraceChunks chunks for producer warp 0, producer warp 11 and issuer 0.
#ifdef WMD_TC_TRACE
constexpr int kTraceChunks = 256, kTraceSlots = 8;
__device__ long long g_trace[3 * kTraceSlots * kTraceChunks];
#define TC_TRACE(role, slot, c) do { if (blockIdx.x == 0 && lane == 0 && (c) < kTraceChunks) \
    g_trace[((role) * kTraceSlots + (slot)) * kTraceChunks + (c)] = clock64(); } while (0)
#else
#define TC_TRACE(role, slot, c) do {} while (0)
#endif
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t id = 0) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < MBAR_SPINS; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  MBAR_FAIL(id);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void producer_barrier() { asm volatile("bar.sync 1, %0;\n" ::"n"(TC_PROD_THREADS) : "memory"); }
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

// D[tmem] (+)= A[tmem] * B[smem]^T : A is 128 lanes x 8 tf32 columns in tensor memory, B a K-major smem tile
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// this thread's TMEM lane (warp quarter base + lane), 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
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

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_rows_tc_kernel(const wmd_conv_desc d, const float* __restrict__ wtc,
                                                                     const int splits, float* __restrict__ partial) {
  using Cfg = TcCfg<BN>;
  constexpr int TC_B_TILE = Cfg::B_TILE;
  constexpr int TC_ISSUERS = Cfg::ISSUERS;
  constexpr int COPIES = Cfg::COPIES;
  constexpr int ACC = Cfg::ACC;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bar_asplit[2];          // split A of the TMEM stage is stored (12 producer warps)
  __shared__ __align__(8) uint64_t bar_mma[2];             // chunk's MMAs done (4 issuers): TMEM A stage + B stage reusable
  __shared__ __align__(8) uint64_t bar_b[TC_B_STAGES];     // weight image of the stage has landed (bulk copy)
  __shared__ __align__(8) uint64_t bar_epoch;              // accumulation epoch complete (4 issuers)
  __shared__ uint32_t tmem_base_slot;

  // warp index through a shuffle: provably warp-uniform, so the role branches and everything the issuer warps
  // compute from it stay on the uniform datapath (see the issuer section)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const bool is_producer = warp < TC_PROD_WARPS;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sA_base = base;
  unsigned char* sB_base = base + TC_A_STAGES * TC_A_TILE;
  uint32_t* tab0 = reinterpret_cast<uint32_t*>(sB_base + TC_B_STAGES * 2 * TC_B_TILE);   // 16-byte-unit offsets into x0
  uint32_t* tab1 = tab0 + 9 * TC_BM;                                                      // ... into x1

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_asplit[s]), TC_PROD_WARPS);
      mbar_init(smem_u32(&bar_mma[s]), TC_ISSUERS);
    }
    for (int s = 0; s < TC_B_STAGES; ++s) mbar_init(smem_u32(&bar_b[s]), 1);
    mbar_init(smem_u32(&bar_epoch), TC_ISSUERS);
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
  const int n_tiles = (d.cout + BN - 1) / BN;
  const long long tiles = static_cast<long long>((rows + TC_BM - 1) / TC_BM) * n_tiles;
  const int Hs = d.H >> d.shift0, Ws = d.W >> d.shift0;
  const bool aligned_rows = (d.taps == 1 && d.map0 == nullptr);
  const uint32_t ld0q = static_cast<uint32_t>(d.ld0 >> 2), ld1q = static_cast<uint32_t>(d.ld1 >> 2);
  // instruction descriptor: D=f32, A=B=tf32, both K-major, N = BN, M = 128
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                              (static_cast<uint32_t>(128 >> 4) << 24);

  // drain / store ownership (all 16 warps): TMEM lane quarter, M half, column half (ACC = BN/2 columns of every copy)
  const int my_q = warp & 3, my_half = (warp >> 2) & 1, my_ch = warp >> 3;
  const int my_row = my_half * 128 + my_q * 32 + lane;          // row within the CTA tile
  const uint32_t lane_field = static_cast<uint32_t>(my_q * 32) << 16;
  const uint32_t my_acc_addr = tmem_acc + lane_field + static_cast<uint32_t>(my_half * COPIES * BN + my_ch * ACC);
  uint32_t mma_rounds = 0;                                       // chunks issued so far by this CTA (all tiles)
  uint32_t epochs = 0;                                           // epoch commits so far

  // accumulators start (and are left by every drain) at zero: every MMA accumulates
#pragma unroll
  for (int j = 0; j < COPIES; ++j)
#pragma unroll
    for (int cc = 0; cc < ACC; cc += 16) tmem_zero16(my_acc_addr + static_cast<uint32_t>(j * BN + cc));
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // Work decomposition.  A "unit" is one 32-channel chunk of one output tile.
  //   splits >= 1 : every tile's reduction is cut into `splits` equal ranges (split-K); splits == 1 = whole tiles.
  //   splits == 0 : BALANCED - the tiles x nchunks units are dealt out to the CTAs in equal contiguous ranges of U
  //                 units (stream-K style), so the SMs finish together however many tiles the (device-side) row
  //                 count yields.  A tile cut by a range boundary has <= 4 segments (U >= nchunks/3).
  // Segments that do not cover a whole tile write raw partial sums to workspace slab `slab`; tc_reduce_kernel sums
  // the slabs in a fixed order and applies bias + activation, so results stay deterministic.
  const bool balanced = (splits == 0);
  const long long total_units = tiles * nchunks;
  const long long U = balanced ? max((total_units + gridDim.x - 1) / gridDim.x, static_cast<long long>((nchunks + 2) / 3)) : 0;
  long long u = balanced ? static_cast<long long>(blockIdx.x) * U : 0;
  const long long u_end = balanced ? min(total_units, u + U) : 0;
  long long item = blockIdx.x;
  const long long items = balanced ? 0 : tiles * splits;
  while (true) {
    long long tile;
    int cb, ce, slab;
    bool whole;
    if (balanced) {
      if (u >= u_end) break;
      tile = u / nchunks;
      cb = static_cast<int>(u - tile * nchunks);
      ce = static_cast<int>(min(static_cast<long long>(nchunks), cb + (u_end - u)));
      slab = static_cast<int>(blockIdx.x - (tile * nchunks) / U);
      whole = (cb == 0 && ce == nchunks);
      u += ce - cb;
    } else {
      if (item >= items) break;
      tile = item / splits;
      slab = static_cast<int>(item - tile * splits);
      cb = static_cast<int>(static_cast<long long>(slab) * nchunks / splits);
      ce = static_cast<int>(static_cast<long long>(slab + 1) * nchunks / splits);
      whole = (splits == 1);
      item += gridDim.x;
    }
    const int len = ce - cb;
    const int m0 = static_cast<int>(tile / n_tiles) * TC_BM;
    const int nt = static_cast<int>(tile % n_tiles);
    const int n0 = nt * BN;

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
    const uint32_t round0 = mma_rounds;

    float acc[ACC];
#pragma unroll
    for (int j = 0; j < ACC; ++j) acc[j] = 0.f;

    // drains this thread's slice (its row, ACC columns of every issuer's copy) with round-to-nearest adds, re-zeroes it
    auto drain = [&]() {
#pragma unroll
      for (int cp = 0; cp < COPIES; ++cp) {
#pragma unroll
        for (int cc = 0; cc < ACC; cc += 16) {
          uint32_t v[16];
          tmem_ld16(my_acc_addr + static_cast<uint32_t>(cp * BN + cc), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[cc + j] += __uint_as_float(v[j]);
          tmem_zero16(my_acc_addr + static_cast<uint32_t>(cp * BN + cc));
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    };

    if (is_producer) {
      // =================================================================================== producers
      // decode chunk c -> gather source (as float4 pointer), channels left, offset table
      auto chunk_src = [&](int c, const float4*& xq, int& cleft) -> const uint32_t* {
        const int tap = c / per_tap;
        const int rr = c - tap * per_tap;
        const bool src1 = rr >= nch0;
        const int ci0 = (src1 ? rr - nch0 : rr) * TC_BK;
        cleft = (src1 ? d.c1 : d.c0) - ci0;
        xq = reinterpret_cast<const float4*>(src1 ? d.x1 : d.x0) + (ci0 >> 2);
        return (src1 ? tab1 : tab0) + tap * TC_BM;
      };
      auto prefetch_chunk = [&](int c) {
        const float4* xq; int cleft;
        const uint32_t* tab = chunk_src(c, xq, cleft);
        if (tid < TC_BM) {                          // one 128-byte line per row
          const uint32_t off = tab[tid];
          if (off != kNoRow) prefetch_l2(xq + off);
        }
      };
      // raw fp32 A tile of chunk c -> shared A stage c&1 (128B-swizzled rows); 2048 pieces over 384 threads
      auto load_a = [&](int c, int st) {
        const float4* xq; int cleft;
        const uint32_t* tab = chunk_src(c, xq, cleft);
        unsigned char* sA = sA_base + st * TC_A_TILE;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int p = tid + i * TC_PROD_THREADS;
          if (p < TC_BM * 8) {
            const int r = p >> 3, j = p & 7;
            const int a_bytes = max(0, min(16, (cleft - j * 4) * 4));
            const uint32_t off = tab[r];
            const bool live = off != kNoRow && a_bytes > 0;
            const float4* src = live ? xq + off + j : xq;
            cp_async16(sA + r * 128 + ((j ^ (r & 7)) << 4), src, live ? a_bytes : 0);
          }
        }
      };

      for (int c = 1; c <= kPrefetchAhead && c < len; ++c) prefetch_chunk(cb + c);
      load_a(cb, 0);
      cp_async_commit();

      const int wt = warp >> 2;                      // 0..2: which of the quarter's three producer warps
      for (int c = 0; c < len; ++c) {
        const uint32_t round = round0 + c;
        const uint32_t tstage = round & 1;
        const int trole = warp == 0 ? 0 : (warp == TC_PROD_WARPS - 1 ? 1 : 3);
        if (trole < 3) TC_TRACE(trole, 0, c);
        cp_async_wait<0>();
        if (trole < 3) TC_TRACE(trole, 1, c);
        producer_barrier();                          // raw A tile of chunk c complete; split reads of chunk c-1 done
        if (trole < 3) TC_TRACE(trole, 2, c);
        if (c + 1 < len) {
          load_a(cb + c + 1, (c + 1) & 1);
          if (c + 1 + kPrefetchAhead < len) prefetch_chunk(cb + c + 1 + kPrefetchAhead);
        }
        cp_async_commit();
        if (trole < 3) TC_TRACE(trole, 3, c);
        if ((c % kFlushChunks) == 0 && c > 0) {      // epoch boundary: everybody drains before the next epoch starts
          mbar_wait(smem_u32(&bar_epoch), (epochs - 1) & 1, 0x10000u + round);
          tc_fence_after();
          drain();
          tc_fence_before();
          __syncthreads();
          tc_fence_after();
        }
        // TMEM A stage free?  It was read by the MMAs of round-2.
        if (round >= 2) mbar_wait(smem_u32(&bar_mma[tstage]), ((round - 2) >> 1) & 1, 0x20000u + round);
        tc_fence_after();
        if (trole < 3) TC_TRACE(trole, 4, c);
        // split my (row, 8 channels) slots: u = half*4 + channel-quarter, u = wt, wt+3, wt+6
        {
          const unsigned char* tile_a = sA_base + (c & 1) * TC_A_TILE;
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            const int slot = wt + 3 * u;
            if (slot < 8) {
              const int h = slot >> 2, kq = slot & 3;
              const int r = h * 128 + my_q * 32 + lane;
              const unsigned char* rowp = tile_a + r * 128;
              const uint4 v0 = *reinterpret_cast<const uint4*>(rowp + (((2 * kq) ^ (r & 7)) << 4));
              const uint4 v1 = *reinterpret_cast<const uint4*>(rowp + (((2 * kq + 1) ^ (r & 7)) << 4));
              const uint32_t raw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
              uint32_t hi[8], lo[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                hi[e] = raw[e] & 0xFFFFE000u;                                       // tf32 by truncation
                lo[e] = __float_as_uint(__uint_as_float(raw[e]) - __uint_as_float(hi[e]));   // exact remainder
              }
              const uint32_t ta = tmem_acc + lane_field + 256u + tstage * 128u + static_cast<uint32_t>(h * 64 + kq * 8);
              tmem_st8(ta, hi);
              tmem_st8(ta + 32u, lo);
            }
          }
          if (trole < 3) TC_TRACE(trole, 5, c);
          asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_asplit[tstage]));
        if (trole < 3) TC_TRACE(trole, 6, c);
        if (((c + 1) % kFlushChunks == 0) || (c == len - 1)) epochs += 1;
      }
    } else {
      // =================================================================================== issuers
      // The whole warp runs the loop and one elected lane issues.  Written this way (warp-uniform control flow,
      // operands derived from warp-uniform values) ptxas keeps the MMA operands in uniform registers and emits the
      // UTCHMMAs back to back: ~78 clk per instruction.  A `lane == 0` branch instead makes it wrap every MMA in an
      // ELECT / R2UR.BROADCAST / BRA.U.ANY loop: ~210 clk (scripts/bench_cu/mma_rate*.cu).
      const int kstep = warp - TC_PROD_WARPS;        // this issuer's index
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_acc, 0);
      const uint32_t sB_u = __shfl_sync(0xffffffffu, smem_u32(sB_base), 0);
      if (kstep == 0 && elect_one()) {
        const unsigned char* w0 = wtile + static_cast<long long>(cb) * (2 * TC_B_TILE);
        bulk_g2s(sB_u + (round0 % TC_B_STAGES) * 2 * TC_B_TILE, w0, 2 * TC_B_TILE, smem_u32(&bar_b[round0 % TC_B_STAGES]));
        if (len > 1)
          bulk_g2s(sB_u + ((round0 + 1) % TC_B_STAGES) * 2 * TC_B_TILE, w0 + 2 * TC_B_TILE, 2 * TC_B_TILE,
                   smem_u32(&bar_b[(round0 + 1) % TC_B_STAGES]));
      }
      __syncwarp();
      // issuer -> (M half, accumulator copy); it alone writes that accumulator
      const int ih = kstep / COPIES, icp = kstep % COPIES;
      const uint32_t dh = tmem_u + static_cast<uint32_t>((ih * COPIES + icp) * BN);
      for (int c = 0; c < len; ++c) {
        const uint32_t round = round0 + c;
        const uint32_t tstage = round & 1;
        const uint32_t bs = round % TC_B_STAGES;
        if ((c % kFlushChunks) == 0 && c > 0) {      // epoch boundary (all lanes: drain is warp-collective)
          mbar_wait(smem_u32(&bar_epoch), (epochs - 1) & 1);
          tc_fence_after();
          drain();
          tc_fence_before();
          __syncthreads();
          tc_fence_after();
        }
        if (kstep < TC_ISSUERS) {
          if (kstep == 0) TC_TRACE(2, 0, c);
          mbar_wait(smem_u32(&bar_asplit[tstage]), (round >> 1) & 1, 0x30000u + round);           // split A of this chunk is in TMEM
          if (kstep == 0) TC_TRACE(2, 1, c);
          mbar_wait(smem_u32(&bar_b[bs]), (round / TC_B_STAGES) & 1, 0x40000u + round);           // weight image has landed
          if (kstep == 0) TC_TRACE(2, 2, c);
          tc_fence_after();
          const uint64_t b0 = umma_desc_sw128(sB_u + bs * 2 * TC_B_TILE);
          const uint32_t ah = tmem_u + 256u + tstage * 128u + static_cast<uint32_t>(ih * 64);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks) {
              if ((ks % COPIES) == icp) {
                const uint64_t bh = b0 + static_cast<uint64_t>(2 * ks);
                const uint64_t bl = bh + static_cast<uint64_t>(TC_B_TILE >> 4);
                const uint32_t a = ah + static_cast<uint32_t>(8 * ks);
                umma_tf32_ts(dh, a + 32u, bh, kIdesc, 1u);     // lo*hi
                umma_tf32_ts(dh, a, bl, kIdesc, 1u);           // hi*lo
                umma_tf32_ts(dh, a, bh, kIdesc, 1u);           // hi*hi
              }
            }
            umma_commit(smem_u32(&bar_mma[tstage]));
            if (((c + 1) % kFlushChunks == 0) || (c == len - 1)) umma_commit(smem_u32(&bar_epoch));
            // weights two chunks ahead: that B stage was last read by round-2, and all of round-2's MMAs are known to
            // be complete - the producers only stored this chunk's A (bar_asplit, awaited above) after bar_mma(round-2).
            // (Waiting on bar_mma here would alias: this round's own commits may already have flipped its phase.)
            if (kstep == 0 && c + 2 < len) {
              const uint32_t ns = (round + 2) % TC_B_STAGES;
              bulk_g2s(sB_u + ns * 2 * TC_B_TILE, wtile + static_cast<long long>(cb + c + 2) * (2 * TC_B_TILE),
                       2 * TC_B_TILE, smem_u32(&bar_b[ns]));
            }
          }
        }
        __syncwarp();
        if (kstep == 0) TC_TRACE(2, 3, c);
        if (((c + 1) % kFlushChunks == 0) || (c == len - 1)) epochs += 1;
      }
    }
    mma_rounds = round0 + static_cast<uint32_t>(len);

    // ---- last epoch + epilogue (all 16 warps): bias, activation, one contiguous 256-byte store per thread
    {
      mbar_wait(smem_u32(&bar_epoch), (epochs - 1) & 1, 0x60000u + mma_rounds);
      tc_fence_after();
      drain();
      const int m = m0 + my_row;
      if (m < rows && !whole) {
        // raw partial sums of this segment (bias / activation are applied by the reduce pass)
        float* pr = partial + (static_cast<long long>(slab) * d.max_rows + m) * d.ldy;
#pragma unroll
        for (int j = 0; j < ACC; j += 4) {
          const int co = n0 + my_ch * ACC + j;       // ldy % 4 == 0: a quad that starts below cout stays inside the row
          if (co < d.cout) *reinterpret_cast<float4*>(pr + co) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
      } else if (m < rows) {
        float* yr = d.y + static_cast<long long>(m) * d.ldy;
        const bool vec_ok = (d.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15) == 0);
#pragma unroll
        for (int j = 0; j < ACC; j += 4) {
          const int co = n0 + my_ch * ACC + j;
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
    __syncthreads();   // accumulators drained + re-zeroed by every warp, tap tables free
    tc_fence_after();
  }

  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_acc),
                 "r"(static_cast<uint32_t>(TC_TMEM_COLS))
                 : "memory");
  }
}

// w (Cout, Cin, taps) fp32 -> per (n-tile, chunk) smem image [tf32 hi: BN x 32 | tf32 lo: BN x 32], K-major,
// 128B-swizzled.  Chunk order = the kernel's: tap-major, then source-0 channel chunks, then source-1 chunks.
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int c0, int c1,
                                      int taps, int BN, long long total) {
  const int nch0 = (c0 + TC_BK - 1) / TC_BK, nch1 = (c1 + TC_BK - 1) / TC_BK;
  const int per_tap = nch0 + nch1;
  const int nchunks = taps * per_tap;
  const int Cin = c0 + c1;
  const long long tile_floats = static_cast<long long>(BN) * TC_BK;
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += step) {
    // i indexes LOGICAL (nt, chunk, hilo, n, kk); the store address applies the swizzle
    long long t = i;
    const int kk = static_cast<int>(t % TC_BK); t /= TC_BK;
    const int n = static_cast<int>(t % BN); t /= BN;
    const int hilo = static_cast<int>(t % 2); t /= 2;
    const int c = static_cast<int>(t % nchunks);
    const int nt = static_cast<int>(t / nchunks);
    const int tap = c / per_tap;
    const int rr = c - tap * per_tap;
    const bool src1 = rr >= nch0;
    const int ci_local = (src1 ? rr - nch0 : rr) * TC_BK + kk;
    const int csrc = src1 ? c1 : c0;
    const int co = nt * BN + n;
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

// y[m, co] = act(bias[co] + sum_s partial[s][m][co]) in a fixed order (deterministic).  splits >= 2: every tile has
// `splits` slabs.  splits == 0 (balanced): the number of slabs of a tile follows from the same unit arithmetic the
// conv kernel used (grid = its CTA count); tiles that one CTA covered entirely were already finished there.
__global__ void tc_reduce_kernel(const float* __restrict__ partial, int splits, int grid, int BN, int nchunks,
                                 const float* __restrict__ bias, float* __restrict__ y, int ldy, int cout,
                                 const int32_t* __restrict__ count, int max_rows, int act, float act_param) {
  const int rows = count ? min(*count, max_rows) : max_rows;
  const long long slab_sz = static_cast<long long>(max_rows) * ldy;
  const int n_tiles = (cout + BN - 1) / BN;
  const long long tiles = static_cast<long long>((rows + TC_BM - 1) / TC_BM) * n_tiles;
  const long long total_units = tiles * nchunks;
  const long long U = splits == 0 ? max((total_units + grid - 1) / grid, static_cast<long long>((nchunks + 2) / 3)) : 0;
  const int quads = BN >> 2;                         // float4 columns of a tile (ldy is a multiple of 4)
  // one CTA per tile per round: the slab arithmetic is per tile, the element loop has no divisions by run-time values
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    int nslabs = splits;
    if (splits == 0) {
      const long long first = (tile * nchunks) / U, last = ((tile + 1) * nchunks - 1) / U;
      if (first == last) continue;                   // whole tile: already written with bias + activation
      nslabs = static_cast<int>(last - first + 1);
    }
    const int m0 = static_cast<int>(tile / n_tiles) * TC_BM;
    const int co0 = static_cast<int>(tile % n_tiles) * BN;
    const int mrows = min(TC_BM, rows - m0);
    for (int e = threadIdx.x; e < mrows * quads; e += blockDim.x) {
      const int r = e / quads, q = e - r * quads;    // quads is a power of two times {8,16,32}: cheap 32-bit division
      const int co = co0 + (q << 2);
      if (co >= cout) continue;
      const long long o = static_cast<long long>(m0 + r) * ldy + co;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) {
        v.x = __ldg(bias + co);
        if (co + 1 < cout) v.y = __ldg(bias + co + 1);
        if (co + 2 < cout) v.z = __ldg(bias + co + 2);
        if (co + 3 < cout) v.w = __ldg(bias + co + 3);
      }
      for (int sidx = 0; sidx < nslabs; ++sidx) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(partial + sidx * slab_sz + o));
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      }
      v.x = activate(v.x, act, act_param); v.y = activate(v.y, act, act_param);
      v.z = activate(v.z, act, act_param); v.w = activate(v.w, act, act_param);
      if (co + 3 < cout) {
        *reinterpret_cast<float4*>(y + o) = v;
      } else {
        y[o] = v.x;
        if (co + 1 < cout) y[o + 1] = v.y;
        if (co + 2 < cout) y[o + 2] = v.z;
      }
    }
  }
}

static int tc_tile_n(int cout) { return cout >= 96 ? 128 : (cout >= 48 ? 64 : 32); }

template <int BN>
static int launch_tc(const wmd_conv_desc& d, int splits, float* partial, cudaStream_t stream) {
  using Cfg = TcCfg<BN>;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    int rc = record(cudaFuncSetAttribute(conv_rows_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(Cfg::SMEM)));
    if (rc != WMD_OK) return rc;
    attr_done[dev] = true;
  }
  const long long tiles = static_cast<long long>(ceil_div(d.max_rows, TC_BM)) * ceil_div(d.cout, BN) * (splits > 0 ? splits : 1);
  const long long cap = sm_count();
  const int grid = splits == 0 ? static_cast<int>(cap) : static_cast<int>(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
  conv_rows_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM, stream>>>(d, d.w, splits, partial);
  int rc = launched();
  if (rc != WMD_OK || splits == 1) return rc;
  const int nchunks = d.taps * ((d.c0 + TC_BK - 1) / TC_BK + (d.c1 + TC_BK - 1) / TC_BK);
  const long long all_tiles = static_cast<long long>(ceil_div(d.max_rows, TC_BM)) * ceil_div(d.cout, BN);
  tc_reduce_kernel<<<static_cast<int>(all_tiles < 8 * cap ? all_tiles : 8 * cap), 256, 0, stream>>>(partial, splits, grid, BN, nchunks, d.bias, d.y, d.ldy, d.cout,
                                                              d.count, d.max_rows, d.act, d.act_param);
  return launched();
}

}  // namespace wmd

#ifdef WMD_TC_TRACE
extern "C" int wmd_debug_tc_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, wmd::g_trace, sizeof(wmd::g_trace)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int wmd_conv_tc_tile_n(int cout) { return wmd::tc_tile_n(cout); }

extern "C" size_t wmd_conv_tc_weight_floats(int cout, int c0, int c1, int taps) {
  using namespace wmd;
  const int bn = tc_tile_n(cout);
  const int nchunks = taps * ((c0 + TC_BK - 1) / TC_BK + (c1 + TC_BK - 1) / TC_BK);
  return static_cast<size_t>(ceil_div(cout, bn)) * nchunks * 2 * bn * TC_BK;
}

extern "C" int wmd_pack_conv_weight_tc_f32(const float* w, float* packed, int Cout, int c0, int c1, int taps,
                                           wmd_stream_t stream) {
  using namespace wmd;
  WMD_REQUIRE(w && packed, WMD_ERR_ARG);
  WMD_REQUIRE(Cout > 0 && c0 > 0 && c1 >= 0 && (taps == 1 || taps == 9), WMD_ERR_SHAPE);
  const long long total = static_cast<long long>(wmd_conv_tc_weight_floats(Cout, c0, c1, taps));
  pack_weight_tc_kernel<<<stride_grid(total, 256), 256, 0, as_stream(stream)>>>(w, packed, Cout, c0, c1, taps,
                                                                               tc_tile_n(Cout), total);
  return launched();
}

extern "C" size_t wmd_conv_tc_splitk_ws_bytes(int max_rows, int ldy, int splits) {
  if (splits == 1) return 0;
  const size_t slabs = splits == 0 ? 4 : static_cast<size_t>(splits);      // balanced mode: <= 4 segments per tile
  return slabs * static_cast<size_t>(max_rows) * static_cast<size_t>(ldy) * sizeof(float);
}

extern "C" int wmd_conv_rows_tc_f32(const wmd_conv_desc* dp, wmd_stream_t stream) {
  return wmd_conv_rows_tc_splitk_f32(dp, 1, nullptr, 0, stream);
}

extern "C" int wmd_conv_rows_tc_splitk_f32(const wmd_conv_desc* dp, int splits, void* ws, size_t ws_bytes,
                                           wmd_stream_t stream) {
  using namespace wmd;
  WMD_REQUIRE(dp, WMD_ERR_ARG);
  WMD_REQUIRE(splits >= 0 && splits <= 16, WMD_ERR_ARG);
  WMD_REQUIRE(splits == 1 || (ws != nullptr && ws_bytes >= wmd_conv_tc_splitk_ws_bytes(dp->max_rows, dp->ldy, splits)),
              WMD_ERR_WORKSPACE);
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
  // the partial-sum passes move float4s
  WMD_REQUIRE(splits == 1 || (d.ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(d.y) & 15) == 0 &&
                              (reinterpret_cast<uintptr_t>(ws) & 15) == 0), WMD_ERR_SHAPE);
  if (d.shift0 == 1) WMD_REQUIRE(d.H % 2 == 0 && d.W % 2 == 0, WMD_ERR_SHAPE);
  if (d.pad_mode == WMD_PAD_REFLECT && d.taps == 9) WMD_REQUIRE(d.H >= 2 && d.W >= 2, WMD_ERR_SHAPE);
  // tap tables hold 32-bit offsets in 16-byte units
  WMD_REQUIRE(static_cast<long long>(d.N) * d.H * d.W * (d.ld0 / 4) < (1ll << 32) &&
                  static_cast<long long>(d.N) * d.H * d.W * (d.ld1 / 4) < (1ll << 32),
              WMD_ERR_UNSUPPORTED);
  if (d.max_rows == 0) return WMD_OK;
  {
    const int nchunks = d.taps * ((d.c0 + TC_BK - 1) / TC_BK + (d.c1 + TC_BK - 1) / TC_BK);
    if (splits > nchunks) splits = nchunks;      // every split needs at least one chunk
    if (splits == 0 && nchunks < 2) splits = 1;   // nothing to balance inside a one-chunk reduction
  }
  float* partial = static_cast<float*>(ws);
  switch (tc_tile_n(d.cout)) {
    case 128: return launch_tc<128>(d, splits, partial, as_stream(stream));
    case 64: return launch_tc<64>(d, splits, partial, as_stream(stream));
    default: return launch_tc<32>(d, splits, partial, as_stream(stream));
  }
}
