// knn.cu -- brute-force Hamming kNN (k = 2) + Lowe ratio test.
//
// Replaces matcher_.knnMatch(d1, d2, knn, 2) + the ratio loop at
// /root/reference/src/feature_processor.cpp:25-40.  Contract: top-2 by
// (distance, trainIdx) ascending, accepted iff two neighbours exist and d0 < ratio * d1 in double.
//
// knn_top2_kernel: a warp owns Q=4 query descriptors in registers; the train set streams through a
// swizzled shared-memory tile (conflict-free 128-bit reads, one train row per lane); distances are
// xor + __popc over the two uint4 halves, with two carry-save adders in front of the population counts (six
// POPC per pair instead of eight: POPC issues at a quarter of the ALU rate); every lane keeps a private top-2 of packed keys
// (dist << 22 | trainIdx), merged at the end by a warp-shuffle top-2 reduction.  The train range
// can be split over blockIdx.y to fill the 148 SMs for a single stream.
// knn_finish_kernel: merges the splits, applies the ratio test and compacts in query order.
#include "context.cuh"
#include <algorithm>

namespace mvo {

#ifndef MVO_KNN_CSA
#define MVO_KNN_CSA 2   // carry-save adders per pair: 2 -> six POPC (217 us per 32 x 2000^2 pairs), 3 -> five POPC (229 us, ALU-bound), 0 -> eight (253 us)
#endif
constexpr int kKnnQ = 4;             // queries per warp
constexpr int kKnnWarps = 8;
constexpr int kKnnThreads = kKnnWarps * 32;
constexpr int kKnnTile = 256;        // train rows per smem tile (8 KB)
constexpr int kKnnMaxSplit = 8;
constexpr uint32_t kInvalidKey = 0xFFFFFFFFu;

__device__ __forceinline__ void top2_insert(uint32_t& k0, uint32_t& k1, uint32_t key) {
  k1 = min(k1, max(k0, key));
  k0 = min(k0, key);
}

__global__ void __launch_bounds__(kKnnThreads)
knn_top2_kernel(const uint8_t* __restrict__ q, const int32_t* __restrict__ nq_dev, int q_stride_rows,
                const uint8_t* __restrict__ t, const int32_t* __restrict__ nt_dev, int t_stride_rows,
                uint2* __restrict__ partial, int max_nq, int nsplit) {
  __shared__ __align__(16) uint4 tile[kKnnTile * 2];
  const int b = blockIdx.z;
  const int nq = min(nq_dev[b], max_nq), nt = nt_dev[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = (blockIdx.x * kKnnWarps + warp) * kKnnQ;
  if (blockIdx.x * kKnnWarps * kKnnQ >= nq) return;

  // this block's slice of the train set, in whole tiles
  const int tiles_total = (nt + kKnnTile - 1) / kKnnTile;
  const int tiles_per = (tiles_total + nsplit - 1) / nsplit;
  const int t_begin = blockIdx.y * tiles_per * kKnnTile;
  const int t_end = min(nt, t_begin + tiles_per * kKnnTile);

  uint32_t qr[kKnnQ][8];
  const uint4* qv = reinterpret_cast<const uint4*>(q + (long long)b * q_stride_rows * 32);
#pragma unroll
  for (int i = 0; i < kKnnQ; ++i) {
    const int qi = min(q0 + i, nq - 1);
    const uint4 a = __ldg(qv + qi * 2), c = __ldg(qv + qi * 2 + 1);
    qr[i][0] = a.x; qr[i][1] = a.y; qr[i][2] = a.z; qr[i][3] = a.w;
    qr[i][4] = c.x; qr[i][5] = c.y; qr[i][6] = c.z; qr[i][7] = c.w;
  }
  uint32_t k0[kKnnQ], k1[kKnnQ];
#pragma unroll
  for (int i = 0; i < kKnnQ; ++i) k0[i] = k1[i] = kInvalidKey;

  const uint4* tv = reinterpret_cast<const uint4*>(t + (long long)b * t_stride_rows * 32);
  for (int base = t_begin; base < t_end; base += kKnnTile) {
    __syncthreads();
    {
      // 256 threads x 2 chunks: row r = threadIdx.x, chunk c stored at slot 2r + (c ^ ((r >> 2) & 1))
      const int r = threadIdx.x;
      const int gr = base + r;
      uint4 a = make_uint4(0, 0, 0, 0), c = a;
      if (gr < t_end) {
        a = __ldg(tv + (long long)gr * 2);
        c = __ldg(tv + (long long)gr * 2 + 1);
      }
      const int sw = (r >> 2) & 1;
      tile[2 * r + sw] = a;
      tile[2 * r + (sw ^ 1)] = c;
    }
    __syncthreads();
    const int rows = min(kKnnTile, t_end - base);
#pragma unroll 2
    for (int r = lane; r < rows; r += 32) {
      const int sw = (r >> 2) & 1;
      const uint4 a = tile[2 * r + sw], c = tile[2 * r + (sw ^ 1)];
      const uint32_t idx = (uint32_t)(base + r);
#pragma unroll
      for (int i = 0; i < kKnnQ; ++i) {
        // Hamming distance of 256 bits with six population counts instead of eight: two carry-save adders (two LOP3
        // each, on the ALU pipe) compress six of the eight xor words into two "ones" words and two "twos" words.
        // POPC issues at a quarter of the ALU rate, so this moves work from the saturated pipe to a less busy one; a
        // third adder (five POPC) makes the ALU pipe the bottleneck.
        const uint32_t x0 = qr[i][0] ^ a.x, x1 = qr[i][1] ^ a.y, x2 = qr[i][2] ^ a.z, x3 = qr[i][3] ^ a.w;
        const uint32_t x4 = qr[i][4] ^ c.x, x5 = qr[i][5] ^ c.y, x6 = qr[i][6] ^ c.z, x7 = qr[i][7] ^ c.w;
        const uint32_t s0 = x0 ^ x1 ^ x2, c0 = (x0 & x1) | (x2 & (x0 | x1));
        const uint32_t s1 = x3 ^ x4 ^ x5, c1 = (x3 & x4) | (x5 & (x3 | x4));
#if MVO_KNN_CSA == 2
        const int d = __popc(s0) + __popc(s1) + __popc(x6) + __popc(x7) + 2 * (__popc(c0) + __popc(c1));
#else
        const uint32_t s2 = s0 ^ s1 ^ x6, c2 = (s0 & s1) | (x6 & (s0 | s1));
        const int d = __popc(s2) + __popc(x7) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
#endif
        top2_insert(k0[i], k1[i], ((uint32_t)d << 22) | idx);
      }
    }
  }
  // warp-shuffle top-2 reduction: merge two sorted pairs per step
#pragma unroll
  for (int i = 0; i < kKnnQ; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t o0 = __shfl_xor_sync(0xffffffffu, k0[i], o);
      const uint32_t o1 = __shfl_xor_sync(0xffffffffu, k1[i], o);
      const uint32_t n1 = min(max(k0[i], o0), min(k1[i], o1));
      k0[i] = min(k0[i], o0);
      k1[i] = n1;
    }
    if (lane == 0 && q0 + i < nq)
      partial[((long long)b * nsplit + blockIdx.y) * max_nq + q0 + i] = make_uint2(k0[i], k1[i]);
  }
}

// ================================================================================================
// knn_mma_kernel: the same top-2 on the 5th-generation tensor cores.
//
// With descriptor bits mapped to +-1, q . t = 256 - 2 * Hamming(q, t): the all-pairs distance matrix is a dense
// {+-8}-valued int8 GEMM (tcgen05.mma kind::i8, accumulators in TMEM, exact: |64 s| <= 2^14).  A CTA owns 128 queries
// and streams its slice of the train set through 128-row tiles:
//   * the packed descriptors (32 B per row) are expanded to the canonical K-major no-swizzle core-matrix layout
//     (8 rows x 16 B atoms, LBO = 128 B between K-adjacent atoms, SBO = 2 KB between 8-row groups) straight into shared
//     memory by the CTA's threads (PRMT sign replication, no table): the expanded operands never exist in global memory
//     (8 x the traffic of the packed form);
//   * one elected thread issues the eight K = 32 MMAs of a tile into one of two 128-column TMEM accumulators and
//     commits to an mbarrier; the expansion of tile j + 1 and the selection epilogue of tile j overlap MMA j + 1;
//   * epilogue: eight warps (TMEM lane quarter x column half) read the 128 x 128 int32 block with tcgen05.ld and keep a
//     per-row top-2 on PACKED 16-bit keys, two columns per instruction: the accumulator 64 s is a multiple of 128, so
//     (64 s) ^ 0x8000 | (127 - column) is an order-preserving 16-bit key (larger = nearer, ties -> lower column).
//     PRMT + LOP + three VIMNMX.U16x2 per column pair.  A tile's winners are decoded into the usual 32-bit key
//     (dist << 22 | trainIdx) only when they beat the row's current second best.
// Results are bit-identical to knn_top2_kernel (tests/test_gpu_knn.py runs both).
constexpr int kMmaM = 128, kMmaN = 128, kMmaK = 256;          // CTA tile: queries x train rows x descriptor bits
constexpr int kMmaThreads = 256;
constexpr int kOperandBytes = kMmaM * kMmaK;                  // 32 KB per expanded tile
constexpr int kKnnMmaSmem = 3 * kOperandBytes + 2 * kMmaM * 8 + 64;

__device__ __forceinline__ uint32_t knn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major, no swizzle: element (row r, byte k) at (r / 8) * 2048 + (k / 16) * 128 + (r % 8) * 16 + k % 16
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(2048u >> 4) << 32) | (1ull << 46);
}
// kind::i8 instruction descriptor: D = s32, A = B = signed 8 bit, both K-major, N = 128, M = 128
constexpr uint32_t kUmmaIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kMmaN >> 3) << 17) | ((uint32_t)(kMmaM >> 4) << 24);

// One packed descriptor half-row (16 bytes = 128 bits) of row r -> 128 operand bytes (+8 for a set bit, -8 for a clear
// one) in eight 16-byte atoms.  No table: a byte whose selector nibble has bit 3 set makes PRMT replicate the sign bit of
// the selected source byte, so PRMT(word << (7 - i), 0xBA98) turns bit i of the four bytes of a word into four 0x00 / 0xFF
// mask bytes; one LOP3 maps them to 0xF8 / 0x08.  Element order inside a row: (word, bit, byte) -- any fixed order works
// as long as queries and train rows share it.
// the sign bit of each byte replicated over the byte (generic PRMT: selector nibbles 8..B; __byte_perm masks that bit off)
__device__ __forceinline__ uint32_t prmt_sign4(uint32_t x) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(d) : "r"(x));
  return d;
}
__device__ __forceinline__ void knn_expand_row(uint8_t* __restrict__ dst, const uint4 v, int r, int h) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint8_t* o = dst + (r >> 3) * 2048 + (8 * h) * 128 + (r & 7) * 16;
#pragma unroll
  for (int wi = 0; wi < 4; ++wi) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = (prmt_sign4(w[wi] << (7 - i)) & 0xF0F0F0F0u) ^ 0xF8F8F8F8u;
    *reinterpret_cast<uint4*>(o + (2 * wi) * 128) = make_uint4(e[0], e[1], e[2], e[3]);
    *reinterpret_cast<uint4*>(o + (2 * wi + 1) * 128) = make_uint4(e[4], e[5], e[6], e[7]);
  }
}
// the 16 bytes thread `tid` expands of tile rows [row0, row0 + 128): row tid / 2, half tid % 2 (zero bits past nrows)
__device__ __forceinline__ uint4 knn_load_row(const uint4* __restrict__ src, int row0, int nrows, int tid) {
  const int r = tid >> 1, h = tid & 1;
  return (row0 + r < nrows) ? __ldg(src + (long long)(row0 + r) * 2 + h) : make_uint4(0, 0, 0, 0);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
      "%25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kMmaThreads, 2)
knn_mma_kernel(const uint8_t* __restrict__ q, const int32_t* __restrict__ nq_dev, int q_stride_rows,
               const uint8_t* __restrict__ t, const int32_t* __restrict__ nt_dev, int t_stride_rows,
               uint2* __restrict__ partial, int max_nq, int nsplit) {
  extern __shared__ __align__(128) uint8_t knn_smem[];
  uint8_t* sA = knn_smem;
  uint8_t* sB = knn_smem + kOperandBytes;                              // two stages
  uint2* merge = reinterpret_cast<uint2*>(knn_smem + 3 * kOperandBytes);   // [2][128] (k0, k1) of the two column halves
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(merge + 2 * kMmaM);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int b = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nq = min(nq_dev[b], max_nq), nt = nt_dev[b];
  const int q0 = blockIdx.x * kMmaM;
  if (q0 >= nq) return;                                                // (uniform: before any barrier / allocation)
  const int tiles_total = (nt + kMmaN - 1) / kMmaN;
  const int tiles_per = (tiles_total + nsplit - 1) / nsplit;
  const int t_begin = blockIdx.y * tiles_per * kMmaN;
  const int t_end = min(nt, t_begin + tiles_per * kMmaN);
  const int ntiles = t_end > t_begin ? (t_end - t_begin + kMmaN - 1) / kMmaN : 0;

  // ---- setup: barriers, TMEM, the query operand ----
  if (tid == 0) {
    for (int s = 0; s < 2; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(knn_smem_u32(&bars[s])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(knn_smem_u32(tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const uint4* qv = reinterpret_cast<const uint4*>(q + (long long)b * q_stride_rows * 32);
  const uint4* tv = reinterpret_cast<const uint4*>(t + (long long)b * t_stride_rows * 32);
  knn_expand_row(sA, knn_load_row(qv, q0, nq, tid), tid >> 1, tid & 1);
  if (ntiles > 0) knn_expand_row(sB, knn_load_row(tv, t_begin, t_end, tid), tid >> 1, tid & 1);
  // the packed rows of the next tile are fetched one iteration ahead of their expansion
  uint4 vnext = ntiles > 1 ? knn_load_row(tv, t_begin + kMmaN, t_end, tid) : make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  auto issue_tile = [&](int stage) {                                   // one thread: eight K = 32 steps, then commit
    const uint32_t a0 = knn_smem_u32(sA), b0 = knn_smem_u32(sB + stage * kOperandBytes);
    const uint32_t d = tmem_base + (uint32_t)(stage * kMmaN);
#pragma unroll
    for (int k = 0; k < kMmaK / 32; ++k) {
      const uint64_t da = umma_desc(a0 + k * 256), db = umma_desc(b0 + k * 256);
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
          "}" ::"r"(d),
          "l"(da), "l"(db), "r"(kUmmaIdesc), "r"(k > 0 ? 1 : 0)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     knn_smem_u32(&bars[stage]))
                 : "memory");
  };
  if (ntiles > 0 && tid == 0) issue_tile(0);

  const int lq = warp & 3, ch = warp >> 2;                             // TMEM lane quarter, column half
  uint32_t k0 = kInvalidKey, k1 = kInvalidKey;                         // this row's top-2 over this warp's columns
  for (int j = 0; j < ntiles; ++j) {
    const int stage = j & 1;
    if (j + 1 < ntiles) {
      knn_expand_row(sB + (stage ^ 1) * kOperandBytes, vnext, tid >> 1, tid & 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (j + 2 < ntiles) vnext = knn_load_row(tv, t_begin + (j + 2) * kMmaN, t_end, tid);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                   // next operand complete, other accumulator drained
    if (j + 1 < ntiles && tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_tile(stage ^ 1);
    }
    // ---- wait for MMA j ----
    {
      const uint32_t parity = (uint32_t)((j >> 1) & 1), bar = knn_smem_u32(&bars[stage]);
      uint32_t ok = 0;
      for (int spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1 << 24)) __trap();                                // (a broken descriptor must not hang the GPU)
      }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- selection epilogue: 64 columns of this thread's row ----
    const int col0 = ch * 64;                                          // first column (inside the tile) of this warp
    const int valid = min(kMmaN, t_end - (t_begin + j * kMmaN)) - col0; // columns of this warp that are train rows
    uint32_t m1 = 0, m2 = 0;                                           // packed (even column | odd column << 16), 0 = none
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(32 * lq) << 16) + (uint32_t)(stage * kMmaN + col0 + 32 * half), acc);
      if (valid >= 64) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const int cc = 32 * half + c;                                // column inside this warp's 64
          // low halves of two accumulators, sign bit flipped, column tags in the seven zero bits
          const uint32_t key = __byte_perm(acc[c], acc[c + 1], 0x5410) ^
                               (0x80008000u | (uint32_t)(127 - cc) | ((uint32_t)(126 - cc) << 16));
          const uint32_t lo = __vminu2(m1, key);
          m1 = __vmaxu2(m1, key);
          m2 = __vmaxu2(m2, lo);
        }
      } else {                                                         // the last tile of the train range: mask the tail
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const int cc = 32 * half + c;
          uint32_t key = __byte_perm(acc[c], acc[c + 1], 0x5410) ^
                         (0x80008000u | (uint32_t)(127 - cc) | ((uint32_t)(126 - cc) << 16));
          key &= (cc < valid ? 0xFFFFu : 0u) | (cc + 1 < valid ? 0xFFFF0000u : 0u);
          const uint32_t lo = __vminu2(m1, key);
          m1 = __vmaxu2(m1, key);
          m2 = __vmaxu2(m2, lo);
        }
      }
    }
    // the tile's two best of this row; decoded only if they can change the row's top-2 (later tiles have larger
    // indices: an equal distance never displaces an earlier entry)
    {
      const uint32_t a_lo = m1 & 0xFFFFu, a_hi = m1 >> 16;
      const uint32_t best = max(a_lo, a_hi);
      const uint32_t second = max(min(a_lo, a_hi), a_lo >= a_hi ? (m2 & 0xFFFFu) : (m2 >> 16));
      const int base_idx = t_begin + j * kMmaN + col0;
      auto decode = [&](uint32_t k16) {
        const int s64 = (int)(short)((k16 & 0xFF80u) ^ 0x8000u);       // 64 * (256 - 2 dist)
        const uint32_t dist = (uint32_t)(128 - (s64 >> 7));
        return (dist << 22) | (uint32_t)(base_idx + 127 - (int)(k16 & 0x7Fu));
      };
      if (best != 0) {
        const uint32_t kb = decode(best);
        if (kb < k1) {
          top2_insert(k0, k1, kb);
          if (second != 0) top2_insert(k0, k1, decode(second));
        }
      }
    }
  }
  // ---- merge the two column halves of a row, write this split's (k0, k1) ----
  merge[ch * kMmaM + 32 * lq + lane] = make_uint2(k0, k1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < kMmaM && q0 + tid < nq) {
    const uint2 a = merge[tid], c = merge[kMmaM + tid];
    uint32_t r0 = a.x, r1 = a.y;
    top2_insert(r0, r1, c.x);
    top2_insert(r0, r1, c.y);
    partial[((long long)b * nsplit + blockIdx.y) * max_nq + q0 + tid] = make_uint2(r0, r1);
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

__global__ void __launch_bounds__(1024)
knn_finish_kernel(const uint2* __restrict__ partial, const int32_t* __restrict__ nq_dev, int max_nq, int nsplit,
                  double ratio, uint2* __restrict__ best, mvo_dmatch* __restrict__ matches,
                  int32_t* __restrict__ nmatch) {
  const int b = blockIdx.x;
  const int nq = min(nq_dev[b], max_nq);
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int q0 = 0; q0 < nq; q0 += 1024) {
    const int qi = q0 + tid;
    uint32_t k0 = kInvalidKey, k1 = kInvalidKey;
    if (qi < nq) {
      for (int s = 0; s < nsplit; ++s) {
        const uint2 p = partial[((long long)b * nsplit + s) * max_nq + qi];
        const uint32_t n1 = min(max(k0, p.x), min(k1, p.y));
        k0 = min(k0, p.x);
        k1 = n1;
      }
      best[(long long)b * max_nq + qi] = make_uint2(k0, k1);
    }
    int ok = 0;
    if (qi < nq && k1 != kInvalidKey) {
      const float d0 = (float)(k0 >> 22), d1 = (float)(k1 >> 22);
      ok = ((double)d0 < ratio * (double)d1) ? 1 : 0;   // float < double * float, as in the reference
    }
    // ordered compaction: block exclusive scan of ok
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int within = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      int v = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
      }
      s_warp[lane] = v;  // inclusive
    }
    __syncthreads();
    const int base = s_base + (warp ? s_warp[warp - 1] : 0);
    if (ok) {
      mvo_dmatch m;
      m.query_idx = qi;
      m.train_idx = (int)(k0 & 0x3FFFFFu);
      m.img_idx = 0;
      m.distance = (float)(k0 >> 22);
      matches[(long long)b * max_nq + base + within] = m;
    }
    __syncthreads();
    if (tid == 0) s_base += s_warp[31];
    __syncthreads();
  }
  if (tid == 0) nmatch[b] = s_base;
}

int knn_prepare(mvo_ctx* c, int maxq) {
  const size_t B = (size_t)c->cfg.batch;
  MVO_CUDA_TRY(c, c->knn_best.alloc(B * (size_t)maxq * 2 * (kKnnMaxSplit + 1)));
  MVO_CUDA_TRY(c, c->knn_matches.alloc(B * (size_t)maxq));
  MVO_CUDA_TRY(c, c->knn_nmatch.alloc(B));
  return MVO_OK;
}

int knn_run(mvo_ctx* c, const uint8_t* q_dev, const int32_t* nq_dev, int q_stride_rows, int max_nq,
            const uint8_t* t_dev, const int32_t* nt_dev, int t_stride_rows, int max_nt, double ratio, int batch) {
  if (max_nt >= (1 << 22)) {
    c->set_error("train set too large (>= 2^22 rows)");
    return MVO_ERR_CAPACITY;
  }
  int rc = knn_prepare(c, max_nq);
  if (rc) return rc;
  uint2* best = reinterpret_cast<uint2*>(c->knn_best.p);
  uint2* partial = best + (size_t)batch * max_nq;
  const int qblocks = std::max(1, (max_nq + kKnnWarps * kKnnQ - 1) / (kKnnWarps * kKnnQ));
  const int tiles = std::max(1, (max_nt + kKnnTile - 1) / kKnnTile);
  int nsplit = (2 * 148 + qblocks * batch - 1) / (qblocks * batch);
  nsplit = std::max(1, std::min(std::min(nsplit, kKnnMaxSplit), tiles));
  if (c->dbg_knn_impl != 1) {
    // tensor-core path: 128-query CTAs, the train range split so that the grid covers the SMs several times
    static bool attr_set = false;
    if (!attr_set) {
      MVO_CUDA_TRY(c, cudaFuncSetAttribute(knn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKnnMmaSmem));
      attr_set = true;
    }
    const int qtiles = std::max(1, (max_nq + kMmaM - 1) / kMmaM);
    const int ttiles = std::max(1, (max_nt + kMmaN - 1) / kMmaN);
    nsplit = (6 * 148 + qtiles * batch - 1) / (qtiles * batch);
    nsplit = std::max(1, std::min(std::min(nsplit, kKnnMaxSplit), ttiles));
    knn_mma_kernel<<<dim3(qtiles, nsplit, batch), kMmaThreads, kKnnMmaSmem, c->stream>>>(
        q_dev, nq_dev, q_stride_rows, t_dev, nt_dev, t_stride_rows, partial, max_nq, nsplit);
  } else {
    dim3 grid(qblocks, nsplit, batch);
    knn_top2_kernel<<<grid, kKnnThreads, 0, c->stream>>>(q_dev, nq_dev, q_stride_rows, t_dev, nt_dev, t_stride_rows,
                                                        partial, max_nq, nsplit);
  }
  c->launches++;
  knn_finish_kernel<<<batch, 1024, 0, c->stream>>>(partial, nq_dev, max_nq, nsplit, ratio, best, c->knn_matches.p,
                                                  c->knn_nmatch.p);
  c->launches++;
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

}  // namespace mvo

// ------------------------------------------------------------------------------------------------
// Measured ceiling of the integer population-count pipe (the bound of knn_top2_kernel): every thread runs eight
// independent xor + popc chains, grid = 148 SMs x 8 CTAs.  Returns 32-bit popc per second.
namespace mvo {
__global__ void __launch_bounds__(256) popc_peak_kernel(uint32_t seed, int iters, uint32_t* out) {
  uint32_t x = seed ^ (blockIdx.x * 256u + threadIdx.x);
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 += __popc(x ^ a7);
    a1 += __popc(x ^ a0);
    a2 += __popc(x ^ a1);
    a3 += __popc(x ^ a2);
    a4 += __popc(x ^ a3);
    a5 += __popc(x ^ a4);
    a6 += __popc(x ^ a5);
    a7 += __popc(x ^ a6);
    x += 0x9e3779b9u;
  }
  if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0xdeadbeefu) out[0] = x;   // keep the chains alive
}
}  // namespace mvo

extern "C" int mvo_measure_popc_peak(mvo_ctx* c, double* popc_per_s) {
  if (!c || !popc_per_s) return MVO_ERR_INVALID;
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  MVO_CUDA_TRY(c, c->knn_counts.alloc(4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = 148 * 8, iters = 1 << 14;
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, c->stream);
    mvo::popc_peak_kernel<<<blocks, 256, 0, c->stream>>>(0x1234567u + rep, iters, reinterpret_cast<uint32_t*>(c->knn_counts.p));
    cudaEventRecord(e1, c->stream);
    MVO_CUDA_TRY(c, cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms > 0) best = std::max(best, (double)blocks * 256.0 * iters * 8.0 / (ms * 1e-3));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  c->launches += 4;
  *popc_per_s = best;
  return MVO_OK;
}
