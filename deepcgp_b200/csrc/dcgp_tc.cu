// Tensor-core path (sm_100a): tcgen05.mma with TMEM accumulators, operands staged by TMA (SWIZZLE_128B) through an
// mbarrier pipeline, warp-specialised (TMA producer / single-thread MMA issuer / 4 epilogue warps), persistent grid.
//
// K-D  `cond_tc_kernel`: the stacked conditional GEMM of SURVEY App. A.4
//        G[t, j] = sum_m K[t, m] * W[j, m]        (t: patch-columns, j: (R+1)*Mp rows of W, m: inducing points)
//      with the per-patch variance aggregation fused into the epilogue (acc[t, blk] = sum_{j in blk} G[t,j]^2 straight out
//      of TMEM, so the reference's [R,M,P,N] tensor of conditionals.py:58 never exists), plus the mean rows.
//      Precision: the reference is float64; fp16 tensor cores with fp32 accumulation are used through a 2-way operand split
//      x = hi + lo (22 mantissa bits), G ~= Kh*Wh + Kh*Wl + Kl*Wh  -> three MMAs per k-step (lo*lo ~ 2^-22 dropped).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "dcgp_tc.cuh"

namespace dcgp {

// ============================================================================================ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// One [rows x 64] k-block tile of a K-major operand: row-major planes [rows, K] (2-D map, coordinates (kb*64, row)) or
// k-blocked planes [K/64][rows][64] (3-D map, coordinates (0, row, kb)) -- the layout of the operands that are transposed in the
// patch index t (a^T, dd^T, patches^T, g_mean^T): a CTA's tile is one contiguous run in HBM instead of 128-byte rows at a
// stride of 2*Tpad bytes, for the kernels that write them as well as for the TMA that reads them.
__device__ __forceinline__ void tma_load_kb(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int kb, int row, int blocked) {
  if (blocked) tma_load_3d(smem_dst, tmap, bar, 0, row, kb);
  else tma_load_2d(smem_dst, tmap, bar, kb * 64, row);
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (K-major, 128 rows = lanes, two 16-bit k elements per 32-bit column) is
// read from tensor memory (cute SM100_MMA_F16BF16_TS).
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> tensor memory: 32 lanes x 16 columns of 32 bits (each thread writes 16 consecutive columns of its own lane)
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i), columns [col, col+32).
// 256-bit global store (sm_100: STG.256), 32-byte aligned
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_256(const void* ptr, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(ptr));
}
// Row-per-lane epilogues (the tcgen05.ld 32x32b layout: lane = tile row) and 16-bit planes [rows, ld]: a 32-column chunk is
// 64 bytes per row.  Instead of 32 rows x 16 bytes per instruction, lane pairs (rows t, t + 1) access half rows -- each
// 256-bit instruction covers 64 contiguous bytes of 16 rows, a quarter of the L1 wavefronts -- and swap halves by shuffle.
// Pointers are to the chunk in the EVEN row of the lane's pair, advanced by 16 columns on the odd lane.
// store: w[0..15] = the lane's 32 values (2 per word)
__device__ __forceinline__ void store_chunk_paired(void* even_row, int ld, const uint32_t* w, bool odd) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t send = odd ? w[i] : w[8 + i], keep = odd ? w[8 + i] : w[i];
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
    a[i] = odd ? recv : keep;        // a: the even row of the pair, b: the odd row
    b[i] = odd ? keep : recv;
  }
  st_global_256(even_row, a);
  st_global_256(reinterpret_cast<uint16_t*>(even_row) + ld, b);
}
// load, in two steps so that several chunks' loads are in flight before the first swap: lo8 / hi8 receive the raw halves,
// load_chunk_swap turns them into the lane's own columns 0..15 (lo8) and 16..31 (hi8)
__device__ __forceinline__ void load_chunk_paired(const void* even_row, int ld, uint32_t* lo8, uint32_t* hi8) {
  ld_global_nc_256(even_row, lo8);
  ld_global_nc_256(reinterpret_cast<const uint16_t*>(even_row) + ld, hi8);
}
__device__ __forceinline__ void load_chunk_swap(uint32_t* lo8, uint32_t* hi8, bool odd) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t a = lo8[i], b = hi8[i];                 // a: my half of the even row, b: my half of the odd row
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? a : b, 1);
    lo8[i] = odd ? recv : a;
    hi8[i] = odd ? b : recv;
  }
}
__device__ __forceinline__ void store_planes_paired(__half* oh, __half* ol, int ld, const __half2* hi, const __half2* lo, bool odd) {
  store_chunk_paired(oh, ld, reinterpret_cast<const uint32_t*>(hi), odd);
  store_chunk_paired(ol, ld, reinterpret_cast<const uint32_t*>(lo), odd);
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Same load without the wait: lets the next chunk's TMEM read overlap the current chunk's arithmetic (tmem_ld_wait() before
// the registers are used).
__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, rows of 128 bytes (64 fp16), 8-row groups 1024 B apart
// (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout_type=SWIZZLE_128B(2) [61,64).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // LBO: unused for swizzled K-major layouts
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// Instruction descriptor (InstrDescriptor in mma_sm100_desc.hpp): fp16 x fp16 -> fp32, both K-major, M=128, N=n.
__host__ __device__ constexpr uint32_t make_idesc_f16(int n) {
  return (1u << 4) /* c_format = F32 */ | (0u << 7) /* a = F16 */ | (0u << 10) /* b = F16 */ | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

constexpr uint32_t kIdescBf16 = (1u << 7) | (1u << 10);   // a_format = b_format = BF16

// ============================================================================================ K-D kernel
constexpr int kBM = 128;      // patch-columns per tile (UMMA M, TMEM lanes)
constexpr int kBK = 64;       // inducing points per pipeline stage (one 128-byte swizzle atom of fp16)
constexpr int kThreads = 192; // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue

template <int BN>
struct CondCfg {
  static constexpr int kStageA = kBM * kBK * 2;                 // one fp16 plane of the A tile
  static constexpr int kStageB = BN * kBK * 2;                  // one fp16 plane of the B tile
  static constexpr int kStageBytes = 2 * kStageA + 2 * kStageB; // hi+lo of both
  static constexpr int kStages = (BN == 256) ? 2 : (BN == 128 ? 3 : 4);
  static constexpr int kTmemCols = 2 * BN;                      // double-buffered accumulator
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

constexpr int MODE_COND = 0;   // conditional GEMM: per-row-block sums of squares + mean rows
constexpr int MODE_GEMM = 1;   // plain batched C = A * B^T: store fp32 and/or accumulate the sum of squares of C
constexpr int MODE_A = 2;      // chained form, first stage: a = K Lm^-T (lower-triangular operand: zero k-blocks skipped),
                               // acc[t, 0] = |a_t|^2 and the split-fp16 planes of a for the second stage
constexpr int MODE_AP = 3;     // MODE_A with the accumulation spread over four TMEM accumulators (BN = 128): the TMEM fp32
                               // accumulator rounds toward zero at every MMA (tools/diag_accum.py), a bias that the cancellation
                               // in Lm^-1 k amplifies.  The two low-order products go to an accumulator of their own (they are
                               // 2^-11 of the sum, so their truncations vanish) and the dominant product to up to three
                               // accumulators over consecutive k-chunks; the epilogue adds the four in fp32 round-to-nearest.
__host__ __device__ constexpr bool mode_is_a(int mode) { return mode == MODE_A || mode == MODE_AP; }
template <int MODE, int BN>
__host__ __device__ constexpr int tmem_cols() { return MODE == MODE_AP ? 4 * BN : 2 * BN; }
// MODE_AP: k-blocks of tile jt that are full width (the rest straddle the diagonal), and the number of k-chunks in use
template <int BN>
__device__ __forceinline__ void ap_chunks(int jt, int kb1, int& nfull, int& nch) {
  nfull = min(kb1, (jt * BN) / 64);
  nch = nfull == 0 ? 1 : min(3, nfull);
}

struct TcParams {
  int n_items;
  int nkb;        // K-dimension blocks of 64
  int nprod;      // split products per k-step: 3 = Al*Bh + Ah*Bl + Ah*Bh (22-bit operands), 2 = Al*Bh + Ah*Bh (A 22 bits, B 11),
                  // 4 = Ah*Bl + Ah*Bh (A 11 bits, B 22), 1 = Ah*Bh (11-bit operands); planes that are not multiplied are not loaded
  int bf16;       // MODE_GEMM: operand planes are bf16 (hi + lo) instead of fp16
  int kblocked;   // MODE_GEMM: both operands are k-blocked planes [K/64][rows][64] (3-D tensor maps)
  // ---- MODE_COND
  int T;          // valid patch-columns
  int Mp;         // padded inducing points (multiple of 64)
  int R;
  int njt;        // Mp / BN  (j-tiles per row block)
  const float* wscal;  // device: [0]=W scale, [1]=1/W scale, [2]=Wmean scale, [3]=1/Wmean scale
  const float* kscal;  // device: [0]=K scale, [1]=1/K scale
  float* acc;     // [T, R+1]
  float* mean;    // [T, R]
  int blk_first;  // MODE_COND: first W block of an item (1 when block 0 was done by the MODE_A stage)
  int tri;        // MODE_COND: 2 = the W blocks are upper triangular (C_r^T): k-blocks below the tile's first row are skipped
  __half* Ah_out; __half* Al_out;   // MODE_A: planes of a [Tpad, Mp]
  const float* ascal;               // MODE_A: {scale, 1/scale} of those planes
  // ---- MODE_GEMM: C[b][i, j] = sum_k A[b*a_batch_rows + i, k] * B[b*b_batch_rows + j, k]
  int m_tiles, n_tiles;
  int a_batch_rows, b_batch_rows;
  int m_valid, n_valid;
  const float* a_scal;   // device {scale, 1/scale} of the A planes
  const float* b_scal;   // device {scale, 1/scale} of the B planes
  float* C;              // may be null
  long long c_batch_stride;
  int ldc;
  double* sq_out;        // may be null: += sum of squares of all valid C entries
  float* absmax_out;     // may be null: atomic max of |C| over valid entries
  int splits;            // split-K: item = (split, batch, it, jn); partial results go to C + split * c_split_stride
  int nkb_split;         // k-blocks per split
  long long c_split_stride;
};

template <int MODE, int BN>
__device__ __forceinline__ int tiles_in_item(const TcParams& p, int item) {
  if (MODE == MODE_COND) { const int per = p.R + 2 - p.blk_first; return (item % per == per - 1) ? 1 : p.njt; }
  if (mode_is_a(MODE)) return p.njt;
  return 1;
}
template <int MODE, int BN>
__device__ __forceinline__ void tile_rows(const TcParams& p, int item, int jt, int& a_row, int& b_row) {
  if (MODE == MODE_COND) {
    const int per = p.R + 2 - p.blk_first;
    const int tt = item / per, blk = p.blk_first + item - tt * per;
    a_row = tt * kBM;
    b_row = blk * p.Mp + jt * BN;
  } else if (mode_is_a(MODE)) {
    a_row = item * kBM;
    b_row = jt * BN;
  } else {
    const int per = p.m_tiles * p.n_tiles;
    const int bb = item / per, rem = item - bb * per;
    const int b = bb % (p.splits > 0 ? (p.n_items / per / p.splits) : 1);
    const int it = rem / p.n_tiles, jn = rem - it * p.n_tiles;
    a_row = b * p.a_batch_rows + it * kBM;
    b_row = b * p.b_batch_rows + jn * BN;
  }
}
// k-block range of an item (split-K for MODE_GEMM)
template <int MODE, int BN>
__device__ __forceinline__ void item_krange(const TcParams& p, int item, int jt, int& kb0, int& kb1) {
  kb0 = 0; kb1 = p.nkb;
  if (mode_is_a(MODE)) kb1 = min(p.nkb, ((jt + 1) * BN + kBK - 1) / kBK);          // Lm^-1[j, m] = 0 for m > j
  if (MODE == MODE_COND && p.tri == 2) {
    const int per = p.R + 2 - p.blk_first;
    if (item % per != per - 1) kb0 = (jt * BN) / kBK;                              // C_r^T[j, i] = 0 for i < j
  }
  if (MODE == MODE_GEMM && p.splits > 1) {
    const int per = p.m_tiles * p.n_tiles;
    const int nbatch = p.n_items / per / p.splits;
    const int sp = (item / per) / nbatch;
    kb0 = sp * p.nkb_split;
    kb1 = min(p.nkb, kb0 + p.nkb_split);
  }
}

// Triangular operands at 64-column granularity.  A k-block that straddles the diagonal of the B operand only multiplies the
// accumulator columns whose B rows are not structurally zero: the MMA is issued with N = n (64 .. BN) on the B rows / TMEM
// columns [col_off, col_off + n), and only those rows are loaded.  The k-blocks of a tile are walked so that the FIRST one is
// full width (it initialises all BN accumulator columns): ascending for MODE_A (Lm^-1 lower triangular), descending for the
// C_r^T blocks of MODE_COND (upper triangular).
template <int MODE, int BN>
__device__ __forceinline__ bool tile_descending(const TcParams& p, int item) {
  if (MODE != MODE_COND || p.tri != 2) return false;
  const int per = p.R + 2 - p.blk_first;
  return item % per != per - 1;
}
template <int MODE, int BN>
__device__ __forceinline__ bool mean_item(const TcParams& p, int item) {   // MODE_COND: the item that carries the mean rows
  if (MODE != MODE_COND) return false;
  const int per = p.R + 2 - p.blk_first;
  return item % per == per - 1;
}
template <int MODE, int BN>
__device__ __forceinline__ void kb_cols(const TcParams& p, bool desc, int jt, int kb, int& col_off, int& n) {
  col_off = 0; n = BN;
  if (mode_is_a(MODE)) { col_off = max(0, kb * kBK - jt * BN); n = BN - col_off; }
  if (MODE == MODE_COND && desc) n = min(BN, kb * kBK + kBK - jt * BN);
}
constexpr int kMaxStages = 8;
// ring geometry for a given number of split products (the ring always owns Cfg::kStages * Cfg::kStageBytes of shared memory)
template <int BN>
__device__ __forceinline__ void ring_geom(int nprod, int& a_planes, int& b_planes, int& stage_bytes, int& n_stages) {
  using Cfg = CondCfg<BN>;
  a_planes = (nprod == 2 || nprod == 3) ? 2 : 1;
  b_planes = nprod >= 3 ? 2 : 1;                 // nprod = 4: A hi only, B hi + lo (Ah*Bl + Ah*Bh)
  stage_bytes = a_planes * Cfg::kStageA + b_planes * Cfg::kStageB;
  n_stages = min(kMaxStages, (Cfg::kStages * Cfg::kStageBytes) / stage_bytes);
}

template <int MODE, int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
          const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
          const __grid_constant__ CUtensorMap tmB64_hi, const __grid_constant__ CUtensorMap tmB64_lo, TcParams p) {
  using Cfg = CondCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024-B alignment
  uint64_t* full_bar = (uint64_t*)(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;             // [2]
  uint32_t* tmem_base_smem = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int a_planes, b_planes, stage_bytes, n_stages;
  ring_geom<BN>(p.nprod, a_planes, b_planes, stage_bytes, n_stages);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
    tma_prefetch_desc(&tmB64_hi); tma_prefetch_desc(&tmB64_lo);
    for (int s = 0; s < n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 128); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_base_smem, (tmem_cols<MODE, BN>()));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int njt = tiles_in_item<MODE, BN>(p, item);
        const bool desc = tile_descending<MODE, BN>(p, item);
        // MODE_COND mean rows (alpha^T a): always the full 22-bit product.  With fewer than four plane slots per stage the
        // k-range is walked three times with single planes in the hi slots: (A_lo, B_hi), (A_hi, B_lo), (A_hi, B_hi).
        const bool mean3 = mean_item<MODE, BN>(p, item) && p.nprod != 3;
        for (int jt = 0; jt < njt; ++jt) {
          int arow, brow, kb0, kb1;
          tile_rows<MODE, BN>(p, item, jt, arow, brow);
          item_krange<MODE, BN>(p, item, jt, kb0, kb1);
          for (int pass = 0; pass < (mean3 ? 3 : 1); ++pass) {
            for (int i = 0; i < kb1 - kb0; ++i) {
              const int kb = desc ? kb1 - 1 - i : kb0 + i;
              int col_off, n;
              kb_cols<MODE, BN>(p, desc, jt, kb, col_off, n);
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* st = smem + stage * stage_bytes;
              uint8_t* sb = st + a_planes * Cfg::kStageA;
              if (mean3) {
                mbar_expect_tx(&full_bar[stage], Cfg::kStageA + Cfg::kStageB);
                tma_load_kb(st, pass == 0 ? &tmA_lo : &tmA_hi, &full_bar[stage], kb, arow, p.kblocked);
                tma_load_kb(sb, pass == 1 ? &tmB_lo : &tmB_hi, &full_bar[stage], kb, brow, p.kblocked);
              } else {
                mbar_expect_tx(&full_bar[stage], a_planes * Cfg::kStageA + b_planes * n * (kBK * 2));
                tma_load_kb(st, &tmA_hi, &full_bar[stage], kb, arow, p.kblocked);
                if (a_planes == 2) tma_load_kb(st + Cfg::kStageA, &tmA_lo, &full_bar[stage], kb, arow, p.kblocked);
                if (n == BN) {
                  tma_load_kb(sb, &tmB_hi, &full_bar[stage], kb, brow, p.kblocked);
                  if (b_planes == 2) tma_load_kb(sb + Cfg::kStageB, &tmB_lo, &full_bar[stage], kb, brow, p.kblocked);
                } else {                                   // partial block: 64-row boxes of the non-zero rows only
                  for (int c = col_off; c < col_off + n; c += 64) {
                    tma_load_2d(sb + c * (kBK * 2), &tmB64_hi, &full_bar[stage], kb * kBK, brow + c);
                    if (b_planes == 2) tma_load_2d(sb + Cfg::kStageB + c * (kBK * 2), &tmB64_lo, &full_bar[stage], kb * kBK, brow + c);
                  }
                }
              }
              if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc0 = (make_idesc_f16(BN) & ~(0x3Fu << 17)) | (p.bf16 ? kIdescBf16 : 0u);
      int stage = 0; uint32_t phase = 0;
      uint32_t tile = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int njt = tiles_in_item<MODE, BN>(p, item);
        if (MODE == MODE_AP) {
          for (int jt = 0; jt < njt; ++jt, ++tile) {
            mbar_wait(&tmem_empty[0], (tile & 1) ^ 1);      // all four accumulators belong to one tile: no MMA / epilogue overlap
            tc_fence_after();
            int kb0, kb1, nfull, nch;
            item_krange<MODE, BN>(p, item, jt, kb0, kb1);
            ap_chunks<BN>(jt, kb1, nfull, nch);
            uint32_t acc_lo = 0, started = 0;
            for (int kb = 0; kb < kb1; ++kb) {
              int col_off, n;
              kb_cols<MODE, BN>(p, false, jt, kb, col_off, n);
              const int c = kb < nfull ? (kb * nch) / nfull : nch - 1;   // a chunk always starts with a full-width block
              const uint32_t idesc = idesc0 | ((uint32_t)(n >> 3) << 17);
              const uint32_t d_lo = tmem_base + col_off, d_hh = tmem_base + (1 + c) * BN + col_off;
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t a_hi = smem_u32(smem + stage * stage_bytes);
              const uint32_t a_lo = a_hi + Cfg::kStageA;
              const uint32_t b_hi = a_hi + 2 * Cfg::kStageA + col_off * (kBK * 2);
              const uint32_t b_lo = b_hi + Cfg::kStageB;
              const uint64_t dah = make_sw128_desc(a_hi), dal = make_sw128_desc(a_lo);
              const uint64_t dbh = make_sw128_desc(b_hi), dbl = make_sw128_desc(b_lo);
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                const uint64_t koff = (uint64_t)((k * 16 * 2) >> 4);
                umma_f16(d_lo, dal + koff, dbh + koff, idesc, acc_lo);
                acc_lo = 1;
                umma_f16(d_lo, dah + koff, dbl + koff, idesc, 1);
                umma_f16(d_hh, dah + koff, dbh + koff, idesc, (started >> c) & 1u);
                started |= 1u << c;
              }
              umma_commit(&empty_bar[stage]);
              if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tmem_full[0]);
          }
          continue;
        }
        const bool desc = tile_descending<MODE, BN>(p, item);
        const bool mean3 = mean_item<MODE, BN>(p, item) && p.nprod != 3;
        const bool two_a = !mean3 && a_planes == 2, two_b = !mean3 && b_planes == 2;
        for (int jt = 0; jt < njt; ++jt, ++tile) {
          const uint32_t buf = tile & 1, use = tile >> 1;
          mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);     // epilogue has drained this accumulator
          tc_fence_after();
          int kb0, kb1;
          item_krange<MODE, BN>(p, item, jt, kb0, kb1);
          uint32_t acc = 0;                               // the first MMA of the tile overwrites the accumulator
          for (int pass = 0; pass < (mean3 ? 3 : 1); ++pass) {
            for (int i = 0; i < kb1 - kb0; ++i) {
              const int kb = desc ? kb1 - 1 - i : kb0 + i;
              int col_off, n;
              kb_cols<MODE, BN>(p, desc, jt, kb, col_off, n);
              const uint32_t idesc = idesc0 | ((uint32_t)(n >> 3) << 17);
              const uint32_t d_tmem = tmem_base + buf * BN + col_off;
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t a_hi = smem_u32(smem + stage * stage_bytes);
              const uint32_t a_lo = a_hi + Cfg::kStageA;
              const uint32_t b_hi = a_hi + a_planes * Cfg::kStageA + col_off * (kBK * 2);
              const uint32_t b_lo = b_hi + Cfg::kStageB;
              const uint64_t dah = make_sw128_desc(a_hi), dal = make_sw128_desc(a_lo);
              const uint64_t dbh = make_sw128_desc(b_hi), dbl = make_sw128_desc(b_lo);
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                const uint64_t koff = (uint64_t)((k * 16 * 2) >> 4);   // advance 32 bytes inside the swizzle atom
                // small cross terms first, dominant term last
                if (two_a) { umma_f16(d_tmem, dal + koff, dbh + koff, idesc, acc); acc = 1; }
                if (two_b) { umma_f16(d_tmem, dah + koff, dbl + koff, idesc, acc); acc = 1; }
                umma_f16(d_tmem, dah + koff, dbh + koff, idesc, acc);
                acc = 1;
              }
              umma_commit(&empty_bar[stage]);                 // smem slot reusable once these MMAs retire
              if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
          }
          umma_commit(&tmem_full[buf]);                     // accumulator complete -> epilogue
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 4 warps <-> 4 TMEM lane quarters
    const int q = warp & 3;                                  // warps 2,3,4,5 -> quarters 2,3,0,1
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t tile = 0;
    if (MODE == MODE_COND) {
      const float inv_w = p.wscal[1], inv_wm = p.wscal[3], inv_k = p.kscal[1];
      const float sq_scale = (inv_w * inv_k) * (inv_w * inv_k);
      const float mean_scale = inv_wm * inv_k;
      const int per = p.R + 2 - p.blk_first;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int tt = item / per, blk = p.blk_first + item - tt * per;
        const bool is_mean = (blk == p.R + 1);
        const int njt = is_mean ? 1 : p.njt;
        const int t = tt * kBM + q * 32 + lane;
        float ssq = 0.f;
        for (int jt = 0; jt < njt; ++jt, ++tile) {
          const uint32_t buf = tile & 1, use = tile >> 1;
          mbar_wait(&tmem_full[buf], use & 1);
          tc_fence_after();
          const uint32_t taddr = tmem_base + lane_base + buf * BN;
          if (is_mean) {
            float v[32];
            tmem_ld_32x32(taddr, v);
            if (t < p.T) {
#pragma unroll
              for (int r = 0; r < 32; ++r)
                if (r < p.R) p.mean[(long long)t * p.R + r] = v[r] * mean_scale;
            }
            if (p.R > 32) {
              tmem_ld_32x32(taddr + 32, v);
              if (t < p.T) {
#pragma unroll
                for (int r = 0; r < 32; ++r)
                  if (32 + r < p.R) p.mean[(long long)t * p.R + 32 + r] = v[r] * mean_scale;
              }
            }
          } else {
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
              float v[32];
              tmem_ld_32x32(taddr + c, v);
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                s0 = fmaf(v[i], v[i], s0); s1 = fmaf(v[i + 1], v[i + 1], s1);
                s2 = fmaf(v[i + 2], v[i + 2], s2); s3 = fmaf(v[i + 3], v[i + 3], s3);
              }
              ssq += (s0 + s1) + (s2 + s3);
            }
          }
          tc_fence_before();
          mbar_arrive(&tmem_empty[buf]);
        }
        if (!is_mean && t < p.T) p.acc[(long long)t * (p.R + 1) + blk] = ssq * sq_scale;
      }
    } else if (mode_is_a(MODE)) {
      const float inv = p.wscal[1] * p.kscal[1];               // accumulator -> a
      const float sa = p.ascal[0];
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const long long t = (long long)item * kBM + q * 32 + lane;   // < Tpad: padding rows hold zeros (K rows are zero)
        float ssq = 0.f;
        for (int jt = 0; jt < p.njt; ++jt, ++tile) {
          const uint32_t buf = MODE == MODE_AP ? 0u : (tile & 1), use = MODE == MODE_AP ? tile : (tile >> 1);
          mbar_wait(&tmem_full[buf], use & 1);
          tc_fence_after();
          const uint32_t taddr = tmem_base + lane_base + buf * BN;
          const bool odd = lane & 1;                       // row pairs (t - odd, t - odd + 1): see store_planes_paired
          __half* oh = p.Ah_out + (t - (odd ? 1 : 0)) * p.Mp + jt * BN + (odd ? 16 : 0);
          __half* ol = p.Al_out + (t - (odd ? 1 : 0)) * p.Mp + jt * BN + (odd ? 16 : 0);
          int nch = 0;
          if (MODE == MODE_AP) {
            int nfull;
            ap_chunks<BN>(jt, min(p.nkb, ((jt + 1) * BN + kBK - 1) / kBK), nfull, nch);
          }
#pragma unroll 1
          for (int c = 0; c < BN; c += 32) {
            float v[32];
            if (MODE == MODE_AP) {     // (chunk 0 + chunk 1 + chunk 2) + low-order products, fp32 round-to-nearest
              float u[32];
              tmem_ld_32x32(taddr + BN + c, v);
              for (int ch = 1; ch < nch; ++ch) {
                tmem_ld_32x32(taddr + (1 + ch) * BN + c, u);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += u[i];
              }
              tmem_ld_32x32(taddr + c, u);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += u[i];
            } else {
              tmem_ld_32x32(taddr + c, v);
            }
            __align__(16) __half2 hi[16];
            __align__(16) __half2 lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a0 = v[2 * i] * inv, a1 = v[2 * i + 1] * inv;
              ssq = fmaf(a0, a0, fmaf(a1, a1, ssq));
              const float s0 = a0 * sa, s1 = a1 * sa;
              hi[i] = __floats2half2_rn(s0, s1);
              const float2 hf = __half22float2(hi[i]);
              lo[i] = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
            }
            store_planes_paired(oh + c, ol + c, p.Mp, hi, lo, odd);
          }
          tc_fence_before();
          mbar_arrive(&tmem_empty[buf]);
        }
        if (t < p.T) p.acc[t * (p.R + 1)] = ssq;
      }
    } else {
      const float inv = p.a_scal[1] * p.b_scal[1];
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++tile) {
        const int per = p.m_tiles * p.n_tiles;
        const int bb = item / per, rem = item - bb * per;
        const int nbatch = p.splits > 0 ? (p.n_items / per / p.splits) : (p.n_items / per);
        const int b = bb % nbatch, sp = bb / nbatch;
        const int it = rem / p.n_tiles, jn = rem - it * p.n_tiles;
        const int row = it * kBM + q * 32 + lane;
        const uint32_t buf = tile & 1, use = tile >> 1;
        mbar_wait(&tmem_full[buf], use & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_base + buf * BN;
        float ssq = 0.f, amax = 0.f;
        float* crow = p.C ? p.C + (long long)sp * p.c_split_stride + (long long)b * p.c_batch_stride + (long long)row * p.ldc : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          float v[32];
          tmem_ld_32x32(taddr + c, v);
          const int col0 = jn * BN + c;
          if (row < p.m_valid && col0 < p.n_valid) {
            if (col0 + 32 <= p.n_valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) { v[i] *= inv; ssq = fmaf(v[i], v[i], ssq); amax = fmaxf(amax, fabsf(v[i])); }
              if (crow) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  *reinterpret_cast<float4*>(crow + col0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.n_valid) {
                  const float x = v[i] * inv;
                  ssq = fmaf(x, x, ssq);
                  amax = fmaxf(amax, fabsf(x));
                  if (crow) crow[col0 + i] = x;
                }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[buf]);
        if (p.sq_out) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
          if (lane == 0) atomicAdd(p.sq_out, (double)ssq);
        }
        if (p.absmax_out) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
          if (lane == 0) atomicMax(reinterpret_cast<int*>(p.absmax_out), __float_as_int(amax));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (tmem_cols<MODE, BN>()));
  }
}

// ============================================================================================ host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}
// 2-D fp16 row-major [rows, cols] tensor, box = [box_rows, 64 cols] (128 bytes inner), SWIZZLE_128B.
static int make_tmap_f16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return DCGP_ERR_CUDA; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box=%u", (int)r, (unsigned long long)rows, (unsigned long long)cols, box_rows); return DCGP_ERR_CUDA; }
  return DCGP_OK;
}

// k-blocked planes [K/64][rows][64] (2-byte elements), box = [box_rows, 64] of one k-block, SWIZZLE_128B: lands in shared
// memory exactly like the 2-D box of make_tmap_f16.
static int make_tmap_f16_blocked(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t kcols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return DCGP_ERR_CUDA; }
  cuuint64_t dims[3] = {64, rows, kcols / 64};
  cuuint64_t strides[2] = {128, rows * 128};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (k-blocked) failed (%d) rows=%llu k=%llu box=%u", (int)r, (unsigned long long)rows, (unsigned long long)kcols, box_rows); return DCGP_ERR_CUDA; }
  return DCGP_OK;
}
static int make_tmap_kmajor(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t kcols, uint32_t box_rows, int blocked) {
  return blocked ? make_tmap_f16_blocked(tm, base, rows, kcols, box_rows) : make_tmap_f16(tm, base, rows, kcols, box_rows);
}

// Optional live kernel timing (bench.py's roofline leg): CUDA events recorded on the launching stream around the
// conditional-GEMM and Kuf kernels when enabled; dcgp_kernel_ms() synchronises on the end event and returns the duration.
static int g_timing = 0;
constexpr int kTimers = 4;   // 0 = conditional GEMM, 1 = Kuf, 2 = dK (+dd) GEMM, 3 = dQ GEMM
static cudaEvent_t g_ev[kTimers][2];
static bool g_ev_init = false, g_ev_used[kTimers] = {false, false, false, false};
void tc_set_timing(int on) {
  g_timing = on;
  if (on && !g_ev_init) {
    for (int i = 0; i < kTimers; ++i) for (int j = 0; j < 2; ++j) cudaEventCreate(&g_ev[i][j]);
    g_ev_init = true;
  }
}
double tc_kernel_ms(int which) {
  if (!g_ev_init || which < 0 || which >= kTimers || !g_ev_used[which]) return -1.0;
  float ms = 0.f;
  cudaEventSynchronize(g_ev[which][1]);
  if (cudaEventElapsedTime(&ms, g_ev[which][0], g_ev[which][1]) != cudaSuccess) return -1.0;
  return (double)ms;
}
static double g_flops[kTimers] = {0, 0, 0, 0};   // executed tensor-pipe flops of the last timed launch(es) of each kind
double tc_kernel_flops(int which) { return (which >= 0 && which < kTimers) ? g_flops[which] : 0.0; }
struct ScopedTimer {
  int which; cudaStream_t st; bool on;
  ScopedTimer(int w, cudaStream_t s) : which(w), st(s), on(g_timing && g_ev_init) { if (on) cudaEventRecord(g_ev[which][0], st); }
  void flops(double f) { if (on) g_flops[which] = f; }
  ~ScopedTimer() { if (on) { cudaEventRecord(g_ev[which][1], st); g_ev_used[which] = true; } }
};

// How many of the three split products each T-sized GEMM family issues (see TcParams::nprod).  The first stage of the
// conditional (a = Lm^-1 k: cancellation) and the small GEMMs always use 3.  Defaults can be overridden with
// DCGP_PROD_COND / DCGP_PROD_DK / DCGP_PROD_DQ or dcgp_set_products().
static TcProducts g_prod = {0, 0, 0};
static int env_prod(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : dflt;
  return (v >= 1 && v <= 4) ? v : dflt;
}
const TcProducts& tc_products() {
  if (!g_prod.cond) {
    g_prod.cond = env_prod("DCGP_PROD_COND", kDefaultProdCond);
    g_prod.dk = env_prod("DCGP_PROD_DK", kDefaultProdDk);
    g_prod.dq = env_prod("DCGP_PROD_DQ", kDefaultProdDq);
  }
  return g_prod;
}
void tc_set_products(int cond, int dk, int dq) {
  tc_products();
  if (cond >= 1 && cond <= 4) g_prod.cond = cond;
  if (dk >= 1 && dk <= 4) g_prod.dk = dk;
  if (dq >= 1 && dq <= 4) g_prod.dq = dq;
}

// Stage 1 of the chained conditional with the accumulation spread over four TMEM accumulators (MODE_AP): 1 (default) =
// whenever M is padded to a multiple of 128, 0 = never, -1 = from kPreciseAutoM inducing points on.  Measured
// (tools/diag_accum.py): +2 % on the conditional at M = 512, +1 % at M = 1024; mean error at cond(Kuu) = 1e4, M = 1024:
// 1.7e-4 -> 3.5e-5.  DCGP_PRECISE_STAGE1 / dcgp_set_precise_stage1().
static int g_precise = -2;
void tc_set_precise_stage1(int mode) { g_precise = mode < 0 ? -1 : (mode ? 1 : 0); }
int tc_get_precise_stage1() {
  if (g_precise == -2) { const char* e = getenv("DCGP_PRECISE_STAGE1"); g_precise = e ? (atoi(e) < 0 ? -1 : (atoi(e) ? 1 : 0)) : 1; }
  return g_precise;
}
bool tc_precise_stage1(int Mp) {
  const int mode = tc_get_precise_stage1();
  return mode > 0 || (mode < 0 && Mp >= kPreciseAutoM);
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// SMs the persistent tensor kernels leave free (grid = SMs - reserve).  A host that runs latency-critical small kernels on
// another stream underneath a long persistent GEMM (grad.TrainStep: a layer's chain rule + optimiser update + next-step
// prepare under the parameter-only GEMMs of the layers above) sets it around those launches: CTAs of a persistent kernel
// never yield, so without free SMs the other stream only advances in the gaps between kernels.
static int g_reserve_sms = 0;
void tc_set_reserved_sms(int n) { g_reserve_sms = n; }
// n < 0: one CTA per work item instead of a persistent grid -- CTAs retire continuously, so a higher-priority stream gets SMs
// at every CTA boundary (a few microseconds apart) without any SM being taken away from this kernel for good.
static int grid_sms() { if (g_reserve_sms < 0) return 1 << 30; const int n = num_sms() - g_reserve_sms; return n < 1 ? 1 : n; }

constexpr int kWPadRows = 256;   // zero rows after the mean block so a BN-row box never leaves the tensor

static size_t w_rows(int Mp, int R) { return (size_t)(R + 1) * Mp + kWPadRows; }

template <int MODE, int BN>
static int launch_tc(const CUtensorMap& tmAh, const CUtensorMap& tmAl, const CUtensorMap& tmBh, const CUtensorMap& tmBl,
                     const CUtensorMap& tmB64h, const CUtensorMap& tmB64l, const TcParams& p, cudaStream_t st) {
  using Cfg = CondCfg<BN>;
  if (p.nprod < 1 || p.nprod > 4) { set_error("tc_kernel: nprod must be 1 .. 4"); return DCGP_ERR_ARG; }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("tc_kernel smem attr: %s", cudaGetErrorString(e)); return DCGP_ERR_CUDA; }
    attr = true;
  }
  const int grid = p.n_items < grid_sms() ? p.n_items : grid_sms();
  if (grid <= 0) return DCGP_OK;
  tc_kernel<MODE, BN><<<grid, kThreads, Cfg::kSmemBytes, st>>>(tmAh, tmAl, tmBh, tmBl, tmB64h, tmB64l, p);
  return check_launch("tc_kernel");
}

template <int BN>
static int launch_cond_tc(const TcPrep& prep, const TcCondWork& w, int T, int Mp, int R, float* acc, float* mean, cudaStream_t st) {
  CUtensorMap tmAh, tmAl, tmBh, tmBl;
  int rc;
  if ((rc = make_tmap_f16(&tmAh, w.Kh, w.Tpad, Mp, kBM))) return rc;
  if ((rc = make_tmap_f16(&tmAl, w.Kl, w.Tpad, Mp, kBM))) return rc;
  if ((rc = make_tmap_f16(&tmBh, prep.Wh, w_rows(Mp, R), Mp, BN))) return rc;
  if ((rc = make_tmap_f16(&tmBl, prep.Wl, w_rows(Mp, R), Mp, BN))) return rc;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.Mp = Mp; p.R = R; p.njt = Mp / BN; p.nkb = Mp / kBK;
  p.n_items = ceil_div(T, kBM) * (R + 2);
  p.wscal = prep.scal; p.kscal = w.kscal; p.acc = acc; p.mean = mean;
  p.nprod = 3;
  ScopedTimer timer(0, st);
  timer.flops(3 * 2.0 * (double)w.Tpad * ((double)(R + 1) * Mp + BN) * Mp);
  return launch_tc<MODE_COND, BN>(tmAh, tmAl, tmBh, tmBl, tmBh, tmBl, p, st);
}

__global__ void set_scale_kernel(float bound, float* __restrict__ scal2);

bool tc_forward_chained() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DCGP_FWD_CHAINED"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}

// Chained conditional (conditionals.py:31-65 as two triangular GEMMs instead of one dense stack):
//   stage 1 (MODE_A):   a = K Lm^-T         -> acc[:, 0] = |a|^2, planes of a        (Lm^-1 lower triangular: 3/4 of the k-blocks)
//   stage 2 (MODE_COND) G_r = a C_r, mean   -> acc[:, r] = |G_r|^2, mean             (C_r^T upper triangular: 3/4 of the k-blocks)
// |a|^2 <= k(x, x) (the conditional variance is non-negative), so sqrt(a_bound) bounds every entry of a: data-independent scale.
// scal2 = scale pair for bound * max(1e-30, max_p |w_p|)   (image-level rows: |a|^2 <= Kdiag <= variance max|w|^2)
__global__ void set_scale_w_kernel(float bound, const double* __restrict__ w, int P, float* __restrict__ scal2) {
  __shared__ float sh[32];
  float m = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) m = fmaxf(m, (float)fabs(w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, sh[i]);
    int e = 0;
    frexpf(bound * fmaxf(m, 1e-30f), &e);
    const float s = ldexpf(1.f, 14 - e);
    scal2[0] = s;
    scal2[1] = 1.f / s;
  }
}

template <int BN>
static int launch_cond_chained(const TcPrep& prep, const TcCondWork& w, int T, int Mp, int R, float a_bound, const double* pw,
                               int P, float* acc, float* mean, cudaStream_t st) {
  CUtensorMap tmKh, tmKl, tmAh, tmAl, tmBh, tmBl, tmB64h, tmB64l;
  int rc;
  if ((rc = make_tmap_f16(&tmKh, w.Kh, w.Tpad, Mp, kBM))) return rc;
  if ((rc = make_tmap_f16(&tmKl, w.Kl, w.Tpad, Mp, kBM))) return rc;
  if ((rc = make_tmap_f16(&tmAh, w.Ah, w.Tpad, Mp, kBM))) return rc;
  if ((rc = make_tmap_f16(&tmAl, w.Al, w.Tpad, Mp, kBM))) return rc;
  if ((rc = make_tmap_f16(&tmBh, prep.Wh, w_rows(Mp, R), Mp, BN))) return rc;
  if ((rc = make_tmap_f16(&tmBl, prep.Wl, w_rows(Mp, R), Mp, BN))) return rc;
  if ((rc = make_tmap_f16(&tmB64h, prep.Wh, w_rows(Mp, R), Mp, 64))) return rc;
  if ((rc = make_tmap_f16(&tmB64l, prep.Wl, w_rows(Mp, R), Mp, 64))) return rc;
  if (pw) set_scale_w_kernel<<<1, 256, 0, st>>>(sqrtf(a_bound), pw, P, w.ascal);
  else set_scale_kernel<<<1, 1, 0, st>>>(sqrtf(a_bound), w.ascal);
  if ((rc = check_launch("set_scale"))) return rc;
  ScopedTimer timer(0, st);
  const int nprod2 = tc_products().cond;
  const int nmul2 = nprod2 == 4 ? 2 : nprod2;
  {   // executed tensor flops of the two launches: 64-column granularity on the triangular operands
    const int nb = Mp / kBK;                                   // 64-blocks per side
    const double tri_blocks = 0.5 * nb * (nb + 1);             // (k-block, 64-column block) pairs that are not structurally zero
    const double per_blk = 2.0 * (double)w.Tpad * kBK * kBK;   // one 64 x 64 block pair over all patch columns
    timer.flops(3 * per_blk * tri_blocks + nmul2 * per_blk * tri_blocks * R + 3 * 2.0 * (double)w.Tpad * BN * Mp);
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.Mp = Mp; p.R = R; p.njt = Mp / BN; p.nkb = Mp / kBK;
  p.n_items = ceil_div(T, kBM);
  p.wscal = prep.scal; p.kscal = w.kscal; p.acc = acc; p.mean = mean;
  p.Ah_out = (__half*)w.Ah; p.Al_out = (__half*)w.Al; p.ascal = w.ascal;
  p.nprod = 3;                                                 // a = Lm^-1 k cancels: always the full 22-bit product
  if (tc_precise_stage1(Mp) && Mp % 128 == 0) {                // four accumulators of 128 columns (MODE_AP)
    CUtensorMap tmB128h = tmBh, tmB128l = tmBl;
    if (BN != 128) {
      if ((rc = make_tmap_f16(&tmB128h, prep.Wh, w_rows(Mp, R), Mp, 128))) return rc;
      if ((rc = make_tmap_f16(&tmB128l, prep.Wl, w_rows(Mp, R), Mp, 128))) return rc;
    }
    p.njt = Mp / 128;
    if ((rc = launch_tc<MODE_AP, 128>(tmKh, tmKl, tmB128h, tmB128l, tmB64h, tmB64l, p, st))) return rc;
    p.njt = Mp / BN;
  } else if ((rc = launch_tc<MODE_A, BN>(tmKh, tmKl, tmBh, tmBl, tmB64h, tmB64l, p, st))) {
    return rc;
  }
  p.n_items = ceil_div(T, kBM) * (R + 1);
  p.blk_first = 1; p.tri = 2; p.kscal = w.ascal;
  p.nprod = nprod2;                                            // G_r = C_r^T a feeds a sum of squares (no cancellation)
  return launch_tc<MODE_COND, BN>(tmAh, tmAl, tmBh, tmBl, tmB64h, tmB64l, p, st);
}

int tc_cond_chained(const TcPrep& prep, const TcCondWork& w, int T, int Mp, int R, float a_bound, float* acc, float* mean,
                    cudaStream_t st, const double* patch_weights, int P) {
  if (R > 64) { set_error("tc_cond: R > 64 unsupported"); return DCGP_ERR_ARG; }
  if (!w.Ah) { set_error("tc_cond_chained: no workspace for the a planes"); return DCGP_ERR_ARG; }
  if (Mp % 256 == 0) return launch_cond_chained<256>(prep, w, T, Mp, R, a_bound, patch_weights, P, acc, mean, st);
  if (Mp % 128 == 0) return launch_cond_chained<128>(prep, w, T, Mp, R, a_bound, patch_weights, P, acc, mean, st);
  return launch_cond_chained<64>(prep, w, T, Mp, R, a_bound, patch_weights, P, acc, mean, st);
}

// Batched C[b] = A[b] * B[b]^T on split-fp16 planes (both K-major, row-stacked batches); rows/K padded by the caller.
int tc_gemm(const TcGemm& g, cudaStream_t st) {
  const int BN = (g.n_pad % 256 == 0) ? 256 : (g.n_pad % 128 == 0 ? 128 : 64);
  CUtensorMap tmAh, tmAl, tmBh, tmBl;
  int rc;
  if ((rc = make_tmap_kmajor(&tmAh, g.Ah, g.a_rows_total, g.k_pad, kBM, g.kblocked))) return rc;
  if ((rc = make_tmap_kmajor(&tmAl, g.Al, g.a_rows_total, g.k_pad, kBM, g.kblocked))) return rc;
  if ((rc = make_tmap_kmajor(&tmBh, g.Bh, g.b_rows_total, g.k_pad, BN, g.kblocked))) return rc;
  if ((rc = make_tmap_kmajor(&tmBl, g.Bl, g.b_rows_total, g.k_pad, BN, g.kblocked))) return rc;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.nkb = g.k_pad / kBK;
  p.bf16 = g.bf16;
  p.kblocked = g.kblocked;
  p.m_tiles = ceil_div(g.m_pad, kBM); p.n_tiles = g.n_pad / BN;   // a 128-row box may run past a batch / the tensor: extra rows are discarded
  p.splits = g.splits > 1 ? g.splits : 1;
  p.nkb_split = ceil_div(p.nkb, p.splits);
  p.splits = ceil_div(p.nkb, p.nkb_split);          // drop empty splits
  p.c_split_stride = g.c_split_stride;
  p.absmax_out = g.absmax_out;
  p.n_items = p.splits * g.batch * p.m_tiles * p.n_tiles;
  p.a_batch_rows = g.a_batch_rows; p.b_batch_rows = g.b_batch_rows;
  p.m_valid = g.m; p.n_valid = g.n;
  p.a_scal = g.a_scal; p.b_scal = g.b_scal;
  p.C = g.C; p.c_batch_stride = g.c_batch_stride; p.ldc = g.ldc; p.sq_out = g.sq_out;
  p.nprod = g.nprod ? g.nprod : 3;
  if (BN == 256) return launch_tc<MODE_GEMM, 256>(tmAh, tmAl, tmBh, tmBl, tmBh, tmBl, p, st);
  if (BN == 128) return launch_tc<MODE_GEMM, 128>(tmAh, tmAl, tmBh, tmBl, tmBh, tmBl, p, st);
  return launch_tc<MODE_GEMM, 64>(tmAh, tmAl, tmBh, tmBl, tmBh, tmBl, p, st);
}

int tc_cond(const TcPrep& prep, const TcCondWork& w, int T, int Mp, int R, float* acc, float* mean, cudaStream_t st) {
  if (R > 64) { set_error("tc_cond: R > 64 unsupported"); return DCGP_ERR_ARG; }
  if (Mp % 256 == 0) return launch_cond_tc<256>(prep, w, T, Mp, R, acc, mean, st);
  if (Mp % 128 == 0) return launch_cond_tc<128>(prep, w, T, Mp, R, acc, mean, st);
  return launch_cond_tc<64>(prep, w, T, Mp, R, acc, mean, st);
}

// ============================================================================================ operand splitting
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

// ---- power-of-two scaling: planes hold x * s with s = 2^(14 - e), max|x| = f * 2^e (f in [0.5,1)), so that the largest
// entry lands in [2^13, 2^14) (fp16 max is 65504) and small entries keep their 22 bits.
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {   // valid for v >= 0 (IEEE order == int order)
  atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}
__device__ __forceinline__ float block_max_256(float m) {
  __shared__ float sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) m = sh[threadIdx.x]; else m = 0.f;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  return m;
}
// mx[slot] = max(mx[slot], max |a[r*ld + c]|) over rows x cols (mx zeroed beforehand); lower_period > 0 masks c > r % period
__global__ void __launch_bounds__(256) maxabs_f64_kernel(const double* __restrict__ a, long long rows, int cols, int ld,
                                                         int lower_period, float* __restrict__ mx) {
  float m = 0.f;
  const long long n = rows * cols;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < n; e += 256LL * gridDim.x) {
    const long long r = e / cols;
    const int c = (int)(e % cols);
    if (lower_period > 0 && c > (int)(r % lower_period)) continue;
    m = fmaxf(m, (float)fabs(a[r * ld + c]));
  }
  m = block_max_256(m);
  if (threadIdx.x == 0) atomic_max_nonneg(mx, m);
}
__global__ void __launch_bounds__(256) maxabs_f32_kernel(const float* __restrict__ a, long long n, float* __restrict__ mx) {
  float m = 0.f;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < n; e += 256LL * gridDim.x) m = fmaxf(m, fabsf(a[e]));
  m = block_max_256(m);
  if (threadIdx.x == 0) atomic_max_nonneg(mx, m);
}
__global__ void scales_from_max_kernel(const float* __restrict__ mx, int first, int count, float* __restrict__ scal) {
  const int i = first + threadIdx.x;
  if (threadIdx.x < count) {
    const float m = mx[i];
    int e = 0;
    if (m > 0.f && isfinite(m)) frexpf(m, &e);
    const float s = ldexpf(1.f, 14 - e);
    scal[2 * i] = s;
    scal[2 * i + 1] = 1.f / s;
  }
}
__global__ void set_pair_kernel(float sc, float* __restrict__ scal2) { scal2[0] = sc; scal2[1] = 1.f / sc; }
__global__ void set_scale_kernel(float bound, float* __restrict__ scal2) {
  int e = 0;
  frexpf(bound, &e);
  const float s = ldexpf(1.f, 14 - e);
  scal2[0] = s;
  scal2[1] = 1.f / s;
}
// scale pair of a plane set from the running maximum of its source (largest entry -> [2^13, 2^14)); the kernel that packs
// the planes computes it itself from the maximum (no separate launch) and publishes {scale, 1/scale} for the GEMM epilogues
// index split in 32-bit arithmetic for the element-wise packing kernels (their element counts are below 2^31; a 64-bit
// division by a run-time divisor costs more than the rest of such a kernel)
__device__ __forceinline__ int div32(long long e, int d, int& rem) {
  const unsigned u = (unsigned)e, q = u / (unsigned)d;
  rem = (int)(u - q * (unsigned)d);
  return (int)q;
}
__device__ __forceinline__ float scale_from_max(float m) {
  int e = 0;
  if (m > 0.f && isfinite(m)) frexpf(m, &e);
  return ldexpf(1.f, 14 - e);
}
__device__ __forceinline__ float pack_scale(const float* __restrict__ mx, float* __restrict__ scal2) {
  if (!mx) return scal2[0];
  const float s = scale_from_max(mx[0]);
  if (blockIdx.x == 0 && threadIdx.x == 0) { scal2[0] = s; scal2[1] = 1.f / s; }
  return s;
}
// Several max|x| scans in ONE launch (blockIdx.y = job): out[job] = max(out[job], max |a|) like maxabs_f64 / maxabs_f32.
struct MaxJob { const void* p; int f64; long long rows; int cols, ld, lower_period; float* out; };
struct MaxJobs { MaxJob j[4]; };
__global__ void __launch_bounds__(256) maxabs_jobs_kernel(MaxJobs J) {
  const MaxJob job = J.j[blockIdx.y];
  float m = 0.f;
  if (job.cols == 1) {                            // flat array: no index arithmetic, 16-byte loads where aligned
    const long long n = job.rows;
    if (job.f64) {
      const double* p = (const double*)job.p;
      for (long long e = blockIdx.x * 256LL + threadIdx.x; e < n; e += 256LL * gridDim.x) m = fmaxf(m, (float)fabs(p[e]));
    } else {
      const float* p = (const float*)job.p;
      const long long n4 = (((uintptr_t)p & 15) == 0) ? n / 4 : 0;
      for (long long e = blockIdx.x * 256LL + threadIdx.x; e < n4; e += 256LL * gridDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p) + e);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
      for (long long e = 4 * n4 + blockIdx.x * 256LL + threadIdx.x; e < n; e += 256LL * gridDim.x) m = fmaxf(m, fabsf(p[e]));
    }
  } else {                                        // one warp per row: one division per row instead of three per element
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < job.rows; r += nwarps) {
      const int cend = job.lower_period > 0 ? min(job.cols, (int)(r % job.lower_period) + 1) : job.cols;
      if (job.f64) {
        const double* row = (const double*)job.p + r * job.ld;
        for (int c = lane; c < cend; c += 32) m = fmaxf(m, (float)fabs(row[c]));
      } else {
        const float* row = (const float*)job.p + r * job.ld;
        for (int c = lane; c < cend; c += 32) m = fmaxf(m, fabsf(row[c]));
      }
    }
  }
  m = block_max_256(m);
  if (threadIdx.x == 0) atomic_max_nonneg(job.out, m);
}
static MaxJob max_f64(const double* a, long long rows, int cols, int ld, int lower_period, float* out) {
  MaxJob j; j.p = a; j.f64 = 1; j.rows = rows; j.cols = cols; j.ld = ld; j.lower_period = lower_period; j.out = out; return j;
}
static MaxJob max_f32(const float* a, long long n, float* out) {      // a flat array: n rows of one column
  MaxJob j; j.p = a; j.f64 = 0; j.rows = n; j.cols = 1; j.ld = 1; j.lower_period = 0; j.out = out; return j;
}
static int maxabs_jobs(const MaxJob* jobs, int n, cudaStream_t st);

static int grid_for(long long n, int per_block) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > 148 * 4) b = 148 * 4;
  return (int)b;
}
static int maxabs_jobs(const MaxJob* jobs, int n, cudaStream_t st) {
  if (n < 1 || n > 4) { set_error("maxabs_jobs: 1..4 jobs"); return DCGP_ERR_ARG; }
  MaxJobs J;
  long long most = 0;
  for (int i = 0; i < n; ++i) { J.j[i] = jobs[i]; const long long e = jobs[i].rows * jobs[i].cols; most = e > most ? e : most; }
  for (int i = n; i < 4; ++i) J.j[i] = jobs[0];
  maxabs_jobs_kernel<<<dim3(grid_for(most, 2048), n), 256, 0, st>>>(J);
  return check_launch("maxabs_jobs");
}
static int maxabs_f64(const double* a, long long rows, int cols, int ld, int lower_period, float* mx, cudaStream_t st) {
  maxabs_f64_kernel<<<grid_for(rows * cols, 2048), 256, 0, st>>>(a, rows, cols, ld, lower_period, mx);
  return check_launch("maxabs_f64");
}
static int maxabs_f32(const float* a, long long n, float* mx, cudaStream_t st) {
  maxabs_f32_kernel<<<grid_for(n, 4096), 256, 0, st>>>(a, n, mx);
  return check_launch("maxabs_f32");
}

// Generic fp64 -> split-fp16 planes: dst[b*rows_pad + i, j] = s * src_b(i, j), where src_b(i,j) = src[b*bstride + i*ld + j]
// (or its transpose src[b*bstride + j*ld + i]); `lower` zeroes STORED entries above the diagonal; zero padding elsewhere.
__global__ void __launch_bounds__(256) pack_planes_f64_kernel(const double* __restrict__ src, int ld, long long bstride, int rows,
                                                              int cols, int transpose, int lower, int batch, int rows_pad,
                                                              int cols_pad, const float* __restrict__ mx,
                                                              float* __restrict__ scal2,
                                                              __half* __restrict__ Ph, __half* __restrict__ Pl) {
  const double s = (double)pack_scale(mx, scal2);
  const long long total = (long long)batch * rows_pad * cols_pad;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    int i, j;
    const int q = div32(e, cols_pad, j), b = div32(q, rows_pad, i);
    double v = 0.0;
    if (i < rows && j < cols) {
      const int sr = transpose ? j : i, sc = transpose ? i : j;
      if (!lower || sc <= sr) v = s * src[b * bstride + (long long)sr * ld + sc];
    }
    const __half hi = __float2half_rn((float)v);
    Ph[e] = hi;
    Pl[e] = __float2half_rn((float)(v - (double)__half2float(hi)));
  }
}
// mx != nullptr: the scale is derived from mx[0] here and published in scal2; else scal2 holds it already
static int pack_planes_f64(const double* src, int ld, long long bstride, int rows, int cols, int transpose, int lower, int batch,
                           int rows_pad, int cols_pad, const float* mx, float* scal2, void* Ph, void* Pl, cudaStream_t st) {
  pack_planes_f64_kernel<<<grid_for((long long)batch * rows_pad * cols_pad, 2048), 256, 0, st>>>(
      src, ld, bstride, rows, cols, transpose, lower, batch, rows_pad, cols_pad, mx, scal2, (__half*)Ph, (__half*)Pl);
  return check_launch("pack_planes_f64");
}

// fp32 -> split-fp16 planes: dst[b*rows_pad + i, j] = s * src[b*bstride + i*ld + j], zero padding elsewhere
// One warp per destination row (cols_pad is a multiple of 64): eight columns per lane and step, 16-byte plane stores; the
// (batch, row) split costs one division per row.
__global__ void __launch_bounds__(256) pack_planes_f32_kernel(const float* __restrict__ src, int ld, long long bstride, int rows,
                                                              int cols, int batch, int rows_pad, int cols_pad,
                                                              const float* __restrict__ mx, float* __restrict__ scal2,
                                                              __half* __restrict__ Ph, __half* __restrict__ Pl) {
  const float s = pack_scale(mx, scal2);
  const int lane = threadIdx.x & 31;
  const long long nrows = (long long)batch * rows_pad, nwarps = (long long)gridDim.x * 8;
  const bool vec = (((uintptr_t)src & 15) == 0) && (ld % 4 == 0) && (bstride % 4 == 0);
  for (long long q = blockIdx.x * 8LL + (threadIdx.x >> 5); q < nrows; q += nwarps) {
    const int i = (int)(q % rows_pad), b = (int)(q / rows_pad);
    const float* row = src + b * bstride + (long long)i * ld;
    for (int j0 = 8 * lane; j0 < cols_pad; j0 += 256) {
      float v[8];
      if (i < rows && vec && j0 + 8 <= cols) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(row + j0)), c = __ldg(reinterpret_cast<const float4*>(row + j0 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i < rows && j0 + u < cols) ? row[j0 + u] : 0.f;
      }
      __align__(16) __half hi[8];
      __align__(16) __half lo[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) split_f16(s * v[u], hi[u], lo[u]);
      *reinterpret_cast<uint4*>(Ph + q * cols_pad + j0) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(Pl + q * cols_pad + j0) = *reinterpret_cast<const uint4*>(lo);
    }
  }
}

// W planes: rows [0, (R+1)*Mp) = blocks (block 0 = Linv, block r = Wr[r-1]), then the mean rows (beta^T), then zero padding.
// Wr comes either as float64 [R, M, M] (Wr64) or as float32 [R*Mp, Mp] (Wr32, the tensor-core product).
__global__ void pack_w_f16_kernel(const double* __restrict__ Linv, int ldl, const double* __restrict__ Wr64,
                                  const float* __restrict__ Wr32, const double* __restrict__ beta, int M, int Mp, int R,
                                  long long rows_total, const float* __restrict__ mx, float* __restrict__ scal,
                                  __half* __restrict__ Wh, __half* __restrict__ Wl) {
  // scal[0..1] = scale pair of the W blocks (from mx[0]), scal[2..3] = of the mean rows (from mx[1])
  const double sw = (double)pack_scale(mx, scal), swm = (double)pack_scale(mx ? mx + 1 : nullptr, scal + 2);
  const long long total = rows_total * Mp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int j;
    const long long row = div32(e, Mp, j);
    double v = 0.0;
    if (j < M) {
      if (row < (long long)(R + 1) * Mp) {
        int i;
        const int blk = div32(row, Mp, i);
        if (i < M) {
          if (blk == 0) v = sw * Linv[(long long)i * ldl + j];
          else if (Wr64) v = sw * Wr64[((long long)(blk - 1) * M + i) * M + j];
          else v = sw * (double)Wr32[((long long)(blk - 1) * Mp + i) * Mp + j];
        }
      } else {
        const int r = (int)(row - (long long)(R + 1) * Mp);
        if (r < R) v = swm * beta[(long long)j * R + r];
      }
    }
    const __half hi = __float2half_rn((float)v);
    // the low part is taken from the float64 value so that hi + lo carries 22 bits of the original
    Wh[e] = hi;
    Wl[e] = __float2half_rn((float)(v - (double)__half2float(hi)));
  }
}

// SP planes [R*Mp + 256, Mp]: row (r-1)*Mp + i = 2 (S_r[i,:] - I[i,:]), S_r = C_r C_r^T float32 (symmetric; Kinv == nullptr) --
// or, in the Q-form, 2 (Q_r - Q_0) with Q_0 = Kinv (float64); the remaining rows are zero.  Scale from the bound
// 4 * max(|S_r|, 1); thread 0 publishes {scale, 1/scale}.  Also writes the planes of alpha [Mp, 64] (`beta` argument) for the
// mean tile of the da GEMM.
__global__ void pack_qp_f16_kernel(const double* __restrict__ Kinv, const float* __restrict__ Qr, const double* __restrict__ beta,
                                   int M, int Mp, int R, const float* __restrict__ mxq, float* __restrict__ scal2,
                                   __half* __restrict__ QPh, __half* __restrict__ QPl, float* __restrict__ beta32,
                                   float* __restrict__ bscal2, __half* __restrict__ BTh, __half* __restrict__ BTl) {
  // beta planes [Mp, 64] (the B operand of the mean-path tile of the dK GEMM), scaled from max|beta| = mxq[1]
  int exb = 0;
  if (mxq[1] > 0.f && isfinite(mxq[1])) frexpf(mxq[1], &exb);
  const double sb = (double)ldexpf(1.f, 14 - exb);
  if (blockIdx.x == 0 && threadIdx.x == 0) { bscal2[0] = (float)sb; bscal2[1] = (float)(1.0 / sb); }
  const float mmax = 4.f * (Kinv ? mxq[0] : fmaxf(mxq[0], 1.f));     // (the identity in S_r - I)
  int ex = 0;
  if (mmax > 0.f && isfinite(mmax)) frexpf(mmax, &ex);
  const double sc = (double)ldexpf(1.f, 14 - ex);
  if (blockIdx.x == 0 && threadIdx.x == 0) { scal2[0] = (float)sc; scal2[1] = (float)(1.0 / sc); }
  const long long rows = (long long)R * Mp + 256;
  const long long total = rows * Mp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int j;
    const long long row = div32(e, Mp, j);
    double v = 0.0;
    if (row < (long long)R * Mp) {
      int i;
      const int r = div32(row, Mp, i);
      if (i < M && j < M) v = 2.0 * sc * ((double)Qr[((long long)r * Mp + i) * Mp + j] - (Kinv ? Kinv[(long long)i * M + j] : (i == j ? 1.0 : 0.0)));
    }
    const __half hi = __float2half_rn((float)v);
    QPh[e] = hi;
    QPl[e] = __float2half_rn((float)(v - (double)__half2float(hi)));
    if (e < (long long)Mp * 64) {
      const int m = (int)(e / 64), rr = (int)(e % 64);
      const double bv = (m < M && rr < R) ? beta[(long long)m * R + rr] : 0.0;
      beta32[e] = (float)bv;
      const __half bh = __float2half_rn((float)(bv * sb));
      BTh[e] = bh;
      BTl[e] = __float2half_rn((float)(bv * sb - (double)__half2float(bh)));
    }
  }
}

// kblocked: the output planes are k-blocked [Mp/64][Tpad rows][64] (see tma_load_kb) instead of row-major [Tpad, Mp]
__global__ void split_rows_kernel(const float* __restrict__ Kt, long long T, int Mp, long long Tpad, const float* __restrict__ mx,
                                  float* __restrict__ kscal, __half* __restrict__ Kh, __half* __restrict__ Kl, int kblocked = 0) {
  // eight consecutive columns per thread (Mp is a multiple of 64): 2 x 16-byte loads, one 16-byte store per plane
  const float s = pack_scale(mx, kscal);
  const int g8 = Mp >> 3;
  const long long total = Tpad * g8;
  const bool vec = ((uintptr_t)Kt & 15) == 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long t = e / g8;
    const int c = (int)(e - t * g8) << 3;
    float v[8];
    if (t < T) {
      const float* src = Kt + t * Mp + c;
      if (vec) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[u];
      }
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = 0.f;
    }
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) split_f16(v[u] * s, hi[u], lo[u]);
    const long long o = kblocked ? ((long long)(c >> 6) * Tpad + t) * 64 + (c & 63) : t * Mp + c;
    *reinterpret_cast<uint4*>(Kh + o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(Kl + o) = *reinterpret_cast<const uint4*>(lo);
  }
}

// ============================================================================================ host plumbing
struct Carve2 {
  char* base; size_t off = 0;
  explicit Carve2(void* p) : base((char*)p) {}
  void* take(size_t bytes) {
    off = align_up(off, 1024);
    void* r = base ? base + off : nullptr;
    off += bytes;
    return r;
  }
};

// Generic batched C[b] = A[b] B[b]^T for float32 operands on the split-fp16 tensor-core GEMM (the R-batched M^3 products of
// the host's M-only chain rule): A [batch or 1, m, k], B [batch or 1, n, k] row-major, C [batch, m, n]; a batch stride of 0
// broadcasts the operand.  ~22-bit products, fp32 accumulation.
struct BgemmWork {
  void *Ah, *Al, *Bh, *Bl;
  float *scal, *mx;
  int m_pad, n_pad, k_pad;
  size_t bytes;
};
static BgemmWork carve_bgemm(int batch, int m, int n, int k, void* buf) {
  BgemmWork w;
  w.m_pad = (int)align_up(m, kBM); w.n_pad = (int)align_up(n, 64); w.k_pad = (int)align_up(k, kBK);
  Carve2 c(buf);
  w.Ah = c.take((size_t)batch * w.m_pad * w.k_pad * 2 + 256 * (size_t)w.k_pad * 2);   // + slack: a 128/256-row TMA box may overrun
  w.Al = c.take((size_t)batch * w.m_pad * w.k_pad * 2 + 256 * (size_t)w.k_pad * 2);
  w.Bh = c.take((size_t)batch * w.n_pad * w.k_pad * 2 + 256 * (size_t)w.k_pad * 2);
  w.Bl = c.take((size_t)batch * w.n_pad * w.k_pad * 2 + 256 * (size_t)w.k_pad * 2);
  w.scal = (float*)c.take(8 * 4);
  w.mx = (float*)c.take(4 * 4);
  w.bytes = align_up(c.off, 1024);
  return w;
}
size_t tc_bgemm_workspace_bytes(int batch, int m, int n, int k) { return carve_bgemm(batch, m, n, k, nullptr).bytes + 1024; }

int tc_bgemm_nt_ld(const float* A, int lda, long long a_bstride, const float* B, int ldb, long long b_bstride, float* C, int ldc,
                   long long c_bstride, int batch, int m, int n, int k, void* ws, cudaStream_t st) {
  BgemmWork w = carve_bgemm(batch, m, n, k, (void*)align_up((size_t)ws, 1024));
  const int ab = a_bstride ? batch : 1, bb = b_bstride ? batch : 1;
  int rc;
  cudaMemsetAsync(w.mx, 0, 4 * sizeof(float), st);
  MaxJob jobs[2];
  jobs[0].p = A; jobs[0].f64 = 0; jobs[0].rows = (long long)(ab - 1) * (a_bstride / (lda > 0 ? lda : 1)) + m; jobs[0].cols = k; jobs[0].ld = lda;
  jobs[0].lower_period = 0; jobs[0].out = w.mx + 0;
  jobs[1].p = B; jobs[1].f64 = 0; jobs[1].rows = (long long)(bb - 1) * (b_bstride / (ldb > 0 ? ldb : 1)) + n; jobs[1].cols = k; jobs[1].ld = ldb;
  jobs[1].lower_period = 0; jobs[1].out = w.mx + 1;
  if (ab > 1 && a_bstride % lda) { set_error("bgemm: batch stride of A must be a multiple of its leading dimension"); return DCGP_ERR_ARG; }
  if (bb > 1 && b_bstride % ldb) { set_error("bgemm: batch stride of B must be a multiple of its leading dimension"); return DCGP_ERR_ARG; }
  if ((rc = maxabs_jobs(jobs, 2, st))) return rc;
  pack_planes_f32_kernel<<<grid_for((long long)ab * w.m_pad * w.k_pad, 2048), 256, 0, st>>>(A, lda, a_bstride, m, k, ab, w.m_pad, w.k_pad,
                                                                                          w.mx + 0, w.scal + 0, (__half*)w.Ah, (__half*)w.Al);
  pack_planes_f32_kernel<<<grid_for((long long)bb * w.n_pad * w.k_pad, 2048), 256, 0, st>>>(B, ldb, b_bstride, n, k, bb, w.n_pad, w.k_pad,
                                                                                          w.mx + 1, w.scal + 2, (__half*)w.Bh, (__half*)w.Bl);
  if ((rc = check_launch("bgemm_pack", 2))) return rc;
  TcGemm g;
  memset(&g, 0, sizeof(g));
  g.Ah = w.Ah; g.Al = w.Al; g.a_rows_total = (long long)ab * w.m_pad; g.a_batch_rows = a_bstride ? w.m_pad : 0;
  g.Bh = w.Bh; g.Bl = w.Bl; g.b_rows_total = (long long)bb * w.n_pad; g.b_batch_rows = b_bstride ? w.n_pad : 0;
  g.batch = batch; g.m = m; g.n = n; g.m_pad = w.m_pad; g.n_pad = w.n_pad; g.k_pad = w.k_pad;
  g.a_scal = w.scal + 0; g.b_scal = w.scal + 2;
  g.C = C; g.c_batch_stride = c_bstride; g.ldc = ldc;
  return tc_gemm(g, st);
}

int tc_bgemm_nt(const float* A, const float* B, float* C, int batch, int m, int n, int k, long long a_bstride, long long b_bstride,
                void* ws, cudaStream_t st) {
  return tc_bgemm_nt_ld(A, k, a_bstride, B, k, b_bstride, C, n, (long long)m * n, batch, m, n, k, ws, st);
}


void tc_carve_prep(TcPrep& t, int M, int Mp, int R, int L, void* buf) {
  memset(&t, 0, sizeof(t));
  t.M = M; t.Mp = Mp; t.R = R; t.L = L; t.Lp = (int)align_up(L, 64); t.LpT = (int)align_up(L + 1, 64);
  Carve2 c(buf);
  const size_t wbytes = w_rows(Mp, R) * Mp * 2;
  t.Wh = c.take(wbytes);
  t.Wl = c.take(wbytes);
  t.Zh = c.take((size_t)Mp * t.Lp * 2);
  t.Zl = c.take((size_t)Mp * t.Lp * 2);
  t.zz = (float*)c.take((size_t)Mp * 4);
  t.scal = (float*)c.take(32 * 4);
  t.mx = (float*)c.take(8 * 4);
  t.QTh = c.take((size_t)R * Mp * Mp * 2);
  t.QTl = c.take((size_t)R * Mp * Mp * 2);
  t.Gh = c.take((size_t)Mp * Mp * 2);
  t.Gl = c.take((size_t)Mp * Mp * 2);
  t.Lph = c.take((size_t)Mp * Mp * 2);
  t.Lpl = c.take((size_t)Mp * Mp * 2);
  t.Wr32 = (float*)c.take((size_t)R * Mp * Mp * 4);
  t.BRh = c.take((size_t)R * Mp * Mp * 2);
  t.BRl = c.take((size_t)R * Mp * Mp * 2);
  t.Br32 = (float*)c.take((size_t)R * Mp * Mp * 4);
  t.Qr32 = (float*)c.take((size_t)R * Mp * Mp * 4);
  t.QBh = c.take(((size_t)R * Mp + 256) * Mp * 2);
  t.QBl = c.take(((size_t)R * Mp + 256) * Mp * 2);
  t.beta32 = (float*)c.take((size_t)Mp * 64 * 4);
  t.ZTh = c.take((size_t)t.LpT * Mp * 2);
  t.ZTl = c.take((size_t)t.LpT * Mp * 2);
  t.BTh = c.take((size_t)Mp * 64 * 2);
  t.BTl = c.take((size_t)Mp * 64 * 2);
  t.LTh = c.take(((size_t)Mp + 256) * Mp * 2);
  t.LTl = c.take(((size_t)Mp + 256) * Mp * 2);
  t.Wmh = t.Wml = nullptr;
  t.bytes = align_up(c.off, 1024);
}

// scal slots: 0 = W blocks, 1 = mean rows, 2 = q_sqrt^T planes, 3 = G planes, 4 = Lp^-1 planes  ({scale, 1/scale} each)
int tc_pack_operands(const TcPrep& t, const double* Linv, int ldl, const double* Wr, const double* beta, int M, int Mp, int R,
                     cudaStream_t st) {
  cudaMemsetAsync(t.mx, 0, 8 * sizeof(float), st);
  int rc;
  const MaxJob jobs[3] = {max_f64(Linv, M, M, ldl, 0, t.mx + 0), max_f64(Wr, (long long)R * M, M, M, 0, t.mx + 0),
                          max_f64(beta, M, R, R, 0, t.mx + 1)};
  if ((rc = maxabs_jobs(jobs, 3, st))) return rc;
  pack_w_f16_kernel<<<num_sms() * 8, 256, 0, st>>>(Linv, ldl, Wr, nullptr, beta, M, Mp, R, (long long)w_rows(Mp, R), t.mx, t.scal,
                                                   (__half*)t.Wh, (__half*)t.Wl);
  return check_launch("tc_pack_operands");
}

// Wr32[(r*Mp + j)*Mp + i] = L_r[i, j] (i >= j), 0 elsewhere: C_r^T for the whitened chained conditional
__global__ void qsqrt_t_f32_kernel(const double* __restrict__ q_sqrt, int M, int Mp, int R, float* __restrict__ out,
                                   const float* __restrict__ mx_in, float* __restrict__ mx_out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) atomic_max_nonneg(mx_out, mx_in[0]);   // max|C_r| = max|q_sqrt|
  const long long total = (long long)R * Mp * Mp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % Mp);
    const long long q = e / Mp;
    const int j = (int)(q % Mp), r = (int)(q / Mp);
    out[e] = (i < M && j < M && i >= j) ? (float)q_sqrt[((long long)r * M + i) * M + j] : 0.f;
  }
}

// out[(r*Mp + i)*Mp + k] = L_r[i, k] (i >= k), 0 elsewhere: C_r for the whitened parameterisation
__global__ void qsqrt_f32_kernel(const double* __restrict__ q_sqrt, int M, int Mp, int R, float* __restrict__ out,
                                 const float* __restrict__ mx_in, float* __restrict__ mx_out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) atomic_max_nonneg(mx_out, mx_in[0]);   // max|C_r| = max|q_sqrt|
  const long long total = (long long)R * Mp * Mp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % Mp);
    const long long q = e / Mp;
    const int i = (int)(q % Mp), r = (int)(q / Mp);
    out[e] = (i < M && k < M && i >= k) ? (float)q_sqrt[((long long)r * M + i) * M + k] : 0.f;
  }
}

// Tensor-core build of the R-batched M-only products (the O(R M^3) part of the step's minibatch-independent work):
//   W_r = L_r^T G            (G = Kuu^-1 symmetric, or Lm^-1 when whitened)                -> fp32, then the W planes
//   trace = sum_r |Lp^-1 L_r|_F^2   (GPflow gauss_kl / DS/layers.py:250)                  -> *trace_out (double)
// Linv/G/Lpinv/beta are float64 (Cholesky-quality); only these products run split-fp16 on tcgen05.
int tc_build_operands(const TcPrep& t, const double* Linv, int ldl, const double* G, int ldg, int g_is_linv,
                      const double* Lpinv, int ldp, const double* q_sqrt, const double* beta, double* trace_out,
                      const double* Kinv, int parts, int chained, const double* alpha, cudaStream_t st) {
  const int M = t.M, Mp = t.Mp, R = t.R;
  int rc;
  if (parts & 1) {
    // ---- part 1: what the forward conditional GEMM needs (W planes)
    cudaMemsetAsync(t.mx, 0, 8 * sizeof(float), st);
    const double* mvec = chained ? alpha : beta;
    {   // every max scan part 1 needs, in one launch (the products' own maxima come out of the GEMM epilogue)
      MaxJob jobs[4];
      int nj = 0;
      jobs[nj++] = max_f64(q_sqrt, (long long)R * M, M, M, M, t.mx + 2);
      if (!chained) jobs[nj++] = max_f64(G, M, M, ldg, 0, t.mx + 3);
      else if (!g_is_linv) jobs[nj++] = max_f64(Linv, M, M, ldl, 0, t.mx + 4);
      jobs[nj++] = max_f64(Linv, M, M, ldl, 0, t.mx + 0);
      jobs[nj++] = max_f64(mvec, M, R, R, 0, t.mx + 1);
      if ((rc = maxabs_jobs(jobs, nj, st))) return rc;
    }
    // QT[r*Mp + i, k] = L_r[k, i]  (transpose of the lower-triangular q_sqrt_r)
    if ((rc = pack_planes_f64(q_sqrt, M, (long long)M * M, M, M, 1, 1, R, Mp, Mp, t.mx + 2, t.scal + 4, t.QTh, t.QTl, st))) return rc;
    if (M != Mp) cudaMemsetAsync(t.Wr32, 0, (size_t)R * Mp * Mp * sizeof(float), st);
    if (chained && g_is_linv) {
      // whitened: C_r = L_r, so the stage-2 operand C_r^T is just the transposed q_sqrt (its maximum is that of q_sqrt)
      qsqrt_t_f32_kernel<<<num_sms() * 8, 256, 0, st>>>(q_sqrt, M, Mp, R, t.Wr32, t.mx + 2, t.mx + 0);
      if ((rc = check_launch("qsqrt_t"))) return rc;
    } else {
      TcGemm g;
      memset(&g, 0, sizeof(g));
      g.Ah = t.QTh; g.Al = t.QTl; g.a_rows_total = (long long)R * Mp; g.a_batch_rows = Mp;
      if (!chained) {
        // W_r = L_r^T G.  B operand: B[j, k] = G[k, j]  (G symmetric when it is Kuu^-1; the transpose of Lm^-1 when whitened)
        if ((rc = pack_planes_f64(G, ldg, 0, M, M, g_is_linv ? 1 : 0, g_is_linv ? 1 : 0, 1, Mp, Mp, t.mx + 3, t.scal + 6, t.Gh, t.Gl, st))) return rc;
        g.Bh = t.Gh; g.Bl = t.Gl; g.b_scal = t.scal + 6;
      } else {
        // C_r^T = L_r^T Lm^-T:  C[(r,j), i] = sum_k L_r[k, j] Lm^-1[i, k].  B operand: the rows of Lm^-1 (in the Lp planes,
        // which part 2 re-packs with the prior's Lp^-1 afterwards)
        if ((rc = pack_planes_f64(Linv, ldl, 0, M, M, 0, 1, 1, Mp, Mp, t.mx + 4, t.scal + 8, t.Lph, t.Lpl, st))) return rc;
        g.Bh = t.Lph; g.Bl = t.Lpl; g.b_scal = t.scal + 8;
      }
      g.b_rows_total = Mp; g.b_batch_rows = 0;
      g.batch = R; g.m = M; g.n = M; g.m_pad = Mp; g.n_pad = Mp; g.k_pad = Mp;
      g.a_scal = t.scal + 4;
      g.C = t.Wr32; g.c_batch_stride = (long long)Mp * Mp; g.ldc = Mp;
      g.absmax_out = t.mx + 0;                  // joins max|Lm^-1| in the W-block slot
      if ((rc = tc_gemm(g, st))) return rc;
    }
    // W planes for the conditional GEMM: block 0 = Lm^-1 (float64), blocks 1..R = W_r or C_r^T (fp32 product), mean rows =
    // beta^T or alpha^T
    pack_w_f16_kernel<<<num_sms() * 8, 256, 0, st>>>(Linv, ldl, nullptr, t.Wr32, mvec, M, Mp, R, (long long)w_rows(Mp, R), t.mx, t.scal,
                                                     (__half*)t.Wh, (__half*)t.Wl);
    if ((rc = check_launch("tc_build_operands"))) return rc;
  }
  if (!(parts & 2)) return DCGP_OK;
  // ---- part 2: KL trace and the backward operands (not needed by the forward conditional)
  cudaMemsetAsync(t.mx + 3, 0, 5 * sizeof(float), st);   // slots 3 (Lm^-1), 4 (prior's Lp^-1), 5 (S_r), 6 (alpha), 7 (C_r)
  {
    MaxJob jobs[3];
    int nj = 0;
    if (Lpinv) jobs[nj++] = max_f64(Lpinv, M, M, ldp, 0, t.mx + 4);
    if (Kinv) { jobs[nj++] = max_f64(Linv, M, M, ldl, 0, t.mx + 3); jobs[nj++] = max_f64(alpha, M, R, R, 0, t.mx + 6); }
    if (nj && (rc = maxabs_jobs(jobs, nj, st))) return rc;
  }
  if (Lpinv) {
    if ((rc = pack_planes_f64(Lpinv, ldp, 0, M, M, 0, 1, 1, Mp, Mp, t.mx + 4, t.scal + 8, t.Lph, t.Lpl, st))) return rc;
    cudaMemsetAsync(trace_out, 0, sizeof(double), st);
    TcGemm h;
    memset(&h, 0, sizeof(h));
    h.Ah = t.Lph; h.Al = t.Lpl; h.a_rows_total = Mp; h.a_batch_rows = 0;
    h.Bh = t.QTh; h.Bl = t.QTl; h.b_rows_total = (long long)R * Mp; h.b_batch_rows = Mp;
    h.batch = R; h.m = M; h.n = M; h.m_pad = Mp; h.n_pad = Mp; h.k_pad = Mp;
    h.a_scal = t.scal + 8; h.b_scal = t.scal + 4;
    h.sq_out = trace_out;
    if ((rc = tc_gemm(h, st))) return rc;
  }
  if (!Kinv) return DCGP_OK;
  // ---- backward operands, in the order of the forward (a = Lm^-1 k first):  C_r = Lm^-1 L_r (or L_r when whitened),
  //      S_r = C_r C_r^T, SP planes [R*Mp + 256, Mp] = 2 (S_r - I), alpha planes, Lm^-T planes.  Br32 keeps C_r for the host's
  //      M-only chain rule.
  {
    if ((rc = pack_planes_f64(Linv, ldl, 0, M, M, 0, 1, 1, Mp, Mp, t.mx + 3, t.scal + 6, t.Gh, t.Gl, st))) return rc;
    cudaMemsetAsync((char*)t.LTh + (size_t)Mp * Mp * 2, 0, (size_t)256 * Mp * 2, st);
    cudaMemsetAsync((char*)t.LTl + (size_t)Mp * Mp * 2, 0, (size_t)256 * Mp * 2, st);
    if ((rc = pack_planes_f64(Linv, ldl, 0, M, M, 1, 1, 1, Mp, Mp, nullptr, t.scal + 6, t.LTh, t.LTl, st))) return rc;
    if (g_is_linv) {          // whitened: C_r = L_r (its maximum is that of q_sqrt, slot 2 of part 1)
      qsqrt_f32_kernel<<<num_sms() * 8, 256, 0, st>>>(q_sqrt, M, Mp, R, t.Br32, t.mx + 2, t.mx + 7);
      if ((rc = check_launch("qsqrt_f32"))) return rc;
    } else {
      TcGemm b1;
      memset(&b1, 0, sizeof(b1));
      b1.Ah = t.Gh; b1.Al = t.Gl; b1.a_rows_total = Mp; b1.a_batch_rows = 0;
      b1.Bh = t.QTh; b1.Bl = t.QTl; b1.b_rows_total = (long long)R * Mp; b1.b_batch_rows = Mp;
      b1.batch = R; b1.m = M; b1.n = M; b1.m_pad = Mp; b1.n_pad = Mp; b1.k_pad = Mp;
      b1.a_scal = t.scal + 6; b1.b_scal = t.scal + 4;
      b1.C = t.Br32; b1.c_batch_stride = (long long)Mp * Mp; b1.ldc = Mp;
      b1.absmax_out = t.mx + 7;
      if (M != Mp) cudaMemsetAsync(t.Br32, 0, (size_t)R * Mp * Mp * sizeof(float), st);
      if ((rc = tc_gemm(b1, st))) return rc;
    }
    split_rows_kernel<<<num_sms() * 8, 256, 0, st>>>(t.Br32, (long long)R * Mp, Mp, (long long)R * Mp, t.mx + 7, t.scal + 14,
                                                     (__half*)t.BRh, (__half*)t.BRl);
    TcGemm b2;
    memset(&b2, 0, sizeof(b2));
    b2.Ah = t.BRh; b2.Al = t.BRl; b2.a_rows_total = (long long)R * Mp; b2.a_batch_rows = Mp;
    b2.Bh = t.BRh; b2.Bl = t.BRl; b2.b_rows_total = (long long)R * Mp; b2.b_batch_rows = Mp;
    b2.batch = R; b2.m = M; b2.n = M; b2.m_pad = Mp; b2.n_pad = Mp; b2.k_pad = Mp;
    b2.a_scal = t.scal + 14; b2.b_scal = t.scal + 14;
    b2.C = t.Qr32; b2.c_batch_stride = (long long)Mp * Mp; b2.ldc = Mp;     // S_r; Br32 keeps C_r for the M-only chain rule
    b2.absmax_out = t.mx + 5;
    if (M != Mp) cudaMemsetAsync(t.Qr32, 0, (size_t)R * Mp * Mp * sizeof(float), st);
    if ((rc = tc_gemm(b2, st))) return rc;
    pack_qp_f16_kernel<<<num_sms() * 8, 256, 0, st>>>(nullptr, t.Qr32, alpha, M, Mp, R, t.mx + 5, t.scal + 10, (__half*)t.QBh, (__half*)t.QBl,
                                                      t.beta32, t.scal + 16, (__half*)t.BTh, (__half*)t.BTl);
    if ((rc = check_launch("tc_build_backward_operands", 2))) return rc;
  }
  return DCGP_OK;
}

int tc_gemm_splits(const TcGemm& g) {
  const int nkb = g.k_pad / kBK;
  const int sp = g.splits > 1 ? g.splits : 1;
  const int per = ceil_div(nkb, sp);
  return ceil_div(nkb, per);
}

// ============================================================================================ K-A: Kuf on tensor cores
// Fused im2col + squared distance + RBF (layers.py:23-32, kernels.py:117-123):
//   D[t, m] = sum_l xs[t, l] * zs[m, l]        (xs = patch / lengthscale gathered straight from the NHWC image, zs = Z / ls)
//   k[t, m] = variance * exp(-0.5 * (|xs_t|^2 + |zs_m|^2 - 2 D[t, m]))        (GPflow's expansion form of the RBF)
// Roles: 4 gather warps (im2col rows -> split fp16 -> SWIZZLE_128B smem, + the TMA loads of the Z planes), 1 MMA warp,
// 4 epilogue warps (TMEM -> exp -> split-fp16 planes of K, the A operand of the conditional GEMM).  Patches never touch HBM.
constexpr float kXScale = 256.f;             // fixed power-of-two pre-scale of xs and zs before the fp16 split
constexpr int kKufGatherWarps = 8, kKufEpiWarps = 8;
constexpr int kKufThreads = 32 * (kKufGatherWarps + 1 + kKufEpiWarps);   // warps 0-7 gather, warp 8 MMA, warps 9-16 epilogue

struct KufParams {
  const float* X;      // [n_rows, HWC]
  View v;
  int T;               // n_rows * P
  int M, Mp;
  int njt;             // Mp / BN
  int n_items;         // t-tiles * njt
  int nkb;             // ceil(L / 64)
  float inv_ls, variance;
  const float* zz;     // [Mp] |zs_m|^2
  const float* kscal;  // {scale, 1/scale} of the output planes
  __half* Kh; __half* Kl;   // [Tpad, Mp]
};

template <int BN, bool PAIR>   // PAIR: patch elements 2i, 2i+1 are adjacent, 8-byte aligned floats (even C): float2 gather
__global__ void __launch_bounds__(kKufThreads, 1)
kuf_tc_kernel(const __grid_constant__ CUtensorMap tmZ_hi, const __grid_constant__ CUtensorMap tmZ_lo, KufParams p) {
  using Cfg = CondCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_smem = (uint32_t*)(tmem_empty + 2);
  float* xx_s = (float*)(tmem_base_smem + 4);     // [4 tiles][128] exponent offsets of the tile's patches (ring over tiles)
  int* off_s = (int*)(xx_s + 4 * 2 * kBM);        // [nkb*64] image offset of patch element l (im2col index math, once per CTA)
  float* zc_s = (float*)(off_s + p.nkb * kBK);    // [Mp] per-column exponent offset log2(ks) - 0.5 log2(e) |zs_m|^2 (-inf: padding)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.v.L;
  // k = ks * exp(-0.5 (xx + zz - 2 D)) = exp2(c2 * acc + zc[m] + xc[t]) with acc = kXScale^2 * D from the tensor core:
  // one FFMA + one FADD + one MUFU.EX2 per element, and the padding (m >= M, t >= T) falls out as exp2(-inf) = 0.
  constexpr float kLog2e = 1.4426950408889634f;
  const float c2 = kLog2e / (kXScale * kXScale);

  for (int l = threadIdx.x; l < p.nkb * kBK; l += kKufThreads) off_s[l] = (l < L) ? p.v.elem_off(l) : 0;
  {
    const float lks = log2f(p.kscal[0] * p.variance);
    for (int m = threadIdx.x; m < p.Mp; m += kKufThreads)
      zc_s[m] = (m < p.M) ? fmaf(-0.5f * kLog2e, p.zz[m], lks) : -INFINITY;
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmZ_hi); tma_prefetch_desc(&tmZ_lo);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], kKufGatherWarps + 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 32 * kKufEpiWarps); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == kKufGatherWarps) tmem_alloc(tmem_base_smem, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp < kKufGatherWarps) {
    // ------------------------------------------------------------------ gather: one WARP per patch row and k-block
    // Lane i owns elements 2i, 2i+1 of the row's 64-element k-block slice, so a warp-level load walks the (dx, c)-contiguous
    // runs of the NHWC image (coalesced), and its 32 x 4-byte shared-memory stores fill exactly one swizzled 128-byte row.
    const int gw = warp;                              // 0..7: rows gw*16 .. gw*16+15 of the tile
    const float sc = p.inv_ls * kXScale;              // patches are staged as xs * kXScale
    int stage = 0; uint32_t phase = 0; uint32_t tile = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++tile) {
      const int tt = item / p.njt, jt = item - tt * p.njt;
      // per-row image base (lanes 0..15 compute the warp's 16 rows, broadcast by shuffle below); -1 marks rows beyond T
      int my_base = -1;
      if (lane < 16) {
        const int t = tt * kBM + gw * 16 + lane;
        if (t < p.T) {
          const int n = t / p.v.P, pp = t - n * p.v.P;
          my_base = n * p.v.HWC + p.v.patch_base(pp);     // < 2^31 (checked by the host)
        }
      }
      float xxp[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) xxp[i] = 0.f;
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int l = kb * kBK + 2 * lane;
        const int o0 = off_s[l], o1 = off_s[l + 1];
        const bool in0 = l < L, in1 = l + 1 < L;
        float x0[16], x1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {                // all 32 loads of the k-block in flight before the slot wait
          const int base = __shfl_sync(0xffffffffu, my_base, i);
          const float* src = p.X + (base < 0 ? 0 : base);
          const float msk = base < 0 ? 0.f : sc;
          if (PAIR) {                                 // one coalesced 256-byte warp load per row and k-block
            const float2 v = in0 ? __ldg(reinterpret_cast<const float2*>(src + o0)) : make_float2(0.f, 0.f);
            x0[i] = v.x * msk;
            x1[i] = v.y * msk;
          } else {
            x0[i] = in0 ? __ldg(src + o0) * msk : 0.f;
            x1[i] = in1 ? __ldg(src + o1) * msk : 0.f;
          }
        }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        if (threadIdx.x == 0) {
          mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageB);
          tma_load_2d(st + 2 * Cfg::kStageA, &tmZ_hi, &full_bar[stage], kb * kBK, jt * BN);
          tma_load_2d(st + 2 * Cfg::kStageA + Cfg::kStageB, &tmZ_lo, &full_bar[stage], kb * kBK, jt * BN);
        }
        // element pair `lane` of row r: 16-byte chunk lane/4 (XOR-swizzled with r mod 8), 4-byte slot lane%4 inside it
        uint8_t* row0 = st + gw * 2048 + ((lane & 3) << 2);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          xxp[i] = fmaf(x0[i], x0[i], fmaf(x1[i], x1[i], xxp[i]));
          const __half2 hi = __floats2half2_rn(x0[i], x1[i]);
          const float2 hf = __half22float2(hi);
          const __half2 lo = __floats2half2_rn(x0[i] - hf.x, x1[i] - hf.y);
          uint8_t* dst = row0 + (i >> 3) * 1024 + (i & 7) * 128 + (((lane >> 2) ^ (i & 7)) << 4);
          *reinterpret_cast<__half2*>(dst) = hi;
          *reinterpret_cast<__half2*>(dst + Cfg::kStageA) = lo;
        }
        if (kb == p.nkb - 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float v = xxp[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            // exponent offset of the row: -0.5 log2(e) |xs_t|^2 (xxp is in kXScale^2 units); rows beyond T give exp2(-inf) = 0
            const int base = __shfl_sync(0xffffffffu, my_base, i);
            if (lane == 0) xx_s[(tile & 3) * kBM + gw * 16 + i] = base < 0 ? -INFINITY : -0.5f * c2 * v;
          }
        }
        fence_proxy_async();                        // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kKufGatherWarps) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BN);
      int stage = 0; uint32_t phase = 0; uint32_t tile = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++tile) {
        const uint32_t buf = tile & 1, use = tile >> 1;
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t a_lo = a_hi + Cfg::kStageA;
          const uint32_t b_hi = a_hi + 2 * Cfg::kStageA;
          const uint32_t b_lo = b_hi + Cfg::kStageB;
          const uint64_t dah = make_sw128_desc(a_hi), dal = make_sw128_desc(a_lo);
          const uint64_t dbh = make_sw128_desc(b_hi), dbl = make_sw128_desc(b_lo);
          int ksteps = (L - kb * kBK + 15) >> 4;
          ksteps = ksteps > 4 ? 4 : ksteps;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t koff = (uint64_t)((k * 16 * 2) >> 4);
            umma_f16(d_tmem, dal + koff, dbh + koff, idesc, (kb | k) != 0);
            umma_f16(d_tmem, dah + koff, dbl + koff, idesc, 1);
            umma_f16(d_tmem, dah + koff, dbh + koff, idesc, 1);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: exp2 + split; 2 warps per lane quarter
    const int ew = warp - kKufGatherWarps - 1;      // 0..7
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int chalf = ew >> 2;                       // which half of the BN columns
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    constexpr int CW = BN / 2, NCH = CW / 32;        // columns per thread, 32-column chunks
    uint32_t tile = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++tile) {
      const int tt = item / p.njt, jt = item - tt * p.njt;
      const int row = q * 32 + lane;
      const long long t = (long long)tt * kBM + row;
      const uint32_t buf = tile & 1, use = tile >> 1;
      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      const float xc = xx_s[(tile & 3) * kBM + row];
      const uint32_t taddr = tmem_base + lane_base + buf * BN + chalf * CW;
      const int m0 = jt * BN + chalf * CW;
      const bool odd = lane & 1;                     // (row pairs: the even lane's row t and t + 1)
      __half* oh = p.Kh + (t - (odd ? 1 : 0)) * p.Mp + m0 + (odd ? 16 : 0);
      __half* ol = p.Kl + (t - (odd ? 1 : 0)) * p.Mp + m0 + (odd ? 16 : 0);
      uint32_t v[2][32];                             // two chunks in flight: the next TMEM read overlaps this chunk's math
      tmem_ld_32x32_nowait(taddr, v[0]);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        tmem_ld_wait();
        if (ch + 1 < NCH) tmem_ld_32x32_nowait(taddr + (ch + 1) * 32, v[(ch + 1) & 1]);
        const uint32_t* vc = v[ch & 1];
        __align__(16) __half2 hi[16];
        __align__(16) __half2 lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 zc2 = *reinterpret_cast<const float2*>(zc_s + m0 + ch * 32 + 2 * i);
          const float k0 = exp2f(fmaf(c2, __uint_as_float(vc[2 * i]), zc2.x + xc));
          const float k1 = exp2f(fmaf(c2, __uint_as_float(vc[2 * i + 1]), zc2.y + xc));
          hi[i] = __floats2half2_rn(k0, k1);
          const float2 hf = __half22float2(hi[i]);
          lo[i] = __floats2half2_rn(k0 - hf.x, k1 - hf.y);
        }
        store_planes_paired(oh + ch * 32, ol + ch * 32, p.Mp, hi, lo, odd);
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kKufGatherWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Z / lengthscale (x kXScale) as split-fp16 planes [Mp, Lp] (Lp = ceil(L/64)*64, zero padded) and |zs_m|^2 in fp32.
__global__ void pack_z_f16_kernel(const double* __restrict__ Z, int M, int Mp, int L, int Lp, double inv_ls,
                                  __half* __restrict__ Zh, __half* __restrict__ Zl, float* __restrict__ zz,
                                  const double* __restrict__ hyp) {
  if (hyp) inv_ls = 1.0 / hyp[1];
  const int m = blockIdx.x;
  double acc = 0.0;
  for (int l = threadIdx.x; l < Lp; l += blockDim.x) {
    double v = 0.0;
    if (m < M && l < L) v = Z[(long long)m * L + l] * inv_ls;
    acc += v * v;
    const double vs = v * (double)kXScale;
    const __half hi = __float2half_rn((float)vs);
    Zh[(long long)m * Lp + l] = hi;
    Zl[(long long)m * Lp + l] = __float2half_rn((float)(vs - (double)__half2float(hi)));
  }
  __shared__ double sh[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) zz[m] = (float)(sh[0] + sh[1] + sh[2] + sh[3]);
}

__global__ void pack_zt_bf16_kernel(const double* __restrict__ Z, int M, int Mp, int L, int Lp, double inv_ls,
                                    __nv_bfloat16* __restrict__ ZTh, __nv_bfloat16* __restrict__ ZTl,
                                    const double* __restrict__ hyp);   // dcgp_tc_bwd.inc

int tc_pack_z(const TcPrep& t, const double* Z, int M, int L, double inv_ls, cudaStream_t st, const double* hyp) {
  pack_z_f16_kernel<<<t.Mp, 128, 0, st>>>(Z, M, t.Mp, L, t.Lp, inv_ls, (__half*)t.Zh, (__half*)t.Zl, t.zz, hyp);
  pack_zt_bf16_kernel<<<grid_for((long long)t.LpT * t.Mp, 2048), 256, 0, st>>>(Z, M, t.Mp, L, t.LpT, inv_ls, (__nv_bfloat16*)t.ZTh,
                                                                               (__nv_bfloat16*)t.ZTl, hyp);
  set_pair_kernel<<<1, 1, 0, st>>>(kXScale, t.scal + 12);
  return check_launch("pack_z_f16", 3);
}

template <int BN>
static int launch_kuf_tc(const TcPrep& prep, const View& v, const float* X, int n_rows, float variance, float inv_ls,
                         const float* kscal, void* Kh, void* Kl, cudaStream_t st) {
  using Cfg = CondCfg<BN>;
  CUtensorMap tmZh, tmZl;
  int rc;
  if ((rc = make_tmap_f16(&tmZh, prep.Zh, prep.Mp, prep.Lp, BN))) return rc;
  if ((rc = make_tmap_f16(&tmZl, prep.Zl, prep.Mp, prep.Lp, BN))) return rc;
  KufParams p;
  p.X = X; p.v = v; p.T = n_rows * v.P; p.M = prep.M; p.Mp = prep.Mp; p.njt = prep.Mp / BN;
  p.n_items = ceil_div(p.T, kBM) * p.njt;
  p.nkb = ceil_div(v.L, kBK);
  p.inv_ls = inv_ls; p.variance = variance; p.zz = prep.zz; p.kscal = kscal; p.Kh = (__half*)Kh; p.Kl = (__half*)Kl;
  const bool pair = (v.C % 2 == 0) && (((uintptr_t)X & 7) == 0);
  if ((((uintptr_t)Kh | (uintptr_t)Kl) & 31) != 0) { set_error("kuf_tc: K planes must be 32-byte aligned"); return DCGP_ERR_ARG; }
  const int smem_bytes = Cfg::kSmemBytes + 8 * kBM * 4 + ceil_div(v.L, kBK) * kBK * 4 + prep.Mp * 4 + 64;
  if (smem_bytes > 227 * 1024) { set_error("kuf_tc: patch length %d too large", v.L); return DCGP_ERR_ARG; }
  static int attr_bytes[2] = {0, 0};
  if (attr_bytes[pair] < smem_bytes) {
    attr_bytes[pair] = smem_bytes;
    cudaError_t e = pair ? cudaFuncSetAttribute(kuf_tc_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)
                         : cudaFuncSetAttribute(kuf_tc_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { set_error("kuf_tc smem attr: %s", cudaGetErrorString(e)); return DCGP_ERR_CUDA; }
  }
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  ScopedTimer timer(1, st);
  if (pair) kuf_tc_kernel<BN, true><<<grid, kKufThreads, smem_bytes, st>>>(tmZh, tmZl, p);
  else kuf_tc_kernel<BN, false><<<grid, kKufThreads, smem_bytes, st>>>(tmZh, tmZl, p);
  return check_launch("kuf_tc");
}

// Planes [Tpad, Mp] of variance * exp(...) * kscal[0] for all n_rows * P patches.
int tc_kuf(const TcPrep& prep, const View& v, const float* X, int n_rows, float variance, float inv_ls, const float* kscal,
           void* Kh, void* Kl, cudaStream_t st) {
  if (prep.Mp % 256 == 0) return launch_kuf_tc<256>(prep, v, X, n_rows, variance, inv_ls, kscal, Kh, Kl, st);
  if (prep.Mp % 128 == 0) return launch_kuf_tc<128>(prep, v, X, n_rows, variance, inv_ls, kscal, Kh, Kl, st);
  return launch_kuf_tc<64>(prep, v, X, n_rows, variance, inv_ls, kscal, Kh, Kl, st);
}

// kernels.py:127-133 on the planes: Kzx[n, m] = (1/P) sum_p w_p K[(n*P+p), m], fp32 (unscaled) [n_rows, Mp]
__global__ void patch_mean_planes_kernel(const __half* __restrict__ Kh, const __half* __restrict__ Kl, int P, int Mp,
                                         const double* __restrict__ w, const float* __restrict__ kscal, float* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (m >= Mp) return;
  float acc = 0.f;
  const long long base = (long long)n * P * Mp + m;
  for (int pp = 0; pp < P; ++pp) {
    const float k = __half2float(Kh[base + (long long)pp * Mp]) + __half2float(Kl[base + (long long)pp * Mp]);
    acc = fmaf(w ? (float)w[pp] : 1.f, k, acc);
  }
  out[(long long)n * Mp + m] = acc * kscal[1] / (float)P;
}

void tc_carve_cond(TcCondWork& w, int M, int Mp, int R, size_t T, void* buf) {
  memset(&w, 0, sizeof(w));
  w.Tpad = align_up(T, kBM);
  Carve2 c(buf);
  w.Kh = c.take(w.Tpad * Mp * 2);
  w.Kl = c.take(w.Tpad * Mp * 2);
  w.kscal = (float*)c.take(8 * 4);
  void* prep_buf = c.take(0);
  tc_carve_prep(w.prep, M, Mp, R, 1, prep_buf);
  c.off += w.prep.bytes;
  w.bytes = align_up(c.off, 1024);
}

int tc_split_rows(const float* Kt, int T, int Mp, const TcCondWork& w, cudaStream_t st) {
  cudaMemsetAsync(w.kscal + 4, 0, sizeof(float), st);
  maxabs_f32_kernel<<<grid_for((long long)T * Mp, 4096), 256, 0, st>>>(Kt, (long long)T * Mp, w.kscal + 4);
  split_rows_kernel<<<num_sms() * 8, 256, 0, st>>>(Kt, T, Mp, (long long)w.Tpad, w.kscal + 4, w.kscal, (__half*)w.Kh, (__half*)w.Kl);
  return check_launch("tc_split_rows", 2);
}

void tc_carve_apply(TcApplyWork& a, int kind, int M, int Mp, int R, int L, size_t Tk, size_t T, void* buf) {
  memset(&a, 0, sizeof(a));
  Carve2 c(buf);
  // conv: the Kuf kernel writes the planes of the [Tk, Mp] kernel matrix the conditional GEMM consumes.
  // svgp: the same planes are averaged over patches into Kzx [T, Mp] (fp32), which is then split into its own planes.
  a.kk.Tpad = align_up(Tk, kBM);
  a.kk.Kh = c.take(a.kk.Tpad * Mp * 2);
  a.kk.Kl = c.take(a.kk.Tpad * Mp * 2);
  a.kk.kscal = (float*)c.take(8 * 4);
  if (kind == DCGP_LAYER_CONV) {   // planes of a = Lm^-1 k for the chained conditional
    a.kk.Ah = c.take(a.kk.Tpad * Mp * 2);
    a.kk.Al = c.take(a.kk.Tpad * Mp * 2);
    a.kk.ascal = (float*)c.take(8 * 4);
  }
  if (kind != DCGP_LAYER_CONV) {
    a.kz.Tpad = align_up(T, kBM);
    a.kz.Kh = c.take(a.kz.Tpad * Mp * 2);
    a.kz.Kl = c.take(a.kz.Tpad * Mp * 2);
    a.kz.kscal = (float*)c.take(8 * 4);
    a.kz.Ah = c.take(a.kz.Tpad * Mp * 2);          // planes of a = Lm^-1 kzx (image-level rows)
    a.kz.Al = c.take(a.kz.Tpad * Mp * 2);
    a.kz.ascal = (float*)c.take(8 * 4);
  }
  a.bytes = align_up(c.off, 1024);
}

int launch_kuf_simt_planes(const float* X, const View& v, int n_rows, const float* zs, int M, float variance, float inv_ls,
                           int ldo, const float* kscal, void* Kh, void* Kl, long long Tpad, cudaStream_t st);

int tc_layer_apply(const dcgp_layer_desc* d, const View& v, const TcPrep& prep, const TcApplyWork& a, const float* zs,
                   float* Kt32, const double* patch_weights, const float* X, int n_rows, float* Kzx, float* acc, float* mean_t,
                   cudaStream_t st) {
  const int Mp = prep.Mp, R = prep.R;
  const float variance = (float)d->variance, inv_ls = (float)(1.0 / d->lengthscale);
  (void)zs; (void)Kt32;
  set_scale_kernel<<<1, 1, 0, st>>>(variance, a.kk.kscal);      // RBF values lie in (0, variance]
  check_launch("set_scale");
  int rc = tc_kuf(prep, v, X, n_rows, variance, inv_ls, a.kk.kscal, a.kk.Kh, a.kk.Kl, st);   // layers.py:112 / kernels.py:123
  if (rc) return rc;
  if (d->kind == DCGP_LAYER_CONV) {
    if (tc_forward_chained()) return tc_cond_chained(prep, a.kk, n_rows * v.P, Mp, R, variance, acc, mean_t, st);
    return tc_cond(prep, a.kk, n_rows * v.P, Mp, R, acc, mean_t, st);
  }
  patch_mean_planes_kernel<<<dim3(ceil_div(Mp, 128), n_rows), 128, 0, st>>>((const __half*)a.kk.Kh, (const __half*)a.kk.Kl, v.P, Mp,
                                                                          patch_weights, a.kk.kscal, Kzx);
  if ((rc = check_launch("patch_mean_planes"))) return rc;
  if ((rc = tc_split_rows(Kzx, n_rows, Mp, a.kz, st))) return rc;
  // |a|^2 = kzx^T Kuu^-1 kzx <= Kdiag <= variance max|w|^2
  if (tc_forward_chained()) return tc_cond_chained(prep, a.kz, n_rows, Mp, R, variance, acc, mean_t, st, patch_weights, v.P);
  return tc_cond(prep, a.kz, n_rows, Mp, R, acc, mean_t, st);
}

#include "dcgp_tc_bwd.inc"

}  // namespace dcgp
