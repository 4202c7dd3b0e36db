// gemm_tcgen05.cu — the dense feature transform of the RGCN layer on the
// 5th-generation tensor cores:  C[M,N] = [A1 | A2][M,K] · B[N,K]^T + bias[N]
//
//   forward : out = [H | x] · [W_live ; root]   (one GEMM instead of upstream's R+1 matmuls + adds,
//             [PyG] RGCNConv.forward called main.py:272,285,298,308); the two A sources are two TMA
//             descriptors, so H and x are never concatenated in memory.  Transform-first layers:
//             Z = x · [W_0 | .. | W_{S-1} | root] with the bias on the root slab.
//   backward: [dH | dx_root] = g · [W_live ; root]^T, written through two output maps so that dH lands
//             contiguous for the transposed aggregation;  dx = dZ · [W | root]^T.
//
// bf16 or fp16 operands (fp16 = what torch.amp.autocast feeds the reference's matmuls, main.py:446,543), fp32
// accumulation in TMEM, bf16 or fp32 output.  Both operands are K-major.  Persistent, warp-specialised kernel
// (one CTA per SM, 6 warps): see gemm_nt_kernel.  The shapes of this path are tall and skinny (M = #nodes,
// N = 64..1280, K = 64..1280): the kernel is bound by streaming A / writing C, which the TMA ring and the
// double-buffered accumulator keep overlapped; B (<= 160 KB) lives in L2.
#include <cuda.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace gmlm {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 320;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue
constexpr int kEpiWarps = 8;
constexpr uint32_t kSpinLimit = 1u << 28;   // a lost barrier traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B apart (SBO);
// LBO is unused for swizzled K-major layouts (canonical value 1); version 1 = sm_100.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);   // start address, 16-byte units
  d |= uint64_t(1) << 16;                       // leading byte offset (unused)
  d |= uint64_t(1024 >> 4) << 32;               // stride byte offset
  d |= uint64_t(1) << 46;                       // descriptor version
  d |= uint64_t(2) << 61;                       // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = F32 (bit 4), A / B element format in bits [7,10) / [10,13) (F16 = 0,
// BF16 = 1), both operands K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int m, int n, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
constexpr int kMaxSources = 4;
struct GemmParams {
  int M, N;
  int N1;                 // output columns [0,N1) go to C1, [N1,N) to C2 (any split; N1 = N: one output)
  int tiles1;             // ceil(N1 / BLOCK_N): tiles [0,tiles1) cover C1, the rest C2 (B rows N1 + ...)
  int n_tiles;            // tiles1 + ceil((N - N1) / BLOCK_N)
  int n_src;              // A = [A_0 | .. | A_{n_src-1}]: one tensor map per source, never concatenated in memory
  int kb_end[kMaxSources];  // cumulative 64-column block counts of the sources (a source's last block may be partial:
  int k_off[kMaxSources];   // TMA zero-fills it); k_off = column of B where the source starts (cumulative TRUE widths)
  int n_route;            // > 0: the N columns go to n_route outputs cut at route_end[] (multiples of 64 columns): every
  int route_end[kMaxSources];  // 64- / 32-column epilogue chunk is routed to its output's tensor map (tiles may straddle)
  const float* bias;      // [N] or nullptr
  const void* addend;     // [M, N] of the output type or nullptr: C = A.B^T + bias + addend (residual folded in)
  int64_t ld_add;
  uint32_t idesc_formats; // a/b element format bits of the instruction descriptor (F16 = 0, BF16 = 1)
};
struct SourceMaps {
  CUtensorMap m[kMaxSources];
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// 256-wide tiles: FOUR 48 KB stages (three left the tensor pipe 65 % active on the K = 960 fusion product: two
// k-blocks of prefetch do not cover the L2 / HBM latency); the shared memory for the fourth stage comes from
// single-buffering the epilogue's staging rows, which have slack (4 us of MMAs per tile against ~2 us of epilogue)
// fp32 operands (3xTF32, see gemm_nt_kernel): a stage also holds the low-part copies of both tiles, tiles <= 128 wide
constexpr int stages_for(int block_n, bool x3 = false) {
  if (x3) return block_n >= 128 ? 3 : (block_n >= 64 ? 3 : 4);
  return block_n >= 256 ? 4 : (block_n >= 128 ? 4 : (block_n >= 64 ? 6 : 7));
}
constexpr int staging_bufs_for(int block_n, bool x3 = false) { return (block_n >= 256 || (x3 && block_n >= 128)) ? 1 : 2; }
constexpr uint32_t tmem_cols_for(int block_n) {
  return 2 * block_n <= 32 ? 32u : (2 * block_n <= 64 ? 64u : (2 * block_n <= 128 ? 128u : (2 * block_n <= 256 ? 256u : 512u)));
}
constexpr int staging_bytes_for(int block_n, bool x3 = false) {   // 8 epilogue warps x 1-2 buffers x (32 rows x 128 B)
  return kEpiWarps * staging_bufs_for(block_n, x3) * 32 * 128;
}
// output columns per staging row: a full 128-byte swizzle row when the tile width allows it, else 64 bytes
template <int BLOCK_N, typename OutT>
constexpr int chunk_cols() {
  return BLOCK_N % (128 / int(sizeof(OutT))) == 0 ? 128 / int(sizeof(OutT)) : 64 / int(sizeof(OutT));
}
template <int BLOCK_N, bool X3 = false>
constexpr size_t smem_bytes_for() {
  // no alignment slack: the dynamic shared memory is declared __align__(1024) (checked at kernel entry); one
  // operand row of a stage is 128 bytes whatever the element type (64 bf16 / fp16 or 32 fp32)
  return size_t(stages_for(BLOCK_N, X3)) * (X3 ? 2 : 1) * (BLOCK_M * 128 + BLOCK_N * 128) +
         staging_bytes_for(BLOCK_N, X3) + 2 * BLOCK_N * sizeof(float) + (3 * stages_for(BLOCK_N, X3) + 4) * 8 + 16;
}

// PERSISTENT kernel: every CTA (one per SM) walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of the
// 128 x BLOCK_N output grid (N tiles of one M block adjacent in t, so the CTAs that run together share their A
// tile through L2).  Three pipelines run concurrently inside a CTA:
//   TMA ring (warp 0)       full/empty mbarriers over STAGES operand stages, k-block counter running across tiles
//   accumulators (warp 1)   TWO TMEM buffers of BLOCK_N fp32 columns: the MMAs of tile i+1 run while the epilogue
//                           drains tile i (tmem_full / tmem_empty mbarriers)
//   epilogue (warps 2-9)    tcgen05.ld -> + bias -> convert -> swizzled staging rows in shared memory -> TMA store
//                           (cp.async.bulk.tensor, clips the M tail), double-buffered per warp; two warps per TMEM
//                           lane quarter split the column chunks (the epilogue, not the MMA, paces K <= 128 shapes)
// The round-1 kernel ran one tile per CTA with one accumulator: on the backward shape (K = 64: ONE k-block per
// tile) load latency, MMA and a 64 KB epilogue were serialised per tile (3.3 ms against 0.88 ms for cuBLAS).
//
// X3 = fp32 operands on the tf32 tensor cores with fp32-grade accuracy ("3xTF32"): TMA lands the fp32 tiles, four
// CONVERTER warps (10-13) split every element in place into hi = the tf32-representable top 19 bits and write
// lo = x - hi (exact) to a second copy of the stage at the same swizzled offsets, and the MMA warp issues
// lo.hi + hi.lo + hi.hi per k-step (the dropped lo.lo term and the tf32 rounding of lo are 2^-22 relative): the
// reference's fp32 eval mode (main.py:603) meets the 1e-5 gate that plain TF32 (2^-11) cannot.
template <int BLOCK_N, typename OutT, bool X3 = false>
__global__ void __launch_bounds__(X3 ? kThreads + 128 : kThreads, 1) gemm_nt_kernel(const __grid_constant__ SourceMaps tma_a,
                                                              const __grid_constant__ CUtensorMap tma_b,
                                                              const __grid_constant__ SourceMaps tma_c,
                                                              const GemmParams p) {
  constexpr int STAGES = stages_for(BLOCK_N, X3);
  constexpr int BK = X3 ? 32 : BLOCK_K;              // elements per k-block: one 128-byte swizzle row
  constexpr uint32_t A_BYTES = BLOCK_M * 128;
  constexpr uint32_t B_BYTES = BLOCK_N * 128;
  // X3: ONE tile in flight, split over all of tensor memory: k-block kb adds its hi.hi product to accumulator
  // kb mod X3_MAIN (3 for 128-wide tiles, 7 for narrower ones), the small lo.hi + hi.lo terms go to one more.  The
  // tensor core adds into an fp32 accumulator with truncation, a bias that grows linearly with the number of
  // accumulations (3 MMAs x K/32 into ONE accumulator measured 1.2e-5 at K = 1500, 2e-6 split over four); the
  // epilogue adds the accumulators with round-to-nearest
  constexpr int X3_MAIN = BLOCK_N >= 128 ? 3 : 7;
  constexpr uint32_t TMEM_COLS = X3 ? uint32_t((X3_MAIN + 1) * BLOCK_N) : tmem_cols_for(BLOCK_N);
  static_assert(!X3 || (TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0), "accumulators exceed tensor memory");
  constexpr int CHUNK = chunk_cols<BLOCK_N, OutT>();      // output columns per staging row (128 or 64 bytes)
  constexpr int ROWB = CHUNK * int(sizeof(OutT));
  constexpr int N_CHUNKS = BLOCK_N / CHUNK;
  static_assert(BLOCK_N % CHUNK == 0 && (ROWB == 128 || ROWB == 64) && (CHUNK == 32 || CHUNK == 64),
                "tile width must be a multiple of one staging row");
  constexpr int STG_BUFS = staging_bufs_for(BLOCK_N, X3);
  constexpr int kStagingBytes = staging_bytes_for(BLOCK_N, X3);
  extern __shared__ __align__(1024) uint8_t smem_nt[];
  // 1024-byte alignment for the 128B-swizzled tiles
  uint8_t* smem = smem_nt;
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + STAGES * A_BYTES;
  uint8_t* smem_alo = smem_b + STAGES * B_BYTES;                      // X3 only: low parts of the A / B tiles
  uint8_t* smem_blo = smem_alo + (X3 ? STAGES * A_BYTES : 0);
  uint8_t* staging = smem_blo + (X3 ? STAGES * B_BYTES : 0);          // 1024-aligned: A_BYTES, B_BYTES are
  float* bias_s = reinterpret_cast<float*>(staging + kStagingBytes);  // [2][BLOCK_N]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 2 * BLOCK_N);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* conv = bars + 2 * STAGES;                                 // X3 only: stage split into hi / lo
  uint64_t* tmem_full = bars + 3 * STAGES;
  uint64_t* tmem_empty = bars + 3 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.kb_end[p.n_src - 1];
  const int n_tiles = p.n_tiles;
  const int total_tiles = ((p.M + BLOCK_M - 1) / BLOCK_M) * n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(empty + s), 1);
      mbar_init(smem_u32(conv + s), 4);                   // one arrival per converter warp
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(tmem_full + b), 1);
      mbar_init(smem_u32(tmem_empty + b), kEpiWarps);     // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      uint32_t kc = 0;                                   // k-blocks issued so far, across tiles
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles) * BLOCK_M;
        const int tn = t % n_tiles;
        const int brow0 = tn < p.tiles1 ? tn * BLOCK_N : p.N1 + (tn - p.tiles1) * BLOCK_N;   // row of B = output column
        int src = 0, kb_first = 0;                       // current A source and its first k-block
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          while (kb >= p.kb_end[src]) kb_first = p.kb_end[src++];
          const int s = kc % STAGES;
          const uint32_t ph = (kc / STAGES) & 1;
          mbar_wait(smem_u32(empty + s), ph ^ 1);
          mbar_expect_tx(smem_u32(full + s), A_BYTES + B_BYTES);
          const int ka = (kb - kb_first) * BK;           // column inside the source; columns past its width and rows
          tma_load_2d(smem_u32(smem_a + s * A_BYTES), &tma_a.m[src], smem_u32(full + s), ka, m0);   // past M: zeros
          tma_load_2d(smem_u32(smem_b + s * B_BYTES), &tma_b, smem_u32(full + s), p.k_off[src] + ka, brow0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer (single thread)
      const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, X3 ? 2u : p.idesc_formats);   // 2 = TF32
      uint32_t kc = 0, it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const uint32_t buf = X3 ? 0u : (it & 1), aph = X3 ? (it & 1) : ((it >> 1) & 1);
        mbar_wait(smem_u32(tmem_empty + buf), aph ^ 1);   // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const int s = kc % STAGES;
          const uint32_t ph = (kc / STAGES) & 1;
          mbar_wait(smem_u32(X3 ? conv + s : full + s), ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + s * A_BYTES));
          const uint64_t db = make_smem_desc(smem_u32(smem_b + s * B_BYTES));
          if constexpr (X3) {
            const uint64_t dal = make_smem_desc(smem_u32(smem_alo + s * A_BYTES));
            const uint64_t dbl = make_smem_desc(smem_u32(smem_blo + s * B_BYTES));
            const uint32_t d_main = tmem_base + uint32_t(kb % X3_MAIN) * BLOCK_N;
            const uint32_t d_small = tmem_base + uint32_t(X3_MAIN) * BLOCK_N;
#pragma unroll
            for (int k = 0; k < 4; ++k) {                  // 8 fp32 per UMMA = 32 bytes = +2 in the address field
              umma_tf32(d_small, dal + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) ? 1u : 0u);
              umma_tf32(d_small, da + uint64_t(2 * k), dbl + uint64_t(2 * k), idesc, 1u);
              umma_tf32(d_main, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb >= X3_MAIN || k) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advancing 16 elements along K inside the swizzle atom = +32 bytes = +2 in the address field
              umma(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(empty + s));      // frees the smem stage once these MMAs have read it
        }
        umma_commit(smem_u32(tmem_full + buf));  // accumulator complete
      }
    }
  } else if (X3 && warp >= 2 + kEpiWarps) {
    // ---------------- converter (fp32 operands): hi = x & 0xffffe000 in place, lo = x - hi into the stage's second
    // copy, same byte offsets (the swizzle is a permutation of 16-byte chunks: element-wise work is layout-blind)
    const int ct = int(threadIdx.x) - 32 * (2 + kEpiWarps);     // 0..127
    uint32_t kc = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb, ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(smem_u32(full + s), ph);
        auto split = [&](uint8_t* hi, uint8_t* lo, int n16) {
#pragma unroll 4
          for (int i = ct; i < n16; i += 128) {
            uint4 v = *reinterpret_cast<uint4*>(hi + 16 * i);
            uint4 h = make_uint4(v.x & 0xffffe000u, v.y & 0xffffe000u, v.z & 0xffffe000u, v.w & 0xffffe000u);
            uint4 l = make_uint4(__float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)),
                                 __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)),
                                 __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)),
                                 __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)));
            *reinterpret_cast<uint4*>(hi + 16 * i) = h;
            *reinterpret_cast<uint4*>(lo + 16 * i) = l;
          }
        };
        split(smem_a + s * A_BYTES, smem_alo + s * A_BYTES, int(A_BYTES / 16));
        split(smem_b + s * B_BYTES, smem_blo + s * B_BYTES, int(B_BYTES / 16));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(conv + s));
      }
    }
  } else {
    // ---------------- epilogue: warp w may read TMEM lanes [32*(w%4), +32) = 32 output rows of the tile; the two
    // warps that share a lane quarter take the even / the odd column chunks.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = int(threadIdx.x) - 64;                  // 0..255 among the epilogue threads
    uint8_t* stg = staging + (warp - 2) * (STG_BUFS * 32 * 128);
    constexpr int LAST0 = ((N_CHUNKS - 1) / 2) * 2;        // last chunk of the even warp
    constexpr int LAST1 = N_CHUNKS >= 2 ? ((N_CHUNKS - 2) / 2) * 2 + 1 : -1;
    const int my_last = half == 0 ? LAST0 : LAST1;
    uint32_t it = 0, sp = 0;                               // tile counter, staging-buffer toggle
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int m0 = (t / n_tiles) * BLOCK_M;
      const int tn = t % n_tiles;
      const bool second = tn >= p.tiles1;
      const int ccol0 = second ? (tn - p.tiles1) * BLOCK_N : tn * BLOCK_N;   // first column inside C1 / C2
      const int n0 = second ? p.N1 + ccol0 : ccol0;                          // ... and of the whole product
      const int n_end = second ? p.N : p.N1;                                 // columns at or past it are clipped
      const uint32_t buf = X3 ? 0u : (it & 1), aph = X3 ? (it & 1) : ((it >> 1) & 1);
      float* bs = bias_s + (it & 1) * BLOCK_N;
      if (p.bias) {
        for (int c = et; c < BLOCK_N; c += 32 * kEpiWarps) bs[c] = n0 + c < n_end ? __ldg(p.bias + n0 + c) : 0.f;
      }
      // every epilogue warp has finished tile it-1 here, hence every read of bias_s[buf] from tile it-2
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(smem_u32(tmem_full + buf), aph);
      tc_fence_after();
      if (my_last < 0) {                                   // nothing to read (one-chunk tiles): stay in step
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(tmem_empty + buf));
        continue;
      }
      const CUtensorMap* cmap = second ? &tma_c.m[1] : &tma_c.m[0];
      const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16) + buf * BLOCK_N;
#pragma unroll 1
      for (int ch = half; ch < N_CHUNKS; ch += 2, sp ^= (STG_BUFS - 1)) {
        uint8_t* sb = stg + sp * (32 * 128);
        // the TMA store last issued from this buffer (two chunks ago; the previous chunk when the tile shape leaves
        // room for one buffer only) must have finished READING it
        if (lane == 0) {
          if constexpr (STG_BUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        float v[CHUNK];
        if constexpr (X3) {                                // CHUNK == 32 (fp32 output): sum the accumulators in use
          static_assert(!X3 || CHUNK == 32, "fp32 operands give fp32 output");
          uint32_t r0[32], r1[32];
          tmem_ld32_nowait(t_lane + uint32_t(X3_MAIN * BLOCK_N + ch * CHUNK), r0);     // small terms first
          tmem_ld32_nowait(t_lane + uint32_t(ch * CHUNK), r1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
          for (int a = 1; a < X3_MAIN && a < num_kb; ++a) {
            tmem_ld32_nowait(t_lane + uint32_t(a * BLOCK_N + ch * CHUNK), r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r1[j]);
          }
        } else {
          uint32_t r0[32];
          tmem_ld32_nowait(t_lane + uint32_t(ch * CHUNK), r0);
          if constexpr (CHUNK == 64) {
            uint32_t r1[32];
            tmem_ld32_nowait(t_lane + uint32_t(ch * CHUNK + 32), r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
          } else {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
          }
        }
        if (ch == my_last) {                               // my share of the accumulator is read: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(tmem_empty + buf));
        }
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(bs + ch * CHUNK);
#pragma unroll
          for (int j = 0; j < CHUNK / 4; ++j) {
            const float4 bb = b4[j];                       // same address in every lane: a broadcast
            v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
          }
        }
        if (p.addend) {
          // residual folded into the epilogue: a lane owns one output row and reads its CHUNK columns of the
          // addend as whole 16-byte pieces (every 32-byte sector it touches is fully used)
          const int row = m0 + quad * 32 + lane;
          constexpr int EPA = 16 / int(sizeof(OutT));
          const OutT* arow = static_cast<const OutT*>(p.addend) + int64_t(row) * p.ld_add + n0 + ch * CHUNK;
#pragma unroll
          for (int j = 0; j < CHUNK / EPA; ++j) {
            if (row < p.M && n0 + ch * CHUNK + (j + 1) * EPA <= n_end) {
              const uint4 q = __ldg(reinterpret_cast<const uint4*>(arow) + j);
              if constexpr (sizeof(OutT) == 2) {
                Pack<__nv_bfloat16, 8> a;
                a.v = q;
                float f[8];
                a.unpack(f);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[EPA * j + e] += f[e];
              } else {
                v[4 * j] += __uint_as_float(q.x); v[4 * j + 1] += __uint_as_float(q.y);
                v[4 * j + 2] += __uint_as_float(q.z); v[4 * j + 3] += __uint_as_float(q.w);
              }
            }
          }
        }
        // one staging row per lane; its 16-byte chunk j goes to position j ^ (row & 7) (128-byte rows) or
        // j ^ ((row >> 1) & 3) (64-byte rows): the SWIZZLE_128B / SWIZZLE_64B pattern of the output tensor map
        // (byte-address bits [4,7) ^= bits [7,10), resp. [4,6) ^= [7,9)), conflict-free for per-lane-row writes
        uint8_t* my_row = sb + lane * ROWB;
        constexpr int EPC = 16 / int(sizeof(OutT));        // elements per 16-byte chunk
#pragma unroll
        for (int j = 0; j < ROWB / 16; ++j) {
          uint4 val;
          if constexpr (sizeof(OutT) == 2) {
            Pack<__nv_bfloat16, 8> o;
            o.pack(v + EPC * j);
            val = o.v;
          } else {
            val = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                             __float_as_uint(v[4 * j + 3]));
          }
          const int pos = ROWB == 128 ? (j ^ (lane & 7)) : (j ^ ((lane >> 1) & 3));
          *reinterpret_cast<uint4*>(my_row + (pos << 4)) = val;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA
        __syncwarp();
        if (lane == 0) {
          if (p.n_route > 0) {                             // route this chunk to the output that owns its columns
            const int col = n0 + ch * CHUNK;
            int o = 0;
            while (o + 1 < p.n_route && col >= p.route_end[o]) ++o;
            tma_store_2d(&tma_c.m[o], smem_u32(sb), col - (o ? p.route_end[o - 1] : 0), m0 + quad * 32);
          } else {
            tma_store_2d(cmap, smem_u32(sb), ccol0 + ch * CHUNK, m0 + quad * 32);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all rows have left shared memory
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// Tensor maps are pure functions of (address, shape, pitch, box, element type): they are cached, so that a
// training loop that calls the same layer with the allocator handing back the same buffers pays for the
// driver's encode once (round 1 encoded three maps on every call).
struct MapKey {
  const void* base;
  int64_t rows, cols, ld;
  int box_rows, box_cols, dtype, dev, swizzle;   // swizzle: 128 / 64 (bytes) or 0 = none
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           box_cols == o.box_cols && dtype == o.dtype && dev == o.dev && swizzle == o.swizzle;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.base);
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(uint64_t(k.rows)); mix(uint64_t(k.cols)); mix(uint64_t(k.ld));
    mix(uint64_t(k.box_rows) << 32 | uint32_t(k.box_cols)); mix(uint64_t(k.dtype) << 8 | uint32_t(k.dev));
    mix(uint64_t(k.swizzle));
    return size_t(h);
  }
};
std::mutex g_map_mutex;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

// row-major [rows, cols] matrix with leading dimension ld (elements); box = box_cols x box_rows; box_cols * element
// size is 128 bytes (SWIZZLE_128B) or 64 bytes (SWIZZLE_64B)
int get_map(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols,
            int dtype, int swizzle = -1) {
  {
    const int e = dtype == GMLM_F32 ? 4 : 2;
    if (swizzle < 0) swizzle = box_cols * e == 128 ? 128 : 64;
  }
  const MapKey key{base, rows, cols, ld, box_rows, box_cols, dtype, current_device(), swizzle};
  {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) { *out = it->second; return GMLM_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(GMLM_ERR_CUDA, "gemm: cuTensorMapEncodeTiled is not available from the driver");
  const int esz = dtype == GMLM_F32 ? 4 : 2;
  const CUtensorMapDataType dt = dtype == GMLM_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == GMLM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld) * esz};
  cuuint32_t box[2] = {cuuint32_t(box_cols), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GMLM_ERR_CUDA, "gemm: cuTensorMapEncodeTiled failed with code %d", int(r));
  std::lock_guard<std::mutex> lock(g_map_mutex);
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache.emplace(key, *out);
  return GMLM_OK;
}

template <int BLOCK_N, typename OutT, bool X3 = false>
int launch(const SourceMaps& a, const CUtensorMap& b, const SourceMaps& c, GemmParams p, cudaStream_t st) {
  constexpr size_t smem = smem_bytes_for<BLOCK_N, X3>();
  static_assert(smem <= 227 * 1024, "tile configuration exceeds the shared memory of one SM");
  auto kern = gemm_nt_kernel<BLOCK_N, OutT, X3>;
  static bool configured[kMaxDevices] = {};          // per device: the opt-in is a per-device function attribute
  const int dev = current_device();
  if (!configured[dev]) {
    GMLM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    configured[dev] = true;
  }
  p.tiles1 = (p.N1 + BLOCK_N - 1) / BLOCK_N;
  p.n_tiles = p.tiles1 + (p.N - p.N1 + BLOCK_N - 1) / BLOCK_N;
  const int64_t tiles = int64_t((p.M + BLOCK_M - 1) / BLOCK_M) * p.n_tiles;
  const unsigned grid = unsigned(std::min<int64_t>(tiles, num_sms()));     // persistent: one CTA per SM
  kern<<<grid, X3 ? kThreads + 128 : kThreads, smem, st>>>(a, b, c, p);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

// tile width: the widest tile whose padding (columns computed past N1 / N and clipped by the store) stays under
// 1/8 of the product; exact divisors first (160 for N = 320, 640: two passes over A through L2 instead of five
// with 64-wide tiles, which made that shape L2-bandwidth-bound)
int pick_block_n(int64_t n1, int64_t n2, bool x3) {
  const int cand[5] = {256, 160, 128, 64, 32};
  const int first = x3 ? 2 : 0;                      // fp32 operands: tiles up to 128 wide (a stage holds hi and lo)
  auto padded = [&](int bn) { return (n1 + bn - 1) / bn * bn + (n2 + bn - 1) / bn * bn; };
  // 256-wide tiles when they pad the product by at most 8 % (N = 960: 1024 computed, 1.19 against 1.07 PFLOP/s for
  // six exact 160-wide tiles: the wider tile reads less shared memory per flop)
  if (!x3 && padded(256) * 100 <= (n1 + n2) * 108) return 256;
  for (int i = first; i < 5; ++i) {
    const int bn = cand[i];
    if (n1 % bn == 0 && n2 % bn == 0 && (bn != 160 || (n1 + n2) % 256 != 0)) return bn;
  }
  for (int i = first; i < 5; ++i)
    if (padded(cand[i]) * 8 <= (n1 + n2) * 9) return cand[i];
  return 32;
}

// outputs: n_out = 1 or 2 with any column split (tiles never straddle: the second output starts a new tile row of B),
// or up to four outputs cut at multiples of 64 columns (tiles straddle, every epilogue chunk is routed)
int gemm_general(int n_src, const void* const* A, const int64_t* lda, const int64_t* Ks, const void* B, int64_t ldb,
                 const float* bias, const void* addend, int64_t ld_add, int64_t M, int n_out, void* const* Cs,
                 const int64_t* ldcs, const int64_t* Ns, int in_dtype, int out_dtype, void* stream) {
  GMLM_REQUIRE(n_out >= 1 && n_out <= kMaxSources && Cs && ldcs && Ns, "gemm: 1..%d outputs", kMaxSources);
  int64_t N = 0;
  for (int i = 0; i < n_out; ++i) {
    GMLM_REQUIRE(Cs[i] != nullptr && Ns[i] > 0, "gemm: output %d is empty", i);
    N += Ns[i];
  }
  const bool route = n_out > 2;
  if (route)
    for (int i = 0; i + 1 < n_out; ++i)
      GMLM_REQUIRE(Ns[i] % 64 == 0, "gemm: with more than two outputs every output but the last must be a multiple of 64 columns wide");
  void* C1 = Cs[0];
  int64_t ldc1 = ldcs[0];
  int64_t N1 = (n_out == 2) ? Ns[0] : N;
  void* C2 = n_out == 2 ? Cs[1] : nullptr;
  int64_t ldc2 = n_out == 2 ? ldcs[1] : 0;
  GMLM_REQUIRE(n_src >= 1 && n_src <= kMaxSources, "gemm: 1..%d A sources", kMaxSources);
  GMLM_REQUIRE(in_dtype == GMLM_BF16 || in_dtype == GMLM_F16 || in_dtype == GMLM_F32,
               "gemm: operands must be GMLM_BF16, GMLM_F16 or GMLM_F32");
  GMLM_REQUIRE(out_dtype == GMLM_F32 || out_dtype == GMLM_BF16, "gemm: out_dtype must be GMLM_F32 or GMLM_BF16");
  const bool x3 = in_dtype == GMLM_F32;               // fp32 operands: the 3xTF32 kernel, fp32 output
  GMLM_REQUIRE(!x3 || out_dtype == GMLM_F32, "gemm: fp32 operands give an fp32 output");
  const int esz = out_dtype == GMLM_F32 ? 4 : 2;
  const int per16 = x3 ? 4 : 8;                       // operand elements per 16 bytes
  const int bk = x3 ? 32 : BLOCK_K;                   // operand elements per 128-byte k-block
  int64_t K = 0;
  for (int i = 0; i < n_src; ++i) {
    GMLM_REQUIRE(A[i] != nullptr && Ks[i] > 0, "gemm: A source %d is empty", i);
    GMLM_REQUIRE(lda[i] >= Ks[i] && lda[i] % per16 == 0, "gemm: lda[%d] must be >= K with a 16-byte row pitch", i);
    GMLM_REQUIRE((reinterpret_cast<uintptr_t>(A[i]) & 15) == 0, "gemm: operands must be 16-byte aligned");
    // a source starts at column sum(K_0..K_{i-1}) of B: the TMA wants that inner coordinate on a 16-byte boundary
    GMLM_REQUIRE(i == n_src - 1 || Ks[i] % per16 == 0, "gemm: every source but the last must be a multiple of 16 bytes wide");
    K += Ks[i];
  }
  GMLM_REQUIRE(M >= 0 && N > 0, "gemm: bad sizes");
  GMLM_REQUIRE(M < (int64_t(1) << 31) && N < (int64_t(1) << 24) && K < (int64_t(1) << 30), "gemm: sizes exceed int32");
  GMLM_REQUIRE(B && C1, "gemm: null pointer");
  GMLM_REQUIRE(ldb >= K && ldb % per16 == 0, "gemm: ldb must be >= K with a 16-byte row pitch");
  GMLM_REQUIRE((reinterpret_cast<uintptr_t>(B) & 15) == 0, "gemm: operands must be 16-byte aligned");
  if (N1 <= 0 || N1 >= N) {
    N1 = N; C2 = C1; ldc2 = ldc1;
  } else {
    GMLM_REQUIRE(C2 != nullptr, "gemm: second output missing");
    GMLM_REQUIRE(addend == nullptr, "gemm: the addend goes with a single output");
  }
  if (!route)
    GMLM_REQUIRE((reinterpret_cast<uintptr_t>(C1) & 15) == 0 && (reinterpret_cast<uintptr_t>(C2) & 15) == 0 &&
                     (ldc1 * esz) % 16 == 0 && (ldc2 * esz) % 16 == 0 && ldc1 >= N1 && ldc2 >= N - N1,
                 "gemm: outputs must be 16-byte aligned with a 16-byte row pitch");
  for (int i = 0; route && i < n_out; ++i)
    GMLM_REQUIRE((reinterpret_cast<uintptr_t>(Cs[i]) & 15) == 0 && (ldcs[i] * esz) % 16 == 0 && ldcs[i] >= Ns[i],
                 "gemm: outputs must be 16-byte aligned with a 16-byte row pitch");
  GMLM_REQUIRE(!(route && addend), "gemm: the addend goes with a single output");
  if (addend)
    GMLM_REQUIRE((reinterpret_cast<uintptr_t>(addend) & 15) == 0 && (ld_add * esz) % 16 == 0 && (N * esz) % 16 == 0 &&
                     ld_add >= N,
                 "gemm: the addend must be 16-byte aligned with a 16-byte row pitch");
  if (M == 0) return GMLM_OK;
  const int bn = pick_block_n(N1, N - N1, x3);
  // cuTensorMapEncodeTiled is a driver entry point: it needs a current context on THIS thread (autograd worker
  // threads may not have touched the runtime of this library yet).  The caller's current device is the device of
  // the operands (the torch layer guards it); it is neither changed nor queried per call.
  {
    static thread_local bool ctx_ready[kMaxDevices] = {};
    const int dev = current_device();
    if (!ctx_ready[dev]) {
      GMLM_CUDA_TRY(cudaFree(nullptr));
      ctx_ready[dev] = true;
    }
  }
  GemmParams p{};
  SourceMaps ma;
  CUtensorMap mb;
  SourceMaps mc;
  int kb = 0, koff = 0;
  for (int i = 0; i < kMaxSources; ++i) {
    if (i < n_src) {
      if (int rc = get_map(&ma.m[i], A[i], M, Ks[i], lda[i], BLOCK_M, bk, in_dtype)) return rc;
      p.k_off[i] = koff;
      koff += int(Ks[i]);
      kb += int((Ks[i] + bk - 1) / bk);
    } else {
      ma.m[i] = ma.m[0];
      p.k_off[i] = koff;
    }
    p.kb_end[i] = kb;
  }
  int rc = get_map(&mb, B, N, K, ldb, bn, bk, in_dtype);
  if (rc) return rc;
  const int chunk = bn % (128 / esz) == 0 ? 128 / esz : 64 / esz;      // chunk_cols<BLOCK_N, OutT>()
  if (route) {
    int end = 0;
    for (int i = 0; i < kMaxSources; ++i) {
      if (i < n_out) {
        if (int rc2 = get_map(&mc.m[i], Cs[i], M, Ns[i], ldcs[i], 32, chunk, out_dtype)) return rc2;
        end += int(Ns[i]);
      } else {
        mc.m[i] = mc.m[0];
      }
      p.route_end[i] = end;
    }
    p.n_route = n_out;
  } else {
    rc = get_map(&mc.m[0], C1, M, N1, ldc1, 32, chunk, out_dtype);
    if (rc) return rc;
    rc = N1 < N ? get_map(&mc.m[1], C2, M, N - N1, ldc2, 32, chunk, out_dtype) : (mc.m[1] = mc.m[0], GMLM_OK);
    if (rc) return rc;
    mc.m[2] = mc.m[3] = mc.m[0];
    p.n_route = 0;
  }
  p.M = int(M); p.N = int(N); p.N1 = int(N1); p.n_src = n_src;
  p.bias = bias;
  p.addend = addend; p.ld_add = ld_add;
  p.idesc_formats = in_dtype == GMLM_BF16 ? 1u : 0u;
  cudaStream_t st = as_stream(stream);
  if (x3) {
    switch (bn) {
      case 128: return launch<128, float, true>(ma, mb, mc, p, st);
      case 64: return launch<64, float, true>(ma, mb, mc, p, st);
      case 32: return launch<32, float, true>(ma, mb, mc, p, st);
    }
    return fail(GMLM_ERR_INVALID, "gemm: unsupported tile");
  }
#define GMLM_GEMM_CASE(BN)                                                                     \
  case BN:                                                                                     \
    return out_dtype == GMLM_F32 ? launch<BN, float>(ma, mb, mc, p, st)                  \
                                 : launch<BN, __nv_bfloat16>(ma, mb, mc, p, st);
  switch (bn) {
    GMLM_GEMM_CASE(256)
    GMLM_GEMM_CASE(160)
    GMLM_GEMM_CASE(128)
    GMLM_GEMM_CASE(64)
    GMLM_GEMM_CASE(32)
  }
#undef GMLM_GEMM_CASE
  return fail(GMLM_ERR_INVALID, "gemm: unsupported tile");
}

// ================================================================== TN product: weight gradients
// D[Ka, N] (fp32) = [A_0 | A_1 | ..]^T[Ka, M] . G[M, N]: the reduction over all M nodes that gives dW = H^T g and
// droot = x^T g of the RGCN layer (autograd of [PyG] RGCNConv.forward's `h @ weight[r]` / `x @ root`, main.py:272;
// SURVEY §8a row A14 "dW_r = h_r^T g") and the weight gradients of the residual / fusion linears.
//
// Both operands are read as they lie in memory — row-major [M, .] activations — i.e. MN-major for the tensor core:
// TMA boxes of 64 columns x 64 rows (128-byte swizzle) ARE the canonical MN-major SWIZZLE_128B layout
// ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units: a shared-memory row is one node (k), eight rows form a
// 1024-byte swizzle atom (SBO), the next 64 channels live one box further (LBO = 8 KB); the instruction descriptor
// sets a_major = b_major = MN.  One UMMA covers 16 nodes = two atoms, so the k-advance inside a stage is +2 KB.
//
// Schedule: output tiles are 128 (channels of A) x BN (channels of G); the M range is cut into `splits` pieces so
// that tiles x splits fill the SMs (the layer shapes have 2..40 tiles but millions of rows).  CTA b works on tile
// b mod tiles of split b / tiles: the CTAs that run together walk the same rows, so A and G are read from HBM once
// and from L2 by the other tiles.  Each CTA accumulates its piece in TMEM and writes fp32 either straight to D
// (splits = 1) or to its slab of the workspace, which a second kernel sums in split order (deterministic).
constexpr int TN_BK = 64;                     // nodes per pipeline stage
constexpr int kTnThreads = 192;               // warp 0: TMA, warp 1: MMA, warps 2-5: epilogue (one per TMEM lane quarter)
constexpr uint32_t TN_BOX_BYTES = 64 * TN_BK * 2;   // one 64-channel x 64-node box

struct TnParams {
  int M, N;
  int n_src;
  int tile_end[kMaxSources];   // cumulative 128-channel tile counts of the A sources
  int k_off[kMaxSources];      // first row of D each source writes (cumulative true widths)
  int k_len[kMaxSources];      // its width
  int ka_tiles, n_tiles, splits, kb_per_split;
  float* out;                  // D (splits == 1) or the workspace [splits][Ka][ldo]
  int64_t ldo, slab;           // row pitch and slab stride (elements)
  uint32_t idesc_formats;
};

// MN-major SWIZZLE_128B operand: LBO = byte distance between 64-channel boxes, SBO = 1024 (8 nodes x 128 B)
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
  d |= uint64_t(TN_BOX_BYTES >> 4) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

template <int BN>
constexpr int tn_stages() { return BN >= 256 ? 4 : (BN >= 128 ? 6 : 8); }
template <int BN>
constexpr size_t tn_smem_bytes() {
  return size_t(tn_stages<BN>()) * (2 * TN_BOX_BYTES + (BN / 64) * TN_BOX_BYTES) + (2 * tn_stages<BN>() + 1) * 8 + 16 + 1024;
}

template <int BN>
__global__ void __launch_bounds__(kTnThreads, 1) gemm_tn_kernel(const __grid_constant__ SourceMaps tma_a,
                                                                const __grid_constant__ CUtensorMap tma_g,
                                                                const TnParams p) {
  constexpr int STAGES = tn_stages<BN>();
  constexpr uint32_t A_BYTES = 2 * TN_BOX_BYTES;
  constexpr uint32_t B_BYTES = (BN / 64) * TN_BOX_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32u : uint32_t(BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.ka_tiles * p.n_tiles;
  const int tile = blockIdx.x % tiles, split = blockIdx.x / tiles;
  const int kt = tile / p.n_tiles, nt = tile % p.n_tiles;
  int src = 0;
  while (kt >= p.tile_end[src]) ++src;
  const int ka0 = (kt - (src ? p.tile_end[src - 1] : 0)) * BLOCK_M;      // first channel inside the source
  const int n0 = nt * BN;
  const int total_kb = (p.M + TN_BK - 1) / TN_BK;
  const int kb_lo = split * p.kb_per_split;
  const int kb_hi = min(total_kb, kb_lo + p.kb_per_split);
  const int num_kb = max(0, kb_hi - kb_lo);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(empty + s), 1);
    }
    mbar_init(smem_u32(acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(smem_u32(empty + s), ph ^ 1);
        mbar_expect_tx(smem_u32(full + s), A_BYTES + B_BYTES);
        const int m0 = (kb_lo + i) * TN_BK;            // rows past M and channels past a source's width: zeros
        tma_load_2d(smem_u32(smem_a + s * A_BYTES), &tma_a.m[src], smem_u32(full + s), ka0, m0);
        tma_load_2d(smem_u32(smem_a + s * A_BYTES + TN_BOX_BYTES), &tma_a.m[src], smem_u32(full + s), ka0 + 64, m0);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_2d(smem_u32(smem_b + s * B_BYTES + j * TN_BOX_BYTES), &tma_g, smem_u32(full + s), n0 + 64 * j, m0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, BN, p.idesc_formats) | (1u << 15) | (1u << 16);   // MN-major A and B
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(smem_u32(full + s), ph);
        tc_fence_after();
        const uint64_t da = make_smem_desc_mn(smem_u32(smem_a + s * A_BYTES));
        const uint64_t db = make_smem_desc_mn(smem_u32(smem_b + s * B_BYTES));
#pragma unroll
        for (int k = 0; k < TN_BK / UMMA_K; ++k) {
          // 16 nodes further = two swizzle atoms = +2048 bytes = +128 in the address field
          umma(tmem_base, da + uint64_t(128 * k), db + uint64_t(128 * k), idesc, (i | k) ? 1u : 0u);
        }
        umma_commit(smem_u32(empty + s));
      }
      umma_commit(smem_u32(acc_full));
    }
  } else {
    // ---------------- epilogue: warp w reads TMEM lanes [32*(w%4), +32) = 32 rows (channels of A) of the tile
    const int quad = warp & 3;
    const int row_in_src = ka0 + quad * 32 + lane;
    const bool row_ok = row_in_src < p.k_len[src];
    float* dst = p.out + int64_t(split) * p.slab + int64_t(p.k_off[src] + row_in_src) * p.ldo + n0;
    if (num_kb > 0) {
      mbar_wait(smem_u32(acc_full), 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      if (num_kb > 0) {
        tmem_ld32_nowait(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(c0), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;          // an empty split still owns (zeroes) its slab
      }
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int col = n0 + c0 + j;
          if (col + 4 <= p.N) {
            *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (col + e < p.N) dst[c0 + j + e] = __uint_as_float(r[j + e]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// D = sum over splits (in order) of the workspace slabs
__global__ void gemm_tn_reduce_kernel(const float* __restrict__ ws, int splits, int64_t slab, int64_t rows, int64_t n4,
                                      int64_t ld_ws, float* __restrict__ out, int64_t ldo) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * n4) return;
  const int64_t r = i / n4, c = (i % n4) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(ws + s * slab + r * ld_ws + c));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(out + r * ldo + c) = acc;
}

struct TnPlan {
  int bn, ka_tiles, n_tiles, splits, kb_per_split;
  int64_t ka, ld_ws;
};
TnPlan tn_plan(int n_src, const int64_t* Ks, int64_t M, int64_t N) {
  TnPlan t{};
  t.bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  for (int i = 0; i < n_src; ++i) {
    t.ka_tiles += int((Ks[i] + BLOCK_M - 1) / BLOCK_M);
    t.ka += Ks[i];
  }
  t.n_tiles = int((N + t.bn - 1) / t.bn);
  const int total_kb = int((M + TN_BK - 1) / TN_BK);
  const int tiles = t.ka_tiles * t.n_tiles;
  // fill the SMs once: the workspace traffic (splits x Ka x N x 8 bytes) stays far below the operand traffic
  int splits = std::max(1, num_sms() / std::max(tiles, 1));
  splits = std::min(splits, std::max(1, total_kb / 8));       // at least 8 k-blocks (512 nodes) per split
  t.kb_per_split = (total_kb + splits - 1) / std::max(splits, 1);
  t.splits = std::max(1, (total_kb + std::max(t.kb_per_split, 1) - 1) / std::max(t.kb_per_split, 1));
  t.ld_ws = (N + 3) / 4 * 4;
  return t;
}

template <int BN>
int launch_tn(const SourceMaps& a, const CUtensorMap& g, const TnParams& p, cudaStream_t st) {
  constexpr size_t smem = tn_smem_bytes<BN>();
  static_assert(smem <= 227 * 1024, "tile configuration exceeds the shared memory of one SM");
  auto kern = gemm_tn_kernel<BN>;
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    GMLM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    configured[dev] = true;
  }
  kern<<<unsigned(p.ka_tiles * p.n_tiles * p.splits), kTnThreads, smem, st>>>(a, g, p);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

}  // namespace

// row-gather tensor map for the aggregation kernels (spmm.cu): [rows, cols] matrix, box = one whole row, no swizzle
// (the box shape `tile::gather4` takes: four arbitrary rows per bulk-tensor op; probed in tools/probes/)
int row_gather_map(void* out_map, const void* base, int64_t rows, int64_t cols, int64_t ld, int dtype) {
  static thread_local bool ctx_ready[kMaxDevices] = {};
  const int dev = current_device();
  if (!ctx_ready[dev]) {
    GMLM_CUDA_TRY(cudaFree(nullptr));
    ctx_ready[dev] = true;
  }
  return get_map(static_cast<CUtensorMap*>(out_map), base, rows, cols, ld, 1, int(cols), dtype, 0);
}
}  // namespace gmlm

using namespace gmlm;

extern "C" int gmlm_gemm_nt_multi(int num_sources, const void* const* A_host, const int64_t* lda_host,
                                  const int64_t* K_host, const void* B, int64_t ldb, const float* bias,
                                  const void* addend, int64_t ld_add, int64_t M, int num_outputs, void* const* C_host,
                                  const int64_t* ldc_host, const int64_t* N_host, int in_dtype, int out_dtype,
                                  void* stream) {
  GMLM_REQUIRE(A_host && lda_host && K_host, "gemm: null source table");
  return gemm_general(num_sources, A_host, lda_host, K_host, B, ldb, bias, addend, ld_add, M, num_outputs, C_host,
                      ldc_host, N_host, in_dtype, out_dtype, stream);
}

extern "C" int gmlm_gemm_nt(const void* A1, int64_t lda1, int64_t K1, const void* A2, int64_t lda2, int64_t K2,
                            const void* B, int64_t ldb, const float* bias, int64_t M, int64_t N, void* C1,
                            int64_t ldc1, int64_t N1, void* C2, int64_t ldc2, int in_dtype, int out_dtype,
                            void* stream) {
  GMLM_REQUIRE(K1 > 0 && K2 >= 0 && A1 && (K2 == 0 || A2) && N > 0, "gemm: bad sizes");
  const void* A[2] = {A1, A2};
  const int64_t lda[2] = {lda1, lda2}, Ks[2] = {K1, K2};
  const bool two = N1 > 0 && N1 < N;
  GMLM_REQUIRE(!two || C2 != nullptr, "gemm: second output missing");
  void* Cs[2] = {C1, C2};
  const int64_t ldcs[2] = {ldc1, ldc2}, Ns[2] = {two ? N1 : N, N - N1};
  return gemm_general(K2 > 0 ? 2 : 1, A, lda, Ks, B, ldb, bias, nullptr, 0, M, two ? 2 : 1, Cs, ldcs, Ns, in_dtype,
                      out_dtype, stream);
}

extern "C" int gmlm_gemm_nt_bf16(const void* A1, int64_t lda1, int64_t K1, const void* A2, int64_t lda2, int64_t K2,
                                 const void* B, int64_t ldb, const float* bias, int64_t M, int64_t N, void* C1,
                                 int64_t ldc1, int64_t N1, void* C2, int64_t ldc2, int out_dtype, void* stream) {
  return gmlm_gemm_nt(A1, lda1, K1, A2, lda2, K2, B, ldb, bias, M, N, C1, ldc1, N1, C2, ldc2, GMLM_BF16, out_dtype,
                      stream);
}

extern "C" size_t gmlm_gemm_tn_workspace_bytes(int num_sources, const int64_t* K_host, int64_t M, int64_t N) {
  if (!K_host || num_sources < 1 || num_sources > kMaxSources || M <= 0 || N <= 0) return 256;
  const TnPlan t = tn_plan(num_sources, K_host, M, N);
  return t.splits > 1 ? size_t(t.splits) * size_t(t.ka) * size_t(t.ld_ws) * sizeof(float) + 256 : 256;
}

extern "C" int gmlm_gemm_tn(int num_sources, const void* const* A_host, const int64_t* lda_host, const int64_t* K_host,
                            const void* G, int64_t ldg, int64_t M, int64_t N, float* D, int64_t ldd, int in_dtype,
                            void* ws, size_t ws_bytes, void* stream) {
  GMLM_REQUIRE(A_host && lda_host && K_host, "gemm_tn: null source table");
  GMLM_REQUIRE(num_sources >= 1 && num_sources <= kMaxSources, "gemm_tn: 1..%d A sources", kMaxSources);
  GMLM_REQUIRE(in_dtype == GMLM_BF16 || in_dtype == GMLM_F16, "gemm_tn: operands must be GMLM_BF16 or GMLM_F16");
  GMLM_REQUIRE(M >= 0 && N > 0 && M < (int64_t(1) << 31) && N < (int64_t(1) << 24), "gemm_tn: bad sizes");
  GMLM_REQUIRE(G && D && ldg >= N && ldg % 8 == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0,
               "gemm_tn: G must be 16-byte aligned with a 16-byte row pitch");
  GMLM_REQUIRE(ldd >= N && ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0,
               "gemm_tn: D must be 16-byte aligned with a 16-byte row pitch");
  int64_t ka = 0;
  for (int i = 0; i < num_sources; ++i) {
    GMLM_REQUIRE(A_host[i] != nullptr && K_host[i] > 0, "gemm_tn: A source %d is empty", i);
    GMLM_REQUIRE(lda_host[i] >= K_host[i] && lda_host[i] % 8 == 0 && (reinterpret_cast<uintptr_t>(A_host[i]) & 15) == 0,
                 "gemm_tn: A sources must be 16-byte aligned with a 16-byte row pitch");
    ka += K_host[i];
  }
  GMLM_REQUIRE(ka < (int64_t(1) << 24), "gemm_tn: too many channels");
  cudaStream_t st = as_stream(stream);
  if (M == 0) {
    GMLM_CUDA_TRY(cudaMemset2DAsync(D, size_t(ldd) * sizeof(float), 0, size_t(N) * sizeof(float), size_t(ka), st));
    return GMLM_OK;
  }
  const TnPlan t = tn_plan(num_sources, K_host, M, N);
  GMLM_REQUIRE(ws_bytes >= gmlm_gemm_tn_workspace_bytes(num_sources, K_host, M, N) && (t.splits == 1 || ws),
               "gemm_tn: workspace too small");
  {
    static thread_local bool ctx_ready[kMaxDevices] = {};
    const int dev = current_device();
    if (!ctx_ready[dev]) {
      GMLM_CUDA_TRY(cudaFree(nullptr));
      ctx_ready[dev] = true;
    }
  }
  TnParams p{};
  SourceMaps ma;
  CUtensorMap mg;
  int tiles = 0, koff = 0;
  for (int i = 0; i < kMaxSources; ++i) {
    if (i < num_sources) {
      if (int rc = get_map(&ma.m[i], A_host[i], M, K_host[i], lda_host[i], TN_BK, 64, in_dtype)) return rc;
      p.k_off[i] = koff;
      p.k_len[i] = int(K_host[i]);
      koff += int(K_host[i]);
      tiles += int((K_host[i] + BLOCK_M - 1) / BLOCK_M);
    } else {
      ma.m[i] = ma.m[0];
      p.k_off[i] = koff;
      p.k_len[i] = 0;
    }
    p.tile_end[i] = tiles;
  }
  if (int rc = get_map(&mg, G, M, N, ldg, TN_BK, 64, in_dtype)) return rc;
  p.M = int(M); p.N = int(N); p.n_src = num_sources;
  p.ka_tiles = t.ka_tiles; p.n_tiles = t.n_tiles; p.splits = t.splits; p.kb_per_split = t.kb_per_split;
  p.idesc_formats = in_dtype == GMLM_BF16 ? 1u : 0u;
  if (t.splits == 1) {
    p.out = D; p.ldo = ldd; p.slab = 0;
  } else {
    p.out = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
    p.ldo = t.ld_ws; p.slab = ka * t.ld_ws;
  }
  int rc = t.bn == 256 ? launch_tn<256>(ma, mg, p, st) : (t.bn == 128 ? launch_tn<128>(ma, mg, p, st) : launch_tn<64>(ma, mg, p, st));
  if (rc) return rc;
  if (t.splits > 1) {
    const int64_t n4 = (N + 3) / 4;
    GMLM_REQUIRE(ldd >= n4 * 4, "gemm_tn: D row pitch must cover N rounded up to 4");
    const int64_t total = ka * n4;
    gemm_tn_reduce_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(p.out, t.splits, p.slab, ka, n4, t.ld_ws, D, ldd);
    GMLM_LAUNCH_CHECK();
  }
  return GMLM_OK;
}
