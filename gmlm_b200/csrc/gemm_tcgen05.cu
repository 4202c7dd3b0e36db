// gemm_tcgen05.cu — the dense feature transform of the RGCN layer on the
// 5th-generation tensor cores:  C[M,N] = [A1 | A2][M,K] · B[N,K]^T + bias[N]
//
//   forward : out = [H | x] · [W_live ; root]   (one GEMM instead of upstream's R+1 matmuls + adds,
//             [PyG] RGCNConv.forward called main.py:272,285,298,308); the two A sources are two TMA
//             descriptors, so H and x are never concatenated in memory.
//   backward: [dH | dx_root] = g · [W_live ; root]^T, written through two output pointers so that
//             dH lands contiguous for the transposed aggregation.
//
// bf16 operands, fp32 accumulation in TMEM, bf16 or fp32 output.  Both operands are K-major.
// Structure (one CTA per 128 x BLOCK_N tile, 6 warps):
//   warp 0   : TMA producer — cp.async.bulk.tensor 2-D tiles (128B swizzle) into a STAGES-deep ring,
//              mbarrier expect_tx / complete_tx
//   warp 1   : TMEM allocation + single-thread tcgen05.mma issue (UMMA 128 x BLOCK_N x 16, cta_group::1),
//              tcgen05.commit releases smem stages and finally signals the epilogue
//   warps 2-5: epilogue — tcgen05.ld (32 lanes x 32 columns), + bias, convert, 128-bit row stores
// The shapes of this path are tall and skinny (M = #nodes, N = 64..512, K = 320..1280): the kernel
// is bound by streaming A from HBM, which the TMA ring keeps in flight; B (<= 160 KB) lives in L2.
#include <cuda.h>

#include "common.cuh"

namespace gmlm {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
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
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M x N
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GemmParams {
  int M, N, K, K1;        // K1: columns served by the first A descriptor (multiple of BLOCK_K)
  int N1;                 // output columns [0,N1) go to C1, [N1,N) to C2 (multiple of BLOCK_N, or N)
  const float* bias;      // [N] or nullptr
  void* C1;
  int64_t ldc1;
  void* C2;
  int64_t ldc2;
};

template <int BLOCK_N, int STAGES, typename OutT>
__global__ void __launch_bounds__(kThreads) gemm_nt_kernel(const __grid_constant__ CUtensorMap tma_a1,
                                                           const __grid_constant__ CUtensorMap tma_a2,
                                                           const __grid_constant__ CUtensorMap tma_b,
                                                           const GemmParams p) {
  constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * BLOCK_N;
  const int num_kb = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(empty + s), 1);
    }
    mbar_init(smem_u32(tmem_full), 1);
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
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(empty + s), ph ^ 1);
        mbar_expect_tx(smem_u32(full + s), A_BYTES + B_BYTES);
        const int k0 = kb * BLOCK_K;
        if (k0 < p.K1) tma_load_2d(smem_u32(smem_a + s * A_BYTES), &tma_a1, smem_u32(full + s), k0, m0);
        else tma_load_2d(smem_u32(smem_a + s * A_BYTES), &tma_a2, smem_u32(full + s), k0 - p.K1, m0);
        tma_load_2d(smem_u32(smem_b + s * B_BYTES), &tma_b, smem_u32(full + s), k0, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer (single thread)
      const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(smem_u32(full + s), ph);
        tc_fence_after();
        const uint64_t da = make_smem_desc(smem_u32(smem_a + s * A_BYTES));
        const uint64_t db = make_smem_desc(smem_u32(smem_b + s * B_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // advancing 16 bf16 along K inside the swizzle atom = +32 bytes = +2 in the address field
          umma(tmem_base, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(smem_u32(empty + s));      // frees the smem stage once these MMAs have read it
      }
      umma_commit(smem_u32(tmem_full));        // accumulator complete
    }
  } else {
    // ---------------- epilogue: warp w owns TMEM lanes [32*(w%4), 32*(w%4)+32) = 32 output rows.
    // TMEM -> registers (one row per lane) -> shared memory (the drained pipeline stages, padded
    // rows: conflict-free 16-byte writes) -> global with fully coalesced 128-bit row segments.
    const int quad = warp & 3;
    mbar_wait(smem_u32(tmem_full), 0);
    tc_fence_after();
    constexpr int ROW_BYTES = BLOCK_N * int(sizeof(OutT));
    constexpr int STRIDE = ROW_BYTES + 16;                 // odd multiple of 16 bytes
    uint8_t* stage = smem + quad * 32 * STRIDE;            // K loop is over: the operand ring is free
    uint8_t* my_row = stage + lane * STRIDE;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(c0), r);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = __uint_as_float(r[j]);
        if (p.bias) v[j] += __ldg(p.bias + n0 + c0 + j);
      }
      if constexpr (sizeof(OutT) == 2) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          Pack<__nv_bfloat16, 8> o;
          o.pack(v + j);
          *reinterpret_cast<uint4*>(my_row + (c0 + j) * 2) = o.v;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(my_row + (c0 + j) * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    __syncwarp();
    const bool second = n0 >= p.N1;
    uint8_t* cbase = second ? static_cast<uint8_t*>(p.C2) + int64_t(n0 - p.N1) * sizeof(OutT)
                            : static_cast<uint8_t*>(p.C1) + int64_t(n0) * sizeof(OutT);
    const int64_t ldc_bytes = (second ? p.ldc2 : p.ldc1) * int64_t(sizeof(OutT));
    constexpr int CHUNKS = ROW_BYTES / 16;                 // 16-byte chunks per output row
    const int row0 = m0 + quad * 32;
#pragma unroll 4
    for (int idx = lane; idx < 32 * CHUNKS; idx += 32) {
      const int rr = idx / CHUNKS;
      const int cc = idx - rr * CHUNKS;
      if (row0 + rr < p.M) {
        const uint4 val = *reinterpret_cast<const uint4*>(stage + rr * STRIDE + cc * 16);
        __stcs(reinterpret_cast<uint4*>(cbase + int64_t(row0 + rr) * ldc_bytes + cc * 16), val);
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

// row-major [rows, cols] bf16 matrix with leading dimension ld (elements); box = 64 cols x box_rows
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(GMLM_ERR_CUDA, "gemm: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
  cuuint32_t box[2] = {cuuint32_t(BLOCK_K), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GMLM_ERR_CUDA, "gemm: cuTensorMapEncodeTiled failed with code %d", int(r));
  return GMLM_OK;
}

template <int BLOCK_N, typename OutT>
int launch(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b, const GemmParams& p, cudaStream_t st) {
  constexpr int STAGES = 4;   // BLOCK_N <= 64: 2 CTAs/SM (prologue/epilogue of one overlap the K loop of the other)
  constexpr size_t smem = size_t(STAGES) * (BLOCK_M * BLOCK_K * 2 + BLOCK_N * BLOCK_K * 2) + (2 * STAGES + 2) * 8 + 1024;
  auto kern = gemm_nt_kernel<BLOCK_N, STAGES, OutT>;
  static bool configured = false;
  if (!configured) {
    GMLM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    configured = true;
  }
  dim3 grid((p.M + BLOCK_M - 1) / BLOCK_M, p.N / BLOCK_N);
  kern<<<grid, kThreads, smem, st>>>(a1, a2, b, p);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

}  // namespace
}  // namespace gmlm

using namespace gmlm;

extern "C" int gmlm_gemm_nt_bf16(const void* A1, int64_t lda1, int64_t K1, const void* A2, int64_t lda2, int64_t K2,
                                 const void* B, int64_t ldb, const float* bias, int64_t M, int64_t N, void* C1,
                                 int64_t ldc1, int64_t N1, void* C2, int64_t ldc2, int out_dtype, void* stream) {
  const int64_t K = K1 + K2;
  GMLM_REQUIRE(M >= 0 && N > 0 && K1 > 0 && K2 >= 0, "gemm: bad sizes");
  GMLM_REQUIRE(M < (int64_t(1) << 31) && N <= 65535 * 256 && K < (int64_t(1) << 31), "gemm: sizes exceed int32");
  GMLM_REQUIRE(K1 % BLOCK_K == 0 && K2 % BLOCK_K == 0, "gemm: K1 and K2 must be multiples of 64");
  GMLM_REQUIRE(N % 32 == 0, "gemm: N must be a multiple of 32");
  GMLM_REQUIRE(out_dtype == GMLM_F32 || out_dtype == GMLM_BF16, "gemm: out_dtype must be GMLM_F32 or GMLM_BF16");
  GMLM_REQUIRE(A1 && B && C1 && (K2 == 0 || A2), "gemm: null pointer");
  GMLM_REQUIRE(lda1 >= K1 && (K2 == 0 || lda2 >= K2) && ldb >= K, "gemm: bad leading dimensions");
  GMLM_REQUIRE(lda1 % 8 == 0 && (K2 == 0 || lda2 % 8 == 0) && ldb % 8 == 0, "gemm: leading dimensions must be multiples of 8");
  if (M == 0) return GMLM_OK;
  int bn = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32));
  if (N1 <= 0 || N1 >= N) { N1 = N; C2 = C1; ldc2 = ldc1; }
  else {
    GMLM_REQUIRE(C2 != nullptr, "gemm: second output missing");
    while (bn > 32 && N1 % bn != 0) bn >>= 1;
    GMLM_REQUIRE(N1 % bn == 0 && (N - N1) % bn == 0, "gemm: N1 must split N on a tile boundary");
  }
  const int esz = out_dtype == GMLM_F32 ? 4 : 2;
  GMLM_REQUIRE((reinterpret_cast<uintptr_t>(C1) & 15) == 0 && (reinterpret_cast<uintptr_t>(C2) & 15) == 0 &&
                   (ldc1 * esz) % 16 == 0 && (ldc2 * esz) % 16 == 0,
               "gemm: outputs must be 16-byte aligned");
  // cuTensorMapEncodeTiled is a driver entry point: it needs a current context on THIS thread.
  // Autograd worker threads may not have one bound yet in this library's runtime instance, so
  // bind the device that owns the operands.
  {
    cudaPointerAttributes attr;
    GMLM_CUDA_TRY(cudaPointerGetAttributes(&attr, A1));
    GMLM_REQUIRE(attr.type == cudaMemoryTypeDevice, "gemm: A1 is not device memory");
    GMLM_CUDA_TRY(cudaSetDevice(attr.device));
    GMLM_CUDA_TRY(cudaFree(nullptr));
  }
  CUtensorMap ma1, ma2, mb;
  int rc = make_map(&ma1, A1, M, K1, lda1, BLOCK_M);
  if (rc) return rc;
  rc = K2 ? make_map(&ma2, A2, M, K2, lda2, BLOCK_M) : make_map(&ma2, A1, M, K1, lda1, BLOCK_M);
  if (rc) return rc;
  rc = make_map(&mb, B, N, K, ldb, bn);
  if (rc) return rc;
  GemmParams p;
  p.M = int(M); p.N = int(N); p.K = int(K); p.K1 = int(K1); p.N1 = int(N1);
  p.bias = bias; p.C1 = C1; p.ldc1 = ldc1; p.C2 = C2; p.ldc2 = ldc2;
  cudaStream_t st = as_stream(stream);
#define GMLM_GEMM_CASE(BN)                                                                        \
  case BN:                                                                                        \
    return out_dtype == GMLM_F32 ? launch<BN, float>(ma1, ma2, mb, p, st)                         \
                                 : launch<BN, __nv_bfloat16>(ma1, ma2, mb, p, st);
  switch (bn) {
    GMLM_GEMM_CASE(256)
    GMLM_GEMM_CASE(128)
    GMLM_GEMM_CASE(64)
    GMLM_GEMM_CASE(32)
  }
#undef GMLM_GEMM_CASE
  return fail(GMLM_ERR_INVALID, "gemm: unsupported tile");
}
