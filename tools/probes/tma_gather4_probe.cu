// Development probe: does cp.async.bulk.tensor ... tile::gather4 gather 4 arbitrary 512-byte rows per op on
// sm_100a, with which box shape in the tensor map, and how fast?   nvcc -arch=sm_100a -o probe ... ; ./probe <box_rows>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// each warp: ring of SLOTS tiles of (ROWS4 gather4 ops x 4 rows x 512 B); lane l issues gather4 op l of a batch
template <int OPS, int SLOTS>
__global__ void __launch_bounds__(128, 1) gather4_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx,
                                                         long n, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t OP_BYTES = 4 * 512, SLOT_BYTES = OPS * OP_BYTES;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint8_t* ring = smem + size_t(warp) * SLOTS * SLOT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(nw) * SLOTS * SLOT_BYTES) + warp * SLOTS;
  if (lane == 0) {
    for (int s = 0; s < SLOTS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bars + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long rows_per_batch = OPS * 4;
  const long n_batches = n / rows_per_batch;            // n is a multiple
  const long first = long(blockIdx.x) * nw + warp, stride = long(gridDim.x) * nw;
  auto issue = [&](long it) {
    const long b = first + it * stride;
    if (b >= n_batches) return;
    const int s = int(it % SLOTS);
    const uint32_t bar = s32(bars + s);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(SLOT_BYTES) : "memory");
    __syncwarp();
    if (lane < OPS) {
      const int* ip = idx + b * rows_per_batch + lane * 4;
      const int r0 = ip[0], r1 = ip[1], r2 = ip[2], r3 = ip[3];
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
          " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
          ::"r"(s32(ring + size_t(s) * SLOT_BYTES + size_t(lane) * OP_BYTES)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar),
            "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
          : "memory");
    }
  };
  constexpr int D = SLOTS - 2;
  for (int it = 0; it < D; ++it) issue(it);
  for (long it = 0;; ++it) {
    const long b = first + it * stride;
    if (b >= n_batches) break;
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    issue(it + D);
    const int s = int(it % SLOTS);
    const uint32_t bar = s32(bars + s), parity = uint32_t((it / SLOTS) & 1);
    uint32_t spins = 0;
    while (true) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
      if (ok) break;
      if (++spins > (1u << 24)) __trap();
    }
    if (lane == 0) {   // the batch is contiguous in `out`: one bulk store
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + b * rows_per_batch * 256),
                   "r"(s32(ring + size_t(s) * SLOT_BYTES)), "r"(SLOT_BYTES) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  const long N = 4000000, M = 4000000;   // source rows, gathered rows
  __nv_bfloat16 *x, *out;
  int* idx;
  CK(cudaMalloc(&x, N * 512));
  CK(cudaMalloc(&out, M * 512));
  CK(cudaMalloc(&idx, M * 4));
  std::vector<uint16_t> hx(size_t(N) * 256);
  for (size_t i = 0; i < hx.size(); ++i) hx[i] = uint16_t((i * 2654435761u) >> 16);
  std::vector<int> hidx(M);
  uint64_t st = 88172645463325252ull;
  for (long i = 0; i < M; ++i) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; hidx[i] = int(st % N); }
  CK(cudaMemcpy(x, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(idx, hidx.data(), M * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(out, 0, M * 512));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  CUtensorMap map;
  cuuint64_t dims[2] = {256, cuuint64_t(N)};
  cuuint64_t strides[1] = {512};
  cuuint32_t box[2] = {256, cuuint32_t(box_rows)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeFn>(fp)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, strides, box, es,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box {256,%d}: CUresult %d\n", box_rows, int(r));
  if (r != CUDA_SUCCESS) return 2;
  constexpr int OPS = 4, SLOTS = 6;       // 16 rows = 8 KB per batch
  const int warps = 4;
  const size_t smem = size_t(warps) * SLOTS * OPS * 2048 + warps * SLOTS * 8 + 1024;
  CK(cudaFuncSetAttribute(gather4_kernel<OPS, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int ctas : {148, 296}) {
    gather4_kernel<OPS, SLOTS><<<ctas, warps * 32, smem>>>(map, idx, M, out);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) gather4_kernel<OPS, SLOTS><<<ctas, warps * 32, smem>>>(map, idx, M, out);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    printf("ctas %d x %d warps: %.3f ms  %.0f GB/s read+write\n", ctas, warps, ms, 2.0 * M * 512 / ms / 1e6);
  }
  std::vector<uint16_t> ho(size_t(4096) * 256);
  CK(cudaMemcpy(ho.data(), out, ho.size() * 2, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (long i = 0; i < 4096; ++i)
    for (int c = 0; c < 256; ++c)
      if (ho[i * 256 + c] != hx[size_t(hidx[i]) * 256 + c]) ++bad;
  printf("mismatches in the first 4096 rows: %ld\n", bad);
  return bad ? 3 : 0;
}
