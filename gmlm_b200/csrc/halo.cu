// halo.cu — row pack / unpack for the halo exchange of a destination-row partitioned graph
// (no reference counterpart — /root/reference is single-process).
//
//   gather_rows      : out[k, :]      = x[ids[k], :]          (pack the rows a peer needs)
//   scatter_add_rows : dst[ids[k], :] += src[k, :]            (return halo gradients to owners)
//
// `ids` must be unique within one call (true for one peer's send list), so the scatter needs no
// atomics and the sum order is fixed by the order of the calls (peer by peer): deterministic.
// HBM-bound: 128-bit accesses, one row handled by feat/VEC consecutive threads.
#include <algorithm>

#include "common.cuh"

namespace gmlm {
int tuning_halo_pull_ctas();     // graph_build.cu: CTA cap of gather_rows_ptr (0 = 16 per SM)
int tuning_halo_pull_threads();  // threads per CTA of gather_rows_ptr (0 = 256)
namespace {

template <typename T, int VEC, bool ADD>
__global__ void __launch_bounds__(256) rows_move_kernel(const T* __restrict__ src, int64_t lds, T* dst,
                                                        int64_t ldd, const int64_t* __restrict__ ids, int64_t n,
                                                        int64_t feat, int64_t src_rows) {
  const int64_t packs = (feat + VEC - 1) / VEC;
  const int64_t total = n * packs;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t k = i / packs;
    const int64_t f = (i - k * packs) * VEC;
    const int64_t id = __ldg(ids + k);
    if (ADD) {  // dst[id] += src[k]
      Pack<T, VEC> a, b;
      a.load(src + k * lds + f);
      b.load_rw(dst + id * ldd + f);          // read-modify-write operand: coherent load, not ld.global.nc
      float fa[VEC], fb[VEC];
      a.unpack(fa);
      b.unpack(fb);
#pragma unroll
      for (int j = 0; j < VEC; ++j) fb[j] += fa[j];
      b.pack(fb);
      b.store(dst + id * ldd + f);
    } else {    // dst[k] = src[id]
      Pack<T, VEC> a;
      a.load(src + id * lds + f);
      a.store(dst + k * ldd + f);
    }
  }
}

// out[k, :] = *(row_ptrs[k])[0:feat]   — every row has its own 64-bit source address, so one
// launch pulls from all NVLink peers at once (addresses in peer-mapped symmetric memory).
template <typename T, int VEC>
__global__ void __launch_bounds__(1024) gather_ptr_kernel(const T* const* __restrict__ row_ptrs,
                                                          const int64_t* __restrict__ out_ids, int64_t n,
                                                          int64_t feat, T* __restrict__ out, int64_t ldo) {
  // eight 16-byte remote loads in flight per thread.  Two launch shapes: alone on the GPU, many 256-thread
  // CTAs; next to an aggregation kernel, a FEW 1024-thread CTAs -- a pull CTA fills its SM's outstanding-load
  // capacity with 3 us NVLink requests and halves the aggregation on that SM (measured), so the pull is
  // confined to a few SMs instead of being sprinkled over all of them
  constexpr int U = 8;
  const int64_t packs = feat / VEC;
  const int64_t total = n * packs;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride * U) {
    Pack<T, VEC> a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = i + u * stride;
      if (j < total) {
        const int64_t k = j / packs;
        a[u].load(row_ptrs[k] + (j - k * packs) * VEC);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = i + u * stride;
      if (j < total) {
        const int64_t k = j / packs;
        const int64_t o = out_ids ? out_ids[k] : k;
        a[u].store(out + o * ldo + (j - k * packs) * VEC);
      }
    }
  }
}


// ------------------------------------------------------------------------------ TMA halo pull
// out[out_ids[k], :] = *(row_ptrs[k])  with the rows moved by the bulk-copy engine (cp.async.bulk, SASS
// UBLKCP) instead of by LDG/STG:  peer memory --bulk load--> shared-memory ring --bulk store--> local HBM.
//
// Why (measured in round 1): an LDG-driven pull that runs NEXT TO the aggregation kernel is
// zero-sum.  Its 3 us NVLink requests sit in the same per-SM load queue as the aggregation's 1 us HBM
// gathers (sprinkled over all SMs it halved the aggregation's throughput), and confined to a few SMs it
// is capped by what one SM's load path keeps in flight (~24 KB: 260 GB/s from 32 SMs).  Bulk copies do
// not pass through the LSU / L1 queue and their bytes in flight are bounded by SHARED MEMORY instead:
// one warp keeps SLOTS-2 batches of 32 rows (16 KB each at 512-byte rows) outstanding, i.e. ~160 KB per
// SM, so a couple of dozen single-warp CTAs cover the NVLink bandwidth-delay product while the
// aggregation's CTAs (which use no shared memory) stay resident on the same SMs.
// Measured on one B200 (profiles/r2_row_gather_tma_vs_ldg.log): ONE warp moves 12.3 M rows/s whatever the row
// width or ring size (6.3 GB/s at 512-byte rows) -- UBLKCP takes its operands from uniform registers, so the
// 32 lanes of a batch are issued one after the other (~75 cycles per bulk op) -- and the rate scales linearly
// with the number of warps (296 single-warp CTAs: 1.8 TB/s).  Hence several warps per CTA, each with its own
// ring, and ~120 warps for an NVLink's worth of 512-byte rows.
//
// Every warp works alone; lane l moves row l of a batch.  Per slot one mbarrier: lane 0 posts the expected
// byte count, every lane issues its row's bulk load against it, the warp waits for the phase, every
// lane issues its row's bulk store and commits it to its own bulk group; a slot is refilled two
// iterations later, after `wait_group.read 1` has shown that the store which read it has drained.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

template <int UNUSED = 0>
__global__ void __launch_bounds__(256, 1) gather_ptr_tma_kernel(const uint64_t* __restrict__ row_ptrs,
                                                                const int64_t* __restrict__ out_ids, int64_t n,
                                                                uint32_t row_bytes, uint8_t* __restrict__ out,
                                                                int64_t ldo_bytes, int slots, int rb) {
  // rb = rows per batch (32, 16 or 8: lanes >= rb idle) -- small batches keep the ring, and with it the bite this
  // kernel takes out of the SM's unified L1 / shared memory, small when it runs under an aggregation kernel
  extern __shared__ __align__(128) uint8_t tma_smem_all[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint32_t slot_bytes = uint32_t(rb) * row_bytes;
  uint8_t* tma_smem = tma_smem_all + size_t(warp) * slots * slot_bytes;              // this warp's ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(tma_smem_all + size_t(nwarps) * slots * slot_bytes) + warp * slots;
  if (lane == 0) {
    for (int s = 0; s < slots; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t n_batches = (n + rb - 1) / rb;
  const int64_t first = int64_t(blockIdx.x) * nwarps + warp, stride = int64_t(gridDim.x) * nwarps;
  const int depth = slots - 2;                       // batches in flight

  auto issue_load = [&](int64_t it) {
    const int64_t b = first + it * stride;
    if (b >= n_batches) return;
    const int s = int(it % slots);
    const int64_t k = b * rb + lane;
    const int cnt = int(min(int64_t(rb), n - b * rb));
    const uint32_t bar = smem_addr(bars + s);
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(uint32_t(cnt) * row_bytes)
                   : "memory");
    __syncwarp();
    if (lane < cnt) {
      const uint64_t src = row_ptrs[k];
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr(tma_smem + size_t(s) * slot_bytes + size_t(lane) * row_bytes)),
                   "l"(src), "r"(row_bytes), "r"(bar)
                   : "memory");
    }
  };

  for (int it = 0; it < depth; ++it) issue_load(it);
  for (int64_t it = 0;; ++it) {
    const int64_t b = first + it * stride;
    if (b >= n_batches) break;
    // the slot refilled now was read by the store of iteration it-2: all but my newest store group are done reading
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    issue_load(it + depth);
    const int s = int(it % slots);
    const uint32_t bar = smem_addr(bars + s);
    const uint32_t parity = uint32_t((it / slots) & 1);
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
      if (ok) break;
      if (++spins > (1u << 26)) __trap();            // a lost transfer traps instead of hanging the GPU
    }
    const int64_t k = b * rb + lane;
    if (lane < rb && k < n) {
      const int64_t o = out_ids ? out_ids[k] : k;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + o * ldo_bytes),
                   "r"(smem_addr(tma_smem + size_t(s) * slot_bytes + size_t(lane) * row_bytes)), "r"(row_bytes)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every row has landed in `out`
}

// dst[row_ids[r], :] += sum_{e in [rowptr[r], rowptr[r+1])} *(entry_ptrs[e])[0:feat]
// fp32 accumulation in entry order (entries of a row are stored in a fixed peer order), one
// rounding at the end; each destination row is owned by one thread group => no atomics.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) reduce_ptr_kernel(T* dst, int64_t ldd,
                                                         const int64_t* __restrict__ row_ids,
                                                         const int32_t* __restrict__ rowptr,
                                                         const T* const* __restrict__ entry_ptrs, int64_t n_rows,
                                                         int64_t feat) {
  const int64_t packs = feat / VEC;
  const int64_t total = n_rows * packs;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / packs;
    const int64_t f = (i - r * packs) * VEC;
    const int32_t b = rowptr[r], e = rowptr[r + 1];
    T* d = dst + row_ids[r] * ldd + f;
    Pack<T, VEC> p;
    p.load_rw(d);                               // read-modify-write operand: coherent load
    float acc[VEC];
    p.unpack(acc);
    int32_t k = b;
    for (; k + 4 <= e; k += 4) {   // four peers' rows in flight, added in order
      Pack<T, VEC> v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u].load(entry_ptrs[k + u] + f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float t[VEC];
        v[u].unpack(t);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += t[j];
      }
    }
    for (; k < e; ++k) {
      Pack<T, VEC> v;
      v.load(entry_ptrs[k] + f);
      float t[VEC];
      v.unpack(t);
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += t[j];
    }
    p.pack(acc);
    p.store(d);
  }
}

template <bool ADD>
int launch(const void* src, int64_t lds, void* dst, int64_t ldd, const int64_t* ids, int64_t n, int64_t feat,
           int dtype, cudaStream_t st) {
  if (n == 0 || feat == 0) return GMLM_OK;
  const int v = dtype == GMLM_F32 ? 4 : 8;
  const bool vec = feat % v == 0 && lds % v == 0 && ldd % v == 0 &&
                   (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
  const int64_t packs = vec ? feat / v : feat;
  int64_t blocks = (n * packs + 255) / 256;
  const int64_t cap = int64_t(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == GMLM_F32) {
    if (vec) rows_move_kernel<float, 4, ADD><<<unsigned(blocks), 256, 0, st>>>(
        static_cast<const float*>(src), lds, static_cast<float*>(dst), ldd, ids, n, feat, 0);
    else rows_move_kernel<float, 1, ADD><<<unsigned(blocks), 256, 0, st>>>(
        static_cast<const float*>(src), lds, static_cast<float*>(dst), ldd, ids, n, feat, 0);
  } else {
    using B = __nv_bfloat16;
    if (vec) rows_move_kernel<B, 8, ADD><<<unsigned(blocks), 256, 0, st>>>(
        static_cast<const B*>(src), lds, static_cast<B*>(dst), ldd, ids, n, feat, 0);
    else rows_move_kernel<B, 1, ADD><<<unsigned(blocks), 256, 0, st>>>(
        static_cast<const B*>(src), lds, static_cast<B*>(dst), ldd, ids, n, feat, 0);
  }
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

}  // namespace
}  // namespace gmlm

using namespace gmlm;

extern "C" int gmlm_gather_rows(const void* x, int dtype, int64_t feat, int64_t ldx, const int64_t* ids, int64_t n,
                                void* out, int64_t ldo, void* stream) {
  GMLM_REQUIRE(dtype == GMLM_F32 || dtype == GMLM_BF16, "gather_rows: dtype must be GMLM_F32 or GMLM_BF16");
  GMLM_REQUIRE(n >= 0 && feat >= 0 && ldx >= feat && ldo >= feat, "gather_rows: bad sizes");
  GMLM_REQUIRE(n == 0 || (x && ids && out), "gather_rows: null pointer");
  return launch<false>(x, ldx, out, ldo, ids, n, feat, dtype, as_stream(stream));
}

extern "C" int gmlm_scatter_add_rows(void* dst, int dtype, int64_t feat, int64_t ldd, const int64_t* ids, int64_t n,
                                     const void* src, int64_t lds, void* stream) {
  GMLM_REQUIRE(dtype == GMLM_F32 || dtype == GMLM_BF16, "scatter_add_rows: dtype must be GMLM_F32 or GMLM_BF16");
  GMLM_REQUIRE(n >= 0 && feat >= 0 && ldd >= feat && lds >= feat, "scatter_add_rows: bad sizes");
  GMLM_REQUIRE(n == 0 || (dst && ids && src), "scatter_add_rows: null pointer");
  return launch<true>(src, lds, dst, ldd, ids, n, feat, dtype, as_stream(stream));
}

extern "C" int gmlm_gather_rows_ptr(const void* const* row_ptrs, const int64_t* out_ids, int dtype, int64_t feat,
                                    int64_t n, void* out, int64_t ldo, void* stream) {
  GMLM_REQUIRE(dtype == GMLM_F32 || dtype == GMLM_BF16, "gather_rows_ptr: dtype must be GMLM_F32 or GMLM_BF16");
  const int v = dtype == GMLM_F32 ? 4 : 8;
  GMLM_REQUIRE(n >= 0 && feat >= 0 && ldo >= feat, "gather_rows_ptr: bad sizes");
  GMLM_REQUIRE(feat % v == 0 && ldo % v == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
               "gather_rows_ptr: rows must be 16-byte multiples and 16-byte aligned");
  if (n == 0 || feat == 0) return GMLM_OK;
  GMLM_REQUIRE(row_ptrs && out, "gather_rows_ptr: null pointer");
  int threads = tuning_halo_pull_threads();
  if (threads <= 0) threads = 256;
  threads = std::min(1024, (threads + 31) / 32 * 32);
  int64_t blocks = (n * (feat / v) + threads - 1) / threads;
  const int tuned = tuning_halo_pull_ctas();
  const int64_t cap = tuned > 0 ? int64_t(tuned) : int64_t(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = as_stream(stream);
  if (dtype == GMLM_F32)
    gather_ptr_kernel<float, 4><<<unsigned(blocks), threads, 0, st>>>(reinterpret_cast<const float* const*>(row_ptrs),
                                                                      out_ids, n, feat, static_cast<float*>(out), ldo);
  else
    gather_ptr_kernel<__nv_bfloat16, 8><<<unsigned(blocks), threads, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16* const*>(row_ptrs), out_ids, n, feat, static_cast<__nv_bfloat16*>(out),
        ldo);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

extern "C" int gmlm_reduce_rows_ptr(void* dst, int dtype, int64_t feat, int64_t ldd, const int64_t* row_ids,
                                    const int32_t* rowptr, const void* const* entry_ptrs, int64_t n_rows,
                                    void* stream) {
  GMLM_REQUIRE(dtype == GMLM_F32 || dtype == GMLM_BF16, "reduce_rows_ptr: dtype must be GMLM_F32 or GMLM_BF16");
  const int v = dtype == GMLM_F32 ? 4 : 8;
  GMLM_REQUIRE(n_rows >= 0 && feat >= 0 && ldd >= feat, "reduce_rows_ptr: bad sizes");
  GMLM_REQUIRE(feat % v == 0 && ldd % v == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0,
               "reduce_rows_ptr: rows must be 16-byte multiples and 16-byte aligned");
  if (n_rows == 0 || feat == 0) return GMLM_OK;
  GMLM_REQUIRE(dst && row_ids && rowptr && entry_ptrs, "reduce_rows_ptr: null pointer");
  int64_t blocks = (n_rows * (feat / v) + 255) / 256;
  const int64_t cap = int64_t(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = as_stream(stream);
  if (dtype == GMLM_F32)
    reduce_ptr_kernel<float, 4><<<unsigned(blocks), 256, 0, st>>>(static_cast<float*>(dst), ldd, row_ids, rowptr,
                                                                  reinterpret_cast<const float* const*>(entry_ptrs),
                                                                  n_rows, feat);
  else
    reduce_ptr_kernel<__nv_bfloat16, 8><<<unsigned(blocks), 256, 0, st>>>(
        static_cast<__nv_bfloat16*>(dst), ldd, row_ids, rowptr,
        reinterpret_cast<const __nv_bfloat16* const*>(entry_ptrs), n_rows, feat);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

// TMA variant of gmlm_gather_rows_ptr: `ctas` CTAs (0 = one per SM) of `warps` warps (0 = 3, at most 8), each
// warp with its own ring of `rows_per_batch`-row batches (0 = 32; 16, 8) inside the CTA's `smem_kb` KiB (0 = 200) of
// shared memory.  Rows must be 16-byte multiples, 16-byte aligned on both sides, and at least three batches per
// warp must fit the ring.
extern "C" int gmlm_gather_rows_ptr_tma(const void* const* row_ptrs, const int64_t* out_ids, int dtype, int64_t feat,
                                        int64_t n, void* out, int64_t ldo, int ctas, int warps, int smem_kb,
                                        int rows_per_batch, void* stream) {
  GMLM_REQUIRE(dtype == GMLM_F32 || dtype == GMLM_BF16, "gather_rows_ptr_tma: dtype must be GMLM_F32 or GMLM_BF16");
  const int esz = dtype == GMLM_F32 ? 4 : 2;
  const int64_t row_bytes = feat * esz;
  GMLM_REQUIRE(n >= 0 && feat >= 0 && ldo >= feat, "gather_rows_ptr_tma: bad sizes");
  GMLM_REQUIRE(row_bytes % 16 == 0 && (ldo * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
               "gather_rows_ptr_tma: rows must be 16-byte multiples and 16-byte aligned");
  if (n == 0 || feat == 0) return GMLM_OK;
  GMLM_REQUIRE(row_ptrs && out, "gather_rows_ptr_tma: null pointer");
  if (smem_kb <= 0) smem_kb = 200;
  GMLM_REQUIRE(smem_kb <= 224, "gather_rows_ptr_tma: at most 224 KiB of shared memory per CTA");
  GMLM_REQUIRE(warps >= 0 && warps <= 8, "gather_rows_ptr_tma: 0..8 warps per CTA");
  const int rb = rows_per_batch <= 0 ? 32 : rows_per_batch;
  GMLM_REQUIRE(rb == 32 || rb == 16 || rb == 8, "gather_rows_ptr_tma: rows_per_batch must be 32, 16 or 8");
  const int64_t slot_bytes = rb * row_bytes;
  const int64_t budget = int64_t(smem_kb) * 1024 - 8 * 64 * 8;
  if (warps == 0) {                       // as many warps (up to 3) as still get a four-slot ring each
    warps = 3;
    while (warps > 1 && budget / (warps * slot_bytes) < 4) --warps;
  }
  const int slots = int(std::min<int64_t>(64, budget / (warps * slot_bytes)));
  GMLM_REQUIRE(slots >= 3, "gather_rows_ptr_tma: rows too wide for the shared-memory ring (three batches per warp)");
  const size_t smem = size_t(warps) * slots * slot_bytes + size_t(warps) * slots * 8;
  if (ctas <= 0) ctas = num_sms();
  const int64_t n_batches = (n + rb - 1) / rb;
  if (int64_t(ctas) * warps > n_batches) ctas = int((n_batches + warps - 1) / warps);
  auto kern = gather_ptr_tma_kernel<0>;
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    GMLM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    configured[dev] = true;
  }
  kern<<<unsigned(ctas), 32 * warps, smem, as_stream(stream)>>>(reinterpret_cast<const uint64_t*>(row_ptrs), out_ids, n,
                                                               uint32_t(row_bytes), static_cast<uint8_t*>(out),
                                                               ldo * esz, slots, rb);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}
