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
__global__ void __launch_bounds__(256) rows_move_kernel(const T* __restrict__ src, int64_t lds, T* __restrict__ dst,
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
      b.load(dst + id * ldd + f);
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

// dst[row_ids[r], :] += sum_{e in [rowptr[r], rowptr[r+1])} *(entry_ptrs[e])[0:feat]
// fp32 accumulation in entry order (entries of a row are stored in a fixed peer order), one
// rounding at the end; each destination row is owned by one thread group => no atomics.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) reduce_ptr_kernel(T* __restrict__ dst, int64_t ldd,
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
    p.load(d);
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
