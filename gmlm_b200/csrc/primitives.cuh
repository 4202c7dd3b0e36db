// primitives.cuh — the three data-parallel primitives the graph build needs, written for this library (round 1 used
// CUB's DeviceScan / DeviceRadixSort / DeviceSelect): an exclusive prefix sum over int32, a STABLE least-significant-
// digit radix sort of (uint32 key, int32 value) pairs, and an order-preserving selection of row ids.  They run once
// per graph (CSR build, transpose, hub plan), are HBM-bound integer work, and are exact: the CSR they produce is
// compared bit for bit with a stable argsort (tests/test_gpu_graph.py, tests/test_gpu_golden.py).
//
//   exclusive_scan_i32   reduce-then-scan over 2048-element tiles, tile sums scanned recursively (3 levels cover 2^31)
//   radix_sort_pairs     8-bit digits; per pass: tile histograms -> exclusive scan of the digit-major [256 x tiles]
//                        table -> stable scatter (ranks inside a tile from __match_any_sync within a warp, a
//                        [warps x 256] count table across warps, and a running per-digit offset across the tile's
//                        rounds); ceil(bits / 8) passes, ping-pong between the caller's two buffer pairs
//   select_rows          per-tile counts -> scan -> every tile rewrites its selected ids in order
#pragma once

#include "common.cuh"

namespace gmlm {
namespace prim {

constexpr int kT = 256;               // threads per block
constexpr int kScanItems = 8;         // elements per thread in the scan kernels
constexpr int kScanTile = kT * kScanItems;
constexpr int kSortRounds = 16;       // rounds of 256 keys per sort tile
constexpr int kSortTile = kT * kSortRounds;
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;

// ---------------------------------------------------------------- block-wide exclusive scan of one value per thread
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums /* [kT / 32] shared */, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kT / 32; ++w) {
    const int s = warp_sums[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();                    // warp_sums may be reused by the caller's next scan
  if (total) *total = tot;
  return base + incl - v;
}

// ---------------------------------------------------------------- exclusive scan
__global__ void __launch_bounds__(kT) scan_tile_sums_kernel(const int32_t* __restrict__ in, int64_t n,
                                                            int32_t* __restrict__ sums) {
  __shared__ int warp_sums[kT / 32];
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  int total;
  block_exclusive_scan(s, warp_sums, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kT) scan_apply_kernel(const int32_t* __restrict__ in, int64_t n,
                                                        const int32_t* __restrict__ tile_offsets /* may be null */,
                                                        int32_t* __restrict__ out) {
  __shared__ int warp_sums[kT / 32];
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = base + i < n ? in[base + i] : 0;
    s += v[i];
  }
  int run = block_exclusive_scan(s, warp_sums, nullptr) + (tile_offsets ? tile_offsets[blockIdx.x] : 0);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;     // in-place safe: a thread reads its items before writing them
    run += v[i];
  }
}

inline size_t scan_temp_bytes(int64_t n) {
  size_t bytes = 0;
  while (n > kScanTile) {
    n = (n + kScanTile - 1) / kScanTile;
    bytes += (size_t(n) * sizeof(int32_t) + 255) & ~size_t(255);
  }
  return bytes + 256;
}

// out[i] = sum_{j<i} in[j]  (out may alias in)
inline int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n <= 0) return GMLM_OK;
  GMLM_REQUIRE(ws_bytes >= scan_temp_bytes(n), "scan: workspace too small");
  const int64_t tiles = (n + kScanTile - 1) / kScanTile;
  if (tiles == 1) {
    scan_apply_kernel<<<1, kT, 0, st>>>(in, n, nullptr, out);
    GMLM_LAUNCH_CHECK();
    return GMLM_OK;
  }
  int32_t* sums = static_cast<int32_t*>(ws);
  const size_t used = (size_t(tiles) * sizeof(int32_t) + 255) & ~size_t(255);
  scan_tile_sums_kernel<<<unsigned(tiles), kT, 0, st>>>(in, n, sums);
  GMLM_LAUNCH_CHECK();
  if (int rc = exclusive_scan_i32(sums, sums, tiles, static_cast<char*>(ws) + used, ws_bytes - used, st)) return rc;
  scan_apply_kernel<<<unsigned(tiles), kT, 0, st>>>(in, n, sums, out);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

// ---------------------------------------------------------------- stable LSD radix sort of (key, value) pairs
// table[d * tiles + t] = number of keys of tile t whose current digit is d
__global__ void __launch_bounds__(kT) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                        int64_t tiles, int32_t* __restrict__ table) {
  __shared__ int hist[kRadix];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = int64_t(blockIdx.x) * kSortTile;
#pragma unroll 4
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = base + r * kT + threadIdx.x;
    if (i < n) atomicAdd(&hist[(keys[i] >> shift) & (kRadix - 1)], 1);
  }
  __syncthreads();
  table[int64_t(threadIdx.x) * tiles + blockIdx.x] = hist[threadIdx.x];
}

// table holds the exclusive scan of the digit-major counts = the first output position of (digit, tile).  A tile is
// consumed in rounds of 256 consecutive keys (thread t takes key r*256 + t): the output order inside one digit is
// round-major, then warp-major, then lane-major = the input order, so the pass is stable.
__global__ void __launch_bounds__(kT) radix_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                           const int32_t* __restrict__ vals_in, int64_t n, int shift,
                                                           int64_t tiles, const int32_t* __restrict__ table,
                                                           uint32_t* __restrict__ keys_out,
                                                           int32_t* __restrict__ vals_out) {
  __shared__ int run[kRadix];                       // next output position of each digit for this tile
  __shared__ int wcount[kT / 32][kRadix];           // keys of each digit per warp in the current round
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  run[threadIdx.x] = table[int64_t(threadIdx.x) * tiles + blockIdx.x];
  const int64_t base = int64_t(blockIdx.x) * kSortTile;
  for (int r = 0; r < kSortRounds; ++r) {
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) wcount[w][threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = base + r * kT + threadIdx.x;
    const bool ok = i < n;
    uint32_t key = 0;
    int32_t val = 0;
    int digit = 0, rank_in_warp = 0;
    if (ok) {
      key = keys_in[i];
      val = vals_in[i];
      digit = int((key >> shift) & (kRadix - 1));
    }
    // lanes of this warp with a valid key of the same digit (an invalid lane matches nobody: distinct pseudo-digits)
    const unsigned peers = __match_any_sync(0xffffffffu, ok ? digit : (kRadix + lane));
    if (ok) {
      rank_in_warp = __popc(peers & ((1u << lane) - 1u));
      if (rank_in_warp == 0) wcount[warp][digit] = __popc(peers);
    }
    __syncthreads();
    // thread d turns the per-warp counts of digit d into exclusive offsets and advances the digit's cursor
    {
      int acc = run[threadIdx.x];
#pragma unroll
      for (int w = 0; w < kT / 32; ++w) {
        const int c = wcount[w][threadIdx.x];
        wcount[w][threadIdx.x] = acc;
        acc += c;
      }
      run[threadIdx.x] = acc;
    }
    __syncthreads();
    if (ok) {
      const int pos = wcount[warp][digit] + rank_in_warp;
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();                                // wcount is cleared at the top of the next round
  }
}

inline int64_t sort_tiles(int64_t n) { return (n + kSortTile - 1) / kSortTile; }

inline size_t sort_temp_bytes(int64_t n) {
  const size_t table = (size_t(sort_tiles(n)) * kRadix * sizeof(int32_t) + 255) & ~size_t(255);
  return table + scan_temp_bytes(sort_tiles(n) * kRadix) + 256;
}

// Sorts n pairs by the low `bits` bits of the key, stable.  (keys_a, vals_a) is the input and is overwritten;
// (keys_b, vals_b) receives the result.
inline int radix_sort_pairs(uint32_t* keys_a, int32_t* vals_a, uint32_t* keys_b, int32_t* vals_b, int64_t n, int bits,
                            void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n <= 0) return GMLM_OK;
  GMLM_REQUIRE(n < (int64_t(1) << 31), "sort: too many elements");
  GMLM_REQUIRE(ws_bytes >= sort_temp_bytes(n), "sort: workspace too small");
  const int64_t tiles = sort_tiles(n);
  int32_t* table = static_cast<int32_t*>(ws);
  const size_t table_bytes = (size_t(tiles) * kRadix * sizeof(int32_t) + 255) & ~size_t(255);
  char* scan_ws = static_cast<char*>(ws) + table_bytes;
  const size_t scan_bytes = ws_bytes - table_bytes;
  const int passes = (bits < 1 ? 1 : bits + kRadixBits - 1) / kRadixBits;
  uint32_t* kin = keys_a;
  int32_t* vin = vals_a;
  uint32_t* kout = keys_b;
  int32_t* vout = vals_b;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * kRadixBits;
    radix_hist_kernel<<<unsigned(tiles), kT, 0, st>>>(kin, n, shift, tiles, table);
    GMLM_LAUNCH_CHECK();
    if (int rc = exclusive_scan_i32(table, table, tiles * kRadix, scan_ws, scan_bytes, st)) return rc;
    radix_scatter_kernel<<<unsigned(tiles), kT, 0, st>>>(kin, vin, n, shift, tiles, table, kout, vout);
    GMLM_LAUNCH_CHECK();
    std::swap(kin, kout);
    std::swap(vin, vout);
  }
  if (kin != keys_b) {                               // an even number of passes left the result in the input pair
    GMLM_CUDA_TRY(cudaMemcpyAsync(keys_b, kin, size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    GMLM_CUDA_TRY(cudaMemcpyAsync(vals_b, vin, size_t(n) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  return GMLM_OK;
}

// ---------------------------------------------------------------- order-preserving selection of row ids
// selected(r) = rowptr[r+1] - rowptr[r] > thresh   (the hub rows of a CSR)
__global__ void __launch_bounds__(kT) select_count_kernel(const int32_t* __restrict__ rowptr, int64_t num_rows,
                                                          int32_t thresh, int32_t* __restrict__ counts) {
  __shared__ int warp_sums[kT / 32];
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int64_t r = base + i;
    if (r < num_rows && rowptr[r + 1] - rowptr[r] > thresh) ++s;
  }
  int total;
  block_exclusive_scan(s, warp_sums, &total);
  if (threadIdx.x == 0) counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kT) select_write_kernel(const int32_t* __restrict__ rowptr, int64_t num_rows,
                                                          int32_t thresh, const int32_t* __restrict__ offsets,
                                                          int64_t max_out, int32_t* __restrict__ out) {
  __shared__ int warp_sums[kT / 32];
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
  bool sel[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int64_t r = base + i;
    sel[i] = r < num_rows && rowptr[r + 1] - rowptr[r] > thresh;
    s += sel[i] ? 1 : 0;
  }
  int64_t pos = int64_t(block_exclusive_scan(s, warp_sums, nullptr)) + offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (sel[i]) {
      if (pos < max_out) out[pos] = int32_t(base + i);
      ++pos;
    }
  }
}

inline size_t select_temp_bytes(int64_t num_rows) {
  const int64_t tiles = (num_rows + kScanTile - 1) / kScanTile;
  return ((size_t(tiles) * sizeof(int32_t) + 255) & ~size_t(255)) + scan_temp_bytes(tiles) + 256;
}

// out[0..] = ascending ids of the rows longer than thresh (at most max_out are written)
inline int select_rows(const int32_t* rowptr, int64_t num_rows, int32_t thresh, int32_t* out, int64_t max_out, void* ws,
                       size_t ws_bytes, cudaStream_t st) {
  if (num_rows <= 0) return GMLM_OK;
  GMLM_REQUIRE(ws_bytes >= select_temp_bytes(num_rows), "select: workspace too small");
  const int64_t tiles = (num_rows + kScanTile - 1) / kScanTile;
  int32_t* counts = static_cast<int32_t*>(ws);
  const size_t used = (size_t(tiles) * sizeof(int32_t) + 255) & ~size_t(255);
  select_count_kernel<<<unsigned(tiles), kT, 0, st>>>(rowptr, num_rows, thresh, counts);
  GMLM_LAUNCH_CHECK();
  if (int rc = exclusive_scan_i32(counts, counts, tiles, static_cast<char*>(ws) + used, ws_bytes - used, st)) return rc;
  select_write_kernel<<<unsigned(tiles), kT, 0, st>>>(rowptr, num_rows, thresh, counts, max_out, out);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

}  // namespace prim
}  // namespace gmlm
