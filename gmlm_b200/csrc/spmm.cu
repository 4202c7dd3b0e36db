// spmm.cu — the message-passing aggregation kernel.
//
//   out[r, :] = reduce_{e in [rowptr[r], rowptr[r+1])} w[e] * x[col[e], :]
//
// Forward (RGCN mean, [PyG] RGCNConv.propagate, called main.py:272,285,298,308): rows are
// (dst,rel) segments, col = source node, reduce = mean.  Backward (autograd of the same):
// rows are source nodes of the transposed CSR, col = forward segment, w = 1/|segment|.
//
// Design (B200, HBM-bound; roofline = E*F*b gather + 4E index + N*F*b write):
//   * A group of LPR lanes owns LPR consecutive CSR rows; each lane owns CH 16-byte packs of
//     the feature row, so one gathered row is LPR coalesced 128-bit loads.  Groups smaller
//     than a warp (LPR = 8/16) serve narrow feature rows without idle lanes.
//   * "flat" variant: the group walks the *concatenated* edge list of its rows.  Indices are
//     fetched LPR at a time with one coalesced streaming load and broadcast by shuffle, U
//     feature-row loads are issued back-to-back before any is consumed (memory-level
//     parallelism independent of row length), and the fp32 accumulator is flushed whenever
//     the edge cursor crosses a row end (a segmented reduction in CSR order).
//   * Determinism: fp32 accumulation strictly in CSR order (stable-sorted = original edge
//     order).  Rows longer than `hub_thresh` are skipped here and reduced by the hub path:
//     fixed chunks of hub_thresh edges -> fp32 partials -> in-order final sum.  No atomics.
#include "common.cuh"

namespace gmlm {

int tuning_spmm_variant();
int tuning_spmm_unroll();

namespace {

struct SpmmParams {
  const void* x;
  int64_t ldx;
  int64_t feat;
  const int32_t* rowptr;     // [num_rows+1]            (direct mode)
  const int32_t* row_beg;    // [num_rows] chunk begins (partial mode) or nullptr
  const int32_t* row_end;    // [num_rows] chunk ends   (partial mode) or nullptr
  const int32_t* col;
  const float* w;
  int64_t num_rows;
  int mean;                  // 1: divide by row length
  int flat;                  // 1: flat segmented walk, 0: row-by-row
  int32_t hub_thresh;        // rows longer than this are left to the hub path (direct mode)
  void* out;                 // T [num_rows, feat] (direct) / float [num_rows, feat] (partial)
  int64_t ldo;
};

template <typename T, int VEC, int CH, int LPR, int U, bool WEIGHTED>
__device__ __forceinline__ void accumulate_range(const T* __restrict__ xf, int64_t ldx, const bool (&fvalid)[CH],
                                                 const int32_t* __restrict__ col, const float* __restrict__ w,
                                                 int beg, int end, unsigned gmask, int gl, float (&acc)[CH][VEC]) {
  for (int base = beg; base < end; base += LPR) {
    const int idx = base + gl;
    const int my_col = idx < end ? ld_stream(col + idx) : 0;
    float my_w = 0.f;
    if (WEIGHTED) my_w = idx < end ? ld_stream(w + idx) : 0.f;
    const int nb = min(LPR, end - base);
    for (int j0 = 0; j0 < nb; j0 += U) {
      Pack<T, VEC> v[U][CH];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j0 + u;
        const int c = __shfl_sync(gmask, my_col, jj, LPR);
        if (jj < nb) {
          const T* rowp = xf + int64_t(c) * ldx;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
            if (fvalid[ch]) v[u][ch].load(rowp + ch * LPR * VEC);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = j0 + u;
        float wgt = 1.f;
        if (WEIGHTED) wgt = __shfl_sync(gmask, my_w, jj, LPR);
        if (jj < nb) {
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            if (fvalid[ch]) {
              float f[VEC];
              v[u][ch].unpack(f);
#pragma unroll
              for (int k = 0; k < VEC; ++k) acc[ch][k] = WEIGHTED ? fmaf(wgt, f[k], acc[ch][k]) : acc[ch][k] + f[k];
            }
          }
        }
      }
    }
  }
}

template <typename T, int VEC, int CH, int LPR>
__device__ __forceinline__ void flush_row(T* __restrict__ dst, const bool (&fvalid)[CH], float (&acc)[CH][VEC],
                                          int len, int mean) {
  const float scale = (mean && len > 1) ? 1.0f / float(len) : 1.0f;
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    if (fvalid[ch]) {
      float f[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) f[k] = acc[ch][k] * scale;
      Pack<T, VEC> p;
      p.pack(f);
      p.store(dst + ch * LPR * VEC);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[ch][k] = 0.f;
  }
}

template <int VEC, int CH, int LPR>
__device__ __forceinline__ void flush_partial(float* __restrict__ dst, const bool (&fvalid)[CH],
                                              float (&acc)[CH][VEC]) {
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    if (fvalid[ch]) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) dst[ch * LPR * VEC + k] = acc[ch][k];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[ch][k] = 0.f;
  }
}

template <typename T, int VEC, int CH, int LPR, bool WEIGHTED>
__global__ void __launch_bounds__(256) spmm_kernel(SpmmParams p) {
  constexpr int GROUPS = 32 / LPR;
  constexpr int U = (CH >= 4) ? 2 : (CH >= 2 ? 4 : 8);
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const int g = lane / LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (g * LPR));
  const int64_t group_id = (int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GROUPS + g;
  const int64_t r0 = group_id * LPR;
  if (r0 >= p.num_rows) return;  // the whole group leaves together
  const int nr = int(min(int64_t(LPR), p.num_rows - r0));

  const int64_t f0 = int64_t(blockIdx.y) * (LPR * VEC * CH) + gl * VEC;
  bool fvalid[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) fvalid[ch] = f0 + ch * LPR * VEC < p.feat;
  const T* __restrict__ xf = static_cast<const T*>(p.x) + f0;
  const bool partial = p.row_beg != nullptr;

  // row extents: lane gl holds [my_beg, my_end) of row r0+gl (empty beyond nr)
  int my_beg, my_end;
  {
    const int64_t r = r0 + min(gl, nr - 1);
    if (partial) {
      my_beg = __ldg(p.row_beg + r);
      my_end = __ldg(p.row_end + r);
    } else {
      my_beg = __ldg(p.rowptr + r);
      my_end = __ldg(p.rowptr + r + 1);
    }
    if (gl >= nr) my_beg = my_end;
  }
  const int my_len = my_end - my_beg;
  const bool has_hub = !partial && (__ballot_sync(gmask, my_len > p.hub_thresh) != 0u);

  float acc[CH][VEC];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[ch][k] = 0.f;

  if (p.flat && !partial && !has_hub) {
    // ---- flat segmented walk over the concatenated edge list of rows r0 .. r0+nr-1
    T* __restrict__ outf = static_cast<T*>(p.out) + r0 * p.ldo + f0;
    const int e0 = __shfl_sync(gmask, my_beg, 0, LPR);
    const int e1 = __shfl_sync(gmask, my_end, nr - 1, LPR);
    int cur = 0;
    int cur_beg = e0;
    int cur_end = __shfl_sync(gmask, my_end, 0, LPR);
    for (int base = e0; base < e1; base += LPR) {
      const int idx = base + gl;
      const int my_col = idx < e1 ? ld_stream(p.col + idx) : 0;
      float my_w = 0.f;
      if (WEIGHTED) my_w = idx < e1 ? ld_stream(p.w + idx) : 0.f;
      const int nb = min(LPR, e1 - base);
      for (int j0 = 0; j0 < nb; j0 += U) {
        Pack<T, VEC> v[U][CH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = j0 + u;
          const int c = __shfl_sync(gmask, my_col, jj, LPR);
          if (jj < nb) {
            const T* rowp = xf + int64_t(c) * p.ldx;
#pragma unroll
            for (int ch = 0; ch < CH; ++ch)
              if (fvalid[ch]) v[u][ch].load(rowp + ch * LPR * VEC);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = j0 + u;
          float wgt = 1.f;
          if (WEIGHTED) wgt = __shfl_sync(gmask, my_w, jj, LPR);
          if (jj < nb) {
            const int ee = base + jj;
            while (ee >= cur_end) {  // crossed one (or several empty) row ends: flush in order
              flush_row<T, VEC, CH, LPR>(outf + int64_t(cur) * p.ldo, fvalid, acc, cur_end - cur_beg, p.mean);
              cur_beg = cur_end;
              ++cur;
              cur_end = __shfl_sync(gmask, my_end, cur, LPR);
            }
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
              if (fvalid[ch]) {
                float f[VEC];
                v[u][ch].unpack(f);
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                  acc[ch][k] = WEIGHTED ? fmaf(wgt, f[k], acc[ch][k]) : acc[ch][k] + f[k];
              }
            }
          }
        }
      }
    }
    while (cur < nr) {  // last row with edges, then trailing empty rows
      flush_row<T, VEC, CH, LPR>(outf + int64_t(cur) * p.ldo, fvalid, acc, cur_end - cur_beg, p.mean);
      cur_beg = cur_end;
      ++cur;
      if (cur < nr) cur_end = __shfl_sync(gmask, my_end, cur, LPR);
    }
    return;
  }

  // ---- row-by-row path (variant 0, groups that contain a hub row, and hub partials)
  for (int i = 0; i < nr; ++i) {
    const int beg = __shfl_sync(gmask, my_beg, i, LPR);
    const int end = __shfl_sync(gmask, my_end, i, LPR);
    const int len = end - beg;
    if (!partial && len > p.hub_thresh) continue;  // written by the hub path
    accumulate_range<T, VEC, CH, LPR, U, WEIGHTED>(xf, p.ldx, fvalid, p.col, p.w, beg, end, gmask, gl, acc);
    if (partial)
      flush_partial<VEC, CH, LPR>(static_cast<float*>(p.out) + (r0 + i) * p.ldo + f0, fvalid, acc);
    else
      flush_row<T, VEC, CH, LPR>(static_cast<T*>(p.out) + (r0 + i) * p.ldo + f0, fvalid, acc, len, p.mean);
  }
}

// final in-order reduction of the hub partials: one warp per hub row, scalar feature loop
template <typename T>
__global__ void __launch_bounds__(256) hub_final_kernel(const float* __restrict__ ws, int64_t feat,
                                                        const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ hub_row,
                                                        const int32_t* __restrict__ hub_chunk_ptr, int64_t n_hub,
                                                        int mean, T* __restrict__ out, int64_t ldo) {
  const int64_t h = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (h >= n_hub) return;
  const int32_t r = hub_row[h];
  const int32_t c0 = hub_chunk_ptr[h], c1 = hub_chunk_ptr[h + 1];
  const int len = rowptr[r + 1] - rowptr[r];
  const float scale = (mean && len > 1) ? 1.0f / float(len) : 1.0f;
  for (int64_t f = lane; f < feat; f += 32) {
    float a = 0.f;
    for (int32_t c = c0; c < c1; ++c) a += ws[int64_t(c) * feat + f];
    out[int64_t(r) * ldo + f] = from_float<T>(a * scale);
  }
}

template <typename T, int VEC, int CH, int LPR>
int launch_geo(const SpmmParams& p, bool weighted, cudaStream_t st) {
  constexpr int GROUPS = 32 / LPR;
  const int64_t rows_per_cta = int64_t(256 / 32) * GROUPS * LPR;  // = 256
  const int64_t gx = (p.num_rows + rows_per_cta - 1) / rows_per_cta;
  const int64_t slab = int64_t(LPR) * VEC * CH;
  const int64_t gy = (p.feat + slab - 1) / slab;
  GMLM_REQUIRE(gx <= 0x7fffffffLL && gy <= 65535, "spmm: grid too large");
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (weighted) spmm_kernel<T, VEC, CH, LPR, true><<<grid, 256, 0, st>>>(p);
  else spmm_kernel<T, VEC, CH, LPR, false><<<grid, 256, 0, st>>>(p);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

template <typename T, int VEC>
int launch_vec(const SpmmParams& p, bool weighted, cudaStream_t st) {
  const int64_t nvec = (p.feat + VEC - 1) / VEC;
  if (nvec <= 8) return launch_geo<T, VEC, 1, 8>(p, weighted, st);
  if (nvec <= 16) return launch_geo<T, VEC, 1, 16>(p, weighted, st);
  if (nvec <= 32) return launch_geo<T, VEC, 1, 32>(p, weighted, st);
  if (nvec <= 64) return launch_geo<T, VEC, 2, 32>(p, weighted, st);
  if (nvec <= 96) return launch_geo<T, VEC, 3, 32>(p, weighted, st);
  return launch_geo<T, VEC, 4, 32>(p, weighted, st);
}

template <typename T, int FULLVEC>
int launch_typed(const SpmmParams& p, bool weighted, bool aligned, cudaStream_t st) {
  if (aligned) return launch_vec<T, FULLVEC>(p, weighted, st);
  return launch_vec<T, 1>(p, weighted, st);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace gmlm

using namespace gmlm;

extern "C" int gmlm_spmm_csr(const void* x, int dtype, int64_t feat, int64_t ldx, const int32_t* rowptr,
                             const int32_t* col, const float* w, int64_t num_rows, int mode, int32_t hub_thresh,
                             int64_t n_hub, int64_t n_chunks, const int32_t* hub_row, const int32_t* hub_chunk_ptr,
                             const int32_t* chunk_beg, const int32_t* chunk_end, float* hub_ws, void* out,
                             int64_t ldo, void* stream) {
  GMLM_REQUIRE(dtype == GMLM_F32 || dtype == GMLM_BF16, "spmm: dtype must be GMLM_F32 or GMLM_BF16");
  GMLM_REQUIRE(mode == GMLM_AGG_SUM || mode == GMLM_AGG_MEAN || mode == GMLM_AGG_WEIGHTED, "spmm: bad mode");
  GMLM_REQUIRE(feat >= 0 && num_rows >= 0 && ldx >= feat && ldo >= feat, "spmm: bad sizes");
  GMLM_REQUIRE(mode != GMLM_AGG_WEIGHTED || w != nullptr, "spmm: weighted mode needs w");
  GMLM_REQUIRE(n_hub >= 0 && n_chunks >= n_hub, "spmm: bad hub plan");
  if (num_rows == 0 || feat == 0) return GMLM_OK;
  GMLM_REQUIRE(x && rowptr && out, "spmm: null pointer");
  GMLM_REQUIRE(n_hub == 0 || (hub_row && hub_chunk_ptr && chunk_beg && chunk_end && hub_ws && hub_thresh >= 1),
               "spmm: hub plan arrays missing");
  cudaStream_t st = as_stream(stream);
  const bool weighted = mode == GMLM_AGG_WEIGHTED;
  const int esz = dtype == GMLM_F32 ? 4 : 2;
  const int fullvec = 16 / esz;
  const bool aligned = aligned16(x) && aligned16(out) && feat % fullvec == 0 && ldx % fullvec == 0 &&
                       ldo % fullvec == 0;

  SpmmParams p;
  p.x = x; p.ldx = ldx; p.feat = feat;
  p.rowptr = rowptr; p.row_beg = nullptr; p.row_end = nullptr;
  p.col = col; p.w = w; p.num_rows = num_rows;
  p.mean = mode == GMLM_AGG_MEAN;
  p.flat = tuning_spmm_variant() != 0;
  p.hub_thresh = n_hub > 0 ? hub_thresh : 0x7fffffff;
  p.out = out; p.ldo = ldo;
  int rc = dtype == GMLM_F32 ? launch_typed<float, 4>(p, weighted, aligned, st)
                             : launch_typed<__nv_bfloat16, 8>(p, weighted, aligned, st);
  if (rc) return rc;
  if (n_hub == 0) return GMLM_OK;

  // hub path: chunk partials (fp32) then the in-order final sum
  SpmmParams q = p;
  q.rowptr = nullptr; q.row_beg = chunk_beg; q.row_end = chunk_end;
  q.num_rows = n_chunks; q.mean = 0; q.flat = 0; q.hub_thresh = 0x7fffffff;
  q.out = hub_ws; q.ldo = feat;
  rc = dtype == GMLM_F32 ? launch_typed<float, 4>(q, weighted, aligned, st)
                         : launch_typed<__nv_bfloat16, 8>(q, weighted, aligned, st);
  if (rc) return rc;
  const int64_t threads = n_hub * 32;
  const unsigned blocks = unsigned((threads + 255) / 256);
  if (dtype == GMLM_F32)
    hub_final_kernel<float><<<blocks, 256, 0, st>>>(hub_ws, feat, rowptr, hub_row, hub_chunk_ptr, n_hub, p.mean,
                                                    static_cast<float*>(out), ldo);
  else
    hub_final_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(hub_ws, feat, rowptr, hub_row, hub_chunk_ptr, n_hub,
                                                            p.mean, static_cast<__nv_bfloat16*>(out), ldo);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}
