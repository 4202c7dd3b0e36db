// gat.cu — edge scores for the two aggregation variants north_star names beside the RGCN mean
// (extensions with no counterpart in /root/reference — their oracle is
// the restatement of upstream GCNConv / GATConv in oracle/pyg_ref.py):
//
//   gcn_edge_weights : w_e = deg[src]^-1/2 * deg[dst]^-1/2      (deg = in-degree incl. self-loops)
//   gat_alpha_fwd    : alpha[e,h] = softmax over the in-edges of dst of leaky_relu(a_src[src,h] + a_dst[dst,h])
//   gat_alpha_bwd    : gradients of the scores through softmax + leaky_relu
//   segment_sum_f32  : out[r,h] = sum_{i in row r} vals[idx[i], h]   (scatter of edge scalars as a gather)
//
// The feature aggregation itself (out_i = sum_e alpha_e z[src_e]) is gmlm_spmm_csr in weighted
// mode with one weight column per head; these kernels only produce its per-edge scalars.
// One warp per destination row, lanes stride over the row's edges, warp-shuffle reductions in
// a fixed order (deterministic).  Scalars are fp32; PyG's softmax epsilon (1e-16) is kept.
#include "common.cuh"

namespace gmlm {
namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

__global__ void __launch_bounds__(256) deg_from_rowptr_rsqrt_kernel(const int32_t* __restrict__ rowptr, int64_t n,
                                                                    float* __restrict__ dis) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t d = rowptr[i + 1] - rowptr[i];
  dis[i] = d > 0 ? 1.0f / sqrtf(float(d)) : 0.f;   // deg^-1/2 with inf -> 0, as upstream gcn_norm
}

// CSR keyed on dst: w[e] = dis[col[e]] * dis[row]
__global__ void __launch_bounds__(256) gcn_weights_kernel(const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ col, int64_t num_rows,
                                                          const float* __restrict__ dis, float* __restrict__ w) {
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= num_rows) return;
  const int32_t b = rowptr[r], e = rowptr[r + 1];
  const float dr = dis[r];
  for (int32_t k = b + lane; k < e; k += 32) w[k] = __fmul_rn(__fmul_rn(dis[col[k]], 1.0f), dr);
}

__global__ void __launch_bounds__(256) gat_alpha_fwd_kernel(const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ col, int64_t num_rows,
                                                            const float* __restrict__ a_src,
                                                            const float* __restrict__ a_dst, int H, float slope,
                                                            float* __restrict__ alpha) {
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= num_rows) return;
  const int32_t b = rowptr[r], e = rowptr[r + 1];
  if (e == b) return;
  for (int h = 0; h < H; ++h) {
    const float ad = a_dst[r * H + h];
    float m = -INFINITY;
    for (int32_t k = b + lane; k < e; k += 32) m = fmaxf(m, lrelu(a_src[int64_t(col[k]) * H + h] + ad, slope));
    m = warp_max(m);
    float l = 0.f;
    for (int32_t k = b + lane; k < e; k += 32) l += expf(lrelu(a_src[int64_t(col[k]) * H + h] + ad, slope) - m);
    l = warp_sum(l) + 1e-16f;
    for (int32_t k = b + lane; k < e; k += 32)
      alpha[int64_t(k) * H + h] = expf(lrelu(a_src[int64_t(col[k]) * H + h] + ad, slope) - m) / l;
  }
}

// d_alpha[e,h] = <g[dst,h,:], z[src,h,:]>;  d_e = alpha (d_alpha - sum_e alpha d_alpha);
// d_score = d_e * lrelu'(raw);  da_dst[dst,h] = sum_e d_score
template <typename T>
__global__ void __launch_bounds__(256) gat_alpha_bwd_kernel(const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ col, int64_t num_rows,
                                                            const T* __restrict__ z, int64_t ldz,
                                                            const T* __restrict__ g, int64_t ldg, int H, int C,
                                                            const float* __restrict__ a_src,
                                                            const float* __restrict__ a_dst,
                                                            const float* __restrict__ alpha, float slope,
                                                            float* __restrict__ d_score, float* __restrict__ da_dst) {
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= num_rows) return;
  const int32_t b = rowptr[r], e = rowptr[r + 1];
  for (int h = 0; h < H; ++h) {
    if (e == b) { if (lane == 0) da_dst[r * H + h] = 0.f; continue; }
    const T* gr = g + r * ldg + int64_t(h) * C;
    // per-edge dot products, lanes over the C features of this head
    for (int32_t k = b; k < e; ++k) {
      const T* zr = z + int64_t(col[k]) * ldz + int64_t(h) * C;
      float d = 0.f;
      for (int c = lane; c < C; c += 32) d = fmaf(to_float(gr[c]), to_float(zr[c]), d);
      d = warp_sum(d);
      if (lane == 0) d_score[int64_t(k) * H + h] = d;   // d_alpha for now
    }
    __syncwarp();
    float t = 0.f;
    for (int32_t k = b + lane; k < e; k += 32) t = fmaf(alpha[int64_t(k) * H + h], d_score[int64_t(k) * H + h], t);
    t = warp_sum(t);
    const float ad = a_dst[r * H + h];
    float s = 0.f;
    for (int32_t k = b + lane; k < e; k += 32) {
      const float al = alpha[int64_t(k) * H + h];
      const float de = al * (d_score[int64_t(k) * H + h] - t);
      const float raw = a_src[int64_t(col[k]) * H + h] + ad;
      const float ds = raw > 0.f ? de : de * slope;
      d_score[int64_t(k) * H + h] = ds;
      s += ds;
    }
    s = warp_sum(s);
    if (lane == 0) da_dst[r * H + h] = s;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) segment_sum_kernel(const float* __restrict__ vals,
                                                          const int64_t* __restrict__ idx,
                                                          const int32_t* __restrict__ rowptr, int64_t num_rows, int H,
                                                          float* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= num_rows * H) return;
  const int64_t r = i / H;
  const int h = int(i - r * H);
  float a = 0.f;
  for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) a += vals[(idx ? idx[k] : int64_t(k)) * H + h];
  out[i] = a;
}

}  // namespace
}  // namespace gmlm

using namespace gmlm;

extern "C" int gmlm_gcn_edge_weights(const int32_t* rowptr, const int32_t* col, int64_t num_rows, float* dis_ws,
                                     float* w, void* stream) {
  GMLM_REQUIRE(num_rows >= 0, "gcn_edge_weights: bad sizes");
  if (num_rows == 0) return GMLM_OK;
  GMLM_REQUIRE(rowptr && col && dis_ws && w, "gcn_edge_weights: null pointer");
  cudaStream_t st = as_stream(stream);
  deg_from_rowptr_rsqrt_kernel<<<unsigned((num_rows + 255) / 256), 256, 0, st>>>(rowptr, num_rows, dis_ws);
  GMLM_LAUNCH_CHECK();
  gcn_weights_kernel<<<unsigned((num_rows * 32 + 255) / 256), 256, 0, st>>>(rowptr, col, num_rows, dis_ws, w);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

extern "C" int gmlm_gat_alpha_fwd(const int32_t* rowptr, const int32_t* col, int64_t num_rows, const float* a_src,
                                  const float* a_dst, int heads, float negative_slope, float* alpha, void* stream) {
  GMLM_REQUIRE(num_rows >= 0 && heads >= 1, "gat_alpha_fwd: bad sizes");
  if (num_rows == 0) return GMLM_OK;
  GMLM_REQUIRE(rowptr && col && a_src && a_dst && alpha, "gat_alpha_fwd: null pointer");
  gat_alpha_fwd_kernel<<<unsigned((num_rows * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
      rowptr, col, num_rows, a_src, a_dst, heads, negative_slope, alpha);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

extern "C" int gmlm_gat_alpha_bwd(const int32_t* rowptr, const int32_t* col, int64_t num_rows, const void* z,
                                  int64_t ldz, const void* g, int64_t ldg, int dtype, int heads, int head_dim,
                                  const float* a_src, const float* a_dst, const float* alpha, float negative_slope,
                                  float* d_score, float* da_dst, void* stream) {
  GMLM_REQUIRE(dtype == GMLM_F32 || dtype == GMLM_BF16, "gat_alpha_bwd: dtype must be GMLM_F32 or GMLM_BF16");
  GMLM_REQUIRE(num_rows >= 0 && heads >= 1 && head_dim >= 1, "gat_alpha_bwd: bad sizes");
  if (num_rows == 0) return GMLM_OK;
  GMLM_REQUIRE(rowptr && col && z && g && a_src && a_dst && alpha && d_score && da_dst, "gat_alpha_bwd: null pointer");
  const unsigned blocks = unsigned((num_rows * 32 + 255) / 256);
  cudaStream_t st = as_stream(stream);
  if (dtype == GMLM_F32)
    gat_alpha_bwd_kernel<float><<<blocks, 256, 0, st>>>(rowptr, col, num_rows, static_cast<const float*>(z), ldz,
                                                        static_cast<const float*>(g), ldg, heads, head_dim, a_src,
                                                        a_dst, alpha, negative_slope, d_score, da_dst);
  else
    gat_alpha_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        rowptr, col, num_rows, static_cast<const __nv_bfloat16*>(z), ldz, static_cast<const __nv_bfloat16*>(g), ldg,
        heads, head_dim, a_src, a_dst, alpha, negative_slope, d_score, da_dst);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

extern "C" int gmlm_segment_sum_f32(const float* vals, const int64_t* idx, const int32_t* rowptr, int64_t num_rows,
                                    int heads, float* out, void* stream) {
  GMLM_REQUIRE(num_rows >= 0 && heads >= 1, "segment_sum: bad sizes");
  if (num_rows == 0) return GMLM_OK;
  GMLM_REQUIRE(vals && rowptr && out, "segment_sum: null pointer");
  segment_sum_kernel<<<unsigned((num_rows * heads + 255) / 256), 256, 0, as_stream(stream)>>>(vals, idx, rowptr,
                                                                                              num_rows, heads, out);
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}
