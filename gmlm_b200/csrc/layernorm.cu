// layernorm.cu — row-wise LayerNorm forward/backward for the tail of the GNN encoder:
// MultiScaleFusion's `self.layer_norm(fused)` (main.py:171,180),
// the last op inside get_graph_embeddings (main.py:320).
//
// HBM-bound streaming over an [N, C] matrix (forward 2·N·C·b, backward 3·N·C·b): one warp owns a
// row, a lane holds NP 16-byte packs of it in registers, so the row is read exactly once per pass
// (mean and the centred variance both come from the register copy: the numerically safe two-pass
// form at one-pass traffic).  Parameter gradients accumulate per lane in registers over the rows a
// warp walks, are written as per-warp partials and summed in fp64 in a fixed order by a second
// kernel: no atomics, bit-identical run to run.
#include <algorithm>

#include "common.cuh"

namespace gmlm {
namespace {

constexpr int kCta = 256;
constexpr int kWarps = kCta / 32;
constexpr int kMaxC = 1024;   // widest row: 32 lanes x 4 packs x 8 bf16 (x 8 packs x 4 fp32)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// y = (x - mu) * rstd * gamma + beta ; mu, var over the row (biased variance, like nn.LayerNorm)
template <typename T, int VEC, int NP>
__global__ void __launch_bounds__(kCta) layernorm_fwd_kernel(const T* __restrict__ x, int64_t num_rows, int64_t C,
                                                             int64_t ldx, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             T* __restrict__ y, int64_t ldy,
                                                             float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = int64_t(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const int64_t n_warps = int64_t(gridDim.x) * kWarps;
  const int packs = int(C / VEC);
  const float inv_c = 1.0f / float(C);
  // per-channel parameters live in shared memory (registers are what bounds the resident warps here)
  __shared__ __align__(16) float sgam[kMaxC], sbet[kMaxC];
  for (int c = threadIdx.x; c < C; c += kCta) { sgam[c] = gamma[c]; sbet[c] = beta[c]; }
  __syncthreads();
  for (int64_t r = warp; r < num_rows; r += n_warps) {
    Pack<T, VEC> p[NP];
    float f[NP][VEC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int pk = i * 32 + lane;
      if (pk < packs) p[i].load(x + r * ldx + int64_t(pk) * VEC); else p[i].zero();
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      p[i].unpack(f[i]);
#pragma unroll
      for (int k = 0; k < VEC; ++k) s += f[i][k];
    }
    const float mu = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      if (i * 32 + lane < packs) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) { const float d = f[i][k] - mu; q = fmaf(d, d, q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int pk = i * 32 + lane;
      if (pk < packs) {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          f[i][k] = fmaf((f[i][k] - mu) * rstd, sgam[pk * VEC + k], sbet[pk * VEC + k]);
        p[i].pack(f[i]);
        p[i].store(y + r * ldy + int64_t(pk) * VEC);
      }
    }
    if (lane == 0) { mean_out[r] = mu; rstd_out[r] = rstd; }
  }
}

// gx = rstd * (g*gamma - mean_c(g*gamma) - xhat * mean_c(g*gamma*xhat));  partial[w] = per-warp sums of
// g*xhat (d gamma) and g (d beta) over the rows warp w walked
template <typename T, int VEC, int NP>
__global__ void __launch_bounds__(kCta, 2) layernorm_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy,
                                                             int64_t num_rows, int64_t C, int64_t ldx, int64_t ldg,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, T* __restrict__ gx,
                                                             int64_t ldgx, float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = int64_t(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const int64_t n_warps = int64_t(gridDim.x) * kWarps;
  const int packs = int(C / VEC);
  const float inv_c = 1.0f / float(C);
  __shared__ __align__(16) float sgam[kMaxC];
  for (int c = threadIdx.x; c < kMaxC; c += kCta) sgam[c] = c < C ? gamma[c] : 0.f;
  __syncthreads();
  float dgam[NP][VEC], dbet[NP][VEC];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) dgam[i][k] = dbet[i][k] = 0.f;
  }
  for (int64_t r = warp; r < num_rows; r += n_warps) {
    Pack<T, VEC> px[NP], pg[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int pk = i * 32 + lane;
      if (pk < packs) {
        px[i].load(x + r * ldx + int64_t(pk) * VEC);
        pg[i].load(gy + r * ldg + int64_t(pk) * VEC);
      } else {
        px[i].zero();
        pg[i].zero();
      }
    }
    const float mu = mean[r], rs = rstd[r];
    float xh[NP][VEC], g[NP][VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      px[i].unpack(xh[i]);
      pg[i].unpack(g[i]);
      const int pk = i * 32 + lane;
      const bool live = pk < packs;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        xh[i][k] = live ? (xh[i][k] - mu) * rs : 0.f;
        dbet[i][k] += g[i][k];
        dgam[i][k] = fmaf(g[i][k], xh[i][k], dgam[i][k]);
        g[i][k] *= sgam[(pk * VEC + k) & (kMaxC - 1)];
        s1 += g[i][k];
        s2 = fmaf(g[i][k], xh[i][k], s2);
      }
    }
    s1 = warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
    if (gx != nullptr) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int pk = i * 32 + lane;
        if (pk < packs) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) g[i][k] = rs * (g[i][k] - s1 - xh[i][k] * s2);
          pg[i].pack(g[i]);
          pg[i].store(gx + r * ldgx + int64_t(pk) * VEC);
        }
      }
    }
  }
  {
    float* row = partial + warp * 2 * C;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int pk = i * 32 + lane;
      if (pk < packs) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          row[pk * VEC + k] = dgam[i][k];
          row[C + pk * VEC + k] = dbet[i][k];
        }
      }
    }
  }
}

// 32 columns x 8 row slices per CTA; slice y sums partial rows y, y+8, ... in fp64, the slices are added in
// a fixed order
__global__ void __launch_bounds__(256) layernorm_param_grads_kernel(const float* __restrict__ partial,
                                                                    int64_t n_partials, int64_t C,
                                                                    float* __restrict__ g_gamma,
                                                                    float* __restrict__ g_beta) {
  __shared__ double sh[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = int64_t(blockIdx.x) * 32 + tx;
  double a = 0.0;
  if (c < 2 * C)
    for (int64_t w = ty; w < n_partials; w += 8) a += double(partial[w * 2 * C + c]);
  sh[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && c < 2 * C) {
    double t = 0.0;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += sh[y][tx];
    if (c < C) { if (g_gamma) g_gamma[c] = float(t); }
    else if (g_beta) g_beta[c - C] = float(t);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <auto Kernel>
int resident_ctas() {
  static int cached = 0;
  if (cached == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, Kernel, kCta, 0) != cudaSuccess || occ < 1) occ = 1;
    cached = occ * num_sms();
  }
  return cached;
}

constexpr int kMaxCtas = 148 * 8;

// calls fn.template operator()<T, VEC, NP>() with the smallest NP that covers the row (<= kMaxC channels)
template <typename Fn>
int dispatch(int dtype, int64_t C, Fn&& fn) {
  const int vec = dtype == GMLM_F32 ? 4 : 8;
  const int np = int((C / vec + 31) / 32);
  if (dtype == GMLM_F32) {
    if (np <= 1) return fn.template operator()<float, 4, 1>();
    if (np <= 2) return fn.template operator()<float, 4, 2>();
    if (np <= 4) return fn.template operator()<float, 4, 4>();
    if (np <= 6) return fn.template operator()<float, 4, 6>();
    return fn.template operator()<float, 4, 8>();
  }
  if (np <= 1) return fn.template operator()<__nv_bfloat16, 8, 1>();
  if (np <= 2) return fn.template operator()<__nv_bfloat16, 8, 2>();
  if (np <= 3) return fn.template operator()<__nv_bfloat16, 8, 3>();
  return fn.template operator()<__nv_bfloat16, 8, 4>();
}

int check_shape(const char* what, int dtype, int64_t N, int64_t C, std::initializer_list<const void*> ptrs,
                std::initializer_list<int64_t> lds) {
  if (dtype != GMLM_F32 && dtype != GMLM_BF16) return fail(GMLM_ERR_INVALID, "%s: dtype must be GMLM_F32 or GMLM_BF16", what);
  const int vec = dtype == GMLM_F32 ? 4 : 8;
  if (N < 0 || C < 1 || C % vec != 0 || C > gmlm_layernorm_max_channels(dtype))
    return fail(GMLM_ERR_INVALID, "%s: channels must be a multiple of %d and <= %lld (got %lld)", what, vec,
                (long long)gmlm_layernorm_max_channels(dtype), (long long)C);
  for (const void* p : ptrs)
    if (!aligned16(p)) return fail(GMLM_ERR_INVALID, "%s: matrices must be 16-byte aligned", what);
  for (int64_t l : lds)
    if (l < C || l % vec != 0) return fail(GMLM_ERR_INVALID, "%s: leading dimension must be >= channels and a multiple of %d", what, vec);
  return GMLM_OK;
}

}  // namespace
}  // namespace gmlm

using namespace gmlm;

extern "C" {

int64_t gmlm_layernorm_max_channels(int /*dtype*/) { return kMaxC; }

int gmlm_layernorm_fwd(const void* x, int dtype, int64_t N, int64_t C, int64_t ldx, const float* gamma,
                       const float* beta, float eps, void* y, int64_t ldy, float* mean_out, float* rstd_out,
                       void* stream) {
  if (int rc = check_shape("layernorm_fwd", dtype, N, C, {x, y}, {ldx, ldy})) return rc;
  if (N == 0) return GMLM_OK;
  GMLM_REQUIRE(x && y && gamma && beta && mean_out && rstd_out, "layernorm_fwd: null pointer");
  cudaStream_t st = as_stream(stream);
  return dispatch(dtype, C, [&]<typename T, int VEC, int NP>() -> int {
    const int64_t want = (N + kWarps - 1) / kWarps;
    const unsigned grid = unsigned(std::min<int64_t>(want, resident_ctas<layernorm_fwd_kernel<T, VEC, NP>>()));
    layernorm_fwd_kernel<T, VEC, NP><<<grid, kCta, 0, st>>>(static_cast<const T*>(x), N, C, ldx, gamma, beta, eps,
                                                            static_cast<T*>(y), ldy, mean_out, rstd_out);
    GMLM_LAUNCH_CHECK();
    return GMLM_OK;
  });
}

size_t gmlm_layernorm_bwd_workspace_bytes(int64_t /*num_rows*/, int64_t channels) {
  return size_t(kMaxCtas) * kWarps * 2 * size_t(channels) * sizeof(float) + 256;
}

int gmlm_layernorm_bwd(const void* x, const void* gy, int dtype, int64_t N, int64_t C, int64_t ldx, int64_t ldg,
                       const float* gamma, const float* mean, const float* rstd, void* gx, int64_t ldgx,
                       float* g_gamma, float* g_beta, void* ws, size_t ws_bytes, void* stream) {
  if (int rc = check_shape("layernorm_bwd", dtype, N, C, {x, gy, gx}, {ldx, ldg, gx ? ldgx : C})) return rc;
  GMLM_REQUIRE(N >= 1, "layernorm_bwd: needs at least one row");
  GMLM_REQUIRE(x && gy && gamma && mean && rstd, "layernorm_bwd: null pointer");
  if (!ws || ws_bytes < gmlm_layernorm_bwd_workspace_bytes(N, C))
    return fail(GMLM_ERR_WORKSPACE, "layernorm_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* partial = static_cast<float*>(ws);
  int64_t n_partials = 0;
  int rc = dispatch(dtype, C, [&]<typename T, int VEC, int NP>() -> int {
    const int64_t want = (N + kWarps - 1) / kWarps;
    const int64_t cap = std::min<int64_t>(resident_ctas<layernorm_bwd_kernel<T, VEC, NP>>(), kMaxCtas);
    const unsigned grid = unsigned(std::min<int64_t>(want, cap));
    n_partials = int64_t(grid) * kWarps;
    layernorm_bwd_kernel<T, VEC, NP><<<grid, kCta, 0, st>>>(static_cast<const T*>(x), static_cast<const T*>(gy), N,
                                                            C, ldx, ldg, gamma, mean, rstd, static_cast<T*>(gx),
                                                            ldgx, partial);
    GMLM_LAUNCH_CHECK();
    return GMLM_OK;
  });
  if (rc) return rc;
  if (g_gamma || g_beta) {
    layernorm_param_grads_kernel<<<unsigned((2 * C + 31) / 32), 256, 0, st>>>(partial, n_partials, C, g_gamma,
                                                                               g_beta);
    GMLM_LAUNCH_CHECK();
  }
  return GMLM_OK;
}

}  // extern "C"
