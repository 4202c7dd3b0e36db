// compose.cu — basis composition of the RGCN weights, forward and backward.
//
//   [PyG] RGCNConv.forward:  weight = (comp @ weight.view(num_bases, -1)).view(R, Fi, Fo)
//   (modules built main.py:189,193,197,201; recomputed on every call, recompute and backward included)
//
// Upstream runs this as a [R x B] . [B x Fi*Fo] matmul, then the per-relation `h @ weight[r]`; the torch path of
// round 1 followed it with an index_select of the populated relations, a concatenation with `root`, a cast to the
// GEMM operand type and a transposition — five passes over the composed weights and, under autocast, a cast of
// the full fp32 basis tensor (1 GB at H = 512) on every call.  Here ONE kernel reads the fp32 bases once
// (the HBM floor of the operation: B * Fi * Fo * 4 bytes) and writes the composed, `root`-extended weights
// directly in the element type and in the (up to) two layouts the tcgen05 GEMMs consume:
//
//   out_n[s * n_slot_stride + i * n_row_stride + o]   "natural"    [K, Fo]:  B operand of [dH | dx] = g . W^T
//   out_t[s * t_slot_stride + o * t_row_stride + i]   "transposed" [Fo, K]:  B operand of out = [H | x] . W
//
//   W_s[i,o] = sum_b comp[rel_of_slot[s], b] * basis[b, i, o]      s < S (populated relations only)
//   slot S   = root[i, o]                                          at n_root_off / t_root_off
//
// Backward (one kernel + a tiny final reduction): given dW in the natural strided layout (fp32),
//   dbasis[b,i,o] = sum_s comp[rel_s, b] * dW_s[i,o]               written once (B * Fi * Fo * 4 bytes)
//   dcomp[rel_s,b] = sum_{i,o} dW_s[i,o] * basis[b,i,o]            per-CTA partials, fixed-order fp64 final sum
// reading the bases once more: 3 passes over the basis tensor per training step in total, the minimum for
// recompute-free fp32 parameter gradients.  Deterministic (fixed grid, no atomics).
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace gmlm {
namespace {

constexpr int kMaxSlots = 8;      // slots per launch (the reference has 4 populated relations)
constexpr int kTileI = 16;        // rows (input channels) per CTA tile
constexpr int kThreads = 256;

template <typename OutT>
__device__ __forceinline__ OutT cvt(float v);
template <>
__device__ __forceinline__ float cvt<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half cvt<__half>(float v) { return __float2half_rn(v); }

struct ComposeParams {
  const float* comp;      // [R, B] or nullptr (no basis decomposition: W_s = basis[rel_of_slot[s]])
  const float* basis;     // [B, Fi, Fo]
  const float* root;      // [Fi, Fo] or nullptr
  int rel_of_slot[kMaxSlots];
  int S, B, Fi, Fo;
  void* out_n;
  int64_t n_slot_stride, n_row_stride, n_root_off;
  void* out_t;
  int64_t t_slot_stride, t_row_stride, t_root_off;
  int slot0;              // first slot of this launch (launches cover kMaxSlots slots each)
  int with_root;          // this launch also writes the root slab
};

// tile = 16 input channels x 16*VEC output channels; thread (tx, ty) owns row ty, VEC adjacent columns.  One row and
// MAXS (4 or 8) slots per thread keep the kernel at <= 64 registers (4 CTAs of 256 threads per SM), and the basis
// loop is unrolled ten deep: 32 warps x 10 independent 16-byte loads per lane in flight per SM (the first version --
// two rows, eight slots, 126 registers, 2 CTAs per SM -- measured 41 % of DRAM peak under ncu: latency-bound).
template <typename OutT, int VEC, int MAXS>
__global__ void __launch_bounds__(kThreads, MAXS == 4 ? 4 : 3) compose_fwd_kernel(const ComposeParams p) {
  constexpr int TILE_O = 16 * VEC;
  extern __shared__ float smem_f[];
  float* comp_s = smem_f;                              // [S][B]
  float* tile = smem_f + kMaxSlots * p.B;              // [TILE_O][kTileI + 1]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int o0 = blockIdx.x * TILE_O + tx * VEC;
  const int i0 = blockIdx.y * kTileI;
  for (int k = threadIdx.x; k < kMaxSlots * p.B; k += kThreads) {
    const int s = k / p.B, b = k % p.B;
    comp_s[k] = s >= p.S ? 0.f
                : (p.comp ? __ldg(p.comp + int64_t(p.rel_of_slot[s]) * p.B + b) : (b == p.rel_of_slot[s] ? 1.f : 0.f));
  }
  __syncthreads();
  float acc[MAXS][VEC];
#pragma unroll
  for (int s = 0; s < MAXS; ++s)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[s][k] = 0.f;
  const bool ok = o0 < p.Fo && i0 + ty < p.Fi;         // Fo is a multiple of VEC: a pack is in or out as a whole
  const int64_t plane = int64_t(p.Fi) * p.Fo;
  const float* src = p.basis + int64_t(i0 + ty) * p.Fo + o0;
  if (ok) {
#pragma unroll 10
    for (int b = 0; b < p.B; ++b) {
      float v[VEC];
      if constexpr (VEC == 4) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(src + b * plane));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
        v[0] = __ldcs(src + b * plane);
      }
#pragma unroll
      for (int s = 0; s < MAXS; ++s) {
        const float c = comp_s[s * p.B + b];             // slots past S hold zeros (never written out)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[s][k] = fmaf(c, v[k], acc[s][k]);
      }
    }
  }
  OutT* out_n = static_cast<OutT*>(p.out_n);
  OutT* out_t = static_cast<OutT*>(p.out_t);
  const int n_slabs = p.S + (p.with_root ? 1 : 0);
#pragma unroll 1
  for (int s = 0; s < n_slabs; ++s) {
    float v[VEC];
    const bool is_root = s == p.S;
    if (is_root) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = ok ? __ldg(p.root + int64_t(i0 + ty) * p.Fo + o0 + k) : 0.f;
    } else {
      // acc is indexed by a loop variable: select with a fully unrolled compare so it stays in registers
#pragma unroll
      for (int q = 0; q < MAXS; ++q)
        if (q == s) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) v[k] = acc[q][k];
        }
    }
    if (out_n && ok) {
      const int64_t base = is_root ? p.n_root_off : int64_t(p.slot0 + s) * p.n_slot_stride;
      OutT* d = out_n + base + int64_t(i0 + ty) * p.n_row_stride + o0;
#pragma unroll
      for (int k = 0; k < VEC; ++k) d[k] = cvt<OutT>(v[k]);
    }
    if (out_t) {
      __syncthreads();                                   // the previous slab has left the tile
#pragma unroll
      for (int k = 0; k < VEC; ++k) tile[(tx * VEC + k) * (kTileI + 1) + ty] = v[k];
      __syncthreads();
      const int64_t base = is_root ? p.t_root_off : int64_t(p.slot0 + s) * p.t_slot_stride;
      const int li = threadIdx.x & (kTileI - 1);         // consecutive threads -> consecutive input channels
      for (int lo = threadIdx.x / kTileI; lo < TILE_O; lo += kThreads / kTileI) {
        const int o = blockIdx.x * TILE_O + lo;
        if (o < p.Fo && i0 + li < p.Fi)
          out_t[base + int64_t(o) * p.t_row_stride + i0 + li] = cvt<OutT>(tile[lo * (kTileI + 1) + li]);
      }
    }
  }
}

struct ComposeBwdParams {
  const float* comp;      // [R, B]
  const float* basis;     // [B, Fi, Fo]
  const float* dw;        // natural strided layout: dw[s * slot_stride + i * row_stride + o]
  int64_t slot_stride, row_stride;
  int rel_of_slot[kMaxSlots];
  int S, B, Fi, Fo;
  float* dbasis;          // [B, Fi, Fo] or nullptr
  float* partial;         // [gridDim.x][S][B] or nullptr (dcomp not needed)
};

// warp w of a CTA owns the bases b = w, w + 8, ...; the 32 lanes own 32 adjacent float4 positions of the
// [Fi, Fo] plane.  Every warp reads the S gradient slabs at its positions (the same lines for the 8 warps of a CTA:
// L1), each basis element is read by exactly one thread.
constexpr int kBwdWarps = kThreads / 32;
constexpr int kMaxBPerWarp = 4;     // bases per warp per launch: 32 bases per launch (the reference has 30)

template <int MAXS>
__global__ void __launch_bounds__(kThreads, MAXS == 4 ? 3 : 2) compose_bwd_kernel(const ComposeBwdParams p, int b_base,
                                                                                 int b_count) {
  extern __shared__ float comp_s[];                    // [S][b_count]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < p.S * b_count; k += kThreads) {
    const int s = k / b_count, b = k % b_count;
    comp_s[k] = __ldg(p.comp + int64_t(p.rel_of_slot[s]) * p.B + b_base + b);
  }
  __syncthreads();
  float acc[kMaxBPerWarp][MAXS];
#pragma unroll
  for (int j = 0; j < kMaxBPerWarp; ++j)
#pragma unroll
    for (int s = 0; s < MAXS; ++s) acc[j][s] = 0.f;
  const int fo4 = p.Fo >> 2;
  const int64_t n4 = int64_t(p.Fi) * fo4;
  const int64_t plane = int64_t(p.Fi) * p.Fo;
  for (int64_t q0 = int64_t(blockIdx.x) * 32; q0 < n4; q0 += int64_t(gridDim.x) * 32) {
    const int64_t q = q0 + lane;
    const bool ok = q < n4;
    const int64_t i = ok ? q / fo4 : 0;
    const int o = ok ? int(q % fo4) * 4 : 0;
    float4 g[MAXS];
#pragma unroll
    for (int s = 0; s < MAXS; ++s) {
      g[s] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s < p.S && ok) g[s] = *reinterpret_cast<const float4*>(p.dw + s * p.slot_stride + i * p.row_stride + o);
    }
#pragma unroll
    for (int j = 0; j < kMaxBPerWarp; ++j) {
      const int b = warp + j * kBwdWarps;
      if (b < b_count && ok) {
        const int64_t off = int64_t(b_base + b) * plane + i * p.Fo + o;
        const float4 w = __ldcs(reinterpret_cast<const float4*>(p.basis + off));
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
          if (s < p.S) {
            const float c = comp_s[s * b_count + b];
            d.x = fmaf(c, g[s].x, d.x); d.y = fmaf(c, g[s].y, d.y); d.z = fmaf(c, g[s].z, d.z); d.w = fmaf(c, g[s].w, d.w);
            acc[j][s] += g[s].x * w.x + g[s].y * w.y + g[s].z * w.z + g[s].w * w.w;
          }
        }
        if (p.dbasis) __stcs(reinterpret_cast<float4*>(p.dbasis + off), d);
      }
    }
  }
  if (p.partial) {
#pragma unroll
    for (int j = 0; j < kMaxBPerWarp; ++j) {
      const int b = warp + j * kBwdWarps;
#pragma unroll
      for (int s = 0; s < MAXS; ++s) {
        float v = acc[j][s];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0 && b < b_count && s < p.S)
          p.partial[(int64_t(blockIdx.x) * p.S + s) * p.B + b_base + b] = v;
      }
    }
  }
}

// dcomp[r, b] = sum over CTAs (fixed order, fp64) of the partials of the slot that maps to relation r; 0 for
// relations without a slot
constexpr int kMaxRelations = 64;
struct SlotOfRel {
  int v[kMaxRelations];
};
// one warp per (relation, basis) pair: lane l sums the partials of CTAs l, l+32, .. in order, the 32 lane sums are
// combined by a fixed xor tree (deterministic)
__global__ void compose_bwd_final_kernel(const float* partial, int n_ctas, int S, int B, int R, const SlotOfRel slot_of_rel,
                                         float* dcomp) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= R * B) return;
  const int r = k / B, b = k % B;
  const int s = slot_of_rel.v[r];
  double sum = 0.0;
  if (s >= 0)
    for (int c = lane; c < n_ctas; c += 32) sum += double(partial[(int64_t(c) * S + s) * B + b]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (lane == 0) dcomp[k] = float(sum);
}

int bwd_grid(int64_t n4) {
  const int64_t want = (n4 + 31) / 32;
  return int(std::max<int64_t>(1, std::min<int64_t>(want, int64_t(num_sms()) * 9)));
}

template <typename OutT>
int launch_fwd(const ComposeParams& p, cudaStream_t st) {
  const size_t smem_vec4 = (size_t(kMaxSlots) * p.B + 64 * (kTileI + 1)) * sizeof(float);
  const size_t smem_vec1 = (size_t(kMaxSlots) * p.B + 16 * (kTileI + 1)) * sizeof(float);
  if (p.Fo % 4 == 0) {
    dim3 grid((p.Fo + 63) / 64, (p.Fi + kTileI - 1) / kTileI);
    if (p.S <= 4) compose_fwd_kernel<OutT, 4, 4><<<grid, kThreads, smem_vec4, st>>>(p);
    else compose_fwd_kernel<OutT, 4, 8><<<grid, kThreads, smem_vec4, st>>>(p);
  } else {
    dim3 grid((p.Fo + 15) / 16, (p.Fi + kTileI - 1) / kTileI);
    compose_fwd_kernel<OutT, 1, 8><<<grid, kThreads, smem_vec1, st>>>(p);
  }
  GMLM_LAUNCH_CHECK();
  return GMLM_OK;
}

}  // namespace
}  // namespace gmlm

using namespace gmlm;

extern "C" int gmlm_basis_compose(const float* comp, const float* basis, const float* root,
                                  const int32_t* rel_of_slot_host, int num_slots, int num_relations, int num_bases,
                                  int64_t in_channels, int64_t out_channels, int out_dtype, void* out_n,
                                  int64_t n_slot_stride, int64_t n_row_stride, int64_t n_root_off, void* out_t,
                                  int64_t t_slot_stride, int64_t t_row_stride, int64_t t_root_off, void* stream) {
  GMLM_REQUIRE(out_dtype == GMLM_F32 || out_dtype == GMLM_BF16 || out_dtype == GMLM_F16, "basis_compose: bad out_dtype");
  GMLM_REQUIRE(num_slots >= 0 && num_relations > 0 && num_bases > 0 && in_channels > 0 && out_channels > 0,
               "basis_compose: bad sizes");
  GMLM_REQUIRE(in_channels < (1 << 30) && out_channels < (1 << 30) && num_bases <= 1024, "basis_compose: sizes too large");
  GMLM_REQUIRE(basis && (num_slots == 0 || rel_of_slot_host), "basis_compose: null pointer");
  GMLM_REQUIRE(out_n || out_t, "basis_compose: no output requested");
  for (int s = 0; s < num_slots; ++s)
    GMLM_REQUIRE(rel_of_slot_host[s] >= 0 && rel_of_slot_host[s] < (comp ? num_relations : num_bases),
                 "basis_compose: rel_of_slot[%d] = %d out of range", s, rel_of_slot_host[s]);
  cudaStream_t st = as_stream(stream);
  bool root_done = root == nullptr;
  for (int s0 = 0; s0 < num_slots || !root_done; s0 += kMaxSlots) {
    ComposeParams p{};
    p.comp = comp; p.basis = basis; p.root = root;
    p.S = std::max(0, std::min(kMaxSlots, num_slots - s0));
    for (int s = 0; s < p.S; ++s) p.rel_of_slot[s] = rel_of_slot_host[s0 + s];
    p.B = num_bases; p.Fi = int(in_channels); p.Fo = int(out_channels);
    p.out_n = out_n; p.n_slot_stride = n_slot_stride; p.n_row_stride = n_row_stride; p.n_root_off = n_root_off;
    p.out_t = out_t; p.t_slot_stride = t_slot_stride; p.t_row_stride = t_row_stride; p.t_root_off = t_root_off;
    p.slot0 = s0;
    p.with_root = (!root_done && s0 + kMaxSlots >= num_slots) ? 1 : 0;
    if (p.with_root) root_done = true;
    int rc = out_dtype == GMLM_F32   ? launch_fwd<float>(p, st)
             : out_dtype == GMLM_BF16 ? launch_fwd<__nv_bfloat16>(p, st)
                                      : launch_fwd<__half>(p, st);
    if (rc) return rc;
  }
  return GMLM_OK;
}

extern "C" size_t gmlm_basis_compose_bwd_workspace_bytes(int num_slots, int num_relations, int num_bases,
                                                         int64_t in_channels, int64_t out_channels) {
  const int64_t n4 = in_channels * (out_channels / 4);
  return size_t(bwd_grid(n4)) * size_t(std::max(num_slots, 1)) * size_t(num_bases) * sizeof(float) +
         512 + 0 * size_t(num_relations);
}

extern "C" int gmlm_basis_compose_bwd(const float* comp, const float* basis, const float* dw, int64_t slot_stride,
                                      int64_t row_stride, const int32_t* rel_of_slot_host, int num_slots,
                                      int num_relations, int num_bases, int64_t in_channels, int64_t out_channels,
                                      float* dbasis, float* dcomp, void* ws, size_t ws_bytes, void* stream) {
  GMLM_REQUIRE(num_slots > 0 && num_slots <= kMaxSlots, "basis_compose_bwd: 1..%d populated relations supported", kMaxSlots);
  GMLM_REQUIRE(num_relations > 0 && num_bases > 0 && in_channels > 0 && out_channels > 0, "basis_compose_bwd: bad sizes");
  GMLM_REQUIRE(out_channels % 4 == 0 && slot_stride % 4 == 0 && row_stride % 4 == 0,
               "basis_compose_bwd: out_channels and the gradient strides must be multiples of 4");
  GMLM_REQUIRE(comp && basis && dw && rel_of_slot_host, "basis_compose_bwd: null pointer");
  GMLM_REQUIRE((reinterpret_cast<uintptr_t>(dw) & 15) == 0 && (reinterpret_cast<uintptr_t>(basis) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dbasis) & 15) == 0,
               "basis_compose_bwd: operands must be 16-byte aligned");
  GMLM_REQUIRE(ws_bytes >= gmlm_basis_compose_bwd_workspace_bytes(num_slots, num_relations, num_bases, in_channels,
                                                                  out_channels),
               "basis_compose_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  ComposeBwdParams p{};
  p.comp = comp; p.basis = basis; p.dw = dw; p.slot_stride = slot_stride; p.row_stride = row_stride;
  p.S = num_slots; p.B = num_bases; p.Fi = int(in_channels); p.Fo = int(out_channels);
  GMLM_REQUIRE(num_relations <= kMaxRelations, "basis_compose_bwd: at most %d relations", kMaxRelations);
  SlotOfRel slot_of_rel;
  for (int r = 0; r < kMaxRelations; ++r) slot_of_rel.v[r] = -1;
  for (int s = 0; s < num_slots; ++s) {
    GMLM_REQUIRE(rel_of_slot_host[s] >= 0 && rel_of_slot_host[s] < num_relations, "basis_compose_bwd: bad relation id");
    p.rel_of_slot[s] = rel_of_slot_host[s];
    slot_of_rel.v[rel_of_slot_host[s]] = s;
  }
  const int64_t n4 = in_channels * (out_channels / 4);
  const int grid = bwd_grid(n4);
  Carver carve(ws);
  p.partial = dcomp ? carve.take<float>(size_t(grid) * num_slots * num_bases) : nullptr;
  p.dbasis = dbasis;
  constexpr int kPerLaunch = kBwdWarps * kMaxBPerWarp;
  for (int b0 = 0; b0 < num_bases; b0 += kPerLaunch) {
    const int cnt = std::min(kPerLaunch, num_bases - b0);
    if (num_slots <= 4) compose_bwd_kernel<4><<<grid, kThreads, size_t(num_slots) * cnt * sizeof(float), st>>>(p, b0, cnt);
    else compose_bwd_kernel<8><<<grid, kThreads, size_t(num_slots) * cnt * sizeof(float), st>>>(p, b0, cnt);
    GMLM_LAUNCH_CHECK();
  }
  if (dcomp) {
    const int n = num_relations * num_bases;
    compose_bwd_final_kernel<<<(n * 32 + 127) / 128, 128, 0, st>>>(p.partial, grid, num_slots, num_bases, num_relations,
                                                              slot_of_rel, dcomp);
    GMLM_LAUNCH_CHECK();
  }
  return GMLM_OK;
}
