// tile.cuh - building blocks shared by the GCN / GraphSAGE layer kernels: staging rows of an
// activation tensor into shared memory with the previous layer's BatchNorm+ReLU+dropout applied,
// per-channel constant staging, Welford statistics epilogue.
#pragma once
#include "common.cuh"

namespace cgnn {

// Per-channel affine constants of an Act into shared memory, padded to C4 with zeros
// (identity -> scale 1, shift 0).  Caller syncs.
__device__ __forceinline__ void stage_affine(const Act& a, int C, int C4, float* s_scale, float* s_shift) {
  for (int c = threadIdx.x; c < C4; c += blockDim.x) {
    float sc = 0.0f, sh = 0.0f;
    if (c < C) { sc = a.scale ? a.scale[c] : 1.0f; sh = a.scale ? a.shift[c] : 0.0f; }
    s_scale[c] = sc;
    s_shift[c] = sh;
  }
}

// Stage `R` rows x C4 channels into dst[r*ld + c]: rows [0, rows) come from t[(grow0 + r)*C + c]
// (global row ids, row stride C) with act applied; rows >= `rows` and channels >= C are zero.
// If raw != nullptr the untransformed values are stored there too (same layout).
// VEC: C % 4 == 0 and 16-byte aligned rows -> float4 loads.
template <bool VEC>
__device__ __forceinline__ void stage_rows(const float* __restrict__ t, long long grow0, int rows, int R, int C,
                                           int C4, int ld, const Act& a, const float* s_scale,
                                           const float* s_shift, float* dst, float* raw) {
  const bool affine = a.scale != nullptr;
  if (VEC) {
    const int q4 = C4 >> 2;
    for (int idx = threadIdx.x; idx < R * q4; idx += blockDim.x) {
      const int r = idx / q4, c = (idx - r * q4) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f), rv = v;
      if (r < rows) {
        rv = *reinterpret_cast<const float4*>(t + (grow0 + r) * C + c);
        const uint32_t rh = a.drop ? drop_row_hash(a, a.row_base + grow0 + r) : 0u;
        v.x = act_fwd(a, affine, rv.x, s_scale[c], s_shift[c], rh, c);
        v.y = act_fwd(a, affine, rv.y, s_scale[c + 1], s_shift[c + 1], rh, c + 1);
        v.z = act_fwd(a, affine, rv.z, s_scale[c + 2], s_shift[c + 2], rh, c + 2);
        v.w = act_fwd(a, affine, rv.w, s_scale[c + 3], s_shift[c + 3], rh, c + 3);
      }
      *reinterpret_cast<float4*>(dst + r * ld + c) = v;
      if (raw) *reinterpret_cast<float4*>(raw + r * ld + c) = rv;
    }
  } else {
    for (int idx = threadIdx.x; idx < R * C4; idx += blockDim.x) {
      const int r = idx / C4, c = idx - r * C4;
      float v = 0.0f, rv = 0.0f;
      if (r < rows && c < C) {
        rv = t[(grow0 + r) * C + c];
        const uint32_t rh = a.drop ? drop_row_hash(a, a.row_base + grow0 + r) : 0u;
        v = act_fwd(a, affine, rv, s_scale[c], s_shift[c], rh, c);
      }
      dst[r * ld + c] = v;
      if (raw) raw[r * ld + c] = rv;
    }
  }
}

// Per-warp running Welford state for HC channels per lane (channel = lane + 32*j).
template <int HC>
struct WarpStats {
  Welford w[HC];
  int rows;
  __device__ __forceinline__ void init() {
    rows = 0;
#pragma unroll
    for (int j = 0; j < HC; ++j) w[j].init();
  }
  __device__ __forceinline__ void begin_row(float& inv) { rows += 1; inv = fast_rcp((float)rows); }
  // Deposit into shared scratch: s_cnt[warp], s_mean[warp*ldc + c], s_m2[...]
  __device__ __forceinline__ void deposit(float* s_cnt, float* s_mean, float* s_m2, int ldc, int C4) const {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_cnt[warp] = (float)rows;
#pragma unroll
    for (int j = 0; j < HC; ++j) {
      const int c = lane + 32 * j;
      if (c < C4) { s_mean[warp * ldc + c] = w[j].mean; s_m2[warp * ldc + c] = w[j].m2; }
    }
  }
};

// ---- one subject's CSR rows, either staged in shared memory or read in place ---------------
// All three arrays are indexed with ABSOLUTE positions (rp[i] are global edge positions, col/w are
// pre-offset by -first_edge when they live in shared memory), so the same loop serves both homes.
struct RowCsr {
  const int* rp;     // [n+1] row i owns edges [rp[i], rp[i+1])
  const int* col;    // global row id of the neighbour
  const float* w;
};

static inline __host__ __device__ int csr_words(int max_nodes, int max_edges) {
  return round_up(max_nodes + 1, 4) + 2 * round_up(max_edges, 4) + round_up(max_nodes, 4);
}

// Start the async copy of a subject's CSR slice (+ one per-row float array) into shared memory.
// s_base layout: [rp: max_nodes+1][col: max_edges][w: max_edges][extra: max_nodes] (each rounded up to 4 words).
__device__ __forceinline__ void stage_csr_async(float* s_base, int max_nodes, int max_edges, const int32_t* rowptr,
                                                const int32_t* col, const float* w, const float* extra, long long nb,
                                                int n, int eb, int m) {
  int* s_rp = reinterpret_cast<int*>(s_base);
  int* s_col = s_rp + round_up(max_nodes + 1, 4);
  float* s_w = reinterpret_cast<float*>(s_col + round_up(max_edges, 4));
  float* s_extra = s_w + round_up(max_edges, 4);
  cp_async_words(s_rp, rowptr + nb, n + 1);
  cp_async_words(s_col, col + eb, m);
  cp_async_words(s_w, w + eb, m);
  if (extra) cp_async_words(s_extra, extra + nb, n);
}
__device__ __forceinline__ RowCsr staged_csr(float* s_base, int max_nodes, int max_edges, int eb, const float** extra) {
  int* s_rp = reinterpret_cast<int*>(s_base);
  int* s_col = s_rp + round_up(max_nodes + 1, 4);
  float* s_w = reinterpret_cast<float*>(s_col + round_up(max_edges, 4));
  *extra = s_w + round_up(max_edges, 4);
  return RowCsr{s_rp, s_col - eb, s_w - eb};
}

// acc[j] (+)= sum over the edges of row i of w_e * tile[(col_e - nb) * ld + lane + 32 j].
// EXACT keeps the reference's rounding (product rounded, then added, in COO order); otherwise FMA.
template <int HC, bool EXACT>
__device__ __forceinline__ void gather_row(const RowCsr& c, int i, long long nb, int n, const float* __restrict__ tile,
                                           int ld, int C4, float (&acc)[HC]) {
  const int lane = threadIdx.x & 31;
  const int e0 = c.rp[i], e1 = c.rp[i + 1];
  for (int e = e0; e < e1; ++e) {
    const int src = (int)(c.col[e] - nb);
    const float w = c.w[e];
    if ((unsigned)src < (unsigned)n) {
#pragma unroll
      for (int j = 0; j < HC; ++j) {
        const int ch = lane + 32 * j;
        if (ch < C4) {
          const float v = tile[src * ld + ch];
          acc[j] = EXACT ? __fadd_rn(acc[j], __fmul_rn(v, w)) : fmaf(v, w, acc[j]);
        }
      }
    }
  }
}

static inline int pick_hc(int C4) { return C4 <= 32 ? 1 : (C4 <= 64 ? 2 : (C4 <= 128 ? 4 : 8)); }

}  // namespace cgnn
