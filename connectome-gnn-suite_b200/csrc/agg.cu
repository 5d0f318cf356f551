// agg.cu - the gather half of the tensor-core generation of layer kernels.
//
//   k_build_agg      CSR arrays -> packed per-subject aggregation blobs (agg.cuh), once per batch and model family
//   k_gather<MODE>   one subject tile [n, C] in shared memory (two CTAs per SM), one warp per output row:
//     GATHER_SAGE_FWD  agg_i = sum_{e: dst=i} w_e u_src / (w_sum_i + 1e-8), u = act(t_in)     (reference models.py:146-149)
//     GATHER_GCN_BWD   dP_j  = sum_{e: src=j} w^_e dz_dst + dinv_j^2 dz_j, dz from z / upstream / BatchNorm backward
//                      on load, plus dbias = column sums of dz                        (autograd of models.py:112-114)
//     GATHER_SAGE_BWD  du_j  = d_u_j + sum_{e: src=j} w_e d_agg_dst / (w_sum_dst + 1e-8), plus the BatchNorm
//                      backward sums of the layer below                               (autograd of models.py:146-149)
// The dense contractions around them live in gemm_tc.cu.
#include "agg.cuh"
#include "tile.cuh"

namespace cgnn {

// ------------------------------------------------------------------------------------------------------------
// blob builder: one CTA per subject, one thread per row
// ------------------------------------------------------------------------------------------------------------
struct BuildAggArgs {
  const int32_t* in_rowptr; const int32_t* in_col; const float* in_w; const float* in_wn;
  const int32_t* out_rowptr; const int32_t* out_col; const float* out_w; const float* out_wn;
  const float* dinv; const float* wsum; const int32_t* meta;
  int kind, max_nodes;
  int32_t* agg_in; int32_t* agg_out; int32_t* row_graph;
};

__global__ void __launch_bounds__(256) k_build_agg(BuildAggArgs p) {
  CGNN_SMEM_DECL;
  __shared__ int s_tot[2][8];
  int* s_pos = reinterpret_cast<int*>(cgnn_smem);   // [2][max_nodes] padded record counts -> exclusive positions
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const long long g = blockIdx.x;
  const int4 m = reinterpret_cast<const int4*>(p.meta)[g];
  const long long nb = m.x, eb = m.z;
  const int n = m.y;
  if (n > p.max_nodes) return;
  const int self = p.kind == AGG_GCN ? 1 : 0;
  for (int dir = 0; dir < 2; ++dir) {
    const int32_t* rp = dir == 0 ? p.in_rowptr : p.out_rowptr;
    // exclusive scan of the padded row lengths: chunked per thread, warp scan, cross-warp carry
    const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    int sum = 0;
    for (int i = lo; i < hi; ++i) {
      const int c = (rp[nb + i + 1] - rp[nb + i] + self + 1) & ~1;
      s_pos[dir * p.max_nodes + i] = c;
      sum += c;
    }
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_tot[dir][warp] = inc;
    __syncthreads();
    int carry = 0;
    for (int w = 0; w < warp; ++w) carry += s_tot[dir][w];
    int run = carry + inc - sum;
    for (int i = lo; i < hi; ++i) { const int c = s_pos[dir * p.max_nodes + i]; s_pos[dir * p.max_nodes + i] = run; run += c; }
    (void)nwarps;
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) p.row_graph[nb + i] = (int32_t)g;
  for (int dir = 0; dir < 2; ++dir) {
    const int32_t* rp = dir == 0 ? p.in_rowptr : p.out_rowptr;
    const int32_t* col = dir == 0 ? p.in_col : p.out_col;
    const float* wv = p.kind == AGG_GCN ? (dir == 0 ? p.in_wn : p.out_wn) : (dir == 0 ? p.in_w : p.out_w);
    int32_t* blob = (dir == 0 ? p.agg_in : p.agg_out) + agg_base_words(nb, eb, g);
    int4* desc = reinterpret_cast<int4*>(blob);
    int2* rec = reinterpret_cast<int2*>(blob + 4 * (long long)n);
    for (int i = tid; i < n; i += blockDim.x) {
      const int e0 = rp[nb + i], e1 = rp[nb + i + 1];
      int pos = s_pos[dir * p.max_nodes + i];
      const int begin = pos;
      for (int e = e0; e < e1; ++e) {
        const int nbr = (int)(col[e] - nb);
        float w = wv[e];
        if (p.kind == AGG_SAGE && dir == 1) w = w / (p.wsum[nb + nbr] + 1e-8f);   // adjoint of the weighted mean
        rec[pos++] = make_int2(nbr, __float_as_int(w));
      }
      if (self) { const float d = p.dinv[nb + i]; rec[pos++] = make_int2(i, __float_as_int(__fmul_rn(d, d))); }
      if (pos & 1) rec[pos++] = make_int2(i, 0);
      const float aux = p.kind == AGG_SAGE ? p.wsum[nb + i] : p.dinv[nb + i];
      desc[i] = make_int4(begin, pos, __float_as_int(aux), 0);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// gather kernels
// ------------------------------------------------------------------------------------------------------------
template <int MODE, int VW>
__global__ void __launch_bounds__(kThreads, 2) k_gather(GatherArgs p) {
  CGNN_SMEM_DECL;
  float* sm = reinterpret_cast<float*>(cgnn_smem);
  const int C = p.C, ld = p.ld;
  float* s_co = sm;                              // [GC_ROWS][ld] per-channel constants
  float* s_red = s_co + GC_ROWS * ld;            // [kWarps][2*32*VW] end-of-kernel reduction
  float* s_tile = s_red + kWarps * 64 * VW;      // [max_nodes][ld] (+32 floats of slack)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool affine = p.act.scale != nullptr;

  stage_affine(p.act, C, ld, s_co + GC_SCALE * ld, s_co + GC_SHIFT * ld);
  for (int c = tid; c < ld; c += kThreads) {
    float bsc = c < C ? 1.0f : 0.0f, mean = 0.0f, rstd = 0.0f, s1n = 0.0f, s2n = 0.0f;
    if (MODE == GATHER_GCN_BWD && c < C && p.has_bn) {
      bsc = p.bn_scale[c]; mean = p.bn_mean[c]; rstd = p.bn_rstd[c];
      if (p.bn_train) { s1n = p.bn_s1[c] * p.inv_count; s2n = p.bn_s2[c] * p.inv_count; }
    }
    if (MODE == GATHER_SAGE_BWD && c < C && p.want_prev) { mean = p.prev_mean[c]; rstd = p.prev_rstd[c]; }
    s_co[GC_BSC * ld + c] = bsc; s_co[GC_MEAN * ld + c] = mean; s_co[GC_RSTD * ld + c] = rstd;
    s_co[GC_S1N * ld + c] = s1n; s_co[GC_S2N * ld + c] = s2n;
  }
  __syncthreads();

  // dz of one element (GCN_BWD): dropout / ReLU backward of the upstream gradient, then BatchNorm backward
  auto dz_of = [&](float t, float up, uint32_t rh, int c) -> float {
    const float dy = act_bwd(p.act, affine, t, s_co[GC_SCALE * ld + c], s_co[GC_SHIFT * ld + c], rh, c, up);
    if (!p.has_bn) return dy;
    if (!p.bn_train) return s_co[GC_BSC * ld + c] * dy;
    const float xh = (t - s_co[GC_MEAN * ld + c]) * s_co[GC_RSTD * ld + c];
    return s_co[GC_BSC * ld + c] * (dy - s_co[GC_S1N * ld + c] - xh * s_co[GC_S2N * ld + c]);
  };

  float colsum[4] = {0.f, 0.f, 0.f, 0.f};          // GCN_BWD: dbias, this thread's channel quad (vec) or channel
  float ps1[VW], ps2[VW];                          // SAGE_BWD: sums of the layer below, channels VW*lane + j
#pragma unroll
  for (int j = 0; j < VW; ++j) { ps1[j] = 0.f; ps2[j] = 0.f; }

  const int4* meta = reinterpret_cast<const int4*>(p.meta);
  for (long long g = blockIdx.x; g < p.B; g += gridDim.x) {
    const int4 m = meta[g];
    const long long nb = m.x, eb = m.z;
    const int n = min(m.y, p.max_nodes);
    // ---- load phase: the subject's tile, transformed on the way in ---------------------------------------
    if (p.vec) {
      const int Q = C >> 2;
      const int total = n * Q;
      const float inv_n = 1.0f / ((float)n + 1e-8f);
      for (int base = tid; base < total; base += 4 * kThreads) {
        float4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * kThreads;
          if (idx < total) {
            const int i = idx / Q, c = (idx - i * Q) << 2;
            a[u] = *reinterpret_cast<const float4*>(p.src + (nb + i) * C + c);
            if (MODE == GATHER_GCN_BWD && p.du) b[u] = *reinterpret_cast<const float4*>(p.du + (nb + i) * C + c);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * kThreads;
          if (idx < total) {
            const int i = idx / Q, c = (idx - i * Q) << 2;
            float4 o = a[u];
            if (MODE == GATHER_SAGE_FWD) {
              const uint32_t rh = p.act.drop ? drop_row_hash(p.act, p.act.row_base + nb + i) : 0u;
              o.x = act_fwd(p.act, affine, a[u].x, s_co[GC_SCALE * ld + c + 0], s_co[GC_SHIFT * ld + c + 0], rh, c + 0);
              o.y = act_fwd(p.act, affine, a[u].y, s_co[GC_SCALE * ld + c + 1], s_co[GC_SHIFT * ld + c + 1], rh, c + 1);
              o.z = act_fwd(p.act, affine, a[u].z, s_co[GC_SCALE * ld + c + 2], s_co[GC_SHIFT * ld + c + 2], rh, c + 2);
              o.w = act_fwd(p.act, affine, a[u].w, s_co[GC_SCALE * ld + c + 3], s_co[GC_SHIFT * ld + c + 3], rh, c + 3);
            } else if (MODE == GATHER_GCN_BWD) {
              if (!p.du) {
                const float4 e = *reinterpret_cast<const float4*>(p.demb + g * C + c);
                b[u] = make_float4(e.x * inv_n, e.y * inv_n, e.z * inv_n, e.w * inv_n);
              }
              const uint32_t rh = p.act.drop ? drop_row_hash(p.act, p.act.row_base + nb + i) : 0u;
              o.x = dz_of(a[u].x, b[u].x, rh, c + 0);
              o.y = dz_of(a[u].y, b[u].y, rh, c + 1);
              o.z = dz_of(a[u].z, b[u].z, rh, c + 2);
              o.w = dz_of(a[u].w, b[u].w, rh, c + 3);
              colsum[0] += o.x; colsum[1] += o.y; colsum[2] += o.z; colsum[3] += o.w;   // quad fixed: kThreads % Q == 0
            }
            *reinterpret_cast<float4*>(s_tile + i * ld + c) = o;
          }
        }
      }
    } else {
      const float inv_n = 1.0f / ((float)n + 1e-8f);
      for (int idx = tid; idx < n * ld; idx += kThreads) {
        const int i = idx / ld, c = idx - i * ld;
        float o = 0.0f;
        if (c < C) {
          const float t = p.src[(nb + i) * C + c];
          const uint32_t rh = p.act.drop ? drop_row_hash(p.act, p.act.row_base + nb + i) : 0u;
          if (MODE == GATHER_SAGE_FWD) o = act_fwd(p.act, affine, t, s_co[GC_SCALE * ld + c], s_co[GC_SHIFT * ld + c], rh, c);
          else if (MODE == GATHER_GCN_BWD) {
            const float up = p.du ? p.du[(nb + i) * C + c] : p.demb[g * C + c] * inv_n;
            o = dz_of(t, up, rh, c);
          } else o = t;
        }
        s_tile[idx] = o;
      }
    }
    __syncthreads();
    // ---- gather phase: one warp per output row --------------------------------------------------------------
    const AggView av = agg_view(p.blob, nb, n, eb, g);
    const float* tile_lane = s_tile + VW * lane;
    const int c0 = VW * lane;
    for (int i = warp; i < n; i += kWarps) {
      float acc[VW];
#pragma unroll
      for (int j = 0; j < VW; ++j) acc[j] = 0.0f;
      const float aux = agg_gather_row<VW>(av, i, tile_lane, ld, acc);
      const long long grow = nb + i;
      if (MODE == GATHER_SAGE_FWD) {
        const float inv = 1.0f / (aux + 1e-8f);
#pragma unroll
        for (int j = 0; j < VW; ++j) acc[j] *= inv;
      }
      if (MODE == GATHER_SAGE_BWD) {
        float dir[VW], raw[VW];
        if (VW == 2) {
          const float2 d2 = *reinterpret_cast<const float2*>(p.direct + grow * C + c0); dir[0] = d2.x; dir[1] = d2.y;
        } else if (VW == 4) {
          const float4 d4 = *reinterpret_cast<const float4*>(p.direct + grow * C + c0);
          dir[0] = d4.x; dir[1] = d4.y; dir[2] = d4.z; dir[3] = d4.w;
        } else {
#pragma unroll
          for (int j = 0; j < VW; ++j) dir[j] = c0 + j < C ? p.direct[grow * C + c0 + j] : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < VW; ++j) acc[j] += dir[j];
        if (p.want_prev) {
          if (VW == 2) {
            const float2 r2 = *reinterpret_cast<const float2*>(p.t_raw + grow * C + c0); raw[0] = r2.x; raw[1] = r2.y;
          } else if (VW == 4) {
            const float4 r4 = *reinterpret_cast<const float4*>(p.t_raw + grow * C + c0);
            raw[0] = r4.x; raw[1] = r4.y; raw[2] = r4.z; raw[3] = r4.w;
          } else {
#pragma unroll
            for (int j = 0; j < VW; ++j) raw[j] = c0 + j < C ? p.t_raw[grow * C + c0 + j] : 0.0f;
          }
          const uint32_t rh = p.act.drop ? drop_row_hash(p.act, p.act.row_base + grow) : 0u;
#pragma unroll
          for (int j = 0; j < VW; ++j) {
            const int c = c0 + j;
            if (c < C) {
              const float dyp = act_bwd(p.act, affine, raw[j], s_co[GC_SCALE * ld + c], s_co[GC_SHIFT * ld + c], rh, c, acc[j]);
              const float xh = (raw[j] - s_co[GC_MEAN * ld + c]) * s_co[GC_RSTD * ld + c];
              ps1[j] += dyp;
              ps2[j] = fmaf(dyp, xh, ps2[j]);
            }
          }
        }
      }
      float* dst = p.out + grow * C + c0;
      if (VW == 2) *reinterpret_cast<float2*>(dst) = make_float2(acc[0], acc[1]);
      else if (VW == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      else {
#pragma unroll
        for (int j = 0; j < VW; ++j) if (c0 + j < C) dst[j] = acc[j];
      }
    }
    __syncthreads();   // the tile is rewritten by the next subject
  }

  // ---- per-CTA partial records ------------------------------------------------------------------------------
  if (MODE == GATHER_GCN_BWD && p.partials) {
    // vec path: thread's quad = tid % Q; sum the kThreads / Q threads that share it
    const int Q = C >> 2;
    float* red = s_tile;   // free now: [kThreads][4]
    *reinterpret_cast<float4*>(red + 4 * tid) = make_float4(colsum[0], colsum[1], colsum[2], colsum[3]);
    __syncthreads();
    for (int c = tid; c < C; c += kThreads) {
      const int q = c >> 2, j = c & 3;
      float s = 0.0f;
      for (int t = q; t < kThreads; t += Q) s += red[4 * t + j];
      p.partials[(size_t)blockIdx.x * p.part_stride + c] = s;
    }
  }
  if (MODE == GATHER_SAGE_BWD && p.partials && p.want_prev) {
#pragma unroll
    for (int j = 0; j < VW; ++j) {
      s_red[warp * 64 * VW + VW * lane + j] = ps1[j];
      s_red[warp * 64 * VW + 32 * VW + VW * lane + j] = ps2[j];
    }
    __syncthreads();
    for (int c = tid; c < 2 * C; c += kThreads) {
      const int which = c / C, ch = c - which * C;
      float s = 0.0f;
      for (int w = 0; w < kWarps; ++w) s += s_red[w * 64 * VW + which * 32 * VW + ch];
      p.partials[(size_t)blockIdx.x * p.part_stride + c] = s;
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
// Launches one gather kernel; returns the grid through *grid_out (the caller reduces `partials` over it).
// Shapes: C a multiple of 32 up to 128, or C <= 32 (one channel per lane).
int launch_gather(int mode, GatherArgs& a, int* grid_out, cudaStream_t stream) {
  const DeviceInfo dev = device_info();
  const int C = a.C;
  if (C <= 0 || C > 128 || (C > 32 && C % 32 != 0)) return -1;
  const int VW = C <= 32 ? 1 : C / 32;
  if (VW == 3) return -1;
  a.ld = round_up(C, 4);
  a.vec = (C % 4 == 0) && ((((uintptr_t)a.src) & 15u) == 0) && (!a.du || (((uintptr_t)a.du) & 15u) == 0) &&
          (!a.demb || (((uintptr_t)a.demb) & 15u) == 0) && (kThreads % (C / 4) == 0);
  if (mode == GATHER_GCN_BWD && !a.vec) return -1;    // dbias column sums need the fixed-quad mapping
  if (a.max_nodes < 1) a.max_nodes = 1;
  const size_t words = (size_t)GC_ROWS * a.ld + (size_t)kWarps * 64 * VW + (size_t)a.max_nodes * a.ld + 32 +
                       (mode == GATHER_GCN_BWD ? 4 * kThreads : 0);
  const size_t smem = words * sizeof(float);
  if (smem > (size_t)dev.smem_optin) return -1;
  int per_sm = (int)((size_t)(228 * 1024) / (smem + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)per_sm * dev.sm_count;
  if (grid > a.B) grid = a.B;
  if (grid < 1) grid = 1;
  *grid_out = (int)grid;
#define CGNN_GATHER(MODE_, VW_)                                                                          \
  {                                                                                                      \
    auto kfn = k_gather<MODE_, VW_>;                                                                     \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    CGNN_LAUNCH(kfn, (unsigned)grid, kThreads, smem, stream, a);                                         \
  }
#define CGNN_GATHER_VW(MODE_)                                                                            \
  { if (VW == 1) CGNN_GATHER(MODE_, 1) else if (VW == 2) CGNN_GATHER(MODE_, 2) else CGNN_GATHER(MODE_, 4) }
  if (mode == GATHER_SAGE_FWD) CGNN_GATHER_VW(GATHER_SAGE_FWD)
  else if (mode == GATHER_GCN_BWD) CGNN_GATHER_VW(GATHER_GCN_BWD)
  else CGNN_GATHER_VW(GATHER_SAGE_BWD)
#undef CGNN_GATHER_VW
#undef CGNN_GATHER
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

}  // namespace cgnn

using namespace cgnn;

extern "C" {

size_t cgnn_agg_words(int64_t rows, int64_t edges, int64_t num_graphs) {
  if (rows < 0 || edges < 0 || num_graphs < 0) return 0;
  return agg_total_words(rows, edges, num_graphs);
}

int cgnn_build_agg(const cgnn_csr_t* csr, int32_t kind, int64_t num_graphs, int64_t rows, int64_t edges,
                   int32_t max_nodes, int32_t* agg_in, int32_t* agg_out, int32_t* row_graph, cgnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!csr || (kind != AGG_GCN && kind != AGG_SAGE) || num_graphs < 0 || rows < 0 || edges < 0 || max_nodes < 0)
    return CGNN_ERR_INVALID_ARG;
  if (num_graphs == 0 || rows == 0) return CGNN_OK;
  if (!agg_in || !agg_out || !row_graph || !csr->in_rowptr || !csr->in_col || !csr->in_w || !csr->in_wn ||
      !csr->out_rowptr || !csr->out_col || !csr->out_w || !csr->out_wn || !csr->dinv || !csr->wsum || !csr->graph_meta)
    return CGNN_ERR_INVALID_ARG;
  BuildAggArgs a;
  a.in_rowptr = csr->in_rowptr; a.in_col = csr->in_col; a.in_w = csr->in_w; a.in_wn = csr->in_wn;
  a.out_rowptr = csr->out_rowptr; a.out_col = csr->out_col; a.out_w = csr->out_w; a.out_wn = csr->out_wn;
  a.dinv = csr->dinv; a.wsum = csr->wsum; a.meta = csr->graph_meta;
  a.kind = kind; a.max_nodes = max_nodes < 1 ? 1 : max_nodes;
  a.agg_in = agg_in; a.agg_out = agg_out; a.row_graph = row_graph;
  const size_t smem = (size_t)2 * a.max_nodes * sizeof(int) + 16;
  const DeviceInfo dev = device_info();
  if (smem > (size_t)dev.smem_optin) return CGNN_ERR_TILE_TOO_LARGE;
  auto kfn = k_build_agg;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  CGNN_LAUNCH(kfn, (unsigned)num_graphs, 256, smem, stream, a);
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

}  // extern "C"
