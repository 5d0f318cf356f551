// gcn_tc.cu - GCN layer forward, tensor-core edition (sm_100a only; the generic kernel in gcn.cu is the
// general-shape path and the on-device cross-check for this one).
//
// Same contract as k_gcn_fwd (reference models.py:84-114 fused with the previous layer's BatchNorm + ReLU +
// dropout), restructured around what ncu showed for the SIMT version (issue-bound: 246k warp instructions
// per subject, FFMA only 20% of them):
//   * projection  P = u W^T  on tcgen05 (3xTF32, fp32 accumulators in TMEM): 128-row tiles of the subject are
//     transformed straight from registers into 128B-swizzled K-major operands, one thread issues the MMAs;
//   * the next 128-row tile (possibly of the NEXT subject) is already in registers while the current one is
//     multiplied, copied out of TMEM and - after the last tile - aggregated;
//   * the subject's CSR arrives by cp.async during the projection; the aggregation loop touches shared memory
//     only, one warp per destination row, VW channels per lane, FMA.
#include "tc05.cuh"
#include "tile.cuh"

namespace cgnn {
#ifndef CGNN_EMU

struct GcnTcArgs {
  const float* t_in; Act act; const float* W; const float* bias;
  const int32_t* in_rowptr; const int32_t* in_col; const float* in_wn; const float* dinv;
  const int32_t* meta; long long B;
  int K, H, ldp, max_nodes, max_edges, vec_in;
  float* z; double* partials;
  uint32_t tmem_cols;
  // byte offsets from the 1024-aligned base
  int o_ahi, o_alo, o_bhi, o_blo, o_p, o_csr, o_scale, o_shift, o_bias, o_st;
};

constexpr int kTcM = 128;

// KB = padded K / 32 (operand blocks), VW = H / 32 (channels per lane in the aggregation)
template <int KB, int VW>
__global__ void __launch_bounds__(kThreads, 1) k_gcn_fwd_tc(GcnTcArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* base = tc::smem_align1024(smem_raw);
  unsigned char* a_hi = base + p.o_ahi;
  unsigned char* a_lo = base + p.o_alo;
  unsigned char* b_hi = base + p.o_bhi;
  unsigned char* b_lo = base + p.o_blo;
  float* s_p = reinterpret_cast<float*>(base + p.o_p);          // [max_nodes][ldp]
  float* s_csr = reinterpret_cast<float*>(base + p.o_csr);
  float* s_scale = reinterpret_cast<float*>(base + p.o_scale);  // [KP]
  float* s_shift = reinterpret_cast<float*>(base + p.o_shift);
  float* s_bias = reinterpret_cast<float*>(base + p.o_bias);    // [H]
  float* s_cnt = reinterpret_cast<float*>(base + p.o_st);
  float* s_mean = s_cnt + kWarps;
  float* s_m2 = s_mean + kWarps * p.H;

  constexpr int KP = 32 * KB;          // padded K
  constexpr int Q = KP / 4;            // float4 quads per operand row
  constexpr int NQ = kTcM * Q / kThreads;   // quads per thread per 128-row tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = p.K, H = p.H, ldp = p.ldp;
  const bool affine = p.act.scale != nullptr;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, p.tmem_cols);
  if (tid == 0) tc::mbar_init(&mbar, 1);
  // W -> hi/lo K-major swizzled operand, zero padded to KP
  for (int idx = tid; idx < H * Q; idx += kThreads) {
    const int n = idx / Q, k = (idx - n * Q) << 2;
    float4 v;
    v.x = k + 0 < K ? p.W[(size_t)n * K + k + 0] : 0.0f;
    v.y = k + 1 < K ? p.W[(size_t)n * K + k + 1] : 0.0f;
    v.z = k + 2 < K ? p.W[(size_t)n * K + k + 2] : 0.0f;
    v.w = k + 3 < K ? p.W[(size_t)n * K + k + 3] : 0.0f;
    tc::store_split4(b_hi, b_lo, n, k, H, v);
  }
  stage_affine(p.act, K, KP, s_scale, s_shift);
  for (int h = tid; h < H; h += kThreads) s_bias[h] = p.bias ? p.bias[h] : 0.0f;
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  const uint32_t a_hi_u = tc::smem_u32(a_hi), a_lo_u = tc::smem_u32(a_lo), b_hi_u = tc::smem_u32(b_hi), b_lo_u = tc::smem_u32(b_lo);

  WarpStats<VW> st;
  st.init();

  const int4* meta = reinterpret_cast<const int4*>(p.meta);
  auto load_meta = [&](long long g) -> int4 {
    int4 m = make_int4(0, 0, 0, 0);
    if (g < p.B) { m = meta[g]; if (m.y > p.max_nodes) m.y = p.max_nodes; }
    return m;
  };
  // one thread's share of a 128-row tile: NQ quads, quad column fixed per thread when Q divides kThreads
  float4 pre[NQ];
  auto load_tile = [&](const int4& m, int r0) {
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int idx = tid + i * kThreads;
      const int r = idx / Q, k = (idx - r * Q) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < m.y) {
        const float* src = p.t_in + ((long long)m.x + r0 + r) * K + k;
        if (p.vec_in) {
          if (k < K) v = *reinterpret_cast<const float4*>(src);
        } else {
          if (k + 0 < K) v.x = src[0];
          if (k + 1 < K) v.y = src[1];
          if (k + 2 < K) v.z = src[2];
          if (k + 3 < K) v.w = src[3];
        }
      }
      pre[i] = v;
    }
  };

  long long g = blockIdx.x;
  int4 cur = load_meta(g);
  load_tile(cur, 0);
  uint32_t phase = 0;

  while (g < p.B) {
    const long long g_next = g + gridDim.x;
    const int4 nxt = load_meta(g_next);
    const long long nb = cur.x;
    const int n = cur.y, eb = cur.z, m = cur.w;
    const bool csr_here = m <= p.max_edges;
    if (csr_here) stage_csr_async(s_csr, p.max_nodes, p.max_edges, p.in_rowptr, p.in_col, p.in_wn, p.dinv, nb, n, eb, m);
    cp_async_commit();

    for (int r0 = 0; r0 < n; r0 += kTcM) {
      const int rows = min(kTcM, n - r0);
      // (1) previous layer's BatchNorm/ReLU/dropout on the register tile, hi/lo split, swizzled operand stores
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int idx = tid + i * kThreads;
        const int r = idx / Q, k = (idx - r * Q) << 2;
        float4 v = pre[i];
        if (r < rows && k < K) {
          const uint32_t rh = p.act.drop ? drop_row_hash(p.act, p.act.row_base + nb + r0 + r) : 0u;
          v.x = act_fwd(p.act, affine, v.x, s_scale[k + 0], s_shift[k + 0], rh, k + 0);
          v.y = k + 1 < K ? act_fwd(p.act, affine, v.y, s_scale[k + 1], s_shift[k + 1], rh, k + 1) : 0.0f;
          v.z = k + 2 < K ? act_fwd(p.act, affine, v.z, s_scale[k + 2], s_shift[k + 2], rh, k + 2) : 0.0f;
          v.w = k + 3 < K ? act_fwd(p.act, affine, v.w, s_scale[k + 3], s_shift[k + 3], rh, k + 3) : 0.0f;
        } else {
          v = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tc::store_split4(a_hi, a_lo, r, k, kTcM, v);
      }
      tc::fence_proxy_async();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();
      // (2) one thread drives the tensor core: 3 x (KP / 8) MMAs of 128 x H x 8
      if (tid == 0) {
        tc::mma_tf32x3(taddr, a_hi_u, a_lo_u, b_hi_u, b_lo_u, kTcM, H, KP);
        tc::mma_commit(&mbar);
      }
      // (3) next tile's global loads fly during the MMA, the epilogue and (after the last tile) the aggregation
      if (r0 + kTcM < n) load_tile(cur, r0 + kTcM);
      else load_tile(nxt, 0);
      // (4) accumulators: TMEM -> registers -> the subject's P tile
      tc::mbar_wait(&mbar, phase);
      phase ^= 1;
      tc::fence_after_sync();
      {
        const int rr = 32 * (warp & 3) + lane;            // row of the tile this thread owns
        const int half = warp >> 2;                       // which half of the columns (H >= 64), else warps 4-7 idle
        const int ncol_w = H >= 64 ? H / 2 : H;
        const int c_begin = H >= 64 ? half * ncol_w : 0;
        const int c_end = ((H >= 64 && half < 2) || half == 0) ? c_begin + ncol_w : 0;
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          float v[32];
          tc::tmem_ld32(taddr + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, v);
          if (rr < rows) {
            float* dst = s_p + (r0 + rr) * ldp + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      tc::fence_before_sync();   // orders the tcgen05.ld above before the next tile's MMAs (after the next barrier)
    }
    if (n == 0) load_tile(nxt, 0);   // empty subject: still hand the register tile on to the next one
    cp_async_wait<0>();
    __syncthreads();

    // ---- aggregation: one warp per destination row, VW consecutive channels per lane ---------------
    const float* dinv_g = p.dinv + nb;
    RowCsr rc{p.in_rowptr + nb, p.in_col, p.in_wn};
    if (csr_here) rc = staged_csr(s_csr, p.max_nodes, p.max_edges, eb, &dinv_g);
    const float* p_lane = s_p + VW * lane;
    float bias_l[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) bias_l[j] = s_bias[VW * lane + j];
    for (int i = warp; i < n; i += kWarps) {
      float acc[VW];
#pragma unroll
      for (int j = 0; j < VW; ++j) acc[j] = 0.0f;
      const int e0 = rc.rp[i], e1 = rc.rp[i + 1];
      for (int e = e0; e < e1; ++e) {
        const int c = (int)(rc.col[e] - nb);
        const float w = rc.w[e];
        const float* src = p_lane + c * ldp;
        if (VW == 2) {
          const float2 v = *reinterpret_cast<const float2*>(src);
          acc[0] = fmaf(v.x, w, acc[0]); acc[1] = fmaf(v.y, w, acc[1]);
        } else if (VW == 4) {
          const float4 v = *reinterpret_cast<const float4*>(src);
          acc[0] = fmaf(v.x, w, acc[0]); acc[1] = fmaf(v.y, w, acc[1]);
          acc[2] = fmaf(v.z, w, acc[2]); acc[3] = fmaf(v.w, w, acc[3]);
        } else {
#pragma unroll
          for (int j = 0; j < VW; ++j) acc[j] = fmaf(src[j], w, acc[j]);
        }
      }
      const float d = dinv_g[i];
      const float wself = d * d;
      float inv;
      st.begin_row(inv);
      float out[VW];
#pragma unroll
      for (int j = 0; j < VW; ++j) {
        out[j] = fmaf(p_lane[i * ldp + j], wself, acc[j]) + bias_l[j];
        st.w[j].push(out[j], inv);
      }
      float* zr = p.z + (nb + i) * H + VW * lane;
      if (VW == 2) *reinterpret_cast<float2*>(zr) = make_float2(out[0], out[1]);
      else if (VW == 4) *reinterpret_cast<float4*>(zr) = make_float4(out[0], out[1], out[2], out[3]);
      else {
#pragma unroll
        for (int j = 0; j < VW; ++j) zr[j] = out[j];
      }
    }
    __syncthreads();   // P tile and staged CSR are rewritten by the next subject
    g = g_next;
    cur = nxt;
  }

  if (p.partials) {
    // WarpStats keeps channel = lane + 32*j; here a lane owns channels VW*lane + j: deposit accordingly
    if (lane == 0) s_cnt[warp] = (float)st.rows;
#pragma unroll
    for (int j = 0; j < VW; ++j) {
      s_mean[warp * H + VW * lane + j] = st.w[j].mean;
      s_m2[warp * H + VW * lane + j] = st.w[j].m2;
    }
    __syncthreads();
    cta_write_stats(s_cnt, s_mean, s_m2, H, H, p.partials + (size_t)blockIdx.x * (1 + 2 * H));
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(taddr, p.tmem_cols);
}

// Returns CGNN_OK when the tensor-core kernel was launched, -1 when the shape is not eligible (caller falls
// back to the generic kernel), or an error status.
int launch_gcn_fwd_tc(const float* t_in, const cgnn_act_t* act, const float* W, const float* bias, const cgnn_csr_t* csr,
                      int64_t num_graphs, int32_t d_in, int32_t H, int32_t max_nodes, int32_t max_edges, float* z,
                      double* partials, int* grid_out, size_t workspace_bytes, cudaStream_t stream) {
  if (H % 32 != 0 || H > 128 || d_in > 128) return -1;
  const int KB = (d_in + 31) / 32, VW = H / 32;
  if (KB == 3 || VW == 3) return -1;   // instantiated for 1, 2, 4 only
  if ((((uintptr_t)z) & 15u) != 0) return -1;
  const DeviceInfo dev = device_info();
  GcnTcArgs a;
  a.t_in = t_in; a.act = make_act(act); a.W = W; a.bias = bias;
  a.in_rowptr = csr->in_rowptr; a.in_col = csr->in_col; a.in_wn = csr->in_wn; a.dinv = csr->dinv;
  a.meta = csr->graph_meta; a.B = num_graphs;
  a.K = d_in; a.H = H; a.ldp = H + 4;
  a.max_nodes = max_nodes < 1 ? 1 : max_nodes;
  a.max_edges = max_edges;
  a.vec_in = (d_in % 4 == 0) && ((((uintptr_t)t_in) & 15u) == 0);
  a.z = z; a.partials = partials;
  a.tmem_cols = 32;
  while (a.tmem_cols < (uint32_t)H) a.tmem_cols <<= 1;
  const int KP = 32 * KB;
  int off = 0;
  a.o_ahi = off; off += KB * kTcM * 128;
  a.o_alo = off; off += KB * kTcM * 128;
  a.o_bhi = off; off += KB * H * 128;
  a.o_blo = off; off += KB * H * 128;
  a.o_p = off; off += round_up(a.max_nodes * a.ldp * 4, 16);
  a.o_csr = off; off += csr_words(a.max_nodes, a.max_edges) * 4;
  a.o_scale = off; off += KP * 4;
  a.o_shift = off; off += KP * 4;
  a.o_bias = off; off += H * 4;
  a.o_st = off; off += round_up((kWarps + 2 * kWarps * H) * 4, 16);
  const size_t smem = (size_t)off + 1024;
  if (smem > (size_t)dev.smem_optin) return -1;
  int grid = persistent_grid(num_graphs, smem, dev, kThreads);
  // tensor memory: 512 columns per SM shared by the resident CTAs
  const int by_tmem = 512 / (int)a.tmem_cols;
  const int per_sm = (grid + dev.sm_count - 1) / dev.sm_count;
  if (per_sm > by_tmem) grid = by_tmem * dev.sm_count;
  if (partials) {
    const size_t rec = (size_t)(1 + 2 * H) * sizeof(double);
    if ((size_t)grid * rec > workspace_bytes) grid = (int)(workspace_bytes / rec);
  }
  *grid_out = grid;
#define CGNN_TC_LAUNCH(KB_, VW_)                                                                    \
  {                                                                                                 \
    auto kfn = k_gcn_fwd_tc<KB_, VW_>;                                                              \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    CGNN_LAUNCH(kfn, grid, kThreads, smem, stream, a);                                              \
  }
  if (KB == 1) { if (VW == 1) CGNN_TC_LAUNCH(1, 1) else if (VW == 2) CGNN_TC_LAUNCH(1, 2) else CGNN_TC_LAUNCH(1, 4) }
  else if (KB == 2) { if (VW == 1) CGNN_TC_LAUNCH(2, 1) else if (VW == 2) CGNN_TC_LAUNCH(2, 2) else CGNN_TC_LAUNCH(2, 4) }
  else { if (VW == 1) CGNN_TC_LAUNCH(4, 1) else if (VW == 2) CGNN_TC_LAUNCH(4, 2) else CGNN_TC_LAUNCH(4, 4) }
#undef CGNN_TC_LAUNCH
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

#endif  // CGNN_EMU
}  // namespace cgnn
