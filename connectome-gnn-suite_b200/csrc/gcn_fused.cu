// gcn_fused.cu - GCN layer forward in one kernel, aggregate-first (sm_100a only).
//
//   z = (A^ u) W^T + b,   u = dropout(relu(bn(t_in)))          reference models.py:84-114 after models.py:208-210
//
// A^ (u W^T) and (A^ u) W^T are the same product; aggregating first keeps the gather in the layer's INPUT width (5
// channels for the first layer instead of 64) and lets the gathered rows go straight from registers into the tensor
// core operand, so the projected tile never exists in shared memory and nothing but t_in, the structure blob and z
// crosses HBM.  One persistent CTA per SM, one subject at a time:
//
//   tile     the subject's input rows, activation applied, fp32, in shared memory (written from registers that were
//            loaded while the previous subject was being processed)
//   blob     the subject's packed in-edge records incl. the self loop (agg.cuh), by cp.async
//   per 128 rows:  gather (32 / LPR rows per warp, float4 lanes) -> registers      | runs while the previous row
//                  wait for the previous MMA, drain its accumulators, bias, store z, statistics            tile's MMA
//                  registers -> hi/lo TF32 split -> K-major swizzled operand, tcgen05.mma x3                is in flight
#include "agg.cuh"
#include "rowtile.cuh"
#include "tile.cuh"

namespace cgnn {
#ifndef CGNN_EMU

struct GcnFusedArgs {
  const float* t_in; Act act; const float* W; const float* bias;
  const int32_t* blob; const int32_t* meta; long long B;
  int K, H, max_nodes, max_edges;
  float* z; double* partials;
  uint32_t tmem_cols;
  int o_tile, o_blob, o_stage, o_rec;   // byte offsets from the 1024-aligned base (after the A and W operands)
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// KB = padded K / 32, HB = H / 32, LPR = float4 lanes per tile row (K / 4 when K is a multiple of 32, else 2: K <= 8)
template <int KB, int HB, int LPR, bool VEC>
__global__ void __launch_bounds__(kThreads, 1) k_gcn_fwd_fused(GcnFusedArgs p) {
  constexpr int KP = 32 * KB, H = 32 * HB, TR = 128;
  constexpr int A_HALF = KB * TR * 128, B_HALF = KB * H * 128;
  constexpr int RP = kThreads / LPR;                    // tile rows per load pass
  constexpr int NPF = (384 * LPR + kThreads - 1) / kThreads;   // register prefetch depth: subjects of <= 384 rows
  constexpr int RPW = 32 / LPR;                         // rows per warp per gather group
  constexpr int NG = (TR / RPW + kWarps - 1) / kWarps;  // gather groups per warp per 128-row tile
  constexpr int QH = H / 4;
  using MH = rt::QuadMap<QH, TR>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* base = tc::smem_align1024(smem_raw);
  unsigned char* a_hi = base;
  unsigned char* a_lo = a_hi + A_HALF;
  unsigned char* b_hi = a_lo + A_HALF;
  unsigned char* b_lo = b_hi + B_HALF;
  float4* s_tile = reinterpret_cast<float4*>(base + p.o_tile);      // [max_nodes][LPR]
  int32_t* s_blob = reinterpret_cast<int32_t*>(base + p.o_blob);
  float4* stage = reinterpret_cast<float4*>(base + p.o_stage);      // [128][H] swizzled; aliases the A operand when it can
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = p.K;
  const int cl = tid % LPR, rl = tid / LPR;
  const int c0 = 4 * cl;
  const bool live_quad = c0 < K;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, p.tmem_cols);
  if (tid == 0) tc::mbar_init(&mbar, 1);
  // W [H][K] -> K-major operand (N = H rows), zero padded to KP
  for (int idx = tid; idx < H * (KP / 4); idx += kThreads) {
    const int n = idx / (KP / 4), q = idx - n * (KP / 4);
    float4 v = rt::mask_quad(rt::ld_quad<false>(p.W, n, K, 4 * q), 4 * q, K);
    float4 h, l;
    rt::split4(v, h, l);
    const uint32_t off = rt::kmajor_quad_offset(n, q, H);
    rt::sts4(b_hi + off, h);
    rt::sts4(b_lo + off, l);
  }
  if (4 * LPR < KP) {   // padded channels of the A operand stay zero (staging does not alias it in this case)
    for (int i = tid; i < 2 * A_HALF / 16; i += kThreads) reinterpret_cast<float4*>(a_hi)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  rt::ChanQuad cq;
  rt::chan_quad_init(cq, p.act, c0, K);
  const rt::RowKey rk = rt::row_key(p.act);
  const int qh = tid % QH, rh = tid / QH;
  float bias4[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bias4[j] = p.bias ? p.bias[4 * qh + j] : 0.0f;
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  const uint32_t a_hi_u = tc::smem_u32(a_hi), a_lo_u = tc::smem_u32(a_lo), b_hi_u = tc::smem_u32(b_hi), b_lo_u = tc::smem_u32(b_lo);
  const uint32_t idesc = tc::idesc_tf32(TR, H);
  const int ksteps = (K + 7) / 8;
  const rt::OperandDescs od = rt::kmajor_descs(a_hi_u, a_lo_u, b_hi_u, b_lo_u);

  const int4* meta = reinterpret_cast<const int4*>(p.meta);
  auto load_meta = [&](long long g) -> int4 {
    int4 m = make_int4(0, 0, 0, 0);
    if (g < p.B) { m = meta[g]; m.y = min(m.y, p.max_nodes); m.w = min(m.w, p.max_edges); }
    return m;
  };
  float4 pre[NPF];
  auto prefetch_tile = [&](const int4& m) {     // this thread's quads of a subject's rows -> registers
#pragma unroll
    for (int i = 0; i < NPF; ++i) {
      const int r = rl + i * RP;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < m.y && live_quad) v = rt::ld_quad<VEC>(p.t_in, (long long)m.x + r, K, c0);
      pre[i] = v;
    }
  };
  auto start_blob = [&](const int4& m, long long g) {
    const int32_t* gb = p.blob + agg_base_words(m.x, m.z, g);
    const int n16 = (agg_copy_words(m.y, m.w) + 3) >> 2;
    for (int i = tid; i < n16; i += kThreads) cp_async_16(s_blob + 4 * i, gb + 4 * i);
    cp_async_commit();
  };
  auto store_tile = [&](const int4& m) {        // registers -> activation -> shared tile (rows beyond NPF passes: direct)
#pragma unroll
    for (int i = 0; i < NPF; ++i) {
      const int r = rl + i * RP;
      if (r < m.y) {
        float4 u = rt::act_fwd4(p.act, cq, pre[i], rk, (uint32_t)(m.x + r));
        if (!VEC) u = rt::mask_quad(u, c0, K);
        s_tile[r * LPR + cl] = u;
      }
    }
    for (int r = rl + NPF * RP; r < m.y; r += RP) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live_quad) v = rt::ld_quad<VEC>(p.t_in, (long long)m.x + r, K, c0);
      s_tile[r * LPR + cl] = rt::mask_quad(rt::act_fwd4(p.act, cq, v, rk, (uint32_t)(m.x + r)), c0, K);
    }
  };

  int cnt = 0;
  const bool want_stats = p.partials != nullptr;
  Welford wf[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) wf[j].init();
  uint32_t phase = 0;
  bool pending = false;
  long long pend_row0 = 0;   // first global row of the row tile whose MMA is in flight
  int pend_rows = 0;
  // drain + epilogue of the row tile whose MMA is in flight
  auto finish_pending = [&]() {
    tc::mbar_wait(&mbar, phase);
    phase ^= 1;
    tc::fence_after_sync();
    rt::drain_rows_to_staging<H>(taddr, stage, warp, lane);
    tc::fence_before_sync();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MH::NQ; ++i) {
      const int r = rh + i * MH::RS;
      if (r < pend_rows) {
        float4 v = stage[rt::stage_index(r, qh, QH)];
        v.x += bias4[0]; v.y += bias4[1]; v.z += bias4[2]; v.w += bias4[3];
        *reinterpret_cast<float4*>(p.z + (pend_row0 + r) * H + 4 * qh) = v;
        if (want_stats) {      // BatchNorm batch statistics: training only
          cnt += 1;
          const float inv = rt::rcp_fast((float)cnt);
          wf[0].push(v.x, inv); wf[1].push(v.y, inv); wf[2].push(v.z, inv); wf[3].push(v.w, inv);
        }
      }
    }
    __syncthreads();   // staging (= A operand) is rewritten next
    pending = false;
  };

  long long g = blockIdx.x;
  int4 cur = load_meta(g);
  prefetch_tile(cur);
  if (g < p.B) start_blob(cur, g);
  while (g < p.B) {
    const long long g_next = g + gridDim.x;
    const int4 nxt = load_meta(g_next);
    // the subject's tile: registers -> shared memory; the next subject's rows start to fly
    store_tile(cur);
    prefetch_tile(nxt);
    if (g_next < p.B) {   // pull the next blob towards L2 while this subject is processed
      const char* nb_ = reinterpret_cast<const char*>(p.blob + agg_base_words(nxt.x, nxt.z, g_next));
      const int lines = (agg_copy_words(nxt.y, nxt.w) * 4 + 127) >> 7;
      for (int i = tid; i < lines; i += kThreads) prefetch_l2(nb_ + 128 * i);
    }
    if (pending) finish_pending();     // the previous subject's last row tile (its MMA ran during the stores above)
    cp_async_wait<0>();
    __syncthreads();
    const long long nb = cur.x;
    const int n = cur.y;
    const int4* s_desc = reinterpret_cast<const int4*>(s_blob);
    const int4* s_rec2 = reinterpret_cast<const int4*>(s_blob + 4 * n);
    for (int r0 = 0; r0 < n; r0 += TR) {
      // (1) gather this row tile into registers
      float4 accs[NG];
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        const int gi = warp + kWarps * k;
        float aux;
        int row;
        accs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gi * RPW < TR) agg_gather_group<LPR>(s_desc, s_rec2, s_tile, r0 + gi * RPW, n, accs[k], aux, row);
      }
      // (2) the previous row tile: accumulators -> z
      if (pending) finish_pending();
      // (3) gathered rows -> hi/lo operand, MMA
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        const int gi = warp + kWarps * k;
        if (gi * RPW < TR && live_quad) {
          const int rt_ = gi * RPW + lane / LPR;     // row of the tile
          float4 h, l;
          rt::split4(accs[k], h, l);
          const uint32_t off = rt::kmajor_quad_offset(rt_, cl, TR);
          rt::sts4(a_hi + off, h);
          rt::sts4(a_lo + off, l);
        }
      }
      tc::fence_proxy_async();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();
      if (tid == 0) {
        rt::issue_kmajor_x3<KP / 8, TR, H>(taddr, od, ksteps, idesc, false);
        tc::mma_commit(&mbar);
      }
      pending = true;
      pend_row0 = nb + r0;
      pend_rows = min(TR, n - r0);
    }
    // all gathers of this subject are done (every warp passed the barrier above): tile and blob may be rewritten
    if (g_next < p.B) start_blob(nxt, g_next);
    g = g_next;
    cur = nxt;
  }
  if (pending) finish_pending();

  if (p.partials) {
    float* rec = reinterpret_cast<float*>(base + p.o_rec);   // [kThreads][9]
    rec[tid * 9] = (float)cnt;
#pragma unroll
    for (int j = 0; j < 4; ++j) { rec[tid * 9 + 1 + j] = wf[j].mean; rec[tid * 9 + 5 + j] = wf[j].m2; }
    __syncthreads();
    double* out = p.partials + (size_t)blockIdx.x * (1 + 2 * H);
    for (int c = tid; c < H; c += kThreads) {
      const int q = c >> 2, j = c & 3;
      double n = 0.0, mean = 0.0, m2 = 0.0;
      for (int th = q; th < kThreads; th += QH) {
        const double nb_ = (double)rec[th * 9];
        if (nb_ <= 0.0) continue;
        const double mb = (double)rec[th * 9 + 1 + j], qb = (double)rec[th * 9 + 5 + j];
        const double nt = n + nb_, delta = mb - mean;
        mean += delta * (nb_ / nt);
        m2 += qb + delta * delta * (n * nb_ / nt);
        n = nt;
      }
      out[1 + c] = mean;
      out[1 + H + c] = m2;
      if (c == 0) out[0] = n;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(taddr, p.tmem_cols);
}

// Returns CGNN_OK when launched (grid in *grid_out: the caller merges `partials`), -1 when the shape is not covered.
int launch_gcn_fwd_fused(const float* t_in, const cgnn_act_t* act, const float* W, const float* bias, const cgnn_csr_t* csr,
                         int64_t num_graphs, int32_t d_in, int32_t H, int32_t max_nodes, int32_t max_edges, float* z,
                         double* partials, int* grid_out, size_t workspace_bytes, cudaStream_t stream) {
  if (!csr->agg_in || csr->agg_kind != AGG_GCN) return -1;
  if (H != 32 && H != 64 && H != 128) return -1;
  int LPR;
  if (d_in == 32 || d_in == 64) LPR = d_in / 4;
  else if (d_in >= 1 && d_in <= 8) LPR = 2;
  else return -1;
  if ((((uintptr_t)z) & 15u) != 0) return -1;
  const int KB = (d_in + 31) / 32, HB = H / 32;
  const DeviceInfo dev = device_info();
  GcnFusedArgs a;
  a.t_in = t_in; a.act = make_act(act); a.W = W; a.bias = bias;
  a.blob = csr->agg_in; a.meta = csr->graph_meta; a.B = num_graphs;
  a.K = d_in; a.H = H;
  a.max_nodes = max_nodes < 1 ? 1 : max_nodes;
  a.max_edges = max_edges;
  a.z = z; a.partials = partials;
  a.tmem_cols = 32;
  while (a.tmem_cols < (uint32_t)H) a.tmem_cols <<= 1;
  const int vec = (d_in % 4 == 0) && ((((uintptr_t)t_in) & 15u) == 0);
  const size_t a_bytes = (size_t)2 * KB * 128 * 128, b_bytes = (size_t)2 * KB * H * 128;
  size_t stage_bytes = (size_t)128 * H * 4;
  if (stage_bytes < (size_t)kThreads * 9 * 4) stage_bytes = (size_t)kThreads * 9 * 4;
  size_t off = a_bytes + b_bytes;
  a.o_tile = (int)off; off += (size_t)a.max_nodes * LPR * 16;
  a.o_blob = (int)off; off += (size_t)agg_smem_words(a.max_nodes, a.max_edges) * 4;
  if (4 * LPR == 32 * KB && stage_bytes <= a_bytes) a.o_stage = 0;     // every operand quad is rewritten per tile: alias
  else { a.o_stage = (int)off; off += stage_bytes; }
  a.o_rec = a.o_stage;
  const size_t smem = off + 1024;
  if (smem > (size_t)dev.smem_optin) return -1;
  long long grid = dev.sm_count;
  if (grid > num_graphs) grid = num_graphs;
  if (partials) {
    const size_t rec = (size_t)(1 + 2 * H) * sizeof(double);
    if ((size_t)grid * rec > workspace_bytes) grid = (long long)(workspace_bytes / rec);
  }
  if (grid < 1) return -1;
  *grid_out = (int)grid;
#define CGNN_GF(KB_, HB_, LPR_, VEC_)                                                                 \
  {                                                                                                   \
    auto kfn = k_gcn_fwd_fused<KB_, HB_, LPR_, VEC_>;                                                 \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                \
    CGNN_LAUNCH(kfn, (unsigned)grid, kThreads, smem, stream, a);                                      \
  }
#define CGNN_GF_H(KB_, LPR_, VEC_) { if (HB == 1) CGNN_GF(KB_, 1, LPR_, VEC_) else if (HB == 2) CGNN_GF(KB_, 2, LPR_, VEC_) else CGNN_GF(KB_, 4, LPR_, VEC_) }
  if (LPR == 2) { if (vec) CGNN_GF_H(1, 2, true) else CGNN_GF_H(1, 2, false) }
  else if (!vec) return -1;
  else if (LPR == 8) CGNN_GF_H(1, 8, true) else CGNN_GF_H(2, 16, true)
#undef CGNN_GF_H
#undef CGNN_GF
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

#endif  // CGNN_EMU
}  // namespace cgnn
