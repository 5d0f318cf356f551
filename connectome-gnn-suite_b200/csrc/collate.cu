// collate.cu - K0: batch collate + per-subject CSR, integer-exact.
//
// Replaces reference graph.py:143-167 (collate_graphs: concatenation, node-id offsets, `batch`,
// `ptr`) and hoists the per-layer structure work of models.py:94-108 (self-loop weights, D^,
// d^-1/2, normalised weights) and models.py:147-148 (w_sum) to once per batch.
//
// One CTA per subject.  The CSR is a *stable* counting sort of the subject's COO edges (order
// inside a row = COO order) so that the sequential fp32 row sums below reproduce the CPU
// scatter_add_ of the reference bit for bit.
#include "common.cuh"
#include "agg.cuh"

namespace cgnn {

// ---- prefix sums of the selected subjects' node / edge counts ---------------------------
__global__ void __launch_bounds__(1024) k_scan_ptrs(const long long* __restrict__ node_ptr,
                                                    const long long* __restrict__ edge_ptr,
                                                    const long long* __restrict__ ids, long long B,
                                                    long long* __restrict__ ptr, long long* __restrict__ eptr) {
  __shared__ long long s_wn[32], s_we[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
  const long long per = (B + nthreads - 1) / nthreads;
  long long lo = (long long)tid * per; if (lo > B) lo = B;
  long long hi = lo + per; if (hi > B) hi = B;
  long long sn = 0, se = 0;
  for (long long g = lo; g < hi; ++g) {
    long long s = ids[g];
    sn += node_ptr[s + 1] - node_ptr[s];
    se += edge_ptr[s + 1] - edge_ptr[s];
  }
  long long in = sn, ie = se;
  for (int o = 1; o < 32; o <<= 1) {
    long long vn = __shfl_up_sync(kFull, in, o), ve = __shfl_up_sync(kFull, ie, o);
    if (lane >= o) { in += vn; ie += ve; }
  }
  if (lane == 31) { s_wn[warp] = in; s_we[warp] = ie; }
  __syncthreads();
  if (warp == 0) {
    int nw = nthreads >> 5;
    long long vn = lane < nw ? s_wn[lane] : 0, ve = lane < nw ? s_we[lane] : 0;
    long long cn = vn, ce = ve;
    for (int o = 1; o < 32; o <<= 1) {
      long long an = __shfl_up_sync(kFull, cn, o), ae = __shfl_up_sync(kFull, ce, o);
      if (lane >= o) { cn += an; ce += ae; }
    }
    s_wn[lane] = cn - vn;
    s_we[lane] = ce - ve;
  }
  __syncthreads();
  long long rn = s_wn[warp] + in - sn, re = s_we[warp] + ie - se;
  if (tid == 0) { ptr[0] = 0; eptr[0] = 0; }
  for (long long g = lo; g < hi; ++g) {
    long long s = ids[g];
    rn += node_ptr[s + 1] - node_ptr[s];
    re += edge_ptr[s + 1] - edge_ptr[s];
    ptr[g + 1] = rn;
    eptr[g + 1] = re;
  }
}

// Subject indices from PINNED HOST memory (dereferenced over the bus: pinned allocations are mapped into the device's
// address space) into a device buffer.  A cudaMemcpyAsync of these few KB would queue behind whatever the host-to-device
// copy engine is doing - with dataset arenas streaming in (StreamingStore) that is a 77 MB upload per step leg, and the
// compute stream would wait for it; a kernel read does not touch the copy engine.
__global__ void __launch_bounds__(256) k_fetch_ids(const long long* __restrict__ host_ids, long long n, long long* __restrict__ ids) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = host_ids[i];
}

// Edge ranges of an already collated COO batch: edges are grouped by subject, so
// eptr[g] = first edge whose source id is >= ptr[g].
__global__ void __launch_bounds__(256) k_edge_ranges(const long long* __restrict__ src, long long E,
                                                     const long long* __restrict__ ptr, long long B,
                                                     long long* __restrict__ eptr) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g > B) return;
  if (g == B) { eptr[B] = E; return; }
  long long key = ptr[g], lo = 0, hi = E;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (src[mid] < key) lo = mid + 1; else hi = mid;
  }
  eptr[g] = lo;
}

struct CollateArgs {
  // source: subject store (FROM_STORE) or collated COO
  cgnn_store_t store;
  const long long* ids;
  const long long* coo;        // [2, E] global ids (from-COO path)
  const float* coo_w;
  long long B, total_rows, total_edges;
  int max_nodes;               // shared memory was sized for this many nodes per subject
  int edge_cap;                // ... and for staging the sorted CSR of subjects with up to this many edges (0: none)
  // reference-visible outputs (FROM_STORE only)
  float* x; long long* edge_index; float* edge_weight; long long* batch; long long* labels;
  const long long* ptr; const long long* eptr;
  cgnn_csr_out_t csr;
};

// In-place exclusive scan of a[0..n) in shared memory by the whole CTA.
__device__ __forceinline__ void block_excl_scan(int* a, int n, int* s_warp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
  const int per = (n + nthreads - 1) / nthreads;
  int lo = min(tid * per, n), hi = min(lo + per, n);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += a[i];
  int inc = s;
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int nw = nthreads >> 5;
    int v = lane < nw ? s_warp[lane] : 0, c = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(kFull, c, o);
      if (lane >= o) c += t;
    }
    s_warp[lane] = c - v;
  }
  __syncthreads();
  int run = s_warp[warp] + inc - s;
  for (int i = lo; i < hi; ++i) { int t = a[i]; a[i] = run; run += t; }
  __syncthreads();
}

// NT threads per CTA: 512 for 360-node subjects; 128 for small subjects (<= 128 nodes), where the per-row phases would
// leave most of a 512-thread CTA waiting at barriers - four times as many CTAs are resident instead.
template <bool FROM_STORE, int NT>
__global__ void __launch_bounds__(NT) k_collate_graph(CollateArgs p) {
  constexpr int kNW = NT / 32;
  CGNN_SMEM_DECL;
  __shared__ int s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long g = blockIdx.x;
  const long long nb = p.ptr[g], eb = p.eptr[g];
  const int n = (int)(p.ptr[g + 1] - nb);
  const int m = (int)(p.eptr[g + 1] - eb);
  if (tid == 0) reinterpret_cast<int4*>(p.csr.graph_meta)[g] = make_int4((int)nb, n, (int)eb, m);
  if (n > p.max_nodes) return;  // host contract violated: never index past the shared arrays

  int* cur_in = reinterpret_cast<int*>(cgnn_smem);   // [n] counts -> row starts -> cursors
  int* cur_out = cur_in + n;                          // [n]
  float* s_dinv = reinterpret_cast<float*>(cur_out + n);  // [n]
  float* s_wsum = s_dinv + n;                             // [n]

  const int32_t* lsrc = nullptr; const int32_t* ldst = nullptr; const float* lw = nullptr;
  const long long* gsrc = nullptr; const long long* gdst = nullptr;
  if (FROM_STORE) {
    const long long sid = p.ids[g];
    const long long sn = p.store.node_ptr[sid], se = p.store.edge_ptr[sid];
    const long long so = p.store.edge_pairs ? (se >> 1) : se;     // pair store: one entry per undirected edge
    lsrc = p.store.src + so; ldst = p.store.dst ? p.store.dst + so : nullptr; lw = p.store.w + so;
    const int F = p.store.num_features;
    const float* sx = p.store.x + sn * F;
    float* dx = p.x + nb * F;
    for (int i0 = tid; i0 < n * F; i0 += 4 * NT) {     // four loads in flight per thread
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = i0 + u * NT < n * F ? sx[i0 + u * NT] : 0.0f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * NT < n * F) dx[i0 + u * NT] = v[u];
    }
    if (p.batch)
      for (int i = tid; i < n; i += NT) p.batch[nb + i] = g;
    if (tid == 0 && p.labels && p.store.label) p.labels[g] = p.store.label[sid];
  } else {
    gsrc = p.coo + eb; gdst = p.coo + p.total_edges + eb; lw = p.coo_w + eb;
  }
  constexpr int kChunks = kNW / 2;
  int* cnt = reinterpret_cast<int*>(s_wsum + n);        // [2][kChunks][n]
  // Subjects that fit are sorted entirely in shared memory: the raw COO list (src | dst << 16, w) is read from
  // global memory once, both sorted lists live next to it, and the row sums, normalised weights and CSR arrays are
  // produced from shared memory with coalesced stores.  Larger subjects re-read global memory (same results).
  const bool staged = m <= p.edge_cap && n <= 65535;
  uint32_t* pk = reinterpret_cast<uint32_t*>(cnt + 2 * kChunks * n);    // [2][m] sorted entries
  float* cw = reinterpret_cast<float*>(pk + 2 * m);                      // [2][m]
  uint32_t* raw_pk = reinterpret_cast<uint32_t*>(cw + 2 * m);            // [m] COO order
  float* raw_w = reinterpret_cast<float*>(raw_pk + m);                   // [m]
  const bool want_agg = staged && p.csr.agg_kind >= 0 && p.csr.agg_in && p.csr.row_graph;   // agg_out is optional
  // lean batch: none of the CSR arrays is wanted (the layer kernels read the blobs only); needs the staged path
  const bool lean = p.csr.in_col == nullptr;
  if (lean && !staged) return;
  // Local endpoints of edge e as stored.  Endpoints outside the subject (malformed hand-built batches)
  // are redirected to a zero-weight self edge on node 0 so the CSR stays consistent.
  const bool pair_store = FROM_STORE && p.store.edge_pairs != 0;
  auto load_edge = [&](int e, int& s, int& d, float& w) {
    if (FROM_STORE) {
      const int ee = pair_store ? (e >> 1) : e;
      const uint32_t v = (uint32_t)lsrc[ee];
      if (ldst) { s = (int)v; d = ldst[ee]; } else { s = (int)(v & 0xffffu); d = (int)(v >> 16); }   // both endpoints in one word
      if (pair_store && (e & 1)) { const int t = s; s = d; d = t; }                                 // the reverse edge of the pair
      w = lw[ee];
    }
    else { s = (int)(gsrc[e] - nb); d = (int)(gdst[e] - nb); w = lw[e]; }
  };
  auto edge = [&](int e, int& s, int& d, float& w) {
    if (staged) {
      const uint32_t v = raw_pk[e];
      s = (int)(v & 0xffffu); d = (int)(v >> 16); w = raw_w[e];
      return;
    }
    load_edge(e, s, d, w);
    if ((unsigned)s >= (unsigned)n || (unsigned)d >= (unsigned)n) { s = 0; d = 0; w = 0.0f; }
  };
  // one pass over the subject's COO list in global memory: reference-visible outputs and the shared-memory copy
  if (FROM_STORE || staged) {
    for (int e0 = tid; e0 < m; e0 += 3 * NT) {           // three edges in flight per thread
      int es[3], ed[3]; float ew[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        es[u] = 0; ed[u] = 0; ew[u] = 0.0f;
        if (e0 + u * NT < m) load_edge(e0 + u * NT, es[u], ed[u], ew[u]);
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int e = e0 + u * NT;
        if (e >= m) continue;
        int s = es[u], d = ed[u]; float w = ew[u];
        if (FROM_STORE && p.edge_index) {
          p.edge_index[eb + e] = (long long)s + nb;
          p.edge_index[p.total_edges + eb + e] = (long long)d + nb;
          p.edge_weight[eb + e] = w;
        }
        if (staged) {
          if ((unsigned)s >= (unsigned)n || (unsigned)d >= (unsigned)n) { s = 0; d = 0; w = 0.0f; }
          raw_pk[e] = (uint32_t)s | ((uint32_t)d << 16);
          raw_w[e] = w;
        }
      }
    }
  }

  // Stable counting sort by destination (in-CSR) and by source (out-CSR), all 16 warps: the COO list is cut into
  // kChunks contiguous chunks; warp (dir, chunk) counts its chunk's keys, a scan over (key, chunk) turns the counts
  // into the first slot of every (chunk, key) pair, then every warp places its chunk 32 edges at a time - lanes that
  // share a key take consecutive slots in lane order, so the order inside a row is the COO order.
  const int dir = warp / kChunks, chunk = warp % kChunks;
  const int clen = (((m + kChunks - 1) / kChunks) + 31) & ~31;
  const int e_lo = min(chunk * clen, m), e_hi = min(e_lo + clen, m);
  int* my_cnt = cnt + (size_t)(dir * kChunks + chunk) * n;
  for (int i = tid; i < 2 * kChunks * n; i += NT) cnt[i] = 0;
  __syncthreads();
  if (n > 0)
    for (int e = e_lo + lane; e < e_hi; e += 32) {
      int s, d; float w; edge(e, s, d, w);
      atomicAdd(&my_cnt[dir == 0 ? d : s], 1);
    }
  __syncthreads();
  for (int idx = tid; idx < 2 * n; idx += NT) {
    const int dd = idx >= n ? 1 : 0, i = idx - dd * n;   // idx < 2n
    int t = 0;
    for (int c = 0; c < kChunks; ++c) t += cnt[(size_t)(dd * kChunks + c) * n + i];
    (dd == 0 ? cur_in : cur_out)[i] = t;
  }
  __syncthreads();
  block_excl_scan(cur_in, n, s_warp);
  block_excl_scan(cur_out, n, s_warp);
  for (int idx = tid; idx < 2 * n; idx += NT) {
    const int dd = idx >= n ? 1 : 0, i = idx - dd * n;   // idx < 2n
    int run = (dd == 0 ? cur_in : cur_out)[i];
    if (!lean) (dd == 0 ? p.csr.in_rowptr : p.csr.out_rowptr)[nb + i] = (int32_t)(eb + run);
    for (int c = 0; c < kChunks; ++c) {
      int* q = &cnt[(size_t)(dd * kChunks + c) * n + i];
      const int t = *q;
      *q = run;
      run += t;
    }
  }
  if (g == p.B - 1 && tid == 0 && !lean) {
    p.csr.in_rowptr[p.total_rows] = (int32_t)p.total_edges;
    p.csr.out_rowptr[p.total_rows] = (int32_t)p.total_edges;
  }
  __syncthreads();
  if (n > 0) {
    int32_t* col = dir == 0 ? p.csr.in_col : p.csr.out_col;
    float* wv = dir == 0 ? p.csr.in_w : p.csr.out_w;
    for (int e0 = e_lo; e0 < e_hi; e0 += 32) {
      const int e = e0 + lane;
      int s = 0, d = 0; float w = 0.0f;
      const bool live = e < e_hi;
      if (live) edge(e, s, d, w);
      const int key = live ? (dir == 0 ? d : s) : -1 - lane;
      const unsigned peers = __match_any_sync(kFull, key);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      int start = 0;
      if (live) start = my_cnt[key];
      __syncwarp();
      if (live) {
        if (staged) {
          pk[dir * m + start + rank] = (uint32_t)s | ((uint32_t)d << 16);
          cw[dir * m + start + rank] = w;
        } else {
          const long long pos = eb + start + rank;
          col[pos] = (int32_t)(nb + (dir == 0 ? s : d));
          wv[pos] = w;
        }
        if (rank == 0) my_cnt[key] = start + __popc(peers);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // cur_in / cur_out hold the exclusive row starts; the passes below want the inclusive ends (start of row i + 1)
  for (int idx = tid; idx < 2 * n; idx += NT) {
    const int dd = idx >= n ? 1 : 0, i = idx - dd * n;   // idx < 2n
    const int* c = cnt + (size_t)(dd * kChunks + kChunks - 1) * n;
    (dd == 0 ? cur_in : cur_out)[i] = c[i];     // the last chunk's cursor of row i ended at the row's end
  }
  __syncthreads();

  if (staged) {
    for (int i = tid; i < n; i += NT) {
      const int o0 = i ? cur_out[i - 1] : 0, o1 = cur_out[i];
      float deg = 0.0f;
      for (int q = o0; q < o1; ++q) deg = __fadd_rn(deg, cw[m + q]);
      deg = __fadd_rn(deg, 1.0f);
      const int i0 = i ? cur_in[i - 1] : 0, i1 = cur_in[i];
      float ws = 0.0f;
      for (int q = i0; q < i1; ++q) ws = __fadd_rn(ws, cw[q]);
      const float dinv = (float)(1.0 / sqrt((double)__fadd_rn(deg, 1e-8f)));
      s_dinv[i] = dinv;
      s_wsum[i] = ws;
      if (!lean) {
        p.csr.deg[nb + i] = deg;
        p.csr.dinv[nb + i] = dinv;
        p.csr.wsum[nb + i] = ws;
      }
    }
    __syncthreads();
    if (want_agg) {
      // packed aggregation blobs of the requested family, straight from the sorted lists (same bits as k_build_agg)
      const int self = p.csr.agg_kind == AGG_GCN ? 1 : 0;
      const int ndir = p.csr.agg_out ? 2 : 1;           // inference batches ask for the by-destination blob only
      int* pos = cnt;                                   // [2][n] padded record counts -> first record of every row
      for (int idx = tid; idx < ndir * n; idx += NT) {
        const int dd = idx >= n ? 1 : 0, i = idx - dd * n;
        const int* ce = dd == 0 ? cur_in : cur_out;
        pos[idx] = ((ce[i] - (i ? ce[i - 1] : 0)) + self + 1) & ~1;
      }
      __syncthreads();
      for (int dd = 0; dd < ndir; ++dd) block_excl_scan(pos + dd * n, n, s_warp);
      for (int i = tid; i < n; i += NT) p.csr.row_graph[nb + i] = (int32_t)g;
      for (int idx = tid; idx < ndir * n; idx += NT) {
        const int dd = idx >= n ? 1 : 0, i = idx - dd * n;
        const int* ce = dd == 0 ? cur_in : cur_out;
        const int q0 = i ? ce[i - 1] : 0, q1 = ce[i];
        int32_t* blob = (dd == 0 ? p.csr.agg_in : p.csr.agg_out) + agg_base_words(nb, eb, g);
        int2* rec = reinterpret_cast<int2*>(blob + 4 * (long long)n);
        int at = pos[idx];
        const int begin = at;
        for (int q = q0; q < q1; ++q) {
          const uint32_t v = pk[dd * m + q];
          const int s = (int)(v & 0xffffu), d = (int)(v >> 16);
          const float w = cw[dd * m + q];
          float wr;
          if (p.csr.agg_kind == AGG_GCN) wr = __fmul_rn(__fmul_rn(s_dinv[s], w), s_dinv[d]);
          else wr = dd == 0 ? w : w / (s_wsum[d] + 1e-8f);      // adjoint of the weighted mean
          rec[at++] = make_int2(dd == 0 ? agg_rec_x(s) : d, __float_as_int(wr));
        }
        const int self_x = dd == 0 ? agg_rec_x(i) : i;
        if (self) { const float dv = s_dinv[i]; rec[at++] = make_int2(self_x, __float_as_int(__fmul_rn(dv, dv))); }
        if (at & 1) rec[at++] = make_int2(self_x, 0);
        const float aux = p.csr.agg_kind == AGG_SAGE ? s_wsum[i] : s_dinv[i];
        reinterpret_cast<int4*>(blob)[i] = make_int4(begin, at, __float_as_int(aux), i);
      }
    }
    if (lean) return;
    for (int q = tid; q < m; q += NT) {
      uint32_t v = pk[q];
      int s = (int)(v & 0xffffu), d = (int)(v >> 16);
      float w = cw[q];
      p.csr.in_col[eb + q] = (int32_t)(nb + s);
      p.csr.in_w[eb + q] = w;
      p.csr.in_wn[eb + q] = __fmul_rn(__fmul_rn(s_dinv[s], w), s_dinv[d]);
      v = pk[m + q];
      s = (int)(v & 0xffffu); d = (int)(v >> 16);
      w = cw[m + q];
      p.csr.out_col[eb + q] = (int32_t)(nb + d);
      p.csr.out_w[eb + q] = w;
      p.csr.out_wn[eb + q] = __fmul_rn(__fmul_rn(s_dinv[s], w), s_dinv[d]);
    }
    return;
  }
  // Row sums in COO order (fp32, sequential): D^ (by source, self-loop weight 1 last) and w_sum.
  for (int i = tid; i < n; i += NT) {
    const int o0 = i ? cur_out[i - 1] : 0, o1 = cur_out[i];
    float deg = 0.0f;
    for (int q = o0; q < o1; ++q) deg = __fadd_rn(deg, p.csr.out_w[eb + q]);
    deg = __fadd_rn(deg, 1.0f);
    const int i0 = i ? cur_in[i - 1] : 0, i1 = cur_in[i];
    float ws = 0.0f;
    for (int q = i0; q < i1; ++q) ws = __fadd_rn(ws, p.csr.in_w[eb + q]);
    const float dinv = (float)(1.0 / sqrt((double)__fadd_rn(deg, 1e-8f)));
    s_dinv[i] = dinv;
    p.csr.deg[nb + i] = deg;
    p.csr.dinv[nb + i] = dinv;
    p.csr.wsum[nb + i] = ws;
  }
  __syncthreads();
  // w^_e = (dinv[src] * w) * dinv[dst]   (reference models.py:108 evaluation order)
  for (int i = tid; i < n; i += NT) {
    const int i0 = i ? cur_in[i - 1] : 0, i1 = cur_in[i];
    for (int q = i0; q < i1; ++q) {
      const int s = (int)(p.csr.in_col[eb + q] - nb);
      p.csr.in_wn[eb + q] = __fmul_rn(__fmul_rn(s_dinv[s], p.csr.in_w[eb + q]), s_dinv[i]);
    }
    const int o0 = i ? cur_out[i - 1] : 0, o1 = cur_out[i];
    for (int q = o0; q < o1; ++q) {
      const int d = (int)(p.csr.out_col[eb + q] - nb);
      p.csr.out_wn[eb + q] = __fmul_rn(__fmul_rn(s_dinv[i], p.csr.out_w[eb + q]), s_dinv[d]);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Lean collate of a PAIR store (one stored entry per undirected edge; entry k stands for the directed edges 2k = u -> v
// and 2k + 1 = v -> u of equal weight).  Every node then meets the same partners, in the same order and with the same
// weights, in its by-destination list as in its by-source list (pair k contributes exactly one half-edge to either
// list of each endpoint, and the lists are ordered by k), so ONE stable counting sort of the 2 * pairs half-edges by
// owner serves both aggregation blobs, the weighted degree D^ and w_sum; the general kernel above sorts twice.  The
// records are then written edge-parallel (consecutive threads, consecutive records) instead of row by row.  Bit for bit
// the blobs of k_collate_graph (tests: lean vs full collate).  NT / 32 warps = NT / 32 chunks of the half-edge list.
template <int NT>
__global__ void __launch_bounds__(NT) k_collate_pairs(CollateArgs p) {
  constexpr int kNW = NT / 32;
  CGNN_SMEM_DECL;
  __shared__ int s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long g = blockIdx.x;
  const long long nb = p.ptr[g], eb = p.eptr[g];
  const int n = (int)(p.ptr[g + 1] - nb);
  const int m = (int)(p.eptr[g + 1] - eb);       // directed edges = half-edges
  if (tid == 0) reinterpret_cast<int4*>(p.csr.graph_meta)[g] = make_int4((int)nb, n, (int)eb, m);
  if (n > p.max_nodes || m > p.edge_cap) return;  // host contract violated: never index past the shared arrays

  // 33 KB for a 360-node subject: six CTAs per SM - the kernel is a chain of short shared-memory phases, occupancy is
  // what hides their latency.  The stored pairs are read from global memory twice (count, place: the second time from
  // L1 / L2) instead of being staged, the per-chunk counters are 16 bits wide.
  int* start = reinterpret_cast<int*>(cgnn_smem);          // [n] packed counts -> first sorted slot (low 16 bits) | first record (high)
  float* s_dinv = reinterpret_cast<float*>(start + n);     // [n]
  float* s_wsum = s_dinv + n;                              // [n]
  uint32_t* pk = reinterpret_cast<uint32_t*>(s_wsum + n);  // [m] sorted half-edges: owner | partner << 16
  float* cw = reinterpret_cast<float*>(pk + m);            // [m]
  uint16_t* cnt = reinterpret_cast<uint16_t*>(cw + m);     // [kNW][n] per-chunk counts -> cursors (all below 2^16)
  uint32_t* cnt32 = reinterpret_cast<uint32_t*>(cnt);      // the same words, for the atomics

  const long long sid = p.ids[g];
  const long long sn = p.store.node_ptr[sid], so = p.store.edge_ptr[sid] >> 1;
  const int32_t* lsrc = p.store.src + so;
  const float* lw = p.store.w + so;
  // stored pair k: u | v << 16; endpoints outside the subject make it a zero-weight self edge on node 0 (as
  // k_collate_graph does, edge by edge)
  auto pair_at = [&](int k, float& w) -> uint32_t {
    uint32_t v = (uint32_t)lsrc[k];
    w = lw[k];
    if ((v & 0xffffu) >= (uint32_t)n || (v >> 16) >= (uint32_t)n) { v = 0u; w = 0.0f; }
    return v;
  };
  {
    const int F = p.store.num_features;
    const float* sx = p.store.x + sn * F;
    float* dx = p.x + nb * F;
    for (int i0 = tid; i0 < n * F; i0 += 4 * NT) {     // four loads in flight per thread
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = i0 + u * NT < n * F ? sx[i0 + u * NT] : 0.0f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * NT < n * F) dx[i0 + u * NT] = v[u];
    }
    if (p.batch)
      for (int i = tid; i < n; i += NT) p.batch[nb + i] = g;
    if (tid == 0 && p.labels && p.store.label) p.labels[g] = p.store.label[sid];
  }
  const int cnt_words = (kNW * n + 1) >> 1;
  for (int i = tid; i < cnt_words; i += NT) cnt32[i] = 0u;
  __syncthreads();

  // half-edge h: owner = (h & 1 ? v : u), partner = the other endpoint of pair h >> 1
  const int clen = (((m + kNW - 1) / kNW) + 31) & ~31;
  const int h_lo = min(warp * clen, m), h_hi = min(h_lo + clen, m);
  const int my_base = warp * n;
  for (int h = h_lo + lane; h < h_hi; h += 32) {
    float w;
    const uint32_t v = pair_at(h >> 1, w);
    const int idx = my_base + (int)((h & 1) ? (v >> 16) : (v & 0xffffu));
    atomicAdd(&cnt32[idx >> 1], 1u << (16 * (idx & 1)));
  }
  __syncthreads();
  const int self = p.csr.agg_kind == AGG_GCN ? 1 : 0;
  for (int i = tid; i < n; i += NT) {
    int t = 0;
    for (int c = 0; c < kNW; ++c) t += cnt[c * n + i];
    start[i] = t | (((t + self + 1) & ~1) << 16);     // both totals stay below 2^16 (the lists fit in shared memory)
  }
  __syncthreads();
  block_excl_scan(start, n, s_warp);
  for (int i = tid; i < n; i += NT) {
    int run = start[i] & 0xffff;
    for (int c = 0; c < kNW; ++c) {
      uint16_t* q = &cnt[c * n + i];
      const int t = *q;
      *q = (uint16_t)run;
      run += t;
    }
  }
  __syncthreads();
  uint16_t* my_cnt = cnt + my_base;
  for (int h0 = h_lo; h0 < h_hi; h0 += 32) {
    const int h = h0 + lane;
    const bool live = h < h_hi;
    uint32_t v = 0u; float w = 0.0f;
    if (live) v = pair_at(h >> 1, w);
    if (h & 1) v = (v >> 16) | (v << 16);             // owner in the low half
    const int key = live ? (int)(v & 0xffffu) : -1 - lane;
    const unsigned peers = __match_any_sync(kFull, key);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int at = 0;
    if (live) at = my_cnt[key];
    __syncwarp();
    if (live) {
      pk[at + rank] = v;
      cw[at + rank] = w;
      if (rank == 0) my_cnt[key] = (uint16_t)(at + __popc(peers));
    }
    __syncwarp();
  }
  __syncthreads();

  // D^ (by-source sum, self-loop weight 1 last) and w_sum (by-destination sum): the same sequence here
  for (int i = tid; i < n; i += NT) {
    const int q0 = start[i] & 0xffff, q1 = i + 1 < n ? (start[i + 1] & 0xffff) : m;
    float s = 0.0f;
    for (int q = q0; q < q1; ++q) s = __fadd_rn(s, cw[q]);
    const float deg = __fadd_rn(s, 1.0f);
    s_dinv[i] = (float)(1.0 / sqrt((double)__fadd_rn(deg, 1e-8f)));
    s_wsum[i] = s;
    p.csr.row_graph[nb + i] = (int32_t)g;
  }
  __syncthreads();

  const bool gcn = p.csr.agg_kind == AGG_GCN;
  int32_t* blob_in = p.csr.agg_in + agg_base_words(nb, eb, g);
  int32_t* blob_out = p.csr.agg_out ? p.csr.agg_out + agg_base_words(nb, eb, g) : nullptr;
  int2* rec_in = reinterpret_cast<int2*>(blob_in + 4 * (long long)n);
  int2* rec_out = blob_out ? reinterpret_cast<int2*>(blob_out + 4 * (long long)n) : nullptr;
  for (int q = tid; q < m; q += NT) {
    const uint32_t v = pk[q];
    const int i = (int)(v & 0xffffu), j = (int)(v >> 16);
    const float w = cw[q];
    const int st = start[i];
    const int at = ((unsigned)st >> 16) + (q - (st & 0xffff));
    const float di = s_dinv[i], dj = s_dinv[j];
    // row i as destination, j the source:  (dinv[src] * w) * dinv[dst]   (reference models.py:108 evaluation order)
    rec_in[at] = make_int2(agg_rec_x(j), __float_as_int(gcn ? __fmul_rn(__fmul_rn(dj, w), di) : w));
    // row i as source, j the destination; GraphSAGE: the adjoint of the weighted mean
    if (rec_out) rec_out[at] = make_int2(j, __float_as_int(gcn ? __fmul_rn(__fmul_rn(di, w), dj) : w / (s_wsum[j] + 1e-8f)));
  }
  for (int i = tid; i < n; i += NT) {
    const int st = start[i];
    const int q0 = st & 0xffff, q1 = i + 1 < n ? (start[i + 1] & 0xffff) : m;
    const int begin = (int)((unsigned)st >> 16);
    const int at0 = begin + (q1 - q0);
    const float dv = s_dinv[i];
    const int self_w = __float_as_int(__fmul_rn(dv, dv));
    int end = at0;
    if (self) ++end;
    const bool pad = end & 1;
    if (pad) ++end;
    {
      int a2 = at0;
      if (self) rec_in[a2++] = make_int2(agg_rec_x(i), self_w);
      if (pad) rec_in[a2++] = make_int2(agg_rec_x(i), 0);
    }
    if (rec_out) {
      int a2 = at0;
      if (self) rec_out[a2++] = make_int2(i, self_w);
      if (pad) rec_out[a2++] = make_int2(i, 0);
    }
    const float aux = gcn ? dv : s_wsum[i];
    reinterpret_cast<int4*>(blob_in)[i] = make_int4(begin, end, __float_as_int(aux), i);
    if (blob_out) reinterpret_cast<int4*>(blob_out)[i] = make_int4(begin, end, __float_as_int(aux), i);
  }
}

}  // namespace cgnn

using namespace cgnn;

// all CSR arrays, or none of them (a lean batch: graph_meta + the aggregation blobs only)
static bool csr_out_full(const cgnn_csr_out_t* c) {
  return c && c->in_rowptr && c->in_col && c->in_w && c->in_wn && c->out_rowptr && c->out_col && c->out_w &&
         c->out_wn && c->deg && c->dinv && c->wsum && c->graph_meta;
}
static bool csr_out_lean(const cgnn_csr_out_t* c) {
  return c && !c->in_rowptr && !c->in_col && !c->in_w && !c->in_wn && !c->out_rowptr && !c->out_col && !c->out_w &&
         !c->out_wn && !c->deg && !c->dinv && !c->wsum && c->graph_meta && c->agg_in && c->row_graph && c->agg_kind >= 0;
}
static bool csr_out_ok(const cgnn_csr_out_t* c) { return csr_out_full(c) || csr_out_lean(c); }

// Dynamic shared memory (two cursor arrays + dinv) is sized for the largest subject of the batch.
static int collate_launch(CollateArgs& a, bool from_store, int max_nodes, int max_edges, cudaStream_t stream) {
  const DeviceInfo dev = device_info();
  if (max_nodes < 1) max_nodes = 1;
  const int nt = (max_nodes <= 128 && a.B > 0 && a.total_edges / a.B <= 2048) ? 128 : kThreads;
  size_t smem = (size_t)max_nodes * (16 + 4 * (nt / 32)) + 16;   // cursors, dinv, wsum, per-chunk counters
  if (smem > (size_t)dev.smem_optin) return CGNN_ERR_TILE_TOO_LARGE;
  // room for the raw and the two sorted edge lists of a typical subject (24 bytes per edge): an eighth above the batch average, as
  // long as two CTAs still fit on an SM; larger subjects take the unstaged path inside the kernel
  a.edge_cap = 0;
  const bool lean = a.csr.in_col == nullptr;
  if (a.B > 0) {
    const long long avg = a.total_edges / a.B;
    long long cap = avg + avg / 8 + 32;
    if (max_edges >= 0 && max_edges <= cap) cap = max_edges;         // every subject staged
    if (cap > a.total_edges) cap = a.total_edges;
    size_t want = smem + (size_t)cap * 24;
    if (want + 1024 <= (size_t)(228 * 1024) / 2) { a.edge_cap = (int)cap; smem = want; }
    if (lean) {
      // a lean batch has nowhere to spill: every subject must be sorted in shared memory (one CTA per SM if need be)
      if (max_edges < 0) return CGNN_ERR_NEED_CSR;
      if (a.edge_cap < max_edges) {
        want = (size_t)max_nodes * (16 + 4 * (nt / 32)) + 16 + (size_t)max_edges * 24;
        if (want > (size_t)dev.smem_optin || max_nodes > 65535) return CGNN_ERR_NEED_CSR;
        a.edge_cap = max_edges; smem = want;
      }
    }
  }
  a.max_nodes = max_nodes;
  if (a.B <= 0) return CGNN_OK;
  // pair store + lean batch: one sort serves both blobs (k_collate_pairs); six CTAs per SM at the 360-node shape
  if (from_store && lean && a.store.edge_pairs && max_edges >= 0 && (max_edges & 1) == 0 && a.csr.agg_kind >= 0 && max_nodes <= 65535) {
    const int ntp = (max_nodes <= 128 && a.total_edges / a.B <= 2048) ? 128 : 256;
    const size_t smem_p = (size_t)max_nodes * (12 + 2 * (ntp / 32)) + (size_t)max_edges * 8 + 16;
    if (smem_p <= (size_t)dev.smem_optin) {
      const int cap = a.edge_cap;
      a.edge_cap = max_edges;
      if (ntp == 128) {
        auto kfn = k_collate_pairs<128>;
        if (smem_p > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
        CGNN_LAUNCH(kfn, (unsigned)a.B, 128, smem_p, stream, a);
      } else {
        auto kfn = k_collate_pairs<256>;
        if (smem_p > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
        CGNN_LAUNCH(kfn, (unsigned)a.B, 256, smem_p, stream, a);
      }
      CGNN_CHECK_LAUNCH();
      (void)cap;
      return CGNN_OK;
    }
  }
#define CGNN_COLLATE(FS_, NT_)                                                                              \
  {                                                                                                         \
    auto kfn = k_collate_graph<FS_, NT_>;                                                                   \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    CGNN_LAUNCH(kfn, (unsigned)a.B, NT_, smem, stream, a);                                                  \
  }
  if (from_store) { if (nt == 128) CGNN_COLLATE(true, 128) else CGNN_COLLATE(true, kThreads) }
  else { if (nt == 128) CGNN_COLLATE(false, 128) else CGNN_COLLATE(false, kThreads) }
#undef CGNN_COLLATE
  CGNN_CHECK_LAUNCH();
  if (!lean && a.csr.agg_kind >= 0 && a.csr.agg_in && a.csr.agg_out && a.csr.row_graph && a.edge_cap < a.total_edges &&
      (max_edges < 0 || a.edge_cap < max_edges)) {
    // subjects the kernel could not stage (more edges than edge_cap) get their blobs from the stand-alone builder,
    // which skips the ones already done
    cgnn_csr_t c{};
    c.in_rowptr = a.csr.in_rowptr; c.in_col = a.csr.in_col; c.in_w = a.csr.in_w; c.in_wn = a.csr.in_wn;
    c.out_rowptr = a.csr.out_rowptr; c.out_col = a.csr.out_col; c.out_w = a.csr.out_w; c.out_wn = a.csr.out_wn;
    c.deg = a.csr.deg; c.dinv = a.csr.dinv; c.wsum = a.csr.wsum; c.graph_meta = a.csr.graph_meta;
    return launch_build_agg(&c, a.csr.agg_kind, a.B, max_nodes, a.csr.agg_in, a.csr.agg_out, a.csr.row_graph, a.edge_cap, stream);
  }
  return CGNN_OK;
}

extern "C" {

int cgnn_collate_csr(const cgnn_store_t* store, const int64_t* subject_ids, int64_t num_graphs,
                     int64_t total_rows, int64_t total_edges, int32_t max_nodes, int32_t max_edges, float* node_features,
                     int64_t* edge_index, float* edge_weight, int64_t* batch, int64_t* labels, int64_t* ptr, int64_t* eptr,
                     const cgnn_csr_out_t* csr, cgnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_graphs == 0) {   // a rank's slice of a short global batch may be empty: an empty batch is a valid batch
    if (!ptr || !eptr || !csr) return CGNN_ERR_INVALID_ARG;
    cudaMemsetAsync(ptr, 0, sizeof(int64_t), stream);
    cudaMemsetAsync(eptr, 0, sizeof(int64_t), stream);
    if (csr->in_rowptr) cudaMemsetAsync(csr->in_rowptr, 0, sizeof(int32_t), stream);
    if (csr->out_rowptr) cudaMemsetAsync(csr->out_rowptr, 0, sizeof(int32_t), stream);
    return CGNN_OK;
  }
  if (!store || !subject_ids || num_graphs < 0 || total_rows < 0 || total_edges < 0 || !ptr || !eptr ||
      !csr_out_ok(csr) || !store->node_ptr || !store->edge_ptr || store->num_features <= 0)
    return CGNN_ERR_INVALID_ARG;
  if (total_rows > 0 && (!node_features || !store->x)) return CGNN_ERR_INVALID_ARG;
  if (total_edges > 0 && (!store->src || !store->w)) return CGNN_ERR_INVALID_ARG;
  if ((edge_index == nullptr) != (edge_weight == nullptr)) return CGNN_ERR_INVALID_ARG;   // the COO fields come together
  if (total_edges >= ((int64_t)1 << 31) || total_rows >= ((int64_t)1 << 31)) return CGNN_ERR_INVALID_ARG;
  if (store->edge_pairs && store->dst) return CGNN_ERR_INVALID_ARG;   // the pair layout exists for the compact store only
  {
    auto kfn = k_scan_ptrs;
    CGNN_LAUNCH(kfn, 1, 1024, 0, stream, (const long long*)store->node_ptr, (const long long*)store->edge_ptr,
                (const long long*)subject_ids, (long long)num_graphs, (long long*)ptr, (long long*)eptr);
    CGNN_CHECK_LAUNCH();
  }
  CollateArgs a;
  a.store = *store;
  a.ids = (const long long*)subject_ids;
  a.coo = nullptr; a.coo_w = nullptr;
  a.B = num_graphs; a.total_rows = total_rows; a.total_edges = total_edges;
  a.x = node_features; a.edge_index = (long long*)edge_index; a.edge_weight = edge_weight;
  a.batch = (long long*)batch; a.labels = (long long*)labels;
  a.ptr = (const long long*)ptr; a.eptr = (const long long*)eptr;
  a.csr = *csr;
  return collate_launch(a, true, max_nodes, max_edges, stream);
}

int cgnn_fetch_ids(const int64_t* pinned_host_ids, int64_t n, int64_t* ids, cgnn_stream_t stream_) {
  if (n < 0 || (n > 0 && (!pinned_host_ids || !ids))) return CGNN_ERR_INVALID_ARG;
  if (n == 0) return CGNN_OK;
  auto kfn = k_fetch_ids;
  CGNN_LAUNCH(kfn, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_, (const long long*)pinned_host_ids, (long long)n,
              (long long*)ids);
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

int cgnn_csr_from_coo(const int64_t* edge_index, const float* edge_weight, const int64_t* ptr,
                      int64_t num_graphs, int64_t total_rows, int64_t total_edges, int32_t max_nodes,
                      int64_t* eptr, const cgnn_csr_out_t* csr, cgnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!ptr || !eptr || num_graphs < 0 || total_rows < 0 || total_edges < 0 || !csr_out_full(csr))
    return CGNN_ERR_INVALID_ARG;
  if (total_edges > 0 && (!edge_index || !edge_weight)) return CGNN_ERR_INVALID_ARG;
  if (total_edges >= ((int64_t)1 << 31) || total_rows >= ((int64_t)1 << 31)) return CGNN_ERR_INVALID_ARG;
  {
    auto kfn = k_edge_ranges;
    long long items = num_graphs + 1;
    CGNN_LAUNCH(kfn, (unsigned)((items + 255) / 256), 256, 0, stream, (const long long*)edge_index,
                (long long)total_edges, (const long long*)ptr, (long long)num_graphs, (long long*)eptr);
    CGNN_CHECK_LAUNCH();
  }
  CollateArgs a;
  a.store = cgnn_store_t{};
  a.ids = nullptr;
  a.coo = (const long long*)edge_index; a.coo_w = edge_weight;
  a.B = num_graphs; a.total_rows = total_rows; a.total_edges = total_edges;
  a.x = nullptr; a.edge_index = nullptr; a.edge_weight = nullptr; a.batch = nullptr; a.labels = nullptr;
  a.ptr = (const long long*)ptr; a.eptr = (const long long*)eptr;
  a.csr = *csr;
  return collate_launch(a, false, max_nodes, -1, stream);
}

}  // extern "C"
