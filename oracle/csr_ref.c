/*
 * ORACLE - TEST INFRASTRUCTURE ONLY (never linked into or loaded by the product).
 *
 * Plain-C restatement of the integer / bit-exact part of the reference hot path:
 *   - collate_graphs                      reference connectome_gnn/graph.py:143-167
 *   - D^ (weighted degree with self loop) reference connectome_gnn/models.py:94-104
 *   - d^-1/2 and w^                       reference connectome_gnn/models.py:105-108
 *   - w_sum                               reference connectome_gnn/models.py:147-148
 * plus the per-subject stable CSR the CUDA kernels use (no reference counterpart: the reference
 * keeps COO and scatter-adds; a stable CSR is the same sums in the same order).
 *
 * Summation order is the observable part: PyTorch's CPU scatter_add_ adds sequentially in COO
 * order per target row, self loops come last because they are concatenated after the real
 * edges.  Every float add below is a separately rounded fp32 add (compile with
 * -ffp-contract=off, no -ffast-math).
 *
 * Pinned by tests/test_oracle.py against tests/golden/collate_*.npz (made from the reference).
 * Build: see oracle/Makefile -> oracle/_build/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Same argument meaning as cgnn_collate_csr in include/cgnn.h, all pointers are HOST pointers.
 * Returns 0, or 1 on an edge endpoint outside its subject, 2 on allocation failure. */
int oracle_collate_csr(
    /* subject store */
    const float* sx, const int32_t* ssrc, const int32_t* sdst, const float* sw, const int64_t* node_ptr,
    const int64_t* edge_ptr, const int64_t* slabel, int32_t num_features,
    /* selection */
    const int64_t* ids, int64_t num_graphs,
    /* reference fields */
    float* x, int64_t* edge_index, float* edge_weight, int64_t* batch, int64_t* labels, int64_t* ptr,
    int64_t* eptr,
    /* CSR */
    int32_t* in_rowptr, int32_t* in_col, float* in_w, float* in_wn, int32_t* out_rowptr, int32_t* out_col,
    float* out_w, float* out_wn, float* deg, float* dinv, float* wsum) {
  int64_t total_rows = 0, total_edges = 0;
  ptr[0] = 0;
  eptr[0] = 0;
  for (int64_t g = 0; g < num_graphs; ++g) {
    const int64_t s = ids[g];
    total_rows += node_ptr[s + 1] - node_ptr[s];
    total_edges += edge_ptr[s + 1] - edge_ptr[s];
    ptr[g + 1] = total_rows;
    eptr[g + 1] = total_edges;
  }
  for (int64_t g = 0; g < num_graphs; ++g) {
    const int64_t s = ids[g];
    const int64_t nb = ptr[g], eb = eptr[g];
    const int64_t n = ptr[g + 1] - nb, m = eptr[g + 1] - eb;
    const int32_t* ls = ssrc + edge_ptr[s];
    const int32_t* ld = sdst + edge_ptr[s];
    const float* lw = sw + edge_ptr[s];

    memcpy(x + nb * num_features, sx + node_ptr[s] * num_features, (size_t)(n * num_features) * sizeof(float));
    for (int64_t i = 0; i < n; ++i) batch[nb + i] = g;
    if (labels && slabel) labels[g] = slabel[s];
    for (int64_t e = 0; e < m; ++e) {
      if (ls[e] < 0 || ls[e] >= n || ld[e] < 0 || ld[e] >= n) return 1;
      edge_index[eb + e] = (int64_t)ls[e] + nb;               /* graph.py:152 */
      edge_index[total_edges + eb + e] = (int64_t)ld[e] + nb;
      edge_weight[eb + e] = lw[e];
    }

    /* stable counting sort by destination (in_*) and by source (out_*) */
    int64_t* cin = (int64_t*)calloc((size_t)(2 * n + 2), sizeof(int64_t));
    if (!cin) return 2;
    int64_t* cout_ = cin + n + 1;
    for (int64_t e = 0; e < m; ++e) { cin[ld[e] + 1]++; cout_[ls[e] + 1]++; }
    for (int64_t i = 0; i < n; ++i) { cin[i + 1] += cin[i]; cout_[i + 1] += cout_[i]; }
    for (int64_t i = 0; i < n; ++i) {
      in_rowptr[nb + i] = (int32_t)(eb + cin[i]);
      out_rowptr[nb + i] = (int32_t)(eb + cout_[i]);
    }
    int64_t* pin = (int64_t*)malloc((size_t)(2 * n + 1) * sizeof(int64_t));
    if (!pin) { free(cin); return 2; }
    int64_t* pout = pin + n;
    for (int64_t i = 0; i < n; ++i) { pin[i] = cin[i]; pout[i] = cout_[i]; }
    for (int64_t e = 0; e < m; ++e) {
      const int64_t a = eb + pin[ld[e]]++, b = eb + pout[ls[e]]++;
      in_col[a] = (int32_t)(nb + ls[e]);  in_w[a] = lw[e];
      out_col[b] = (int32_t)(nb + ld[e]); out_w[b] = lw[e];
    }

    /* D^ by source row, self loop (weight 1) last; w_sum by destination row */
    for (int64_t i = 0; i < n; ++i) {
      volatile float d = 0.0f;
      for (int64_t q = cout_[i]; q < cout_[i + 1]; ++q) d = d + out_w[eb + q];
      d = d + 1.0f;
      volatile float ws = 0.0f;
      for (int64_t q = cin[i]; q < cin[i + 1]; ++q) ws = ws + in_w[eb + q];
      volatile float shifted = d + 1e-8f;
      deg[nb + i] = d;
      wsum[nb + i] = ws;
      dinv[nb + i] = (float)(1.0 / sqrt((double)shifted));
    }
    /* w^ = (dinv[src] * w) * dinv[dst]    (models.py:108, left-to-right) */
    for (int64_t i = 0; i < n; ++i) {
      for (int64_t q = cin[i]; q < cin[i + 1]; ++q) {
        volatile float t = dinv[in_col[eb + q]] * in_w[eb + q];
        in_wn[eb + q] = t * dinv[nb + i];
      }
      for (int64_t q = cout_[i]; q < cout_[i + 1]; ++q) {
        volatile float t = dinv[nb + i] * out_w[eb + q];
        out_wn[eb + q] = t * dinv[out_col[eb + q]];
      }
    }
    free(pin);
    free(cin);
  }
  in_rowptr[total_rows] = (int32_t)total_edges;
  out_rowptr[total_rows] = (int32_t)total_edges;
  return 0;
}

/* D^ and w_sum straight from COO, exactly as the reference computes them (no CSR involved):
 * used to cross-check the CSR-ordered sums above. */
void oracle_degrees_from_coo(const int64_t* edge_index, const float* edge_weight, int64_t num_edges,
                             int64_t num_rows, float* deg, float* wsum) {
  for (int64_t i = 0; i < num_rows; ++i) { deg[i] = 0.0f; wsum[i] = 0.0f; }
  for (int64_t e = 0; e < num_edges; ++e) {
    volatile float a = deg[edge_index[e]] + edge_weight[e];
    deg[edge_index[e]] = a;
    volatile float b = wsum[edge_index[num_edges + e]] + edge_weight[e];
    wsum[edge_index[num_edges + e]] = b;
  }
  for (int64_t i = 0; i < num_rows; ++i) { volatile float a = deg[i] + 1.0f; deg[i] = a; }
}
