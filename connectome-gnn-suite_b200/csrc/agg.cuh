// agg.cuh - packed per-subject aggregation blobs and the warp-per-row gather over them.
//
// The generic kernels walk {rowptr, col, w} arrays: three dependent scalar loads per edge.  Measured on B200
// (tools/ubench/ldtm_gather.cu) every LSU instruction costs ~1.9 issue cycles per SM whatever its width, so the
// tensor-core generation of kernels reads the structure as 16-byte records instead:
//   row descriptor  int4 {rec_begin, rec_end, aux bits, 0}          one broadcast load per row
//   record pair     int4 {nbr0, w0 bits, nbr1, w1 bits}             one broadcast load per two edges
// Rows are padded to an even number of records (zero weight), GCN rows end with their self-loop record.
#pragma once
#include "common.cuh"

namespace cgnn {

enum { AGG_GCN = 0, AGG_SAGE = 1 };

// word offset of subject g's blob: 8 words per row + 2 per edge (edge base rounded up to even) + 4 per subject;
// a blob uses at most 4n (descriptors) + 2(m + 2n) (records incl. self-loop and padding) words.
static inline __host__ __device__ long long agg_base_words(long long nb, long long eb, long long g) {
  return 8 * nb + 2 * (eb + (eb & 1)) + 4 * g;
}
static inline size_t agg_total_words(long long rows, long long edges, long long graphs) {
  return (size_t)(8 * rows + 2 * (edges + 1) + 4 * graphs + 8);
}

struct AggView {
  const int4* desc;   // [n]
  const int4* rec2;   // record pairs; desc.x / desc.y index single records (even)
};
static __device__ __forceinline__ AggView agg_view(const int32_t* blob, long long nb, int n, long long eb, long long g) {
  const int32_t* b = blob + agg_base_words(nb, eb, g);
  AggView v;
  v.desc = reinterpret_cast<const int4*>(b);
  v.rec2 = reinterpret_cast<const int4*>(b + 4 * (long long)n);
  return v;
}

// acc[j] += sum over the records of row i of w * tile[nbr * ld + VW * lane + j]; returns the row's aux word.
// `tile` lives in shared memory, the records are read through the read-only path (all lanes the same address).
template <int VW>
static __device__ __forceinline__ float agg_gather_row(const AggView& a, int i, const float* __restrict__ tile_lane, int ld,
                                                       float (&acc)[VW]) {
  const int4 d = __ldg(a.desc + i);
  for (int e = d.x; e < d.y; e += 2) {
    const int4 r = __ldg(a.rec2 + (e >> 1));
    const float w0 = __int_as_float(r.y), w1 = __int_as_float(r.w);
    const float* p0 = tile_lane + r.x * ld;
    const float* p1 = tile_lane + r.z * ld;
    if (VW == 1) {
      const float v0 = p0[0], v1 = p1[0];
      acc[0] = fmaf(v0, w0, acc[0]);
      acc[0] = fmaf(v1, w1, acc[0]);
    } else if (VW == 2) {
      const float2 v0 = *reinterpret_cast<const float2*>(p0), v1 = *reinterpret_cast<const float2*>(p1);
      acc[0] = fmaf(v0.x, w0, acc[0]); acc[1] = fmaf(v0.y, w0, acc[1]);
      acc[0] = fmaf(v1.x, w1, acc[0]); acc[1] = fmaf(v1.y, w1, acc[1]);
    } else {
#pragma unroll
      for (int q = 0; q < VW; q += 4) {
        const float4 v0 = *reinterpret_cast<const float4*>(p0 + q), v1 = *reinterpret_cast<const float4*>(p1 + q);
        acc[q + 0] = fmaf(v0.x, w0, acc[q + 0]); acc[q + 1] = fmaf(v0.y, w0, acc[q + 1]);
        acc[q + 2] = fmaf(v0.z, w0, acc[q + 2]); acc[q + 3] = fmaf(v0.w, w0, acc[q + 3]);
        acc[q + 0] = fmaf(v1.x, w1, acc[q + 0]); acc[q + 1] = fmaf(v1.y, w1, acc[q + 1]);
        acc[q + 2] = fmaf(v1.z, w1, acc[q + 2]); acc[q + 3] = fmaf(v1.w, w1, acc[q + 3]);
      }
    }
  }
  return __int_as_float(d.z);
}

enum { GATHER_SAGE_FWD = 0, GATHER_GCN_BWD = 1, GATHER_SAGE_BWD = 2 };
enum { GC_SCALE = 0, GC_SHIFT, GC_BSC, GC_MEAN, GC_RSTD, GC_S1N, GC_S2N, GC_ROWS };   // per-channel constant rows

struct GatherArgs {
  const int32_t* meta; long long B; const int32_t* blob;
  int C, ld, vec, max_nodes;
  // tile source
  const float* src;        // SAGE_FWD: t_in   GCN_BWD: z   SAGE_BWD: d_agg
  Act act;                 // SAGE_FWD: act on load   GCN_BWD: act_out (its backward)   SAGE_BWD: act_in (for the sums)
  const float* du; const float* demb;                                    // GCN_BWD upstream
  const float* bn_scale; const float* bn_mean; const float* bn_rstd; const float* bn_s1; const float* bn_s2;
  float inv_count; int bn_train, has_bn;
  // SAGE_BWD
  const float* direct;     // d_u [rows, C]
  const float* t_raw;      // this layer's stored input (for the sums of the layer below)
  const float* prev_mean; const float* prev_rstd; int want_prev;
  float* out;              // [rows, C]
  float* partials; int part_stride;   // GCN_BWD: dbias [C]; SAGE_BWD: prev sums [2C]
};


// agg.cu: one gather kernel; CGNN_OK when launched (grid in *grid_out: the caller reduces `partials` over it),
// -1 when the shape is not covered.
int launch_gather(int mode, GatherArgs& a, int* grid_out, cudaStream_t stream);
#ifndef CGNN_EMU
// gemm_tc.cu: tensor-core contractions; same return convention.
int launch_sage_fwd_gemm(const float* t_in, const cgnn_act_t* act, const float* agg, const float* W, const float* bias,
                         int64_t rows, int32_t C, int32_t H, float* z, double* partials, int* grid_out,
                         size_t workspace_bytes, cudaStream_t stream);
int launch_gcn_bwd_gemm(const float* dP, const float* t_in, const cgnn_act_t* act_in, const float* W, int64_t rows,
                        int32_t d_in, int32_t H, float* du_in, const float* prev_mean, const float* prev_rstd, int want_prev,
                        float* partials, int part_stride, int o_pprev, int* grid_out, size_t partial_bytes,
                        cudaStream_t stream);
#endif

}  // namespace cgnn
