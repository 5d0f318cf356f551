// ingest.cu - dense connectivity matrices -> thresholded COO subjects on the device.
//
// The reference documents one other way into ConnectomeGraph besides its generator: a dense N x N structural
// connectivity matrix per subject (84 x 84 or 360 x 360), thresholded at its own 90th percentile and listed in both
// directions (reference README.md:145-179, `hcp_matrix_to_graph`):
//
//     threshold = A.flatten().quantile(0.90)                 # linear interpolation between two order statistics
//     A_thresh  = (A > threshold).float() * A
//     src, dst  = torch.where(A_thresh > 0)                  # row-major order
//     edge_index  = [cat(src, dst); cat(dst, src)]           # first every (i -> j), then every (j -> i)
//     edge_weight = cat(w, w),  w = A_thresh[src, dst]       # (the README writes A_thresh[src]: a row lookup, a slip)
//     node_features = deg / (deg.max() + 1e-8),  deg = A_thresh.sum(dim=1, keepdim=True)
//
// Two kernels, one CTA per subject:
//   k_ingest_threshold   exact order statistics by a 4-pass radix select over the N^2 values (re-read from L2), torch's
//                        fp32 rank / lerp arithmetic, and the number of selected entries
//   k_ingest_emit        per-row counts -> scan -> stable, row-major emission of both directions with warp ballots;
//                        weighted row sums and the normalised feature
// Edges and weights are bit-exact against the recipe; the feature is a plain fp32 row sum (torch's vectorised sum uses
// another association order: equal to ~1e-7 relative, tests/test_ingest.py).
#include "common.cuh"

namespace cgnn {

constexpr int kIngestThreads = 1024;

__device__ __forceinline__ uint32_t ingest_key(float f) {     // order-preserving map float -> unsigned
  const uint32_t b = (uint32_t)__float_as_int(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ingest_unkey(uint32_t k) {
  const uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __int_as_float((int)b);
}

// block-wide sum of one int per thread (fixed order); s_tmp has 32 entries
__device__ __forceinline__ int ingest_block_sum(int v, int* s_tmp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  __syncthreads();
  if (lane == 0) s_tmp[warp] = v;
  __syncthreads();
  int t = 0;
  for (int i = 0; i < nw; ++i) t += s_tmp[i];
  return t;
}

// k-th smallest key (0-based) of the n values; *below = number of keys strictly smaller, *equal = multiplicity of the result
__device__ __forceinline__ uint32_t ingest_select(const float* __restrict__ A, int n, int k, int* s_hist, int* s_pick,
                                                  int* below_out, int* equal_out) {
  uint32_t prefix = 0u;
  int below = 0, equal = 0;
  for (int pass = 3; pass >= 0; --pass) {
    const int shift = 8 * pass;
    const uint32_t hmask = pass == 3 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = ingest_key(A[i]);
      if ((key & hmask) == (prefix & hmask)) atomicAdd(&s_hist[(key >> shift) & 0xffu], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int run = 0, b = 0;
      for (; b < 256; ++b) {
        if (run + s_hist[b] > k) break;
        run += s_hist[b];
      }
      if (b > 255) b = 255;          // k < n is a host contract
      s_pick[0] = b; s_pick[1] = run; s_pick[2] = s_hist[b];
    }
    __syncthreads();
    prefix |= (uint32_t)s_pick[0] << shift;
    below += s_pick[1];
    k -= s_pick[1];
    equal = s_pick[2];
    __syncthreads();
  }
  *below_out = below;
  *equal_out = equal;
  return prefix;
}

__global__ void __launch_bounds__(kIngestThreads) k_ingest_threshold(const float* __restrict__ mats, int N, float q,
                                                                     float* __restrict__ thr, int* __restrict__ nnz) {
  __shared__ int s_hist[256];
  __shared__ int s_pick[4];
  __shared__ int s_tmp[32];
  __shared__ uint32_t s_min;
  const long long s = blockIdx.x;
  const int n = N * N;
  const float* A = mats + s * (long long)n;
  // torch.quantile (aten Sorting.cpp quantile_compute): rank = q * (n - 1) in the input's dtype, values at floor / ceil,
  // result = lerp(below, above, rank - floor(rank))
  const float rank = __fmul_rn(q, (float)(n - 1));
  const int lo = (int)floorf(rank), hi = (int)ceilf(rank);
  const float wgt = __fsub_rn(rank, floorf(rank));
  int below = 0, equal = 0;
  const uint32_t key_lo = ingest_select(A, n, lo, s_hist, s_pick, &below, &equal);
  uint32_t key_hi = key_lo;
  if (hi > lo && below + equal <= hi) {       // the next order statistic is the smallest key above key_lo
    if (threadIdx.x == 0) s_min = 0xffffffffu;
    __syncthreads();
    uint32_t mine = 0xffffffffu;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = ingest_key(A[i]);
      if (key > key_lo && key < mine) mine = key;
    }
    for (int o = 16; o > 0; o >>= 1) { const uint32_t t = __shfl_xor_sync(kFull, mine, o); if (t < mine) mine = t; }
    if ((threadIdx.x & 31) == 0) atomicMin(&s_min, mine);
    __syncthreads();
    key_hi = s_min;
  }
  const float v_lo = ingest_unkey(key_lo), v_hi = ingest_unkey(key_hi);
  // aten Lerp.h: |w| < 0.5 ? a + w (b - a) : b - (b - a) (1 - w)
  const float diff = __fsub_rn(v_hi, v_lo);
  const float t = fabsf(wgt) < 0.5f ? __fadd_rn(v_lo, __fmul_rn(wgt, diff)) : __fsub_rn(v_hi, __fmul_rn(diff, __fsub_rn(1.0f, wgt)));
  int cnt = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = A[i];
    cnt += (a > t && a > 0.0f) ? 1 : 0;
  }
  const int total = ingest_block_sum(cnt, s_tmp);
  if (threadIdx.x == 0) { thr[s] = t; nnz[s] = total; }
}

// edge_ptr[s] = first DIRECTED edge of subject s (2 * selected entries per subject); node rows at s * N
__global__ void __launch_bounds__(kIngestThreads) k_ingest_emit(const float* __restrict__ mats, int N, const float* __restrict__ thr,
                                                                const long long* __restrict__ edge_ptr, int32_t* __restrict__ src,
                                                                int32_t* __restrict__ dst, float* __restrict__ w,
                                                                float* __restrict__ x) {
  CGNN_SMEM_DECL;
  __shared__ int s_tmp[32];
  __shared__ float s_max[32];
  int* rowstart = reinterpret_cast<int*>(cgnn_smem);            // [N] counts -> exclusive starts
  float* deg = reinterpret_cast<float*>(rowstart + N);          // [N]
  const long long s = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
  const float* A = mats + s * (long long)N * N;
  const float t = thr[s];
  const long long e0 = edge_ptr[s];
  const int half = (int)((edge_ptr[s + 1] - e0) >> 1);
  for (int i = warp; i < N; i += nw) {
    int c = 0;
    float d = 0.0f;
    for (int j = lane; j < N; j += 32) {
      const float a = A[(long long)i * N + j];
      if (a > t) { d += a; c += a > 0.0f ? 1 : 0; }            // A_thresh = (A > thr) * A; edges where A_thresh > 0
    }
    for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(kFull, c, o); d += __shfl_xor_sync(kFull, d, o); }
    if (lane == 0) { rowstart[i] = c; deg[i] = d; }
  }
  __syncthreads();
  if (warp == 0) {                     // exclusive scan of the row counts, max of the row sums: one warp
    const int per = (N + 31) / 32;
    const int a = min(lane * per, N), b = min(a + per, N);
    int sum = 0;
    float mx = -INFINITY;
    for (int i = a; i < b; ++i) { sum += rowstart[i]; mx = fmaxf(mx, deg[i]); }
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += v; }
    int run = inc - sum;
    for (int i = a; i < b; ++i) { const int c = rowstart[i]; rowstart[i] = run; run += c; }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    if (lane == 0) s_max[0] = mx;
  }
  __syncthreads();
  const float denom = s_max[0] + 1e-8f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) x[s * (long long)N + i] = deg[i] / denom;
  for (int i = warp; i < N; i += nw) {
    int at = rowstart[i];
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      float a = 0.0f;
      if (j < N) a = A[(long long)i * N + j];
      const bool pick = j < N && a > t && a > 0.0f;
      const unsigned m = __ballot_sync(kFull, pick);
      if (pick) {
        const long long e = e0 + at + __popc(m & ((1u << lane) - 1u));
        src[e] = i; dst[e] = j; w[e] = a;                       // (i -> j)
        src[e + half] = j; dst[e + half] = i; w[e + half] = a;  // (j -> i), second block
      }
      at += __popc(m);
    }
  }
  (void)s_tmp;
}

}  // namespace cgnn

using namespace cgnn;

extern "C" {

int cgnn_ingest_threshold(const float* matrices, int64_t num_subjects, int32_t N, float q, float* threshold, int32_t* selected,
                          cgnn_stream_t stream_) {
  if (num_subjects < 0 || N <= 0 || N > 32767 || !(q >= 0.0f && q <= 1.0f)) return CGNN_ERR_INVALID_ARG;
  if (num_subjects == 0) return CGNN_OK;
  if (!matrices || !threshold || !selected) return CGNN_ERR_INVALID_ARG;
  auto kfn = k_ingest_threshold;
  CGNN_LAUNCH(kfn, (unsigned)num_subjects, kIngestThreads, 0, (cudaStream_t)stream_, matrices, (int)N, q, threshold, (int*)selected);
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

int cgnn_ingest_emit(const float* matrices, int64_t num_subjects, int32_t N, const float* threshold, const int64_t* edge_ptr,
                     int32_t* src, int32_t* dst, float* weight, float* node_features, cgnn_stream_t stream_) {
  if (num_subjects < 0 || N <= 0 || N > 32767) return CGNN_ERR_INVALID_ARG;
  if (num_subjects == 0) return CGNN_OK;
  if (!matrices || !threshold || !edge_ptr || !src || !dst || !weight || !node_features) return CGNN_ERR_INVALID_ARG;
  const size_t smem = (size_t)N * 8;
  if (smem > (size_t)device_info().smem_optin) return CGNN_ERR_TILE_TOO_LARGE;
  auto kfn = k_ingest_emit;
#ifndef CGNN_EMU
  if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  CGNN_LAUNCH(kfn, (unsigned)num_subjects, kIngestThreads, smem, (cudaStream_t)stream_, matrices, (int)N, threshold,
              (const long long*)edge_ptr, src, dst, weight, node_features);
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

}  // extern "C"
