// norm.cu - K3: BatchNorm1d bookkeeping around the layer kernels (reference models.py:191-193,
// 208, 260 -> nn.BatchNorm1d defaults: eps 1e-5, momentum 0.1, affine, running stats) and the
// standalone BatchNorm-backward reduction used for the top layer.
#include "tile.cuh"

namespace cgnn {

__global__ void __launch_bounds__(256) k_bn_finalize(const double* __restrict__ stats, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, int C, float eps, float momentum,
                                                     float* running_mean, float* running_var,
                                                     long long* num_batches_tracked, float* __restrict__ scale,
                                                     float* __restrict__ shift, float* __restrict__ mean_out,
                                                     float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) num_batches_tracked[0] += 1;
  if (c >= C) return;
  const double n = stats[0];
  const double mean = stats[1 + c];
  const double var_b = n > 0.0 ? stats[1 + C + c] / n : 0.0;            // biased: normalisation
  const double var_u = n > 1.0 ? stats[1 + C + c] / (n - 1.0) : var_b;  // unbiased: running estimate
  const float rstd = (float)(1.0 / sqrt(var_b + (double)eps));
  const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
  const float sc = g * rstd;
  scale[c] = sc;
  shift[c] = b - (float)mean * sc;
  mean_out[c] = (float)mean;
  rstd_out[c] = rstd;
  if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)var_u;
}

__global__ void __launch_bounds__(256) k_bn_eval_affine(const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ running_mean,
                                                        const float* __restrict__ running_var, int C, float eps,
                                                        float* __restrict__ scale, float* __restrict__ shift,
                                                        float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float rstd = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
  const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
  const float sc = g * rstd;
  scale[c] = sc;
  shift[c] = b - running_mean[c] * sc;
  if (mean_out) mean_out[c] = running_mean[c];
  if (rstd_out) rstd_out[c] = rstd;
}

struct BnSumsArgs {
  const float* z; Act act; const float* mean; const float* rstd;
  const float* du; const float* demb; const long long* ptr; long long B;
  int C, C4;
  float* partials;  // [grid][2*C4]
};

// s1 = sum dy, s2 = sum dy * xhat with dy = d act(z) * du.  Warp per row, lanes over channels.
template <int CC>
__global__ void __launch_bounds__(kThreads) k_bn_bwd_sums(BnSumsArgs p) {
  CGNN_SMEM_DECL;
  float* sm = reinterpret_cast<float*>(cgnn_smem);
  const int C = p.C, C4 = p.C4;
  float* s_c = sm;               // [4][C4] scale, shift, mean, rstd
  float* s_red = sm + 4 * C4;    // [kWarps][2*C4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool affine = p.act.scale != nullptr;
  stage_affine(p.act, C, C4, s_c, s_c + C4);
  for (int c = tid; c < C4; c += kThreads) {
    s_c[2 * C4 + c] = c < C ? p.mean[c] : 0.0f;
    s_c[3 * C4 + c] = c < C ? p.rstd[c] : 0.0f;
  }
  __syncthreads();
  float s1[CC], s2[CC];
#pragma unroll
  for (int j = 0; j < CC; ++j) { s1[j] = 0.0f; s2[j] = 0.0f; }
  for (long long g = blockIdx.x; g < p.B; g += gridDim.x) {
    const long long nb = p.ptr[g];
    const int n = (int)(p.ptr[g + 1] - nb);
    const float inv_n = 1.0f / ((float)n + 1e-8f);
    for (int i = warp; i < n; i += kWarps) {
      const long long grow = nb + i;
      const uint32_t rh = p.act.drop ? drop_row_hash(p.act, p.act.row_base + grow) : 0u;
#pragma unroll
      for (int j = 0; j < CC; ++j) {
        const int ch = lane + 32 * j;
        if (ch < C) {
          const float t = p.z[grow * C + ch];
          const float up = p.du ? p.du[grow * C + ch] : p.demb[g * C + ch] * inv_n;
          const float dy = act_bwd(p.act, affine, t, s_c[ch], s_c[C4 + ch], rh, ch, up);
          const float xh = (t - s_c[2 * C4 + ch]) * s_c[3 * C4 + ch];
          s1[j] += dy;
          s2[j] = fmaf(dy, xh, s2[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CC; ++j) {
    const int ch = lane + 32 * j;
    if (ch < C4) { s_red[warp * 2 * C4 + ch] = s1[j]; s_red[warp * 2 * C4 + C4 + ch] = s2[j]; }
  }
  __syncthreads();
  float* part = p.partials + (size_t)blockIdx.x * 2 * C4;
  for (int c = tid; c < 2 * C4; c += kThreads) {
    float s = 0.0f;
    for (int w = 0; w < kWarps; ++w) s += s_red[w * 2 * C4 + c];
    part[c] = s;
  }
}

}  // namespace cgnn

using namespace cgnn;

extern "C" {

int cgnn_bn_finalize(const double* stats, const float* gamma, const float* beta, int32_t C, float eps,
                     float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                     float* scale, float* shift, float* mean, float* rstd, cgnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!stats || C <= 0 || !scale || !shift || !mean || !rstd) return CGNN_ERR_INVALID_ARG;
  auto kfn = k_bn_finalize;
  CGNN_LAUNCH(kfn, (C + 255) / 256, 256, 0, stream, stats, gamma, beta, (int)C, eps, momentum, running_mean,
              running_var, (long long*)num_batches_tracked, scale, shift, mean, rstd);
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

int cgnn_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                        int32_t C, float eps, float* scale, float* shift, float* mean, float* rstd,
                        cgnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!running_mean || !running_var || C <= 0 || !scale || !shift) return CGNN_ERR_INVALID_ARG;
  auto kfn = k_bn_eval_affine;
  CGNN_LAUNCH(kfn, (C + 255) / 256, 256, 0, stream, gamma, beta, running_mean, running_var, (int)C, eps, scale,
              shift, mean, rstd);
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

int cgnn_bn_bwd_sums(const float* z, const cgnn_act_t* act, const float* mean, const float* rstd, const float* du,
                     const float* demb, const int64_t* ptr, int64_t num_graphs, int64_t rows, int32_t C,
                     float* sums, void* workspace, size_t workspace_bytes, cgnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!sums || C <= 0 || num_graphs < 0 || rows < 0) return CGNN_ERR_INVALID_ARG;
  if (num_graphs == 0 || rows == 0) {
    cudaMemsetAsync(sums, 0, (size_t)2 * C * sizeof(float), stream);
    return CGNN_OK;
  }
  if ((du == nullptr) == (demb == nullptr)) return CGNN_ERR_INVALID_ARG;
  if (!z || !mean || !rstd || !ptr || !workspace) return CGNN_ERR_INVALID_ARG;
  const DeviceInfo dev = device_info();
  BnSumsArgs a;
  a.z = z; a.act = make_act(act); a.mean = mean; a.rstd = rstd; a.du = du; a.demb = demb;
  a.ptr = (const long long*)ptr; a.B = num_graphs; a.C = C; a.C4 = round_up(C, 4);
  if (a.C4 > 256) return CGNN_ERR_TILE_TOO_LARGE;
  const size_t smem = (size_t)(4 * a.C4 + kWarps * 2 * a.C4) * sizeof(float);
  int grid = persistent_grid(num_graphs, smem, dev, kThreads);
  const size_t rec = (size_t)2 * a.C4 * sizeof(float);
  if (workspace_bytes < rec) return CGNN_ERR_WORKSPACE;
  if ((size_t)grid * rec > workspace_bytes) grid = (int)(workspace_bytes / rec);
  a.partials = (float*)workspace;
  const int cc = pick_hc(a.C4);
#define CGNN_BN_SUMS(CC_)                                                  \
  {                                                                        \
    auto kfn = k_bn_bwd_sums<CC_>;                                         \
    CGNN_LAUNCH(kfn, grid, kThreads, smem, stream, a);                     \
  }
  if (cc == 1) CGNN_BN_SUMS(1) else if (cc == 2) CGNN_BN_SUMS(2) else if (cc == 4) CGNN_BN_SUMS(4) else CGNN_BN_SUMS(8)
#undef CGNN_BN_SUMS
  CGNN_CHECK_LAUNCH();
  return launch_reduce_partials(a.partials, grid, 2 * a.C4, 2, C, a.C4, sums, stream);
}

}  // extern "C"
