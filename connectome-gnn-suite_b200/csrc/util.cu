// util.cu - status strings, device query, deterministic partial reductions.
#include "common.cuh"


namespace cgnn {

static thread_local int g_last_cuda_error = 0;
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
static int g_use_tensor_cores = 1;
bool tensor_cores_enabled() { return g_use_tensor_cores != 0; }
void set_tensor_cores(int on) { g_use_tensor_cores = on; }
void set_cuda_error(int err) { g_last_cuda_error = err; }
void set_tensor_cores(int on);

DeviceInfo device_info() {
  static thread_local int cached_dev = -1;
  static thread_local DeviceInfo cached = {148, 227 * 1024};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached;
  if (dev != cached_dev) {
    int sms = 148, smem = 227 * 1024;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cached.sm_count = sms;
    cached.smem_optin = smem;
    cached_dev = dev;
  }
  return cached;
}

// out[r*out_ld + c] = sum_g partials[g*stride + r*ld + c]; fp64, fixed order: a CTA owns 32 consecutive outputs
// (128-byte coalesced loads), its 8 warps sum one eighth of the records each, warp 0 adds the eight slices in order.
// Up to four such reductions (the weight, bias and BatchNorm-sum partials of one backward call) share ONE launch: the
// segments are laid end to end over the grid.  Every output is summed exactly as a launch of its own would.
constexpr int kReduceSlices = 8;
__global__ void __launch_bounds__(256) k_reduce_partials(ReduceBatch b) {
  __shared__ double s_part[kReduceSlices][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  int sg = 0;
  while (sg + 1 < b.n && (int)blockIdx.x >= b.seg[sg + 1].first_block) ++sg;
  const ReduceSeg q = b.seg[sg];
  const int i = ((int)blockIdx.x - q.first_block) * 32 + lane;
  const bool live = i < q.rows * q.cols;
  double s = 0.0;
  int r = 0, c = 0;
  if (live) {
    r = i / q.cols; c = i - r * q.cols;
    const float* p = q.partials + (size_t)r * q.ld + c;
    const int per = (q.G + kReduceSlices - 1) / kReduceSlices;
    const int g0 = slice * per, g1 = (g0 + per < q.G) ? g0 + per : q.G;
#pragma unroll 4
    for (int g = g0; g < g1; ++g) s += (double)p[(size_t)g * q.stride];
  }
  s_part[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && live) {
    double t = 0.0;
    for (int k = 0; k < kReduceSlices; ++k) t += s_part[k][lane];
    if (q.out) q.out[(size_t)r * q.out_ld + c] = (float)t;
    if (q.out64) q.out64[(size_t)r * q.cols + c] = t;
  }
}

void ReduceQueue::add(const float* partials, int G, int stride, int rows, int cols, int ld, float* out, int out_ld, double* out64) {
  if (rows * cols <= 0) return;
  if (b.n == kReduceMaxSegs) { const int rc = flush(); if (rc && !status) status = rc; }
  ReduceSeg& q = b.seg[b.n++];
  q.partials = partials; q.G = G; q.stride = stride; q.rows = rows; q.cols = cols; q.ld = ld; q.out = out;
  q.out_ld = out_ld > 0 ? out_ld : cols;
  q.out64 = out64;
  q.first_block = blocks;
  blocks += (rows * cols + 31) / 32;
}

int ReduceQueue::flush() {
  if (b.n == 0) return status;
  auto kfn = k_reduce_partials;
  CGNN_LAUNCH(kfn, (unsigned)blocks, 256, 0, stream, b);
  b.n = 0; blocks = 0;
  CGNN_CHECK_LAUNCH();
  return status;
}

int launch_reduce_partials(const float* partials, int G, int stride, int rows, int cols, int ld, float* out,
                           cudaStream_t stream, int out_ld, double* out64) {
  ReduceQueue q(stream);
  q.add(partials, G, stride, rows, cols, ld, out, out_ld, out64);
  return q.flush();
}

// One warp per channel: merge `parts` records {count, mean[C], M2[C]} (doubles) exactly:
//   N = sum n_i;  mean = sum n_i mean_i / N;  M2 = sum M2_i + n_i (mean_i - mean)^2
__global__ void __launch_bounds__(256) k_stats_merge(const double* __restrict__ parts, int parts_n, int C,
                                                     double* __restrict__ stats) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= C) return;
  const int rec = 1 + 2 * C;
  double n = 0.0, s = 0.0;
  for (int g = lane; g < parts_n; g += 32) {
    double ng = parts[(size_t)g * rec];
    n += ng;
    s += ng * parts[(size_t)g * rec + 1 + warp];
  }
  n = warp_sum(n);
  s = warp_sum(s);
  double mean = n > 0.0 ? s / n : 0.0;
  double q = 0.0;
  for (int g = lane; g < parts_n; g += 32) {
    double ng = parts[(size_t)g * rec];
    if (ng > 0.0) {
      double d = parts[(size_t)g * rec + 1 + warp] - mean;
      q += parts[(size_t)g * rec + 1 + C + warp] + ng * d * d;
    }
  }
  q = warp_sum(q);
  if (lane == 0) {
    stats[1 + warp] = mean;
    stats[1 + C + warp] = q;
    if (warp == 0) stats[0] = n;
  }
}

int launch_stats_merge(const double* parts, int parts_n, int C, double* stats, cudaStream_t stream) {
  auto kfn = k_stats_merge;
  int warps_per_block = 8;
  CGNN_LAUNCH(kfn, (C + warps_per_block - 1) / warps_per_block, 256, 0, stream, parts, parts_n, C, stats);
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

}  // namespace cgnn

extern "C" {

const char* cgnn_status_string(int status) {
  switch (status) {
    case CGNN_OK: return "ok";
    case CGNN_ERR_INVALID_ARG: return "invalid argument";
    case CGNN_ERR_TILE_TOO_LARGE: return "subject tile or weight matrix does not fit in shared memory";
    case CGNN_ERR_WORKSPACE: return "workspace too small";
    case CGNN_ERR_CUDA: return "CUDA runtime error";
    case CGNN_ERR_NEED_CSR: return "lean batch: this code path needs the CSR arrays";
    case CGNN_ERR_UNSUPPORTED: return "shape not covered by the fused entry point";
    default: return "unknown status";
  }
}

int cgnn_abi_version(void) { return CGNN_ABI_VERSION; }
int cgnn_last_cuda_error(void) { return cgnn::g_last_cuda_error; }
size_t cgnn_workspace_bytes(void) { return (size_t)64 << 20; }   // 148 CTAs x a 256 x 256 fp32 partial + the small records
uint64_t cgnn_kernel_launches(void) { return (uint64_t)cgnn::launches(); }
int cgnn_set_option(int32_t key, int32_t value) {
  if (key == CGNN_OPT_TENSOR_CORES) { cgnn::set_tensor_cores(value); return CGNN_OK; }
  return CGNN_ERR_INVALID_ARG;
}

int cgnn_bn_merge_stats(const double* stats_parts, int32_t parts, int32_t C, double* stats, cgnn_stream_t stream) {
  if (!stats_parts || !stats || parts <= 0 || C <= 0) return CGNN_ERR_INVALID_ARG;
  return cgnn::launch_stats_merge(stats_parts, parts, C, stats, (cudaStream_t)stream);
}

}  // extern "C"
