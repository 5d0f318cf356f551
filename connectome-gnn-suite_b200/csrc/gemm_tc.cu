// gemm_tc.cu - the dense half of the tensor-core generation of layer kernels (sm_100a only).
//
// Row-tile streaming kernels: 128 rows of the batch at a time, operands split into TF32 hi/lo parts and written as
// 128B-swizzled canonical tiles, products issued by one thread as tcgen05.mma kind::tf32 (three terms: lo*hi, hi*lo,
// hi*hi; fp32 accumulators in tensor memory), results read back with tcgen05.ld.  The subject structure is gone
// at this point - the gathers around these contractions live in agg.cu.
//
//   k_sage_fwd_gemm   z = relu([u || agg] W^T + b), BatchNorm partial statistics        (reference models.py:151-152)
//   k_gcn_bwd_gemm    du_in = dP W,  dW += dP^T u,  BatchNorm-backward sums of the layer below
//                                                                            (autograd of reference models.py:111)
//   k_sage_bwd_gemm   [d_u || d_agg] = dz W,  dW += dz^T [u || agg],  dbias    (autograd of reference models.py:151-152)
//
// X B products (contraction over channels) read K-major 128B-swizzled tiles, X^T Y products (contraction over the
// rows of the tile) read MN-major SWIZZLE_128B_BASE32B tiles - the only MN-major layout for tf32; see tc05.cuh.
#include "tc05.cuh"
#include "tile.cuh"

namespace cgnn {
#ifndef CGNN_EMU

static_assert(kThreads == 512, "the drain / staging maps below assume 16 warps");
constexpr int kRows = 128;

// 128 lanes x NC columns of an accumulator -> padded staging tile [128][NC + 4] in shared memory
template <int NC>
__device__ __forceinline__ void drain_to_staging(uint32_t taddr, float* stage, int warp, int lane) {
  constexpr int CW = NC / 4;   // columns per warp
  const int q = warp & 3, cg = warp >> 2;
  float v[CW];
  tc::tmem_ld_cols<CW>(taddr + ((uint32_t)(32 * q) << 16) + (uint32_t)(cg * CW), v);
  float* dst = stage + (32 * q + lane) * (NC + 4) + cg * CW;
#pragma unroll
  for (int j = 0; j < CW; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}

__device__ __forceinline__ float f4_get(const float4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

// ================================================================================================================
// GraphSAGE forward contraction
// ================================================================================================================
struct SageFwdGemmArgs {
  const float* t_in; Act act; const float* agg; const float* W; const float* bias;
  long long rows; int C, H, vec;
  float* z; double* partials;
  uint32_t tmem_cols; int o_stage;
};

template <int KB, int HB>
__global__ void __launch_bounds__(kThreads, 1) k_sage_fwd_gemm(SageFwdGemmArgs p) {
  constexpr int KP = 32 * KB, H = 32 * HB;
  constexpr int QA = KP / 4, NQA = kRows * QA / kThreads;     // operand quads per thread per tile
  constexpr int QH = H / 4, NQH = kRows * QH / kThreads;      // output quads per thread per tile
  constexpr int A_HALF = KB * kRows * 128, B_HALF = KB * H * 128;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_hi = base;
  unsigned char* a_lo = a_hi + A_HALF;
  unsigned char* b_hi = a_lo + A_HALF;
  unsigned char* b_lo = b_hi + B_HALF;
  float* s_scale = reinterpret_cast<float*>(b_lo + B_HALF);   // [KP]
  float* s_shift = s_scale + KP;
  float* s_bias = s_shift + KP;                                // [H]
  float* stage = reinterpret_cast<float*>(base + p.o_stage);   // [128][H + 4], aliases the A operand when it fits
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, K = 2 * C;
  const bool affine = p.act.scale != nullptr;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, p.tmem_cols);
  if (tid == 0) tc::mbar_init(&mbar, 1);
  for (int idx = tid; idx < H * QA; idx += kThreads) {
    const int n = idx / QA, k = (idx - n * QA) << 2;
    float4 v;
    v.x = k + 0 < K ? p.W[(size_t)n * K + k + 0] : 0.0f;
    v.y = k + 1 < K ? p.W[(size_t)n * K + k + 1] : 0.0f;
    v.z = k + 2 < K ? p.W[(size_t)n * K + k + 2] : 0.0f;
    v.w = k + 3 < K ? p.W[(size_t)n * K + k + 3] : 0.0f;
    tc::store_split4(b_hi, b_lo, n, k, H, v);
  }
  stage_affine(p.act, C, KP, s_scale, s_shift);
  for (int h = tid; h < H; h += kThreads) s_bias[h] = p.bias ? p.bias[h] : 0.0f;
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  const uint32_t a_hi_u = tc::smem_u32(a_hi), a_lo_u = tc::smem_u32(a_lo), b_hi_u = tc::smem_u32(b_hi), b_lo_u = tc::smem_u32(b_lo);

  const long long ntiles = (p.rows + kRows - 1) / kRows;
  float4 pre[NQA];
  auto load_tile = [&](long long t) {
    const long long r0 = t * kRows;
#pragma unroll
    for (int i = 0; i < NQA; ++i) {
      const int idx = tid + i * kThreads;
      const int r = idx / QA, k = (idx - r * QA) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < ntiles && r0 + r < p.rows) {
        const long long row = r0 + r;
        if (p.vec) {
          if (k < C) v = *reinterpret_cast<const float4*>(p.t_in + row * C + k);
          else if (k < K) v = *reinterpret_cast<const float4*>(p.agg + row * C + (k - C));
        } else {
          float e[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kk = k + j;
            e[j] = kk < C ? p.t_in[row * C + kk] : (kk < K ? p.agg[row * C + (kk - C)] : 0.0f);
          }
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      pre[i] = v;
    }
  };

  int cnt = 0;
  Welford wf[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) wf[j].init();

  long long t = blockIdx.x;
  load_tile(t);
  uint32_t phase = 0;
  for (; t < ntiles; t += gridDim.x) {
    const long long r0 = t * kRows;
    // (1) previous layer's BatchNorm / dropout on the u half, hi/lo split, swizzled operand stores
#pragma unroll
    for (int i = 0; i < NQA; ++i) {
      const int idx = tid + i * kThreads;
      const int r = idx / QA, k = (idx - r * QA) << 2;
      float4 v = pre[i];
      if (r0 + r < p.rows && k < C) {
        const uint32_t rh = p.act.drop ? drop_row_hash(p.act, p.act.row_base + r0 + r) : 0u;
        v.x = act_fwd(p.act, affine, v.x, s_scale[k + 0], s_shift[k + 0], rh, k + 0);
        if (k + 1 < C) v.y = act_fwd(p.act, affine, v.y, s_scale[k + 1], s_shift[k + 1], rh, k + 1);
        if (k + 2 < C) v.z = act_fwd(p.act, affine, v.z, s_scale[k + 2], s_shift[k + 2], rh, k + 2);
        if (k + 3 < C) v.w = act_fwd(p.act, affine, v.w, s_scale[k + 3], s_shift[k + 3], rh, k + 3);
      }
      tc::store_split4(a_hi, a_lo, r, k, kRows, v);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (tid == 0) {
      const uint32_t id = tc::idesc_tf32(kRows, H);
#pragma unroll 1
      for (int ks = 0; ks < KP / 8; ++ks) {
        const uint32_t kb = (uint32_t)(ks >> 2), ko = (uint32_t)((ks & 3) * 32);
        const uint32_t ao = kb * (uint32_t)kRows * 128u + ko, bo = kb * (uint32_t)H * 128u + ko;
        tc::mma_tf32x3_step(taddr, tc::smem_desc_sw128(a_hi_u + ao), tc::smem_desc_sw128(a_lo_u + ao),
                            tc::smem_desc_sw128(b_hi_u + bo), tc::smem_desc_sw128(b_lo_u + bo), id, ks > 0 ? 1u : 0u);
      }
      tc::mma_commit(&mbar);
    }
    load_tile(t + gridDim.x);          // next tile's loads fly during the MMA and the epilogue
    tc::mbar_wait(&mbar, phase);
    phase ^= 1;
    tc::fence_after_sync();
    drain_to_staging<H>(taddr, stage, warp, lane);
    tc::fence_before_sync();
    __syncthreads();
    // (2) bias + ReLU, coalesced store, BatchNorm statistics (fixed channel quad per thread)
#pragma unroll
    for (int i = 0; i < NQH; ++i) {
      const int idx = tid + i * kThreads;
      const int r = idx / QH, c = (idx - r * QH) << 2;
      if (r0 + r < p.rows) {
        float4 v = *reinterpret_cast<const float4*>(stage + r * (H + 4) + c);
        v.x = fmaxf(v.x + s_bias[c + 0], 0.0f);
        v.y = fmaxf(v.y + s_bias[c + 1], 0.0f);
        v.z = fmaxf(v.z + s_bias[c + 2], 0.0f);
        v.w = fmaxf(v.w + s_bias[c + 3], 0.0f);
        *reinterpret_cast<float4*>(p.z + (r0 + r) * H + c) = v;
        cnt += 1;
        const float inv = 1.0f / (float)cnt;
        wf[0].push(v.x, inv); wf[1].push(v.y, inv); wf[2].push(v.z, inv); wf[3].push(v.w, inv);
      }
    }
    __syncthreads();   // staging (= operand tile) is rewritten by the next tile
  }

  if (p.partials) {
    float* rec = stage;   // [kThreads][9]
    rec[tid * 9] = (float)cnt;
#pragma unroll
    for (int j = 0; j < 4; ++j) { rec[tid * 9 + 1 + j] = wf[j].mean; rec[tid * 9 + 5 + j] = wf[j].m2; }
    __syncthreads();
    double* out = p.partials + (size_t)blockIdx.x * (1 + 2 * H);
    for (int c = tid; c < H; c += kThreads) {
      const int q = c >> 2, j = c & 3;
      double n = 0.0, mean = 0.0, m2 = 0.0;
      for (int th = q; th < kThreads; th += QH) {
        const double nb = (double)rec[th * 9];
        if (nb <= 0.0) continue;
        const double mb = (double)rec[th * 9 + 1 + j], qb = (double)rec[th * 9 + 5 + j];
        const double nt = n + nb, delta = mb - mean;
        mean += delta * (nb / nt);
        m2 += qb + delta * delta * (n * nb / nt);
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

int launch_sage_fwd_gemm(const float* t_in, const cgnn_act_t* act, const float* agg, const float* W, const float* bias,
                         int64_t rows, int32_t C, int32_t H, float* z, double* partials, int* grid_out,
                         size_t workspace_bytes, cudaStream_t stream) {
  if (H % 32 != 0 || H > 128 || H == 96 || C <= 0 || 2 * C > 128) return -1;
  const int KB = (2 * C + 31) / 32, HB = H / 32;
  if (KB == 3) return -1;
  if ((((uintptr_t)z) & 15u) != 0) return -1;
  const DeviceInfo dev = device_info();
  SageFwdGemmArgs a;
  a.t_in = t_in; a.act = make_act(act); a.agg = agg; a.W = W; a.bias = bias;
  a.rows = rows; a.C = C; a.H = H;
  a.vec = (C % 4 == 0) && ((((uintptr_t)t_in) | ((uintptr_t)agg)) & 15u) == 0;
  a.z = z; a.partials = partials;
  a.tmem_cols = 32;
  while (a.tmem_cols < (uint32_t)H) a.tmem_cols <<= 1;
  const int KP = 32 * KB;
  const size_t a_bytes = (size_t)2 * KB * kRows * 128, b_bytes = (size_t)2 * KB * H * 128;
  const size_t c_bytes = (size_t)(2 * KP + H) * 4;
  size_t stage_bytes = (size_t)kRows * (H + 4) * 4;
  if (stage_bytes < (size_t)kThreads * 9 * 4) stage_bytes = (size_t)kThreads * 9 * 4;
  size_t total = a_bytes + b_bytes + c_bytes;
  if (stage_bytes <= a_bytes) a.o_stage = 0;
  else { a.o_stage = (int)((total + 15) & ~(size_t)15); total = a.o_stage + stage_bytes; }
  const size_t smem = total + 1024;
  if (smem > (size_t)dev.smem_optin) return -1;
  const long long ntiles = (rows + kRows - 1) / kRows;
  long long grid = dev.sm_count;
  if (grid > ntiles) grid = ntiles;
  if (partials) {
    const size_t rec = (size_t)(1 + 2 * H) * sizeof(double);
    if ((size_t)grid * rec > workspace_bytes) grid = (long long)(workspace_bytes / rec);
  }
  if (grid < 1) return -1;
  *grid_out = (int)grid;
#define CGNN_SF(KB_, HB_)                                                                             \
  {                                                                                                   \
    auto kfn = k_sage_fwd_gemm<KB_, HB_>;                                                             \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                \
    CGNN_LAUNCH(kfn, (unsigned)grid, kThreads, smem, stream, a);                                      \
  }
#define CGNN_SF_H(KB_) { if (HB == 1) CGNN_SF(KB_, 1) else if (HB == 2) CGNN_SF(KB_, 2) else CGNN_SF(KB_, 4) }
  if (KB == 1) CGNN_SF_H(1) else if (KB == 2) CGNN_SF_H(2) else CGNN_SF_H(4)
#undef CGNN_SF_H
#undef CGNN_SF
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

// ================================================================================================================
// GCN backward contractions
// ================================================================================================================
struct GcnBwdGemmArgs {
  const float* dP; const float* t_in; Act act_in; const float* W;
  long long rows; int H, Kin, vec_dp, vec_in;
  float* du_in; const float* prev_mean; const float* prev_rstd; int want_prev;
  float* partials; int part_stride, o_pprev;   // per CTA: [dW H x Kin][prev 2 x Kin]
  uint32_t tmem_cols;
};

template <int HB, int KB>
__global__ void __launch_bounds__(kThreads, 1) k_gcn_bwd_gemm(GcnBwdGemmArgs p) {
  constexpr int H = 32 * HB, KP = 32 * KB;
  constexpr int Q1 = H / 4, NQ1 = kRows * Q1 / kThreads;      // dP quads per thread per tile
  constexpr int Q2 = KP / 4, NQ2 = kRows * Q2 / kThreads;     // layer-input quads per thread per tile
  constexpr int A1_HALF = HB * kRows * 128, A2_HALF = KB * kRows * 128, B1_HALF = HB * KP * 128;
  // M of the dW product: always 128 (an M = 64 instruction costs the same tensor-pipe time); for H < 128 the upper
  // MN groups of the A view run into the neighbouring operand tiles and fill accumulator rows that are never read.
  constexpr int MM = 128;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a1k_hi = base;                 // dP [128 rows][H], K-major (for du_in = dP W)
  unsigned char* a1k_lo = a1k_hi + A1_HALF;
  unsigned char* a1_hi = a1k_lo + A1_HALF;      // dP again, MN-major (for dW = dP^T u)
  unsigned char* a1_lo = a1_hi + A1_HALF;
  unsigned char* a2_hi = a1_lo + A1_HALF;       // u  [128 rows][KP], MN-major
  unsigned char* a2_lo = a2_hi + A2_HALF;
  unsigned char* b1_hi = a2_lo + A2_HALF;       // W^T [KP rows (input channel)][H]
  unsigned char* b1_lo = b1_hi + B1_HALF;
  float* s_ci = reinterpret_cast<float*>(b1_lo + B1_HALF);    // [4][KP]: scale, shift, mean, rstd of the layer below
  float* stage = reinterpret_cast<float*>(base);               // [128][KP + 4] / end-of-kernel scratch; aliases a1 | a2
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Kin = p.Kin;
  const bool aff_in = p.act_in.scale != nullptr;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, p.tmem_cols);
  if (tid == 0) tc::mbar_init(&mbar, 1);
  for (int idx = tid; idx < KP * Q1; idx += kThreads) {
    const int n = idx / Q1, h = (idx - n * Q1) << 2;          // operand row n = input channel, K index = h
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < Kin) v = make_float4(p.W[(size_t)(h + 0) * Kin + n], p.W[(size_t)(h + 1) * Kin + n],
                                 p.W[(size_t)(h + 2) * Kin + n], p.W[(size_t)(h + 3) * Kin + n]);
    tc::store_split4(b1_hi, b1_lo, n, h, KP, v);
  }
  stage_affine(p.act_in, Kin, KP, s_ci, s_ci + KP);
  for (int c = tid; c < KP; c += kThreads) {
    const bool ok = c < Kin && p.want_prev;
    s_ci[2 * KP + c] = ok ? p.prev_mean[c] : 0.0f;
    s_ci[3 * KP + c] = ok ? p.prev_rstd[c] : 0.0f;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  const uint32_t a1h = tc::smem_u32(a1_hi), a1l = tc::smem_u32(a1_lo), a2h = tc::smem_u32(a2_hi), a2l = tc::smem_u32(a2_lo);
  const uint32_t a1kh = tc::smem_u32(a1k_hi), a1kl = tc::smem_u32(a1k_lo);
  const uint32_t b1h = tc::smem_u32(b1_hi), b1l = tc::smem_u32(b1_lo);

  const long long ntiles = (p.rows + kRows - 1) / kRows;
  float4 dpn[NQ1], tin[NQ2], tcur[NQ2];
  auto load_tile = [&](long long t) {
    const long long r0 = t * kRows;
#pragma unroll
    for (int i = 0; i < NQ1; ++i) {
      const int idx = tid + i * kThreads;
      const int r = idx / Q1, k = (idx - r * Q1) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < ntiles && r0 + r < p.rows) v = *reinterpret_cast<const float4*>(p.dP + (r0 + r) * H + k);
      dpn[i] = v;
    }
#pragma unroll
    for (int i = 0; i < NQ2; ++i) {
      const int idx = tid + i * kThreads;
      const int r = idx / Q2, k = (idx - r * Q2) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < ntiles && r0 + r < p.rows && k < Kin) {
        const float* src = p.t_in + (r0 + r) * Kin + k;
        if (p.vec_in) v = *reinterpret_cast<const float4*>(src);
        else {
          v.x = src[0];
          if (k + 1 < Kin) v.y = src[1];
          if (k + 2 < Kin) v.z = src[2];
          if (k + 3 < Kin) v.w = src[3];
        }
      }
      tin[i] = v;
    }
  };

  float ps1[4] = {0.f, 0.f, 0.f, 0.f}, ps2[4] = {0.f, 0.f, 0.f, 0.f};
  long long t = blockIdx.x;
  load_tile(t);
  uint32_t phase = 0;
  bool first = true;
  for (; t < ntiles; t += gridDim.x) {
    const long long r0 = t * kRows;
    // (1) operands: dP as is, u = BatchNorm / ReLU / dropout of the stored layer input
#pragma unroll
    for (int i = 0; i < NQ1; ++i) {
      const int idx = tid + i * kThreads;
      const int r = idx / Q1, k = (idx - r * Q1) << 2;
      if (p.du_in) tc::store_split4(a1k_hi, a1k_lo, r, k, kRows, dpn[i]);
      tc::store_split4_mn32(a1_hi, a1_lo, r, k, kRows, dpn[i]);
    }
#pragma unroll
    for (int i = 0; i < NQ2; ++i) {
      const int idx = tid + i * kThreads;
      const int r = idx / Q2, k = (idx - r * Q2) << 2;
      const float4 raw = tin[i];
      tcur[i] = raw;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < p.rows && k < Kin) {
        const uint32_t rh = p.act_in.drop ? drop_row_hash(p.act_in, p.act_in.row_base + r0 + r) : 0u;
        v.x = act_fwd(p.act_in, aff_in, raw.x, s_ci[k + 0], s_ci[KP + k + 0], rh, k + 0);
        if (k + 1 < Kin) v.y = act_fwd(p.act_in, aff_in, raw.y, s_ci[k + 1], s_ci[KP + k + 1], rh, k + 1);
        if (k + 2 < Kin) v.z = act_fwd(p.act_in, aff_in, raw.z, s_ci[k + 2], s_ci[KP + k + 2], rh, k + 2);
        if (k + 3 < Kin) v.w = act_fwd(p.act_in, aff_in, raw.w, s_ci[k + 3], s_ci[KP + k + 3], rh, k + 3);
      }
      tc::store_split4_mn32(a2_hi, a2_lo, r, k, kRows, v);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (tid == 0) {
      if (p.du_in) {   // du_in tile = dP W : M = 128 rows, N = KP, K = H
        const uint32_t id1 = tc::idesc_tf32(kRows, KP);
#pragma unroll 1
        for (int ks = 0; ks < H / 8; ++ks) {
          const uint32_t kb = (uint32_t)(ks >> 2), ko = (uint32_t)((ks & 3) * 32);
          const uint32_t ao = kb * (uint32_t)kRows * 128u + ko, bo = kb * (uint32_t)KP * 128u + ko;
          tc::mma_tf32x3_step(taddr, tc::smem_desc_sw128(a1kh + ao), tc::smem_desc_sw128(a1kl + ao),
                              tc::smem_desc_sw128(b1h + bo), tc::smem_desc_sw128(b1l + bo), id1, ks > 0 ? 1u : 0u);
        }
      }
      // dW += dP^T u : M = H (MN-major view of the dP tile), N = KP (MN-major view of the u tile), K = 128 rows
      const uint32_t id2 = tc::idesc_tf32(MM, KP, 1, 1);
      const uint32_t lbo = (uint32_t)kRows * 128u;
#pragma unroll 1
      for (int ks = 0; ks < kRows / 8; ++ks) {
        const uint32_t off = (uint32_t)ks * 1024u;
        tc::mma_tf32x3_step(taddr + (uint32_t)KP, tc::smem_desc_mn32(a1h + off, lbo, 512u), tc::smem_desc_mn32(a1l + off, lbo, 512u),
                            tc::smem_desc_mn32(a2h + off, lbo, 512u), tc::smem_desc_mn32(a2l + off, lbo, 512u), id2,
                            (first && ks == 0) ? 0u : 1u);
      }
      tc::mma_commit(&mbar);
    }
    first = false;
    load_tile(t + gridDim.x);
    tc::mbar_wait(&mbar, phase);
    phase ^= 1;
    tc::fence_after_sync();
    if (p.du_in) {
      drain_to_staging<KP>(taddr, stage, warp, lane);
      tc::fence_before_sync();
      __syncthreads();
      // (2) coalesced du_in store + BatchNorm-backward sums of the layer below (fixed channel quad per thread)
#pragma unroll
      for (int i = 0; i < NQ2; ++i) {
        const int idx = tid + i * kThreads;
        const int r = idx / Q2, c = (idx - r * Q2) << 2;
        if (r0 + r < p.rows && c < Kin) {
          const float4 d = *reinterpret_cast<const float4*>(stage + r * (KP + 4) + c);
          float* dst = p.du_in + (r0 + r) * Kin + c;
          if (p.vec_in) *reinterpret_cast<float4*>(dst) = d;
          else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (c + j < Kin) dst[j] = f4_get(d, j);
          }
          if (p.want_prev) {
            const uint32_t rh = p.act_in.drop ? drop_row_hash(p.act_in, p.act_in.row_base + r0 + r) : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ch = c + j;
              if (ch < Kin) {
                const float t0 = f4_get(tcur[i], j);
                const float dyp = act_bwd(p.act_in, aff_in, t0, s_ci[ch], s_ci[KP + ch], rh, ch, f4_get(d, j));
                const float xh = (t0 - s_ci[2 * KP + ch]) * s_ci[3 * KP + ch];
                ps1[j] += dyp;
                ps2[j] = fmaf(dyp, xh, ps2[j]);
              }
            }
          }
        }
      }
    }
    __syncthreads();   // operand tiles / staging are rewritten by the next tile
  }

  // ---- per-CTA partial record ---------------------------------------------------------------------------------
  float* part = p.partials + (size_t)blockIdx.x * p.part_stride;
  {
    constexpr int CW = KP / 4;
    const int q = warp & 3, cg = warp >> 2;
    float v[CW];
    tc::tmem_ld_cols<CW>(taddr + ((uint32_t)(32 * q) << 16) + (uint32_t)(KP + cg * CW), v);
    const int h = MM == 128 ? 32 * q + lane : (lane < 16 ? 16 * q + lane : -1);
    if (h >= 0 && h < H) {
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const int c = cg * CW + j;
        if (c < Kin) part[(size_t)h * Kin + c] = v[j];
      }
    }
  }
  if (p.want_prev) {
    float* red = stage;   // [kThreads][8]
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[tid * 8 + j] = ps1[j]; red[tid * 8 + 4 + j] = ps2[j]; }
    __syncthreads();
    for (int c = tid; c < 2 * Kin; c += kThreads) {
      const int which = c / Kin, ch = c - which * Kin;
      const int q = ch >> 2, j = ch & 3;
      float s = 0.0f;
      for (int th = q; th < kThreads; th += Q2) s += red[th * 8 + which * 4 + j];
      part[p.o_pprev + c] = s;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(taddr, p.tmem_cols);
}

// Returns CGNN_OK when launched (grid in *grid_out; the caller reduces the partial records), -1 when the shape is
// not covered (the caller falls back to the generic kernel).
int launch_gcn_bwd_gemm(const float* dP, const float* t_in, const cgnn_act_t* act_in, const float* W, int64_t rows,
                        int32_t d_in, int32_t H, float* du_in, const float* prev_mean, const float* prev_rstd, int want_prev,
                        float* partials, int part_stride, int o_pprev, int* grid_out, size_t partial_bytes,
                        cudaStream_t stream) {
  if (H % 32 != 0 || H > 128 || H == 96 || d_in <= 0 || d_in > 128) return -1;
  const int HB = H / 32, KB = (d_in + 31) / 32;
  if (KB == 3) return -1;
  if ((((uintptr_t)dP) & 15u) != 0) return -1;
  const DeviceInfo dev = device_info();
  GcnBwdGemmArgs a;
  a.dP = dP; a.t_in = t_in; a.act_in = make_act(act_in); a.W = W;
  a.rows = rows; a.H = H; a.Kin = d_in; a.vec_dp = 1;
  a.vec_in = (d_in % 4 == 0) && ((((uintptr_t)t_in) & 15u) == 0) && (!du_in || (((uintptr_t)du_in) & 15u) == 0);
  a.du_in = du_in; a.prev_mean = prev_mean; a.prev_rstd = prev_rstd; a.want_prev = want_prev;
  a.partials = partials; a.part_stride = part_stride; a.o_pprev = o_pprev;
  const int KP = 32 * KB;
  a.tmem_cols = 32;
  while (a.tmem_cols < (uint32_t)(2 * KP)) a.tmem_cols <<= 1;
  size_t total = (size_t)4 * HB * kRows * 128 + (size_t)2 * KB * kRows * 128 + (size_t)2 * HB * KP * 128 + (size_t)4 * KP * 4;
  const size_t a_view = (size_t)3 * HB * kRows * 128 + (size_t)4 * kRows * 128;   // the M = 128 view of the MN-major dP lo tile ends here
  if (total < a_view) total = a_view;
  const size_t smem = total + 1024;
  if (smem > (size_t)dev.smem_optin) return -1;
  const long long ntiles = (rows + kRows - 1) / kRows;
  long long grid = dev.sm_count;
  if (grid > ntiles) grid = ntiles;
  const size_t rec = (size_t)part_stride * sizeof(float);
  if ((size_t)grid * rec > partial_bytes) grid = (long long)(partial_bytes / rec);
  if (grid < 1) return -1;
  *grid_out = (int)grid;
#define CGNN_GB(HB_, KB_)                                                                             \
  {                                                                                                   \
    auto kfn = k_gcn_bwd_gemm<HB_, KB_>;                                                              \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                \
    CGNN_LAUNCH(kfn, (unsigned)grid, kThreads, smem, stream, a);                                      \
  }
#define CGNN_GB_K(HB_) { if (KB == 1) CGNN_GB(HB_, 1) else if (KB == 2) CGNN_GB(HB_, 2) else CGNN_GB(HB_, 4) }
  if (HB == 1) CGNN_GB_K(1) else if (HB == 2) CGNN_GB_K(2) else CGNN_GB_K(4)
#undef CGNN_GB_K
#undef CGNN_GB
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

#endif  // CGNN_EMU
}  // namespace cgnn
