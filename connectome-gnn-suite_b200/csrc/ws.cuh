// ws.cuh - primitives of the warp-specialised layer kernels (engine.cu): mbarrier pipelines, TMA loads
// (cp.async.bulk / cp.async.bulk.tensor), tensor memory (tcgen05.alloc/st/ld) and the TF32 tcgen05.mma forms with the
// A operand in tensor memory.
//
// Two implementations of the same interface:
//   * device (nvcc, sm_100a): inline PTX;
//   * -DCGNN_EMU (g++, tests/emu): a functional model - mbarriers are phase/count words polled by fibers, TMA
//     copies complete at issue, tensor memory is a [128][512] array, tcgen05.mma is a TF32-truncating triple loop
//     that decodes the SAME shared-memory descriptors.  It checks the protocol (phases, counts, lane quadrants,
//     indexing, layouts as this file states them), not the hardware; the layouts themselves were pinned on a B200
//     in round 1 (tc05.cuh, tests/test_gpu_parity.py).
#pragma once
#include "common.cuh"

#ifndef CGNN_EMU
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "tc05.cuh"
#endif

namespace cgnn {
namespace ws {

// ------------------------------------------------------------------------------------------------------------------
// shared-memory addresses
// ------------------------------------------------------------------------------------------------------------------
#ifdef CGNN_EMU
// "shared address" = byte offset into the block's dynamic shared memory (what descriptors carry)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)(reinterpret_cast<const unsigned char*>(p) - cgnn_emu::g_dyn_smem);
}
__device__ __forceinline__ unsigned char* smem_align1024(unsigned char* raw) {
  return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
}
#else
using tc::smem_u32;
using tc::smem_align1024;
#endif

// ------------------------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------------------------
#ifdef CGNN_EMU
// word layout of the model: [0,20) pending arrivals | [20,40) arrival count per phase | [40,63) pending tx bytes | 63 phase
namespace emu {
inline void settle(uint64_t* bar) {
  uint64_t v = *bar;
  const uint64_t pending = v & 0xFFFFFu, init = (v >> 20) & 0xFFFFFu, tx = (v >> 40) & 0x7FFFFFu;
  if (pending == 0 && tx == 0) {
    v = (v & (1ull << 63)) ^ (1ull << 63);
    v |= init | (init << 20);
    *bar = v;
  }
  cgnn_emu::note_event();
}
inline void complete_tx(uint64_t* bar, uint32_t bytes) {
  const uint64_t tx = (*bar >> 40) & 0x7FFFFFu;
  if (tx < bytes) { fprintf(stderr, "[cuda_emu] complete_tx of %u bytes with %llu expected\n", bytes, (unsigned long long)tx); abort(); }
  *bar -= (uint64_t)bytes << 40;
  settle(bar);
}
}  // namespace emu
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { *bar = (uint64_t)count | ((uint64_t)count << 20); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  if ((*bar & 0xFFFFFu) == 0) { fprintf(stderr, "[cuda_emu] mbarrier over-arrival (thread %u)\n", threadIdx.x); abort(); }
  *bar -= 1;
  emu::settle(bar);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  *bar += (uint64_t)bytes << 40;
  mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (((*bar >> 63) & 1u) == (uint64_t)(parity & 1u)) cgnn_emu::poll_yield();
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void fence_mbar_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void named_sync(int id, int nthreads) { cgnn_emu::named_barrier(id, nthreads); }
#else
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { tc::mbar_wait(bar, parity); }
// (a poll loop with __nanosleep between the try_waits was measured in the fused eval kernel: 400 M of its 680 M warp
// instructions were the sleeping loop and every hand-over gained its latency; the plain try_wait suspends in hardware)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) { tc::mbar_wait(bar, parity); }
__device__ __forceinline__ void fence_proxy_async() { tc::fence_proxy_async(); }
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif

// ------------------------------------------------------------------------------------------------------------------
// TMA: 1-D bulk copies and 2-D tiled loads of a row-major fp32 matrix with the 128-byte swizzle
// ------------------------------------------------------------------------------------------------------------------
// A box is 32 fp32 columns (128 B) x BOX_ROWS rows; in shared memory row r of the box sits at r * 128 and its 16-byte
// chunk j at position j ^ (r & 7) (CU_TENSOR_MAP_SWIZZLE_128B; the destination must be 1024-byte aligned).
// Rows / columns outside the matrix arrive as zeros.
constexpr int kBoxCols = 32;
__device__ __host__ __forceinline__ uint32_t box_chunk_offset(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

#ifdef CGNN_EMU
struct TensorMap {
  const float* base; long long rows; int cols, box_rows;
};
static inline int make_tensor_map(TensorMap* m, const float* base, long long rows, int cols, int box_rows) {
  m->base = base; m->rows = rows; m->cols = cols; m->box_rows = box_rows;
  return CGNN_OK;
}
__device__ __forceinline__ void tma_load_box(void* dst, const TensorMap* m, int col0, long long row0, uint64_t* bar) {
  if (smem_u32(dst) & 1023u) { fprintf(stderr, "[cuda_emu] TMA box destination not 1024-byte aligned\n"); abort(); }
  unsigned char* d = reinterpret_cast<unsigned char*>(dst);
  for (int r = 0; r < m->box_rows; ++r)
    for (int j = 0; j < 8; ++j) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const long long gr = row0 + r;
      for (int e = 0; e < 4; ++e) {
        const int c = col0 + 4 * j + e;
        if (gr >= 0 && gr < m->rows && c < m->cols) v[e] = m->base[gr * m->cols + c];
      }
      memcpy(d + box_chunk_offset(r, j), v, 16);
    }
  emu::complete_tx(bar, (uint32_t)m->box_rows * 128u);
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  if ((bytes & 15u) || (smem_u32(dst) & 15u) || (((uintptr_t)src) & 15u)) { fprintf(stderr, "[cuda_emu] bulk copy alignment\n"); abort(); }
  memcpy(dst, src, bytes);
  emu::complete_tx(bar, bytes);
}
__device__ __forceinline__ void prefetch_tensor_map(const TensorMap*) {}
#else
using TensorMap = CUtensorMap;
// host: encode a [rows, cols] fp32 row-major matrix with 32-column x box_rows boxes (defined in engine.cu)
int make_tensor_map(TensorMap* m, const float* base, long long rows, int cols, int box_rows);
__device__ __forceinline__ void tma_load_box(void* dst, const TensorMap* m, int col0, long long row0, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(col0), "r"((int)row0), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void prefetch_tensor_map(const TensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
#endif

// ------------------------------------------------------------------------------------------------------------------
// tensor memory + MMA.  taddr = (lane << 16) | column; a warp reaches lanes [32 * (warp % 4), +32) only.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
// byte offset of element (row, k) of a K-major 128B-swizzled operand with `rows` rows per 32-wide K block
__device__ __host__ __forceinline__ uint32_t kmajor_offset(int row, int k, int rows) {
  return (uint32_t)((k >> 5) * rows * 128 + (row >> 3) * 1024 + (row & 7) * 128 + (((((k & 31) >> 2) ^ (row & 7))) << 4) + ((k & 3) << 2));
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (same bits as tc::smem_desc_sw128)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#ifdef CGNN_EMU
namespace emu {
inline int my_quadrant() { return (int)((threadIdx.x >> 5) & 3u); }
inline float* cell(uint32_t taddr, int lane_off, int col_off) {
  const int lane = (int)(taddr >> 16) + lane_off, col = (int)(taddr & 0xffffu) + col_off;
  if (lane < 0 || lane >= 128 || col < 0 || col >= 512) { fprintf(stderr, "[cuda_emu] tensor memory access out of range (lane %d col %d)\n", lane, col); abort(); }
  return cgnn_emu::tmem() + lane * 512 + col;
}
inline void check_quadrant(uint32_t taddr) {
  if ((int)(taddr >> 16) != 32 * my_quadrant()) {
    fprintf(stderr, "[cuda_emu] warp %u touches tensor-memory lane base %u outside its quadrant\n", threadIdx.x >> 5, taddr >> 16);
    abort();
  }
}
inline float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
}  // namespace emu
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t cols) { (void)cols; *smem_result = 0; }
__device__ __forceinline__ void tmem_dealloc(uint32_t, uint32_t) {}
__device__ __forceinline__ void fence_before_sync() {}
__device__ __forceinline__ void fence_after_sync() {}
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[N]) {
  emu::check_quadrant(taddr);
  for (int i = 0; i < N; ++i) *emu::cell(taddr, (int)(threadIdx.x & 31u), i) = v[i];
}
__device__ __forceinline__ void tmem_st_wait() {}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]) {
  emu::check_quadrant(taddr);
  for (int i = 0; i < N; ++i) v[i] = *emu::cell(taddr, (int)(threadIdx.x & 31u), i);
}
__device__ __forceinline__ void tmem_ld_wait() {}
// D[128 x N] (+)= A[tmem: lane = row, 8 columns = K] * B[smem, K-major SW128, N rows x 8 K]^T
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3Fu) << 3, M = (int)((idesc >> 24) & 0x1Fu) << 4;
  if (M != 128) { fprintf(stderr, "[cuda_emu] mma: M = %d\n", M); abort(); }
  const uint32_t start = (uint32_t)(b_desc & 0x3FFFu) << 4;
  const uint32_t sbo = (uint32_t)((b_desc >> 32) & 0x3FFFu) << 4;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = accumulate ? *emu::cell(d_tmem, m, n) : 0.0f;
      for (int kk = 0; kk < 8; ++kk) {
        uint32_t addr = start + (uint32_t)(n >> 3) * sbo + (uint32_t)(n & 7) * 128u + 4u * (uint32_t)kk;
        addr ^= ((addr >> 7) & 7u) << 4;      // the 128-byte swizzle is a function of the address bits
        float b;
        memcpy(&b, cgnn_emu::g_dyn_smem + addr, 4);
        acc += emu::trunc_tf32(*emu::cell(a_tmem, m, kk)) * emu::trunc_tf32(b);
      }
      *emu::cell(d_tmem, m, n) = acc;
    }
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) { mbar_arrive(bar); }
#else
using tc::tmem_alloc;
using tc::tmem_dealloc;
using tc::fence_before_sync;
using tc::fence_after_sync;
using tc::mma_tf32_ts;
using tc::mma_commit;
using tc::tmem_st_wait;
using tc::tmem_ld_wait;
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_st<8>(uint32_t taddr, const float (&v)[8]) { tc::tmem_st8(taddr, v); }
template <>
__device__ __forceinline__ void tmem_st<16>(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
// N columns of this thread's lane WITHOUT the wait (call tmem_ld_wait() before using v)
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  tc::tmem_ld16_nowait(taddr, r);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  tc::tmem_ld8_nowait(taddr, r);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
#endif

// One K = 8 step of the fp32-grade product with A = (a_hi, a_lo) in tensor memory and B = (b_hi, b_lo) descriptors:
// small terms first, the dominant hi * hi term last.
__device__ __forceinline__ void mma_tf32x3_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                                              uint32_t accumulate) {
  mma_tf32_ts(d_tmem, a_lo, b_hi, idesc, accumulate);
  mma_tf32_ts(d_tmem, a_hi, b_lo, idesc, 1u);
  mma_tf32_ts(d_tmem, a_hi, b_hi, idesc, 1u);
}

__device__ __forceinline__ bool elect_lane0() { return (threadIdx.x & 31u) == 0u; }

}  // namespace ws
}  // namespace cgnn
