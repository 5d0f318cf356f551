// cuda_emu.cpp - fiber scheduler behind tests/emu/cuda_emu.h (TEST-ONLY, see that header).
#include "cuda_emu.h"

namespace cgnn_emu {

uint3_ g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;

namespace {

constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;
constexpr uint64_t kDead = 0xDEADDEADDEADDEADull;
constexpr size_t kGuard = 256;

struct Fiber {
  ucontext_t ctx;
  bool done = true;
  int wait = 0;  // 0 runnable, 1 at block barrier, 2 at warp barrier, 3 polling (runnable, but proves no progress)
};
uint64_t g_events = 0;                  // bumped by note_event()
std::vector<float> g_tmem;              // 128 x 512 tensor-memory words of the running block
int g_named_cnt[16];
uint64_t g_named_gen[16];

std::vector<Fiber> fibers;
std::vector<char*> stacks;
ucontext_t main_ctx;
int cur = -1;
int nthreads = 0;
const std::function<void()>* body = nullptr;
uint64_t warp_slots[kMaxThreads / 32][32];

void set_thread_index(int t) {
  g_threadIdx.x = t % g_blockDim.x;
  g_threadIdx.y = (t / g_blockDim.x) % g_blockDim.y;
  g_threadIdx.z = t / (g_blockDim.x * g_blockDim.y);
}

void trampoline() {
  (*body)();
  fibers[cur].done = true;
  warp_slots[cur / 32][cur % 32] = kDead;
  // returning follows uc_link back to the scheduler
}

void yield() { swapcontext(&fibers[cur].ctx, &main_ctx); }

void run_block(size_t smem_bytes) {
  std::vector<unsigned char> smem(smem_bytes + kGuard);
  memset(smem.data(), 0xCB, smem_bytes);          // poison: uninitialised reads give garbage
  memset(smem.data() + smem_bytes, 0xA5, kGuard);  // canary: catches writes past the end
  g_dyn_smem = smem.data();

  for (auto& w : warp_slots) for (auto& s : w) s = kDead;
  g_tmem.assign(128 * 512, 0.0f);
  for (int i = 0; i < 16; ++i) { g_named_cnt[i] = 0; g_named_gen[i] = 0; }
  int idle_sweeps = 0;
  for (int t = 0; t < nthreads; ++t) {
    Fiber& f = fibers[t];
    f.done = false;
    f.wait = 0;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = stacks[t];
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = &main_ctx;
    makecontext(&f.ctx, trampoline, 0);
  }

  int alive = nthreads;
  while (alive > 0) {
    bool progressed = false;
    const uint64_t events0 = g_events;
    for (int t = 0; t < nthreads; ++t) {
      Fiber& f = fibers[t];
      if (f.done || (f.wait != 0 && f.wait != 3)) continue;
      const bool polling = f.wait == 3;
      f.wait = 0;
      cur = t;
      set_thread_index(t);
      swapcontext(&main_ctx, &f.ctx);
      if (!polling || f.done || f.wait != 3) progressed = true;   // a poller that only polled again proves nothing
    }
    if (g_events != events0) progressed = true;
    alive = 0;
    int at_block = 0;
    for (int t = 0; t < nthreads; ++t)
      if (!fibers[t].done) { ++alive; at_block += fibers[t].wait == 1; }
    bool released = false;
    if (alive > 0 && at_block == alive) {
      for (int t = 0; t < nthreads; ++t) fibers[t].wait = 0;
      released = true;
    }
    for (int w = 0; w * 32 < nthreads; ++w) {
      int live = 0, at_warp = 0;
      for (int l = 0; l < 32 && w * 32 + l < nthreads; ++l) {
        const Fiber& f = fibers[w * 32 + l];
        if (!f.done) { ++live; at_warp += f.wait == 2; }
      }
      if (live > 0 && at_warp == live) {
        for (int l = 0; l < 32 && w * 32 + l < nthreads; ++l)
          if (fibers[w * 32 + l].wait == 2) fibers[w * 32 + l].wait = 0;
        released = true;
      }
    }
    if (progressed || released) idle_sweeps = 0; else ++idle_sweeps;
    if (alive > 0 && idle_sweeps >= 3) {
      fprintf(stderr, "[cuda_emu] DEADLOCK in block (%u,%u,%u): %d live threads;", g_blockIdx.x,
              g_blockIdx.y, g_blockIdx.z, alive);
      int shown = 0;
      for (int t = 0; t < nthreads && shown < 16; ++t)
        if (!fibers[t].done) { fprintf(stderr, " t%d:%s", t, fibers[t].wait == 1 ? "block" : fibers[t].wait == 2 ? "warp" : "poll"); ++shown; }
      fprintf(stderr, "\n");
      abort();
    }
  }
  for (size_t i = 0; i < kGuard; ++i)
    if (smem[smem_bytes + i] != 0xA5) {
      fprintf(stderr, "[cuda_emu] dynamic shared memory overrun (+%zu B past %zu) in block %u\n", i,
              smem_bytes, g_blockIdx.x);
      abort();
    }
  g_dyn_smem = nullptr;
}

}  // namespace

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn) {
  nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || nthreads > kMaxThreads) { fprintf(stderr, "[cuda_emu] bad block size %d\n", nthreads); abort(); }
  if (smem > 227 * 1024) { fprintf(stderr, "[cuda_emu] %zu B dynamic smem exceeds 227 KB\n", smem); abort(); }
  if ((int)fibers.size() < nthreads) fibers.resize(nthreads);
  while ((int)stacks.size() < nthreads) stacks.push_back((char*)malloc(kStackBytes));
  body = &fn;
  g_blockDim = block;
  g_gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx = uint3_{bx, by, bz};
        run_block(smem);
      }
  body = nullptr;
}

void sync_threads() {
  fibers[cur].wait = 1;
  yield();
}

int lane_id() { return cur % 32; }

const uint64_t* warp_gather(uint64_t v) {
  warp_slots[cur / 32][cur % 32] = v;
  fibers[cur].wait = 2;
  yield();
  return warp_slots[cur / 32];
}

void warp_release() {
  fibers[cur].wait = 2;
  yield();
}

void poll_yield() {
  fibers[cur].wait = 3;
  yield();
}

void note_event() { ++g_events; }

float* tmem() { return g_tmem.data(); }

void named_barrier(int id, int count) {
  if (id < 0 || id >= 16) { fprintf(stderr, "[cuda_emu] named barrier id %d\n", id); abort(); }
  const uint64_t gen = g_named_gen[id];
  if (++g_named_cnt[id] == count) {
    g_named_cnt[id] = 0;
    ++g_named_gen[id];
    note_event();
    return;
  }
  while (g_named_gen[id] == gen) poll_yield();
}

}  // namespace cgnn_emu
