// Runtime of the SIMT emulator (see cuda_emu.h) -- TEST INFRASTRUCTURE, never part of the product.
//
// Fibers: one context (own stack) per CUDA thread of the running block, scheduled round-robin; a fiber
// runs until it reaches a barrier / collective whose other participants have not arrived,
// or until its kernel body returns.  Also here: the handful of CUDA runtime entry points the
// library calls, implemented on host memory.
#include "cuda_emu.h"

#include <sys/mman.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <algorithm>
#include <chrono>
#include <vector>

// Context switch.  x86-64: callee-saved registers pushed on the old stack, stack pointers
// swapped (glibc's swapcontext makes a sigprocmask system call per switch -- 20x slower);
// elsewhere ucontext.
#if defined(__x86_64__)
extern "C" void mtn_emu_switch(void** save_sp, void* to_sp);
asm(R"(
.text
.globl mtn_emu_switch
.type mtn_emu_switch,@function
mtn_emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size mtn_emu_switch,.-mtn_emu_switch
)");
struct Ctx {
  void* sp = nullptr;
};
static inline void ctx_switch(Ctx* from, Ctx* to) { mtn_emu_switch(&from->sp, to->sp); }
static inline void ctx_make(Ctx* c, void* stack, size_t bytes, void (*entry)()) {
  // after the six pops `ret` enters `entry` with rsp = top - 8: the alignment of a fresh call
  uintptr_t top = ((uintptr_t)stack + bytes) & ~(uintptr_t)15;
  void** sp = (void**)top;
  *--sp = nullptr;        // entry's (never used) return address
  *--sp = (void*)entry;
  for (int i = 0; i < 6; ++i) *--sp = nullptr;
  c->sp = sp;
}
#else
struct Ctx {
  ucontext_t uc;
};
static inline void ctx_switch(Ctx* from, Ctx* to) { swapcontext(&from->uc, &to->uc); }
static inline void ctx_make(Ctx* c, void* stack, size_t bytes, void (*entry)()) {
  getcontext(&c->uc);
  c->uc.uc_stack.ss_sp = stack;
  c->uc.uc_stack.ss_size = bytes;
  c->uc.uc_link = nullptr;
  makecontext(&c->uc, entry, 0);
}
#endif

namespace mtn_emu {

Idx3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

namespace {

constexpr size_t STACK_BYTES = 256 << 10;

struct WarpState {
  unsigned alive = 0;     // lanes whose fiber has not returned
  unsigned arrived = 0;   // lanes that deposited into the collective in progress
  unsigned drained = 0;   // lanes that read it back
  unsigned mask = 0;      // mask of the collective in progress
  bool draining = false;
  uint64_t slot[32];
};

struct Fiber {
  Ctx ctx;
  void* stack = nullptr;
  bool done = false;
  Idx3 tid;
};

struct Block {
  std::vector<Fiber> fibers;
  std::vector<WarpState> warps;
  int n_threads = 0;
  int alive = 0;
  int bar_arrived = 0;
  unsigned bar_gen = 0;
  int current = -1;
  unsigned long progress = 0;  // bumped on every state change; a full idle round = deadlock
  const std::function<void()>* body = nullptr;
  Ctx sched;
};

Block* g_blk = nullptr;
std::vector<unsigned char> g_dyn_smem;
int g_violations = 0;
int g_sched_mode = 0;  // 0 forward, 1 reverse, 2 shuffled every round
uint64_t g_sched_rng = 0x9E3779B97F4A7C15ull;
std::vector<void*> g_stack_pool;

void* get_stack() {
  if (!g_stack_pool.empty()) {
    void* s = g_stack_pool.back();
    g_stack_pool.pop_back();
    return s;
  }
  void* s = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (s == MAP_FAILED) {
    perror("mtn_emu: mmap");
    abort();
  }
  return s;
}

void fiber_main() {
  Block* b = g_blk;
  Fiber& f = b->fibers[b->current];
  (*b->body)();
  // the thread has exited: it no longer takes part in barriers or collectives
  f.done = true;
  const int t = b->current;
  b->warps[t >> 5].alive &= ~(1u << (t & 31));
  b->alive -= 1;
  b->progress++;
  ctx_switch(&f.ctx, &b->sched);
  abort();  // a finished fiber is never resumed
}

}  // namespace

void yield() {
  Block* b = g_blk;
  Fiber& f = b->fibers[b->current];
  ctx_switch(&f.ctx, &b->sched);
  g_threadIdx = f.tid;  // (the scheduler also restores it; kept for clarity)
}

unsigned char* dyn_smem() { return g_dyn_smem.data(); }
void set_schedule(int mode, unsigned long long seed) {
  g_sched_mode = mode;
  g_sched_rng = seed ? seed : 0x9E3779B97F4A7C15ull;
}
int violations() { return g_violations; }

void block_barrier() {
  Block* b = g_blk;
  const unsigned gen = b->bar_gen;
  b->bar_arrived += 1;
  b->progress++;
  for (;;) {
    if (b->bar_gen != gen) return;
    if (b->bar_arrived >= b->alive) {  // exited threads count as arrived
      b->bar_arrived = 0;
      b->bar_gen += 1;
      b->progress++;
      return;
    }
    yield();
  }
}

void warp_exchange(unsigned mask, uint64_t val, uint64_t out[32], unsigned* present) {
  Block* b = g_blk;
  const int t = b->current, lane = t & 31;
  WarpState& w = b->warps[t >> 5];
  if (!((mask >> lane) & 1u)) {
    fprintf(stderr, "mtn_emu: lane %d calls a collective whose mask %08x excludes it\n", lane, mask);
    abort();
  }
  while (w.draining) yield();  // the previous collective is still being read
  if (w.arrived == 0) {
    w.mask = mask;
  } else if (w.mask != mask) {
    fprintf(stderr, "mtn_emu: lanes of one warp meet in collectives with masks %08x / %08x\n", w.mask, mask);
    abort();
  }
  w.slot[lane] = val;
  w.arrived |= 1u << lane;
  b->progress++;
  for (;;) {
    if (w.draining) break;
    const unsigned need = mask & w.alive;
    if ((w.arrived & need) == need) {
      if (need != mask) g_violations++;  // a named lane has already exited
      w.draining = true;
      b->progress++;
      break;
    }
    yield();
  }
  const unsigned need = w.arrived;
  for (int l = 0; l < 32; ++l) out[l] = w.slot[l];
  *present = need;
  w.drained |= 1u << lane;
  b->progress++;
  if ((w.drained & (need & w.alive)) == (need & w.alive)) {
    w.arrived = w.drained = 0;
    w.draining = false;
  }
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  const int nt = (int)(block.x * block.y * block.z);
  if (nt <= 0 || nt > 1024) {
    fprintf(stderr, "mtn_emu: bad block size %d\n", nt);
    abort();
  }
  if (g_dyn_smem.size() < smem_bytes + 128) g_dyn_smem.resize(smem_bytes + 128);
  g_gridDim = Idx3{grid.x, grid.y, grid.z};
  g_blockDim = Idx3{block.x, block.y, block.z};
  Block blk;
  blk.fibers.resize(nt);
  for (auto& f : blk.fibers) f.stack = get_stack();
  // the order in which the runnable threads of a block get their turn: results must not
  // depend on it (a missing barrier usually shows up under "reverse" or "shuffle")
  std::vector<int> order(nt);
  for (int t = 0; t < nt; ++t) order[t] = g_sched_mode == 1 ? nt - 1 - t : t;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blk.n_threads = blk.alive = nt;
        blk.bar_arrived = 0;
        blk.bar_gen = 0;
        blk.progress = 0;
        blk.body = &body;
        blk.warps.assign((nt + 31) / 32, WarpState());
        for (int t = 0; t < nt; ++t) {
          Fiber& f = blk.fibers[t];
          f.done = false;
          f.tid = Idx3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
          blk.warps[t >> 5].alive |= 1u << (t & 31);
          ctx_make(&f.ctx, f.stack, STACK_BYTES, fiber_main);
        }
        g_blk = &blk;
        g_blockIdx = Idx3{bx, by, bz};
        while (blk.alive > 0) {
          const unsigned long before = blk.progress;
          if (g_sched_mode == 2) {  // a fresh pseudo-random order every round (xorshift, seeded)
            for (int i = nt - 1; i > 0; --i) {
              g_sched_rng ^= g_sched_rng << 13;
              g_sched_rng ^= g_sched_rng >> 7;
              g_sched_rng ^= g_sched_rng << 17;
              std::swap(order[i], order[g_sched_rng % (uint64_t)(i + 1)]);
            }
          }
          for (int k = 0; k < nt; ++k) {
            const int t = order[k];
            Fiber& f = blk.fibers[t];
            if (f.done) continue;
            blk.current = t;
            g_threadIdx = f.tid;
            ctx_switch(&blk.sched, &f.ctx);
          }
          if (blk.progress == before && blk.alive > 0) {
            fprintf(stderr, "mtn_emu: deadlock in block (%u,%u,%u): %d threads wait forever\n", bx, by, bz,
                    blk.alive);
            abort();
          }
        }
        g_blk = nullptr;
      }
  for (auto& f : blk.fibers) g_stack_pool.push_back(f.stack);
}

}  // namespace mtn_emu

// ------------------------------------------------------------------ fake CUDA runtime
// Streams are ignored (everything is synchronous); "device" memory is host memory.
extern "C" {

int mtn_emu_violations(void) { return mtn_emu::violations(); }
void mtn_emu_set_schedule(int mode, unsigned long long seed) { mtn_emu::set_schedule(mode, seed); }

cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr attr, int) {
  switch (attr) {
    case cudaDevAttrMultiProcessorCount: *v = 2; break;  // a 2-SM "device": small persistent grids
    case cudaDevAttrComputeCapabilityMajor: *v = 10; break;
    case cudaDevAttrComputeCapabilityMinor: *v = 0; break;
    default: *v = 0;
  }
  return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "mtn_emu: no error text"; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memmove(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpyToSymbol(const void* sym, const void* src, size_t n, size_t off, cudaMemcpyKind) {
  memcpy((char*)const_cast<void*>(sym) + off, src, n);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t n) {
  *p = malloc(n);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) {
  free(p);
  return cudaSuccess;
}

struct CUevent_st {
  std::chrono::steady_clock::time_point t;
};
cudaError_t cudaEventCreate(cudaEvent_t* e) {
  *e = new CUevent_st();
  return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  e->t = std::chrono::steady_clock::now();
  return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}

}  // extern "C"
