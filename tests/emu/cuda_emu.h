// SIMT emulator for the CPU test suite -- TEST INFRASTRUCTURE, never part of the product.
//
// Force-included (g++ -include) in front of martini_b200/csrc/api.cu when tests/emu/build.py
// compiles the CUDA sources as plain C++ into tests/emu/libmartini_emu.so.  The kernels run
// unmodified: every CUDA thread of a block is a fiber (ucontext), blocks run one after the
// other, __syncthreads / warp collectives / mbarrier waits are scheduling points, "device"
// memory is host memory.  This checks the kernels' LOGIC (indexing, predicates, barriers,
// arithmetic) against the oracle without a GPU; it says nothing about performance or about
// hardware memory-ordering, and the product (martini_b200/_lib.py) never loads it.
#pragma once
#define MTN_HOST_EMU 1

#include <cuda_runtime.h>  // types and the runtime API's declarations only; no libcudart
#include <math_constants.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ------------------------------------------------------------------------- qualifiers
#undef __shared__
#define __shared__ static  // blocks run one at a time: one static copy is the block's copy
#undef __launch_bounds__
#define __launch_bounds__(...)

namespace mtn_emu {

struct Idx3 {
  unsigned x = 0, y = 0, z = 0;
};
extern Idx3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;  // of the running fiber

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);
unsigned char* dyn_smem();
void yield();

// warp collective: every lane of `mask` deposits `val`, gets all 32 deposits back
void warp_exchange(unsigned mask, uint64_t val, uint64_t out[32], unsigned* present);
void block_barrier();
void set_schedule(int mode, unsigned long long seed);
int violations();  // collectives entered with exited lanes in their mask, etc.

}  // namespace mtn_emu

#define threadIdx (mtn_emu::g_threadIdx)
#define blockIdx (mtn_emu::g_blockIdx)
#define blockDim (mtn_emu::g_blockDim)
#define gridDim (mtn_emu::g_gridDim)

#define MTN_EMU_LAUNCH(kernel, grid, block, smem, stream, ...) \
  mtn_emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); })

// --------------------------------------------------------------------- barriers, collectives
inline void __syncthreads() { mtn_emu::block_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  uint64_t o[32];
  unsigned present;
  mtn_emu::warp_exchange(mask, 0, o, &present);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  uint64_t o[32];
  unsigned present, r = 0;
  mtn_emu::warp_exchange(mask, pred ? 1 : 0, o, &present);
  for (int l = 0; l < 32; ++l)
    if (((present >> l) & 1u) && o[l]) r |= 1u << l;
  return r;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  uint64_t o[32];
  unsigned present, r = 0;
  mtn_emu::warp_exchange(mask, v, o, &present);
  for (int l = 0; l < 32; ++l)
    if ((present >> l) & 1u) r |= (unsigned)o[l];
  return r;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
  uint64_t o[32];
  unsigned present, r = 0;
  mtn_emu::warp_exchange(mask, v, o, &present);
  for (int l = 0; l < 32; ++l)
    if ((present >> l) & 1u) r += (unsigned)o[l];
  return r;
}
inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
  uint64_t o[32];
  unsigned present, r = 0xffffffffu;
  mtn_emu::warp_exchange(mask, v, o, &present);
  for (int l = 0; l < 32; ++l)
    if (((present >> l) & 1u) && (unsigned)o[l] < r) r = (unsigned)o[l];
  return r;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  uint64_t o[32];
  unsigned present, r = 0;
  mtn_emu::warp_exchange(mask, v, o, &present);
  for (int l = 0; l < 32; ++l)
    if (((present >> l) & 1u) && (unsigned)o[l] > r) r = (unsigned)o[l];
  return r;
}
inline unsigned __match_any_sync(unsigned mask, unsigned v) {
  uint64_t o[32];
  unsigned present, r = 0;
  mtn_emu::warp_exchange(mask, v, o, &present);
  for (int l = 0; l < 32; ++l)
    if (((present >> l) & 1u) && (unsigned)o[l] == v) r |= 1u << l;
  return r;
}
namespace mtn_emu {
template <typename T>
inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffles move at most 8 bytes");
  uint64_t b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <typename T>
inline T from_bits(uint64_t b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
template <typename T>
inline T shfl_from(unsigned mask, T v, int src, bool in_range) {
  uint64_t o[32];
  unsigned present;
  warp_exchange(mask, to_bits(v), o, &present);
  if (!in_range || !((present >> (src & 31)) & 1u)) return v;
  return from_bits<T>(o[src & 31]);
}
}  // namespace mtn_emu
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src) {
  return mtn_emu::shfl_from(mask, v, src & 31, true);
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned d) {
  const int lane = (int)(threadIdx.x & 31u);
  return mtn_emu::shfl_from(mask, v, lane - (int)d, lane - (int)d >= 0);
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned d) {
  const int lane = (int)(threadIdx.x & 31u);
  return mtn_emu::shfl_from(mask, v, lane + (int)d, lane + (int)d < 32);
}

// ------------------------------------------------------------------------------ atomics
// fibers are cooperative and blocks sequential: plain read-modify-write is atomic
template <typename T, typename U>
inline T atomicAdd(T* p, U v) {
  const T old = *p;
  *p = (T)(old + (T)v);
  return old;
}

// --------------------------------------------------------------------------- intrinsics
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __double2hiint(double d) {
  uint64_t b;
  memcpy(&b, &d, 8);
  return (int)(b >> 32);
}
inline int __double2loint(double d) {
  uint64_t b;
  memcpy(&b, &d, 8);
  return (int)(uint32_t)b;
}
inline double __longlong_as_double(long long v) {
  double d;
  memcpy(&d, &v, 8);
  return d;
}
inline long long __double_as_longlong(double d) {
  long long v;
  memcpy(&v, &d, 8);
  return v;
}
inline double __hiloint2double(int hi, int lo) {
  const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  memcpy(&d, &b, 8);
  return d;
}
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline void __trap() {
  fprintf(stderr, "mtn_emu: __trap()\n");
  abort();
}
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
using std::isnan;

// CUDA's mixed-type min / max overloads
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
inline double min(double a, double b) { return fmin(a, b); }
inline double max(double a, double b) { return fmax(a, b); }

// cudaFuncSetAttribute on a kernel (the runtime's template overload exists only under nvcc)
template <typename... A>
inline cudaError_t cudaFuncSetAttribute(void (*)(A...), cudaFuncAttribute, int) { return cudaSuccess; }

// ----------------------------------------------------------- mbarrier + bulk copy (project.cuh)
// A bulk copy is asynchronous on the GPU: its bytes land at some time between issue and the
// completion of the mbarrier phase it reports to.  The emulator makes both ends of that window
// bite: at issue the destination is POISONED (0xff bytes: NaN as a double, -1 as an integer),
// so a thread still reading the buffer being refilled, or reading it before it has observed
// the barrier, gets garbage; the bytes land as late as possible, when a waiter first polls the
// barrier.  The barrier completes a phase once its pending arrivals and its transaction bytes
// are both zero -- the PTX mbarrier contract.
namespace mtn {
struct EmuBar {
  int32_t tx;       // outstanding transaction bytes (may go negative before expect_tx)
  uint16_t pending; // arrivals still expected in this phase
  uint8_t count;    // arrivals per phase
  uint8_t phase;    // parity of the phase in progress
};
static_assert(sizeof(EmuBar) == 8, "lives in the kernel's uint64_t mbarrier slot");
struct EmuBulkCopy {
  void* dst;
  const void* src;
  uint32_t bytes;
  uint64_t* bar;
};
inline std::vector<EmuBulkCopy>& emu_bulk_in_flight() {
  static std::vector<EmuBulkCopy> v;  // blocks run one after the other: one list
  return v;
}
inline void emu_bar_check(EmuBar* b) {
  if (b->pending == 0 && b->tx == 0) {
    b->phase ^= 1;
    b->pending = b->count;
  }
}
inline void emu_bulk_land(uint64_t* bar) {  // the copies reporting to `bar` arrive now
  auto& v = emu_bulk_in_flight();
  EmuBar* b = reinterpret_cast<EmuBar*>(bar);
  size_t keep = 0;
  for (size_t i = 0; i < v.size(); ++i) {
    if (v[i].bar == bar) {
      memcpy(v[i].dst, v[i].src, v[i].bytes);
      b->tx -= (int32_t)v[i].bytes;
    } else {
      v[keep++] = v[i];
    }
  }
  v.resize(keep);
  emu_bar_check(b);
}
inline void mbar_init(uint64_t* bar, uint32_t count) {
  EmuBar* b = reinterpret_cast<EmuBar*>(bar);
  b->tx = 0;
  b->pending = (uint16_t)count;
  b->count = (uint8_t)count;
  b->phase = 0;
  auto& v = emu_bulk_in_flight();  // a new block: nothing of the previous one is in flight
  size_t keep = 0;
  for (size_t i = 0; i < v.size(); ++i)
    if (v[i].bar != bar) v[keep++] = v[i];
  v.resize(keep);
}
inline void mbar_fence_init() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  EmuBar* b = reinterpret_cast<EmuBar*>(bar);
  b->tx += (int32_t)bytes;
  b->pending -= 1;
  emu_bar_check(b);
}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  EmuBar* b = reinterpret_cast<EmuBar*>(bar);
  if (b->phase != (uint8_t)parity) return true;  // the phase with that parity has completed
  emu_bulk_land(bar);
  if (b->phase != (uint8_t)parity) return true;
  mtn_emu::yield();
  return false;
}
inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  if (bytes % 16 || (uintptr_t)dst_smem % 16 || (uintptr_t)src_gmem % 16) {
    fprintf(stderr, "mtn_emu: cp.async.bulk needs 16-byte aligned size and addresses\n");
    abort();
  }
  memset(dst_smem, 0xff, bytes);  // in flight: whoever reads it now reads garbage
  emu_bulk_in_flight().push_back(EmuBulkCopy{dst_smem, src_gmem, bytes, bar});
}
}  // namespace mtn
