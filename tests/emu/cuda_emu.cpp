// Fiber scheduler of the CUDA-on-CPU shim (see cuda_emu.h) — TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <time.h>

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#if !defined(__x86_64__)
#error "tests/emu: the fiber switch is written for x86-64"
#endif

// save callee-saved registers on the current stack, publish its top, adopt the other stack
extern "C" void emu_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
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
.size emu_switch,.-emu_switch
)");

namespace emu {

thread_local uint3 t_threadIdx;
thread_local uint3 t_blockIdx;
thread_local dim3 t_blockDim;
thread_local dim3 t_gridDim;
thread_local uint8_t* t_dyn_smem;

namespace {

constexpr size_t kStackBytes = 256 * 1024;
enum State : int { FRESH = 0, RUNNABLE = 1, WAIT_BLOCK = 2, WAIT_WARP = 3, DONE = 4 };

struct Fiber {
  void* sp = nullptr;
  uint8_t* stack = nullptr;
  int state = FRESH;
  unsigned gen = 0;        // generation of the barrier the fiber waits on
  unsigned shfl = 0;       // shuffles executed by this thread (selects the slot set)
};

struct Worker {
  std::vector<Fiber> fibers;
  std::vector<uint8_t> smem;
  void* sched_sp = nullptr;
  int cur = -1;
  int nthreads = 0;
  const std::function<void()>* fn = nullptr;
  // block barrier
  int live = 0, bar_arrived = 0;
  unsigned bar_gen = 0;
  // warp barriers
  int warp_live[32], warp_arrived[32];
  unsigned warp_gen[32];
  uint64_t slots[32][64];
  dim3 bdim;
  bool reverse = false;    // MAED_EMU_ORDER=reverse: threads are scheduled from the last to the first (see README)
};

thread_local Worker* t_worker = nullptr;

void fiber_main() {
  Worker& w = *t_worker;
  (*w.fn)();
  Fiber& f = w.fibers[w.cur];
  f.state = DONE;
  // a thread that exits counts as arrived for every later barrier (CUDA: exited threads do not take part)
  const int warp = w.cur >> 5;
  --w.live;
  --w.warp_live[warp];
  if (w.live > 0 && w.bar_arrived >= w.live) { w.bar_arrived = 0; ++w.bar_gen; }
  if (w.warp_live[warp] > 0 && w.warp_arrived[warp] >= w.warp_live[warp]) { w.warp_arrived[warp] = 0; ++w.warp_gen[warp]; }
  void* dummy;
  emu_switch(&dummy, w.sched_sp);
  abort();     // a finished fiber is never resumed
}

void prepare(Fiber& f) {
  // initial frame: six zero registers + the return address emu_switch `ret`s to; the ABI wants rsp % 16 == 8 on entry
  uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
  void** sp = (void**)(top - 8);
  *--sp = (void*)&fiber_main;
  for (int i = 0; i < 6; ++i) *--sp = nullptr;
  f.sp = sp;
  f.state = RUNNABLE;
}

void set_thread_index(Worker& w, int t) {
  t_threadIdx.x = t % w.bdim.x;
  t_threadIdx.y = (t / w.bdim.x) % w.bdim.y;
  t_threadIdx.z = t / (w.bdim.x * w.bdim.y);
}

void yield_to_scheduler(Worker& w) {
  Fiber& f = w.fibers[w.cur];
  emu_switch(&f.sp, w.sched_sp);
}

void run_block(Worker& w) {
  const int n = w.nthreads;
  const int nwarps = (n + 31) / 32;
  w.live = n;
  w.bar_arrived = 0;
  for (int i = 0; i < nwarps; ++i) {
    w.warp_live[i] = (i == nwarps - 1) ? n - 32 * i : 32;
    w.warp_arrived[i] = 0;
  }
  for (int t = 0; t < n; ++t) { w.fibers[t].state = FRESH; w.fibers[t].shfl = 0; }
  int done = 0;
  while (done < n) {
    bool progress = false;
    for (int i = 0; i < n; ++i) {
      const int t = w.reverse ? n - 1 - i : i;
      Fiber& f = w.fibers[t];
      if (f.state == DONE) continue;
      if (f.state == WAIT_BLOCK && f.gen == w.bar_gen) continue;
      if (f.state == WAIT_WARP && f.gen == w.warp_gen[t >> 5]) continue;
      if (f.state == FRESH) prepare(f);
      f.state = RUNNABLE;
      w.cur = t;
      set_thread_index(w, t);
      emu_switch(&w.sched_sp, f.sp);
      progress = true;
      if (f.state == DONE) ++done;
    }
    if (!progress) {
      fprintf(stderr, "emu: deadlock in block (%u,%u,%u): %d of %d threads alive, none runnable "
              "(divergent __syncthreads / warp shuffle?)\n", t_blockIdx.x, t_blockIdx.y, t_blockIdx.z, n - done, n);
      abort();
    }
  }
}

}  // namespace

int lane_id() { return t_worker->cur & 31; }
int warp_lanes() {
  const Worker& w = *t_worker;
  const int warp = w.cur >> 5;
  const int rem = w.nthreads - 32 * warp;
  return rem < 32 ? rem : 32;
}
uint64_t* warp_slots() { return t_worker->slots[t_worker->cur >> 5]; }
unsigned shfl_parity() { return t_worker->fibers[t_worker->cur].shfl++; }

void block_sync() {
  Worker& w = *t_worker;
  Fiber& f = w.fibers[w.cur];
  if (++w.bar_arrived >= w.live) {
    w.bar_arrived = 0;
    ++w.bar_gen;
    return;
  }
  f.state = WAIT_BLOCK;
  f.gen = w.bar_gen;
  yield_to_scheduler(w);
}

void warp_sync() {
  Worker& w = *t_worker;
  const int warp = w.cur >> 5;
  Fiber& f = w.fibers[w.cur];
  if (++w.warp_arrived[warp] >= w.warp_live[warp]) {
    w.warp_arrived[warp] = 0;
    ++w.warp_gen[warp];
    return;
  }
  f.state = WAIT_WARP;
  f.gen = w.warp_gen[warp];
  yield_to_scheduler(w);
}

static int worker_count() {
  static int n = 0;
  if (!n) {
    const char* v = getenv("MAED_EMU_THREADS");
    n = v ? atoi(v) : (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if (n > 64) n = 64;
  }
  return n;
}

static std::mutex g_smem_mu;
static std::map<const void*, int> g_smem_optin;
void set_smem_optin(const void* kernel, int bytes) {
  std::lock_guard<std::mutex> lk(g_smem_mu);
  g_smem_optin[kernel] = bytes;
}
void check_smem_optin(const void* kernel, size_t bytes) {
  size_t limit = 48 * 1024;
  {
    std::lock_guard<std::mutex> lk(g_smem_mu);
    auto it = g_smem_optin.find(kernel);
    if (it != g_smem_optin.end()) limit = (size_t)it->second;
  }
  if (bytes > limit) {
    fprintf(stderr, "emu: launch with %zu bytes of dynamic shared memory, but the kernel's limit is %zu (48 KB unless "
            "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) raised it)\n", bytes, limit);
    abort();
  }
}

double g_prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // seconds: [0] fiber kernels, [1..] stubs (tc_stubs.cpp)
static double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
struct ProfScope { double t0; int slot; ProfScope(int s) : t0(now_s()), slot(s) {} ~ProfScope() { g_prof[slot] += now_s() - t0; } };

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& thread_fn) {
  ProfScope prof(0);
  const long long nblocks = (long long)grid.x * grid.y * grid.z;
  const int nthreads = (int)(block.x * block.y * block.z);
  if (nblocks <= 0 || nthreads <= 0 || nthreads > 1024 || grid.y > 65535 || grid.z > 65535 || grid.x > 2147483647u ||
      block.z > 64) {
    fprintf(stderr, "emu: invalid launch configuration grid=(%u,%u,%u) block=(%u,%u,%u)\n", grid.x, grid.y, grid.z, block.x,
            block.y, block.z);
    abort();
  }
  if (smem > 227 * 1024) { fprintf(stderr, "emu: %zu bytes of dynamic shared memory exceed 227 KB\n", smem); abort(); }
  std::atomic<long long> next{0};
  auto body = [&]() {
    Worker w;
    w.fibers.resize(nthreads);
    // one lazily committed arena holds the stacks of all fibers of this worker
    void* arena = mmap(nullptr, (size_t)nthreads * kStackBytes, PROT_READ | PROT_WRITE,
                       MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (arena == MAP_FAILED) { perror("emu: mmap"); abort(); }
    for (int t = 0; t < nthreads; ++t) w.fibers[t].stack = (uint8_t*)arena + (size_t)t * kStackBytes;
    w.smem.assign(smem + 128, 0);
    w.nthreads = nthreads;
    w.fn = &thread_fn;
    w.bdim = block;
    { const char* o = getenv("MAED_EMU_ORDER"); w.reverse = o && o[0] == 'r'; }
    t_worker = &w;
    t_blockDim = block;
    t_gridDim = grid;
    t_dyn_smem = (uint8_t*)(((uintptr_t)w.smem.data() + 127) & ~(uintptr_t)127);
    for (;;) {
      const long long b = next.fetch_add(1);
      if (b >= nblocks) break;
      t_blockIdx.x = (unsigned)(b % grid.x);
      t_blockIdx.y = (unsigned)((b / grid.x) % grid.y);
      t_blockIdx.z = (unsigned)(b / ((long long)grid.x * grid.y));
      run_block(w);
    }
    munmap(arena, (size_t)nthreads * kStackBytes);
    t_worker = nullptr;
  };
  int nw = worker_count();
  if (nw > nblocks) nw = (int)nblocks;
  if (nw <= 1) { body(); return; }
  std::vector<std::thread> pool;
  for (int i = 1; i < nw; ++i) pool.emplace_back(body);
  body();
  for (auto& t : pool) t.join();
}

}  // namespace emu

namespace emu {
double prof_now() { return now_s(); }
namespace {
struct KernelProf {
  std::mutex mu;
  std::map<const void*, std::pair<double, long>> t;
  ~KernelProf() {
    if (!getenv("MAED_EMU_PROFILE")) return;
    std::vector<std::pair<double, const void*>> v;
    for (auto& kv : t) v.push_back({kv.second.first, kv.first});
    std::sort(v.rbegin(), v.rend());
    for (auto& e : v) {
      Dl_info info;
      const char* name = (dladdr(e.second, &info) && info.dli_sname) ? info.dli_sname : "?";
      fprintf(stderr, "emu-profile %8.2fs %6ld launches  %s\n", e.first, t[e.second].second, name);
    }
  }
} g_kprof;
}  // namespace
void prof_kernel(const void* fn, double seconds) {
  std::lock_guard<std::mutex> lk(g_kprof.mu);
  auto& e = g_kprof.t[fn];
  e.first += seconds;
  e.second += 1;
}
}  // namespace emu

extern "C" void emu_profile(double* out8, int reset) {
  for (int i = 0; i < 8; ++i) { out8[i] = emu::g_prof[i]; if (reset) emu::g_prof[i] = 0; }
}
