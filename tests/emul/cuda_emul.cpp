// cuda_emul.cpp -- TEST INFRASTRUCTURE ONLY: the execution engine behind cuda_emul.h.
#include "cuda_emul.h"

thread_local dim3 threadIdx;
dim3 blockIdx, blockDim, gridDim;
long long emul_graph_replays = 0;
static emul_graph* g_capture = nullptr;  // non-null while a stream capture is open
static bool g_replaying = false;

cudaError_t cudaStreamBeginCapture(cudaStream_t, int) {
  if (g_capture) return 1;
  g_capture = new emul_graph();
  return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) {
  *g = g_capture;
  g_capture = nullptr;
  return *g ? cudaSuccess : 1;
}

#include <map>
static std::mutex g_reg_mutex;
static std::map<uintptr_t, size_t> g_reg;
void emul_register(const void* p, size_t bytes) { std::lock_guard<std::mutex> lk(g_reg_mutex); g_reg[(uintptr_t)p] = bytes; }
void emul_unregister(const void* p) { std::lock_guard<std::mutex> lk(g_reg_mutex); g_reg.erase((uintptr_t)p); }
bool emul_is_device(const void* p) {
  std::lock_guard<std::mutex> lk(g_reg_mutex);
  auto it = g_reg.upper_bound((uintptr_t)p);
  if (it == g_reg.begin()) return false;
  --it;
  return (uintptr_t)p < it->first + it->second;
}

#ifdef EMUL_PTHREADS  // one OS thread per CUDA thread (the AddressSanitizer build: ASan and swapcontext do not mix well)
namespace {

// barrier whose participants may leave for good (a CUDA thread that returns stops counting)
struct DropBarrier {
  std::mutex m;
  std::condition_variable cv;
  int expected = 0, arrived = 0;
  unsigned gen = 0;
  void reset(int n) { expected = n; arrived = 0; }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++arrived >= expected) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
  void drop() {
    std::unique_lock<std::mutex> lk(m);
    --expected;
    if (expected > 0 && arrived >= expected) { arrived = 0; ++gen; cv.notify_all(); }
  }
};

constexpr int kMaxWarps = 32;
DropBarrier g_block_bar;
DropBarrier g_warp_bar[kMaxWarps];
double g_warp_buf[kMaxWarps][32];
char* g_dyn = nullptr;
thread_local int t_warp = 0, t_lane = 0;

struct Launch {
  const std::function<void()>* body;
  int nthreads;
  pthread_barrier_t start, done;
  bool stop = false;
};

struct WorkerArg { Launch* L; int tid; };

void* worker(void* p) {
  WorkerArg* a = static_cast<WorkerArg*>(p);
  Launch* L = a->L;
  threadIdx = dim3((unsigned)a->tid, 0, 0);
  t_warp = a->tid / 32;
  t_lane = a->tid % 32;
  for (;;) {
    pthread_barrier_wait(&L->start);
    if (L->stop) break;
    (*L->body)();
    g_warp_bar[t_warp].drop();
    g_block_bar.drop();
    pthread_barrier_wait(&L->done);
  }
  return nullptr;
}

}  // namespace

void __syncthreads() { g_block_bar.wait(); }
void __syncwarp(unsigned) { g_warp_bar[t_warp].wait(); }
void* emul_dyn_smem() { return g_dyn; }

void emul_warp_allgather(double v, double (&all)[32]) {
  g_warp_buf[t_warp][t_lane] = v;
  g_warp_bar[t_warp].wait();
  for (int i = 0; i < 32; ++i) all[i] = g_warp_buf[t_warp][i];
  g_warp_bar[t_warp].wait();
}

void emul_launch_impl(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body) {
  if (block.y != 1 || block.z != 1 || block.x == 0 || block.x > 32u * kMaxWarps) {
    std::fprintf(stderr, "[emul] unsupported block shape %u x %u x %u\n", block.x, block.y, block.z);
    std::abort();
  }
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return;
  if (g_capture) {  // record: the body owns copies of the kernel arguments
    const std::function<void()> copy = body;
    g_capture->launches.push_back([=] { g_replaying = true; emul_launch_impl(grid, block, dyn_smem, copy); g_replaying = false; });
    return;
  }
  if (g_replaying) ++emul_graph_replays;
  const int nthreads = (int)block.x;
  gridDim = grid;
  blockDim = block;
  void* dyn = nullptr;
  if (posix_memalign(&dyn, 128, dyn_smem + 128) != 0) std::abort();
  g_dyn = static_cast<char*>(dyn);
  Launch L;
  L.body = &body;
  L.nthreads = nthreads;
  pthread_barrier_init(&L.start, nullptr, (unsigned)nthreads + 1);
  pthread_barrier_init(&L.done, nullptr, (unsigned)nthreads + 1);
  std::vector<pthread_t> th(nthreads);
  std::vector<WorkerArg> args(nthreads);
  pthread_attr_t at;
  pthread_attr_init(&at);
  pthread_attr_setstacksize(&at, 512 * 1024);
  for (int i = 0; i < nthreads; ++i) {
    args[i] = WorkerArg{&L, i};
    if (pthread_create(&th[i], &at, worker, &args[i]) != 0) { std::perror("pthread_create"); std::abort(); }
  }
  const int nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx = dim3(bx, by, bz);
        std::memset(g_dyn, 0xEE, dyn_smem);
        g_block_bar.reset(nthreads);
        for (int w = 0; w < nwarps; ++w) g_warp_bar[w].reset(std::min(32, nthreads - 32 * w));
        pthread_barrier_wait(&L.start);
        pthread_barrier_wait(&L.done);
      }
  L.stop = true;
  pthread_barrier_wait(&L.start);
  for (int i = 0; i < nthreads; ++i) pthread_join(th[i], nullptr);
  pthread_attr_destroy(&at);
  pthread_barrier_destroy(&L.start);
  pthread_barrier_destroy(&L.done);
  std::free(dyn);
  g_dyn = nullptr;
}
#else  // default engine: the threads of a block are fibers (ucontext) on the calling OS thread, scheduled round-robin; a
       // barrier yields to the scheduler, which releases it once every thread that has not returned is waiting at it.
       // No futex traffic: an order of magnitude faster than one OS thread per CUDA thread, and deterministic.
#include <ucontext.h>

namespace {

enum { kRunnable = 0, kWaitBlock = 1, kWaitWarp = 2, kDone = 3 };
struct Fiber {
  ucontext_t ctx;
  int state;
};
constexpr size_t kStackBytes = 256 * 1024;
ucontext_t g_sched;
std::vector<Fiber> g_fibers;
char* g_stacks = nullptr;
size_t g_stacks_n = 0;
int g_cur = 0;
const std::function<void()>* g_body = nullptr;
double g_warp_buf[32][32];
char* g_dyn = nullptr;

void yield_as(int state) {
  g_fibers[g_cur].state = state;
  swapcontext(&g_fibers[g_cur].ctx, &g_sched);
}

void fiber_entry() {
  (*g_body)();
  g_fibers[g_cur].state = kDone;
  // returning resumes uc_link = the scheduler
}

void run_block(int nthreads) {
  for (int i = 0; i < nthreads; ++i) {
    Fiber& f = g_fibers[i];
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = g_stacks + (size_t)i * kStackBytes;
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = &g_sched;
    makecontext(&f.ctx, fiber_entry, 0);
    f.state = kRunnable;
  }
  const int nwarps = (nthreads + 31) / 32;
  for (;;) {
    bool ran = false;
    for (int i = 0; i < nthreads; ++i) {
      if (g_fibers[i].state != kRunnable) continue;
      g_cur = i;
      threadIdx = dim3((unsigned)i, 0, 0);
      swapcontext(&g_sched, &g_fibers[i].ctx);
      ran = true;
    }
    // release the barriers that are complete
    int done = 0, at_block = 0;
    for (int i = 0; i < nthreads; ++i) { done += g_fibers[i].state == kDone; at_block += g_fibers[i].state == kWaitBlock; }
    if (done == nthreads) return;
    bool released = false;
    for (int w = 0; w < nwarps; ++w) {
      int live = 0, waiting = 0;
      for (int i = 32 * w; i < std::min(nthreads, 32 * w + 32); ++i) { live += g_fibers[i].state != kDone; waiting += g_fibers[i].state == kWaitWarp; }
      if (waiting > 0 && waiting == live) {
        for (int i = 32 * w; i < std::min(nthreads, 32 * w + 32); ++i) if (g_fibers[i].state == kWaitWarp) g_fibers[i].state = kRunnable;
        released = true;
      }
    }
    if (at_block > 0 && at_block == nthreads - done) {
      for (int i = 0; i < nthreads; ++i) if (g_fibers[i].state == kWaitBlock) g_fibers[i].state = kRunnable;
      released = true;
    }
    if (!ran && !released) {
      std::fprintf(stderr, "[emul] deadlock in block (%u,%u,%u): %d threads at __syncthreads, %d returned, the rest at warp barriers\n",
                   blockIdx.x, blockIdx.y, blockIdx.z, at_block, done);
      std::abort();
    }
  }
}

}  // namespace

void __syncthreads() { yield_as(kWaitBlock); }
void __syncwarp(unsigned) { yield_as(kWaitWarp); }
void* emul_dyn_smem() { return g_dyn; }

void emul_warp_allgather(double v, double (&all)[32]) {
  const int warp = g_cur / 32, lane = g_cur % 32;
  g_warp_buf[warp][lane] = v;
  yield_as(kWaitWarp);
  for (int i = 0; i < 32; ++i) all[i] = g_warp_buf[warp][i];
  yield_as(kWaitWarp);
}

void emul_launch_impl(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body) {
  if (block.y != 1 || block.z != 1 || block.x == 0 || block.x > 1024) {
    std::fprintf(stderr, "[emul] unsupported block shape %u x %u x %u\n", block.x, block.y, block.z);
    std::abort();
  }
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return;
  if (g_capture) {  // record: the body owns copies of the kernel arguments
    const std::function<void()> copy = body;
    g_capture->launches.push_back([=] { g_replaying = true; emul_launch_impl(grid, block, dyn_smem, copy); g_replaying = false; });
    return;
  }
  if (g_replaying) ++emul_graph_replays;
  if (g_body != nullptr) { std::fprintf(stderr, "[emul] nested or concurrent launches are not supported\n"); std::abort(); }
  const int nthreads = (int)block.x;
  gridDim = grid;
  blockDim = block;
  if (g_stacks_n < (size_t)nthreads) {
    std::free(g_stacks);
    if (posix_memalign(reinterpret_cast<void**>(&g_stacks), 4096, (size_t)nthreads * kStackBytes) != 0) std::abort();
    g_stacks_n = (size_t)nthreads;
  }
  g_fibers.resize(nthreads);
  void* dyn = nullptr;
  if (posix_memalign(&dyn, 128, dyn_smem + 128) != 0) std::abort();
  g_dyn = static_cast<char*>(dyn);
  g_body = &body;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx = dim3(bx, by, bz);
        std::memset(g_dyn, 0xEE, dyn_smem);
        run_block(nthreads);
      }
  g_body = nullptr;
  std::free(dyn);
  g_dyn = nullptr;
}
#endif
