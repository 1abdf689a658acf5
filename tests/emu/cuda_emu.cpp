// TEST INFRASTRUCTURE ONLY.  Fiber scheduler behind tests/emu/cuda_emu.h: one fiber per CUDA thread of the running
// block, round-robin until every fiber has finished; __syncthreads / __syncwarp / shuffles park a fiber until its
// block / warp has arrived.
#include "cuda_emu.h"

#include <sys/mman.h>

#include <vector>

extern "C" void emu_switch(void** from_sp, void* to_sp);
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

enum { READY = 0, DONE = 1, WAIT_BLOCK = 2, WAIT_WARP = 3 };
constexpr size_t STACK_BYTES = 512 * 1024;

struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  int state = DONE;
  uint3 tid{0, 0, 0};
  int lin = 0;
};

uint3 g_blockIdx{0, 0, 0}, g_blockDim{1, 1, 1}, g_gridDim{1, 1, 1};
uint3 g_threadIdx_dummy{0, 0, 0};
static std::vector<Fiber> fibers;          // stacks are kept across launches
static Fiber* cur = nullptr;
static void* sched_sp = nullptr;
static const std::function<void()>* body_fn = nullptr;
static std::vector<char> dyn;
static std::vector<uint64_t> warp_slot;    // [nwarps][32]
static int nthreads = 0;
static const bool reverse_order = [] { const char* e = std::getenv("EMU_ORDER"); return e && e[0] == 'r'; }();

uint3& cur_tid() { return cur ? cur->tid : g_threadIdx_dummy; }
void* dyn_smem() { return dyn.data(); }
int lane_id() { return cur->lin & 31; }

static void yield_to_scheduler() { emu_switch(&cur->sp, sched_sp); }

static void fiber_main() {
  (*body_fn)();
  cur->state = DONE;
  yield_to_scheduler();
  std::fprintf(stderr, "[emu] finished fiber resumed\n");
  std::abort();
}

static void prepare(Fiber& f) {
  if (!f.stack) {
    void* p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { std::perror("[emu] mmap"); std::abort(); }
    f.stack = (char*)p;
  }
  uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
  void** base = (void**)(top - 64);      // 6 callee-saved registers, return address, 8 bytes of padding
  for (int i = 0; i < 6; i++) base[i] = nullptr;
  base[6] = (void*)&fiber_main;
  f.sp = base;
  f.state = READY;
}

void sync_block() { cur->state = WAIT_BLOCK; yield_to_scheduler(); }
void sync_warp() { cur->state = WAIT_WARP; yield_to_scheduler(); }

uint64_t warp_exchange(uint64_t bits, int src_lane) {
  const int w = cur->lin >> 5;
  warp_slot[(size_t)w * 32 + (cur->lin & 31)] = bits;
  sync_warp();
  const uint64_t r = warp_slot[(size_t)w * 32 + src_lane];
  sync_warp();
  return r;
}
unsigned warp_vote(bool pred) {
  const int w = cur->lin >> 5;
  warp_slot[(size_t)w * 32 + (cur->lin & 31)] = pred ? 1u : 0u;
  sync_warp();
  unsigned m = 0;
  const int lo = w * 32, hi = lo + 32 < nthreads ? lo + 32 : nthreads;
  for (int t = lo; t < hi; t++) if (fibers[t].state != DONE && warp_slot[(size_t)w * 32 + (t & 31)]) m |= 1u << (t & 31);
  sync_warp();
  return m;
}

static void run_block() {
  for (int t = 0; t < nthreads; t++) prepare(fibers[t]);
  for (;;) {
    bool progressed = false, any_live = false;
    for (int t0 = 0; t0 < nthreads; t0++) {
      // EMU_ORDER=reverse runs the threads of a block (hence the lanes of a warp) in descending order between
      // rendezvous points: code that is correct under the CUDA model gives the same answer either way
      const int t = reverse_order ? nthreads - 1 - t0 : t0;
      Fiber& f = fibers[t];
      if (f.state != READY) { if (f.state != DONE) any_live = true; continue; }
      any_live = true;
      progressed = true;
      cur = &f;
      emu_switch(&sched_sp, f.sp);
      cur = nullptr;
    }
    if (!any_live) break;
    // release complete warp rendezvous
    const int nw = (nthreads + 31) / 32;
    for (int w = 0; w < nw; w++) {
      const int lo = w * 32, hi = lo + 32 < nthreads ? lo + 32 : nthreads;
      bool all = true, some = false;
      for (int t = lo; t < hi; t++) {
        if (fibers[t].state == WAIT_WARP) some = true;
        else if (fibers[t].state != DONE) all = false;
      }
      if (some && all) { for (int t = lo; t < hi; t++) if (fibers[t].state == WAIT_WARP) fibers[t].state = READY; progressed = true; }
    }
    // release a complete block barrier
    {
      bool all = true, some = false;
      for (int t = 0; t < nthreads; t++) {
        if (fibers[t].state == WAIT_BLOCK) some = true;
        else if (fibers[t].state != DONE) all = false;
      }
      if (some && all) { for (int t = 0; t < nthreads; t++) if (fibers[t].state == WAIT_BLOCK) fibers[t].state = READY; progressed = true; }
    }
    if (!progressed) {
      std::fprintf(stderr, "[emu] deadlock: block (%u,%u,%u), states:", g_blockIdx.x, g_blockIdx.y, g_blockIdx.z);
      for (int t = 0; t < nthreads && t < 64; t++) std::fprintf(stderr, " %d", fibers[t].state);
      std::fprintf(stderr, "\n");
      std::abort();
    }
  }
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  if (cur) { std::fprintf(stderr, "[emu] nested launch\n"); std::abort(); }
  nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || grid.x * grid.y * grid.z == 0) return;
  if ((int)fibers.size() < nthreads) fibers.resize(nthreads);
  warp_slot.assign((size_t)((nthreads + 31) / 32) * 32, 0);
  dyn.assign(smem + 64, 0);
  body_fn = &body;
  g_blockDim = uint3{block.x, block.y, block.z};
  g_gridDim = uint3{grid.x, grid.y, grid.z};
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        g_blockIdx = uint3{bx, by, bz};
        int lin = 0;
        for (unsigned tz = 0; tz < block.z; tz++)
          for (unsigned ty = 0; ty < block.y; ty++)
            for (unsigned tx = 0; tx < block.x; tx++, lin++) { fibers[lin].tid = uint3{tx, ty, tz}; fibers[lin].lin = lin; }
        run_block();
      }
  body_fn = nullptr;
}

}  // namespace emu
