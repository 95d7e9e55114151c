// cuda_emul.cpp -- fiber scheduler behind tests/emul/cuda_emul.h (TEST-ONLY, see that header).
#include "cuda_emul.h"

#include <stdio.h>
#include <algorithm>
#include <vector>

namespace napa_emul {

uint3_ threadIdx_, blockIdx_;
dim3 blockDim_, gridDim_;
unsigned char *dyn_smem = nullptr;
int shfl_slot[64];

namespace {
constexpr size_t kStack = 96 * 1024;
enum FState { READY, AT_BARRIER, DONE };
struct Fiber { ucontext_t ctx; FState st; unsigned char *stack; };
std::vector<Fiber> fibers;
ucontext_t sched_ctx;
int current = -1;
const std::function<void()> *cur_body = nullptr;

void trampoline() {
    (*cur_body)();
    fibers[current].st = DONE;
    swapcontext(&fibers[current].ctx, &sched_ctx);
}
}  // namespace

namespace {
unsigned bar_arrived[16];
unsigned long long bar_gen[16];
unsigned long long progress = 0;

// Order in which the live fibers get their turn in one scheduling round.  Fibers only switch at barriers and waits, so a
// MISSING barrier between a write and another thread's read shows up as a wrong result for the orders in which the
// reader runs first: the parity tests are therefore also run with NAPA_EMUL_ORDER=reverse and =shuffle (a poor man's
// racecheck for barrier placement; compute-sanitizer is not available on the GPU pool).
enum Order { ORDER_FORWARD, ORDER_REVERSE, ORDER_SHUFFLE };
Order sched_order() {
    const char *e = getenv("NAPA_EMUL_ORDER");
    if (!e) return ORDER_FORWARD;
    return e[0] == 'r' ? ORDER_REVERSE : (e[0] == 's' ? ORDER_SHUFFLE : ORDER_FORWARD);
}
unsigned long long rng_state = 0x9e3779b97f4a7c15ull;
void fill_order(std::vector<size_t> &idx, size_t n, Order o) {
    idx.resize(n);
    for (size_t i = 0; i < n; i++) idx[i] = o == ORDER_REVERSE ? n - 1 - i : i;
    if (o == ORDER_SHUFFLE)
        for (size_t i = n; i > 1; i--) {
            rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
            std::swap(idx[i - 1], idx[(size_t)((rng_state >> 33) % i)]);
        }
}
}

static void yield_fiber() {
    fibers[current].st = AT_BARRIER;
    swapcontext(&fibers[current].ctx, &sched_ctx);
}

// named barriers are per CTA: in cooperative mode every CTA has its own set (coop_named below)
static unsigned *named_arrived(int id);
static unsigned long long *named_gen(int id);
static void yield_current();

void note_progress() { progress++; }

void bar_arrive(int id, unsigned count) {
    progress++;
    if (++*named_arrived(id) >= count) { *named_arrived(id) = 0; ++*named_gen(id); }
}

void bar_sync(int id, unsigned count) {
    progress++;
    if (++*named_arrived(id) >= count) { *named_arrived(id) = 0; ++*named_gen(id); return; }
    unsigned long long *gen = named_gen(id);
    const unsigned long long g = *gen;
    while (*gen == g) yield_current();
}

void syncthreads();

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    // (plain launches: CTAs one after the other, see launch_cooperative below for the all-alive mode)
    const unsigned nthreads = block.x * block.y * block.z;
    if (fibers.size() < nthreads) {
        size_t old = fibers.size();
        fibers.resize(nthreads);
        for (size_t i = old; i < nthreads; i++) fibers[i].stack = (unsigned char *)malloc(kStack);
    }
    void *sm = nullptr;
    if (posix_memalign(&sm, 128, smem ? smem : 16)) abort();
    dyn_smem = (unsigned char *)sm;
    blockDim_ = block; gridDim_ = grid;
    cur_body = &body;
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx_ = { bx, by, bz };
        for (unsigned t = 0; t < nthreads; t++) {
            Fiber &f = fibers[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &sched_ctx;
            makecontext(&f.ctx, trampoline, 0);
            f.st = READY;
        }
        for (int i = 0; i < 16; i++) { bar_arrived[i] = 0; }
        const Order order = sched_order();
        std::vector<size_t> turn;
        for (;;) {
            unsigned live = 0;
            const unsigned long long before = progress;
            fill_order(turn, nthreads, order);
            for (unsigned k = 0; k < nthreads; k++) {
                const unsigned t = (unsigned)turn[k];
                Fiber &f = fibers[t];
                if (f.st == DONE) continue;
                current = (int)t;
                threadIdx_ = { t % block.x, (t / block.x) % block.y, t / (block.x * block.y) };
                f.st = READY;
                swapcontext(&sched_ctx, &f.ctx);
                if (f.st != DONE) live++; else progress++;
            }
            if (live == 0) break;
            if (progress == before) { fprintf(stderr, "napa_emul: deadlock (threads wait on a barrier nobody completes)\n"); abort(); }
        }
    }
    free(sm);
    dyn_smem = nullptr;
    cur_body = nullptr;
}


// ---- cooperative launch: every CTA of the grid is alive at once (group barriers between CTAs) ----
// One fiber per thread of the WHOLE grid, scheduled round-robin; __syncthreads() is per CTA, the
// dynamic shared memory is per CTA and swapped in with the fiber, and a fiber that spins on another
// CTA's progress gives way through yield_spin().
namespace {
struct CoopBlock { unsigned arrived; unsigned long long gen; unsigned char *smem; unsigned n_arrived[16]; unsigned long long n_gen[16]; };
std::vector<CoopBlock> coop_blocks;
std::vector<Fiber> coop_fibers;
bool coop_mode = false;
unsigned coop_threads = 0;
int coop_current = -1;
void coop_trampoline() {
    (*cur_body)();
    coop_fibers[coop_current].st = DONE;
    swapcontext(&coop_fibers[coop_current].ctx, &sched_ctx);
}
void coop_yield() {
    coop_fibers[coop_current].st = AT_BARRIER;
    swapcontext(&coop_fibers[coop_current].ctx, &sched_ctx);
}
}  // namespace

void yield_spin() {
    if (coop_mode) coop_yield();
}
static void yield_current() {
    if (coop_mode) coop_yield(); else yield_fiber();
}
void yield_any() { yield_current(); }
static unsigned *named_arrived(int id) {
    return coop_mode ? &coop_blocks[(size_t)coop_current / coop_threads].n_arrived[id] : &bar_arrived[id];
}
static unsigned long long *named_gen(int id) {
    return coop_mode ? &coop_blocks[(size_t)coop_current / coop_threads].n_gen[id] : &bar_gen[id];
}

void syncthreads() {
    if (!coop_mode) { bar_sync(0, blockDim_.x * blockDim_.y * blockDim_.z); return; }
    CoopBlock &cb = coop_blocks[(size_t)coop_current / coop_threads];
    progress++;
    if (++cb.arrived >= coop_threads) { cb.arrived = 0; cb.gen++; return; }
    const unsigned long long g = cb.gen;
    while (cb.gen == g) coop_yield();
}

void launch_cooperative(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    const unsigned nthreads = block.x * block.y * block.z, nblocks = grid.x * grid.y * grid.z;
    const size_t total = (size_t)nthreads * nblocks;
    if (coop_fibers.size() < total) {
        size_t old = coop_fibers.size();
        coop_fibers.resize(total);
        for (size_t i = old; i < total; i++) coop_fibers[i].stack = (unsigned char *)malloc(kStack);
    }
    coop_blocks.assign(nblocks, CoopBlock{});
    for (unsigned b = 0; b < nblocks; b++) {
        void *sm = nullptr;
        if (posix_memalign(&sm, 128, smem ? smem : 16)) abort();
        coop_blocks[b].smem = (unsigned char *)sm;
    }
    blockDim_ = block; gridDim_ = grid;
    cur_body = &body;
    coop_mode = true; coop_threads = nthreads;
    for (size_t i = 0; i < total; i++) {
        Fiber &f = coop_fibers[i];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &sched_ctx;
        makecontext(&f.ctx, coop_trampoline, 0);
        f.st = READY;
    }
    const Order order = sched_order();
    std::vector<size_t> turn;
    for (;;) {
        size_t live = 0;
        const unsigned long long before = progress;
        fill_order(turn, total, order);
        for (size_t k = 0; k < total; k++) {
            const size_t i = turn[k];
            Fiber &f = coop_fibers[i];
            if (f.st == DONE) continue;
            const unsigned b = (unsigned)(i / nthreads), t = (unsigned)(i % nthreads);
            coop_current = (int)i;
            blockIdx_ = { b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y) };
            threadIdx_ = { t % block.x, (t / block.x) % block.y, t / (block.x * block.y) };
            dyn_smem = coop_blocks[b].smem;
            f.st = READY;
            swapcontext(&sched_ctx, &f.ctx);
            if (f.st != DONE) live++; else progress++;
        }
        if (live == 0) break;
        if (progress == before) { fprintf(stderr, "napa_emul: cooperative launch made no progress (deadlock)\n"); abort(); }
    }
    for (unsigned b = 0; b < nblocks; b++) free(coop_blocks[b].smem);
    coop_blocks.clear();
    coop_mode = false;
    dyn_smem = nullptr;
    cur_body = nullptr;
}

}  // namespace napa_emul
