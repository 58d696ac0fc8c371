// TEST INFRASTRUCTURE ONLY -- fiber scheduler of the SIMT emulator (see simt_runtime.h).
#include "simt_runtime.h"

#include <sys/mman.h>

#include <memory>

#if !defined(__x86_64__)
#error "the emulator's context switch is written for x86-64 (the CPU test tier of this repo runs there)"
#endif

// callee-saved registers + stack pointer; everything else is dead across a call by the SysV ABI
extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
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
.size simt_switch,.-simt_switch
)");

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace simt {

thread_local Thread* cur = nullptr;
thread_local Graph* capturing = nullptr;

static std::atomic<long long> n_launches{0}, n_switches{0};
long long coll_count[kCollEnd] = {0};
long long counters(int which) { return which == 0 ? n_launches.load() : which == 1 ? n_switches.load() : which < kCollEnd ? coll_count[which] : 0; }

void submit(std::function<void()> op) {
    if (capturing) capturing->ops.push_back(std::move(op));
    else op();
}

unsigned long long sched_seed = 0, sched_state = 0;

namespace {
constexpr size_t kStackBytes = 256 * 1024;

struct StackPool {
    std::vector<char*> stacks;
    char* get(size_t i) {
        while (stacks.size() <= i) {
            void* p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (p == MAP_FAILED) { std::perror("simt: mmap"); std::abort(); }
            stacks.push_back((char*)p);
        }
        return stacks[i];
    }
};
thread_local StackPool pool;
thread_local void* sched_sp = nullptr;
thread_local const std::function<void()>* body_fn = nullptr;
thread_local long long local_switches = 0;
thread_local int live_fibers = 0;

void release_if_complete(Thread* t) {
    // a thread that exits no longer takes part in barriers (whole warps return early in the env-step kernel)
    Warp* w = t->warp;
    if (--w->alive > 0 && w->arrived >= w->alive) { w->arrived = 0; w->gen++; }
    Block* b = t->block;
    if (--b->alive > 0 && b->arrived >= b->alive) { b->arrived = 0; b->or_acc[(b->gen + 1) & 1] = 0; b->gen++; }
    Cluster* c = t->cluster;
    if (--c->alive > 0 && c->arrived >= c->alive) { c->arrived = 0; c->gen++; }
}

void fiber_entry() {
    Thread* t = cur;
    (*body_fn)();
    t->done = true;
    --live_fibers;
    release_if_complete(t);
    void* dummy;
    simt_switch(&dummy, sched_sp);
    std::abort();   // a finished fiber is never resumed
}
}  // namespace

void yield_until_changed(const volatile unsigned* gen, unsigned val) {
    Thread* t = cur;
    t->wait_gen = gen;
    t->wait_val = val;
    ++local_switches;
    // fast path (default schedule only): a lane that waits hands the processor straight to the next runnable lane of its own
    // warp -- the common case inside warp collectives -- instead of going through the scheduler loop
    if (!sched_seed && t->warp_lanes > 1) {
        const int nl = t->warp_lanes;
        for (int k = 1; k < nl; ++k) {
            Thread& c = t->warp_base[(t->lane + k) % nl];
            if (c.done || (c.wait_gen && *c.wait_gen == c.wait_val)) continue;
            c.wait_gen = nullptr;
            cur = &c;
            threadIdx = c.tid;
            simt_switch(&t->sp, c.sp);
            return;
        }
    }
    // ... or to the next runnable thread of the group (CTA-wide and cluster-wide barriers)
    if (!sched_seed) {
        const int ng = t->group_size;
        for (int k = 1; k < ng; ++k) {
            int i = t->group_index + k;
            if (i >= ng) i -= ng;
            Thread& c = t->group_base[i];
            if (c.done || (c.wait_gen && *c.wait_gen == c.wait_val)) continue;
            c.wait_gen = nullptr;
            cur = &c;
            threadIdx = c.tid; blockIdx = c.bid;
            simt_switch(&t->sp, c.sp);
            return;
        }
    }
    simt_switch(&t->sp, sched_sp);
}

void run_grid(dim3 grid, dim3 block, size_t dyn_smem_bytes, int cluster_size, const std::function<void()>& body) {
    if (cur) { std::fprintf(stderr, "simt: nested launch\n"); std::abort(); }
    n_launches++;
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nblocks = (int)(grid.x * grid.y * grid.z);
    if (cluster_size < 1) cluster_size = 1;
    if (grid.x % cluster_size) { std::fprintf(stderr, "simt: grid.x %u not a multiple of the cluster size %d\n", grid.x, cluster_size); std::abort(); }
    const int nwarps = (nthreads + 31) / 32;
    blockDim = block; gridDim = grid;
    body_fn = &body;
    const int group_threads = nthreads * cluster_size;
    std::vector<Thread> threads(group_threads);
    std::vector<Warp> warps((size_t)nwarps * cluster_size);
    std::vector<Block> blocks(cluster_size);
    std::vector<std::unique_ptr<char[]>> smem(cluster_size);
    for (int b = 0; b < cluster_size; ++b) smem[b].reset(new char[dyn_smem_bytes + 64]);
    for (int first = 0; first < nblocks; first += cluster_size) {
        Cluster cl;
        cl.nblocks = cluster_size; cl.blocks = blocks.data(); cl.alive = group_threads;
        for (int b = 0; b < cluster_size; ++b) {
            blocks[b] = Block();
            blocks[b].alive = nthreads;
            blocks[b].rank_in_cluster = b;
            blocks[b].dyn_smem = (char*)(((uintptr_t)smem[b].get() + 63) & ~(uintptr_t)63);
            const int lin = first + b;
            dim3 bid(lin % grid.x, (lin / grid.x) % grid.y, lin / (grid.x * grid.y));
            for (int w = 0; w < nwarps; ++w) {
                Warp& wp = warps[(size_t)b * nwarps + w];
                wp = Warp();
                wp.alive = std::min(32, nthreads - 32 * w);
            }
            for (int t = 0; t < nthreads; ++t) {
                Thread& th = threads[(size_t)b * nthreads + t];
                th = Thread();
                th.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                th.bid = bid;
                th.lane = t & 31;
                th.warp = &warps[(size_t)b * nwarps + t / 32];
                th.block = &blocks[b];
                th.cluster = &cl;
                th.warp_base = &threads[(size_t)b * nthreads + (t / 32) * 32];
                th.warp_lanes = std::min(32, nthreads - (t / 32) * 32);
                th.group_base = threads.data(); th.group_size = group_threads; th.group_index = b * nthreads + t;
                char* top = pool.get((size_t)b * nthreads + t) + kStackBytes;
                void** sp = (void**)(((uintptr_t)top & ~(uintptr_t)15));
                *--sp = nullptr;                    // fake return address of fiber_entry (keeps rsp = 8 mod 16 at entry)
                *--sp = (void*)&fiber_entry;
                for (int r = 0; r < 6; ++r) *--sp = nullptr;
                th.sp = sp;
            }
        }
        live_fibers = group_threads;
        while (live_fibers > 0) {
            bool progressed = false;
            // SIMT_SCHED_SEED: visit the fibers in a different rotation / direction on every pass.  Between two collectives the
            // lanes of a warp then run in varying orders, so a shared-memory hand-off that lacks its __syncwarp / __syncthreads
            // shows up as a result that depends on the seed (tests compare seeds bit for bit).
            int start = 0, dir = 1;
            if (sched_seed) {
                sched_state = sched_state * 6364136223846793005ull + 1442695040888963407ull;
                start = (int)((sched_state >> 33) % (unsigned long long)group_threads);
                dir = ((sched_state >> 20) & 1) ? 1 : -1;
            }
            for (int k = 0; k < group_threads; ++k) {
                const int i = ((start + dir * k) % group_threads + group_threads) % group_threads;
                Thread& th = threads[i];
                if (th.done) continue;
                if (th.wait_gen && *th.wait_gen == th.wait_val) continue;
                th.wait_gen = nullptr;
                cur = &th;
                threadIdx = th.tid; blockIdx = th.bid;
                simt_switch(&sched_sp, th.sp);
                cur = nullptr;
                progressed = true;
            }
            if (!progressed) {
                std::fprintf(stderr, "simt: deadlock -- a collective / barrier was not reached by every live thread "
                                     "(block %d, %d threads still waiting)\n", first, live_fibers);
                for (int i = 0; i < group_threads && i < 64; ++i)
                    if (!threads[i].done) std::fprintf(stderr, "  thread %d (block rank %d) waits\n", i % nthreads, i / nthreads);
                std::abort();
            }
        }
    }
    body_fn = nullptr;
    n_switches += local_switches;
    local_switches = 0;
}

}  // namespace simt

extern "C" void simt_set_sched_seed(unsigned long long seed) { simt::sched_seed = seed; simt::sched_state = seed; }
extern "C" long long simt_counter(int which) { return simt::counters(which); }
