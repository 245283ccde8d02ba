// cuda_emu.cpp -- runtime of the test-only CUDA execution-model stand-in (see cuda_emu.h).
#include "cuda_emu.h"
#include <mutex>
#include <sys/mman.h>

thread_local AzEmuFiber* az_emu_cur = nullptr;

namespace {
constexpr size_t kStack = 128 * 1024;
struct Worker {
    std::vector<char*> stacks;
    ~Worker() { for (char* p : stacks) munmap(p, kStack); }
    char* stack(int i)
    {
        while ((int)stacks.size() <= i) {
            void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (p == MAP_FAILED) { perror("emu mmap"); abort(); }
            stacks.push_back((char*)p);
        }
        return stacks[i];
    }
};
thread_local Worker tl_worker;
thread_local AzEmuBlock* tl_block = nullptr;

void fiber_entry()
{
    AzEmuFiber* f = az_emu_cur;
    (*f->blk->body)();
    f->done = true;
    AzEmuBlock* b = f->blk;
    b->live--;
    AzEmuWarp& w = b->warps[f->linear >> 5];
    w.expected--;
    // threads that left early must not dead-lock the ones waiting at a barrier
    if (b->live > 0 && b->bar_arrived >= b->live) { b->bar_arrived = 0; b->bar_gen++; }
    int p = w.gen & 1;
    if (w.expected > 0 && w.arrived[p] >= w.expected) { w.arrived[p] = 0; w.gen++; }
    swapcontext(&f->ctx, &b->sched);
}

void run_block(AzEmuBlock& b)
{
    tl_block = &b;
    int n = b.nthreads;
    b.live = n; b.bar_arrived = 0; b.bar_gen = 0;
    b.fibers.resize(n);
    b.warps.assign((n + 31) / 32, AzEmuWarp());
    for (size_t w = 0; w < b.warps.size(); w++) {
        memset(&b.warps[w], 0, sizeof(AzEmuWarp));
        b.warps[w].expected = std::min(32, n - 32 * (int)w);
    }
    for (int t = 0; t < n; t++) {
        AzEmuFiber& f = b.fibers[t];
        f.linear = t; f.done = false; f.blk = &b;
        f.tid.x = t % b.bdim.x; f.tid.y = (t / b.bdim.x) % b.bdim.y; f.tid.z = t / (b.bdim.x * b.bdim.y);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = tl_worker.stack(t);
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, fiber_entry, 0);
    }
    long guard = 0;
    while (b.live > 0) {
        for (int t = 0; t < n; t++) {
            AzEmuFiber& f = b.fibers[t];
            if (f.done) continue;
            az_emu_cur = &f;
            swapcontext(&b.sched, &f.ctx);
        }
        if (++guard > 50000000L) { fprintf(stderr, "cuda_emu: block (%u) appears dead-locked\n", b.bid.x); abort(); }
    }
    az_emu_cur = nullptr;
}
} // namespace

void az_emu_yield()
{
    AzEmuFiber* f = az_emu_cur;
    swapcontext(&f->ctx, &f->blk->sched);
    az_emu_cur = f;
}

void az_emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body)
{
    long nblocks = (long)grid.x * grid.y * grid.z;
    if (nblocks <= 0) return;
    int nthreads = (int)(block.x * block.y * block.z);
    static int nworkers = [] {
        const char* e = getenv("AZ_EMU_THREADS");
        int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        return std::max(1, std::min(n, 64));
    }();
    std::atomic<long> next(0);
    auto work = [&]() {
        AzEmuBlock b;
        b.bdim = block; b.gdim = grid; b.nthreads = nthreads; b.body = &body;
        std::vector<char> dyn(smem + 64);
        b.dyn_smem = (void*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
        for (;;) {
            long i = next.fetch_add(1);
            if (i >= nblocks) break;
            b.bid.x = (unsigned)(i % grid.x); b.bid.y = (unsigned)((i / grid.x) % grid.y); b.bid.z = (unsigned)(i / ((long)grid.x * grid.y));
            run_block(b);
        }
    };
    int nw = (int)std::min<long>(nworkers, nblocks);
    if (nw <= 1) { work(); return; }
    std::vector<std::thread> th;
    for (int k = 0; k < nw; k++) th.emplace_back(work);
    for (auto& t : th) t.join();
}
