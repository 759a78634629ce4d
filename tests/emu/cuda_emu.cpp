// Runtime of the test-only SIMT emulator (see cuda_emu.h).
#include "cuda_emu.h"

#include <omp.h>

namespace emu {

thread_local Block* g_block = nullptr;
thread_local uint3 g_threadIdx, g_blockIdx;
thread_local dim3 g_blockDim, g_gridDim;

static const size_t kStack = 256 * 1024;

void yield_to_scheduler(State s)
{
    Block* b = g_block;
    Fiber& f = b->fibers[b->cur];
    f.st = s;
    swapcontext(&f.ctx, &b->sched);
}

static void trampoline()
{
    Block* b = g_block;
    b->body();
    b->fibers[b->cur].st = DONE;
    swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

static void run_block(Block& b, dim3 block)
{
    const int n = (int)(block.x * block.y * block.z);
    g_block = &b;
    for (int t = 0; t < n; ++t) {
        Fiber& f = b.fibers[t];
        f.st = RUNNABLE;
        f.tid.x = t % block.x;
        f.tid.y = (t / block.x) % block.y;
        f.tid.z = t / (block.x * block.y);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &b.sched;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    const int nwarps = (n + 31) / 32;
    for (;;) {
        bool progressed = false;
        int done = 0;
        for (int t = 0; t < n; ++t) {
            Fiber& f = b.fibers[t];
            if (f.st == DONE) { ++done; continue; }
            if (f.st != RUNNABLE) continue;
            b.cur = t;
            g_threadIdx = f.tid;
            swapcontext(&b.sched, &f.ctx);
            progressed = true;
        }
        if (done == n) break;
        // release warps whose live threads all wait at a warp barrier
        for (int w = 0; w < nwarps; ++w) {
            int live = 0, at = 0;
            for (int t = w * 32; t < n && t < w * 32 + 32; ++t) {
                if (b.fibers[t].st == DONE) continue;
                ++live;
                if (b.fibers[t].st == AT_WARP) ++at;
            }
            if (live && at == live) {
                for (int t = w * 32; t < n && t < w * 32 + 32; ++t)
                    if (b.fibers[t].st == AT_WARP) b.fibers[t].st = RUNNABLE;
                progressed = true;
            }
        }
        // release the block barrier when every live thread waits at it
        int live = 0, at = 0;
        for (int t = 0; t < n; ++t) {
            if (b.fibers[t].st == DONE) continue;
            ++live;
            if (b.fibers[t].st == AT_BLOCK) ++at;
        }
        if (live && at == live) {
            for (int t = 0; t < n; ++t)
                if (b.fibers[t].st == AT_BLOCK) b.fibers[t].st = RUNNABLE;
            progressed = true;
        }
        if (!progressed) {
            int nb = 0, nw = 0;
            for (int t = 0; t < n; ++t) { nb += b.fibers[t].st == AT_BLOCK; nw += b.fibers[t].st == AT_WARP; }
            std::fprintf(stderr, "cuda_emu: DEADLOCK in block (%u,%u,%u): %d threads at __syncthreads, %d at a "
                                 "warp barrier, %d exited of %d -- divergent barrier\n",
                         g_blockIdx.x, g_blockIdx.y, g_blockIdx.z, nb, nw, done, n);
            std::abort();
        }
    }
    g_block = nullptr;
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body)
{
    const long nblocks = (long)grid.x * grid.y * grid.z;
    const int n = (int)(block.x * block.y * block.z);
#pragma omp parallel
    {
        Block b;
        b.fibers.resize(n);
        for (int t = 0; t < n; ++t) b.fibers[t].stack = (char*)std::malloc(kStack);
        b.dyn_smem = (char*)std::aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
        b.body = body;
#pragma omp for schedule(dynamic, 1)
        for (long i = 0; i < nblocks; ++i) {
            g_gridDim = grid;
            g_blockDim = block;
            g_blockIdx.x = (unsigned)(i % grid.x);
            g_blockIdx.y = (unsigned)((i / grid.x) % grid.y);
            g_blockIdx.z = (unsigned)(i / ((long)grid.x * grid.y));
            std::memset(b.dyn_smem, 0xAB, smem);  // poison: uninitialised shared memory shows up
            run_block(b, block);
        }
        for (int t = 0; t < n; ++t) std::free(b.fibers[t].stack);
        std::free(b.dyn_smem);
    }
}

}  // namespace emu
