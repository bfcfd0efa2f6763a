// cuda_shim.cpp -- TEST INFRASTRUCTURE ONLY: the fiber scheduler behind cuda_shim.hpp (see there).
#include "cuda_shim.hpp"
#include <stdio.h>
#include <sys/mman.h>

namespace emu {

thread_local BlockCtx* g_blk = nullptr;
static constexpr size_t STACK = 256 << 10;
static thread_local std::vector<char*> g_stacks;       // reused across launches

void yield() { BlockCtx& B = *g_blk; swapcontext(&B.fibers[B.cur].ctx, &B.sched); }

static void fiber_main() {
    BlockCtx& B = *g_blk;
    B.body();
    Fiber& f = B.fibers[B.cur];
    f.done = true; B.alive--;
    B.warps[f.tid >> 5].exited |= 1u << (f.tid & 31);
    if (B.sync_arrived && B.sync_arrived >= B.alive) { B.sync_arrived = 0; B.sync_gen++; }      // the others were waiting for this thread at a barrier
    swapcontext(&f.ctx, &B.sched);
}

uint64_t collective(unsigned mask, uint64_t v, int op, int arg) {
    BlockCtx& B = *g_blk; const unsigned tid = B.fibers[B.cur].tid, lane = tid & 31, bit = 1u << lane;
    Warp& w = B.warps[tid >> 5];
    const unsigned in_warp = std::min(32u, B.block.x - (tid & ~31u)), valid = in_warp == 32 ? 0xFFFFFFFFu : ((1u << in_warp) - 1u);
    while (w.arrived & bit) yield();                       // the previous rendezvous is still being read by slower lanes
    w.slot[lane] = v; w.arrived |= bit;
    // lanes of the mask that are expected: all of it except lanes that left the kernel WITHOUT taking part (a lane that took
    // part and then finished still counts: its arrived bit stays until everybody has read the slots)
    unsigned need = mask & valid & ~(w.exited & ~w.arrived);
    while ((w.arrived & need) != need) { yield(); need = mask & valid & ~(w.exited & ~w.arrived); }
    const unsigned part = w.arrived;                       // the participants of this rendezvous
    uint64_t r = 0;
    switch (op) {
    case 0: for (unsigned l = 0; l < 32; l++) if ((part >> l) & 1) r |= (uint64_t)(w.slot[l] & 1) << l; break;
    case 1: { unsigned s = (unsigned)arg & 31; r = ((part >> s) & 1) ? w.slot[s] : v; break; }
    case 2: { unsigned s = lane ^ ((unsigned)arg & 31); r = ((part >> s) & 1) ? w.slot[s] : v; break; }
    case 3: { int s = (int)lane - arg; r = (s >= 0 && ((part >> s) & 1)) ? w.slot[s] : v; break; }
    case 4: for (unsigned l = 0; l < 32; l++) if ((part >> l) & 1) r |= w.slot[l]; break;
    case 5: r = ~0ull; for (unsigned l = 0; l < 32; l++) if ((part >> l) & 1) r = std::min(r, w.slot[l]); break;
    case 6: r = 1; for (unsigned l = 0; l < 32; l++) if (((part >> l) & 1) && !(w.slot[l] & 1)) r = 0; break;
    }
    w.departed |= bit;
    if (w.departed == w.arrived) { w.arrived = 0; w.departed = 0; }
    return r;
}

void syncthreads() {
    BlockCtx& B = *g_blk; const unsigned gen = B.sync_gen;
    if (++B.sync_arrived >= B.alive) { B.sync_arrived = 0; B.sync_gen++; return; }
    while (B.sync_gen == gen) yield();
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    if (g_blk) { fprintf(stderr, "cuda_shim: nested launch\n"); abort(); }
    const unsigned nt = block.x, nb = grid.x;
    if (!nt || !nb) return;
    BlockCtx B; B.grid = grid; B.block = block; B.body = body;
    std::vector<uint8_t> dyn(smem + 64, 0xAB);            // dynamic shared memory is not zero-initialised on the device either
    B.dyn = dyn.data();
    while (g_stacks.size() < nt) { void* p = mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0); if (p == MAP_FAILED) { perror("cuda_shim: mmap"); abort(); } g_stacks.push_back((char*)p); }
    B.fibers.resize(nt);
    g_blk = &B;
    for (unsigned bx = 0; bx < nb; bx++) {
        B.bx = bx; B.warps.assign((nt + 31) / 32, Warp{}); B.sync_arrived = 0; B.sync_gen = 0; B.alive = nt;
        for (unsigned t = 0; t < nt; t++) {
            Fiber& f = B.fibers[t]; f.done = false; f.tid = t; f.stack = g_stacks[t];
            getcontext(&f.ctx); f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = &B.sched;
            makecontext(&f.ctx, (void (*)())fiber_main, 0);
        }
        unsigned long spins = 0;
        while (B.alive) {
            for (unsigned t = 0; t < nt; t++) if (!B.fibers[t].done) { B.cur = t; swapcontext(&B.sched, &B.fibers[t].ctx); }
            if (++spins > 200000000ul) { fprintf(stderr, "cuda_shim: block %u does not finish (deadlocked rendezvous?)\n", bx); abort(); }
        }
    }
    g_blk = nullptr;
}

}  // namespace emu
