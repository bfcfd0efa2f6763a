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

// ------------------------------------------------------------------------------------------------ NCCL stand-in
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
namespace {
struct Group {
    int world = 0; std::mutex mu; std::condition_variable cv; int waiting = 0; unsigned long gen = 0;
    std::vector<const void*> pub; std::vector<size_t> pub_n;                       // per rank: published buffer of the running collective
    std::map<std::pair<int, int>, std::vector<std::vector<uint8_t>>> mail;         // (src, dst) -> queued messages
    void barrier() { std::unique_lock<std::mutex> l(mu); unsigned long g = gen; if (++waiting == world) { waiting = 0; gen++; cv.notify_all(); } else cv.wait(l, [&] { return gen != g; }); }
};
struct Comm { Group* g; int rank; };
std::mutex g_reg_mu; std::map<std::string, Group*> g_groups; unsigned long g_uid_counter = 0;
size_t dsize(int dtype) { return dtype == 3 ? 4 : dtype == 5 ? 8 : 1; }           // ncclUint32 = 3, ncclUint64 = 5 (the two the library uses)
struct PendingOp { bool send; const void* sbuf; void* rbuf; size_t bytes; int peer; Comm* c; };
thread_local std::vector<PendingOp> t_ops; thread_local int t_group_depth = 0;
int flush_ops() {
    if (t_ops.empty()) return 0;
    Group* g = t_ops[0].c->g; int me = t_ops[0].c->rank;
    { std::lock_guard<std::mutex> l(g->mu); for (auto& o : t_ops) if (o.send) g->mail[{me, o.peer}].emplace_back((const uint8_t*)o.sbuf, (const uint8_t*)o.sbuf + o.bytes); }
    for (auto& o : t_ops) if (!o.send) {        // blocking receive: wait until the peer has posted
        std::unique_lock<std::mutex> l(g->mu);
        for (;;) { auto it = g->mail.find({o.peer, me}); if (it != g->mail.end() && !it->second.empty()) { auto msg = std::move(it->second.front()); it->second.erase(it->second.begin()); l.unlock(); if (msg.size() != o.bytes) return 5; memcpy(o.rbuf, msg.data(), o.bytes); break; } l.unlock(); std::this_thread::yield(); l.lock(); }
    }
    t_ops.clear();
    return 0;
}
}  // namespace
int emu_ncclGetUniqueId(void* uid128) { std::lock_guard<std::mutex> l(g_reg_mu); memset(uid128, 0, 128); unsigned long v = ++g_uid_counter; memcpy(uid128, "EMUNCCL", 8); memcpy((char*)uid128 + 8, &v, sizeof v); return 0; }
int emu_ncclCommInitRank(void** comm, int world, const void* uid128, int rank) {
    Group* g;
    { std::lock_guard<std::mutex> l(g_reg_mu); std::string key((const char*)uid128, 128); auto it = g_groups.find(key); if (it == g_groups.end()) { g = new Group(); g->world = world; g->pub.assign(world, nullptr); g->pub_n.assign(world, 0); g_groups[key] = g; } else g = it->second; }
    if (g->world != world || rank < 0 || rank >= world) return 4;
    *comm = new Comm{g, rank};
    g->barrier();                    // ncclCommInitRank is collective
    return 0;
}
int emu_ncclCommDestroy(void* comm) { delete (Comm*)comm; return 0; }
int emu_ncclAllGather(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t) {
    Comm* c = (Comm*)comm; Group* g = c->g; size_t nb = count * dsize(dtype);
    g->pub[c->rank] = send; g->barrier();
    std::vector<uint8_t> tmp(nb * g->world); for (int r = 0; r < g->world; r++) memcpy(tmp.data() + r * nb, g->pub[r], nb);
    g->barrier(); memcpy(recv, tmp.data(), tmp.size()); g->barrier();
    return 0;
}
int emu_ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int, void* comm, cudaStream_t) {
    Comm* c = (Comm*)comm; Group* g = c->g;
    g->pub[c->rank] = send; g->barrier();
    std::vector<uint8_t> tmp(count * dsize(dtype), 0);
    for (int r = 0; r < g->world; r++) { if (dtype == 3) { const uint32_t* s = (const uint32_t*)g->pub[r]; uint32_t* d = (uint32_t*)tmp.data(); for (size_t i = 0; i < count; i++) d[i] += s[i]; } else { const uint64_t* s = (const uint64_t*)g->pub[r]; uint64_t* d = (uint64_t*)tmp.data(); for (size_t i = 0; i < count; i++) d[i] += s[i]; } }
    g->barrier(); memcpy(recv, tmp.data(), tmp.size()); g->barrier();
    return 0;
}
int emu_ncclSend(const void* send, size_t count, int dtype, int peer, void* comm, cudaStream_t) { t_ops.push_back({true, send, nullptr, count * dsize(dtype), peer, (Comm*)comm}); return t_group_depth ? 0 : flush_ops(); }
int emu_ncclRecv(void* recv, size_t count, int dtype, int peer, void* comm, cudaStream_t) { t_ops.push_back({false, nullptr, recv, count * dsize(dtype), peer, (Comm*)comm}); return t_group_depth ? 0 : flush_ops(); }
int emu_ncclGroupStart() { t_group_depth++; return 0; }
int emu_ncclGroupEnd() { if (--t_group_depth == 0) return flush_ops(); return 0; }
const char* emu_ncclGetErrorString(int e) { return e ? "emulated NCCL error" : "no error"; }
