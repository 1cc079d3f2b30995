// tests/emu/simt_emu.cpp -- fiber scheduler behind simt_emu.h.  TEST INFRASTRUCTURE ONLY.
#include "simt_emu.h"
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <unordered_map>

namespace emu {
ThreadCtx *cur = nullptr;
uint8_t *dyn_smem = nullptr;
int lane_order_desc = 0;

namespace {
enum Hint { H_WARP, H_BLOCK, H_DONE };
struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false;
    ThreadCtx tc;
    int wkind = 0;      // 0 running, 1 warp collective, 2 block barrier
    unsigned wmask = 0; // mask or barrier id
    unsigned long wgen = 0;
};
// one rendezvous state per (warp, participation mask): lanes named by a mask all execute the
// same sequence of collectives for that mask, so their generation counters stay in step.
struct MaskState {
    uint64_t val[2][32];
    unsigned arrived[2] = {0, 0};
    unsigned readers[2] = {0, 0};
    unsigned gen[32] = {0};
};
struct WarpState {
    std::unordered_map<unsigned, MaskState> by_mask;
};
struct Barrier {
    int count = 0;
    unsigned gen = 0;
    int orv = 0;
    int or_result[2] = {0, 0};
};
std::vector<Fiber> fibers;
std::vector<WarpState> warps;
Barrier barriers[17];
ucontext_t main_ctx;
int cur_idx = 0;
Hint hint = H_WARP;
unsigned long progress = 0;
const std::function<void()> *body_fn = nullptr;
const size_t kStack = 256 * 1024;

void fiber_entry() {
    (*body_fn)();
    fibers[cur_idx].done = true;
    hint = H_DONE;
    progress++;
    swapcontext(&fibers[cur_idx].ctx, &main_ctx);
}
void yield(Hint h) {
    hint = h;
    swapcontext(&fibers[cur_idx].ctx, &main_ctx);
}
}  // namespace

uint64_t warp_exchange(unsigned mask, uint64_t v, uint64_t *all, unsigned *arrived_mask) {
    int tid = cur_idx;
    int lane = tid & 31;
    MaskState &w = warps[tid >> 5].by_mask[mask];
    unsigned g = w.gen[lane]++;
    int b = g & 1;
    unsigned bit = 1u << lane;
    if (!(mask & bit)) { fprintf(stderr, "emu: lane %d not in its own mask %08x\n", lane, mask); abort(); }
    w.val[b][lane] = v;
    w.arrived[b] |= bit;
    fibers[tid].wkind = 1; fibers[tid].wmask = mask; fibers[tid].wgen = g;
    unsigned long spins = 0;
    while ((w.arrived[b] & mask) != mask) {
        yield(H_WARP);
        if (++spins > 100000000ul) { fprintf(stderr, "emu: warp collective deadlock (tid %d mask %08x arrived %08x)\n", tid, mask, w.arrived[b]); abort(); }
    }
    fibers[tid].wkind = 0;
    for (int i = 0; i < 32; i++) all[i] = w.val[b][i];
    *arrived_mask = w.arrived[b];
    w.readers[b] |= bit;
    if ((w.readers[b] & mask) == mask) { w.arrived[b] &= ~mask; w.readers[b] &= ~mask; progress++; }
    return v;
}

void block_barrier(int id, int nthreads) {
    Barrier &B = barriers[id];
    unsigned my = B.gen;
    B.count++;
    if (B.count == nthreads) { B.count = 0; B.gen++; progress++; return; }
    unsigned long spins = 0;
    fibers[cur_idx].wkind = 2; fibers[cur_idx].wmask = (unsigned)id; fibers[cur_idx].wgen = my;
    while (B.gen == my) {
        yield(H_BLOCK);
        if (++spins > 100000000ul) { fprintf(stderr, "emu: block barrier %d deadlock (count %d of %d)\n", id, B.count, nthreads); abort(); }
    }
    fibers[cur_idx].wkind = 0;
}
int block_barrier_or(int pred) {
    Barrier &B = barriers[16];
    unsigned my = B.gen;
    B.orv |= pred ? 1 : 0;
    B.count++;
    int n = (int)fibers.size();
    if (B.count == n) { B.or_result[my & 1] = B.orv; B.orv = 0; B.count = 0; B.gen++; progress++; return B.or_result[my & 1]; }
    fibers[cur_idx].wkind = 2; fibers[cur_idx].wmask = 16; fibers[cur_idx].wgen = my;
    while (B.gen == my) yield(H_BLOCK);
    fibers[cur_idx].wkind = 0;
    return B.or_result[my & 1];
}

void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body) {
    body_fn = &body;
    uint8_t *smem = (uint8_t *)aligned_alloc(1024, ((smem_bytes + 1023) / 1024 + 1) * 1024);
    for (unsigned b = 0; b < grid; b++) {
        memset(smem, 0xA5, smem_bytes);  // poison: shared memory is NOT zero-initialised on the device
        dyn_smem = smem;
        fibers.assign(block, Fiber());
        warps.assign((block + 31) / 32, WarpState());
        for (auto &B : barriers) B = Barrier();
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = fibers[t];
            f.stack = (char *)malloc(kStack);
            f.tc.tid = {t, 0, 0}; f.tc.bid = {b, 0, 0}; f.tc.bdim = {block, 1, 1}; f.tc.gdim = {grid, 1, 1};
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = kStack; f.ctx.uc_link = &main_ctx;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        unsigned remaining = block;
        int idx = lane_order_desc ? 31 < (int)block ? 31 : (int)block - 1 : 0;
        unsigned long last_progress = progress, idle = 0;
        while (remaining) {
            // run fiber idx
            if (!fibers[idx].done) {
                cur_idx = idx; cur = &fibers[idx].tc;
                swapcontext(&main_ctx, &fibers[idx].ctx);
                if (fibers[idx].done) remaining--;
            }
            if (!remaining) break;
            // choose next
            int n = (int)block;
            if (hint == H_WARP) {
                int wbase = idx & ~31, lane = idx & 31, next = -1;
                for (int k = 1; k <= 32; k++) {
                    int l = lane_order_desc ? (lane - k) & 31 : (lane + k) & 31;
                    int c = wbase + l;
                    if (c < n && !fibers[c].done) { next = c; break; }
                }
                if (next < 0) hint = H_BLOCK; else idx = next;
            }
            if (hint != H_WARP) {
                int next = -1;
                for (int k = 1; k <= n; k++) { int c = (idx + k) % n; if (!fibers[c].done) { next = c; break; } }
                if (next < 0) break;
                idx = next;
            }
            if (progress == last_progress) {
                if (++idle > 2000000ul) {
                    fprintf(stderr, "emu: no progress (deadlock?)\n");
                    for (int t = 0; t < (int)block; t++) {
                        Fiber &f = fibers[t];
                        if ((t & 31) == 0 || f.wkind != fibers[t - 1].wkind || f.wmask != fibers[t - 1].wmask || f.wgen != fibers[t-1].wgen || f.done != fibers[t-1].done)
                            fprintf(stderr, "  tid %d: done=%d wait=%d mask/id=%08x gen=%lu\n", t, (int)f.done, f.wkind, f.wmask, f.wgen);
                    }
                    abort();
                }
            } else { last_progress = progress; idle = 0; }
        }
        for (auto &f : fibers) free(f.stack);
    }
    free(smem);
    fibers.clear();
    cur = nullptr;
}
}  // namespace emu
