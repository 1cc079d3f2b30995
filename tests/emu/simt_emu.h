// tests/emu/simt_emu.h -- a minimal single-threaded SIMT emulator (fibers) so the
// CUDA kernel sources under compress_b200/csrc can be compiled with g++ and
// exercised on the GPU-less dev box.
//
// TEST INFRASTRUCTURE ONLY.  It is never built into libb200comp.so, never
// loaded by compress_b200/, and is not a CPU fallback: the product fails loudly
// without a GPU.  It exists because kernels cannot be run in the dev container;
// every CUDA thread becomes a ucontext fiber, warp collectives and barriers
// are rendezvous points, and lanes are run in either ascending or descending
// order between rendezvous so order-dependent (racy) code shows up as a diff.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <functional>

#define B2C_EMU 1

namespace emu {
struct Dim3 { unsigned x, y, z; };
struct ThreadCtx { Dim3 tid; Dim3 bid; Dim3 bdim; Dim3 gdim; };
extern ThreadCtx *cur;
extern uint8_t *dyn_smem;
extern int lane_order_desc;  // 0: lanes scheduled ascending, 1: descending

void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body);

// rendezvous primitives
uint64_t warp_exchange(unsigned mask, uint64_t v, uint64_t *all /*[32]*/, unsigned *arrived_mask);
void block_barrier(int id, int nthreads);
int block_barrier_or(int pred);
}  // namespace emu

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::cur->bid)
#define blockDim (emu::cur->bdim)
#define gridDim (emu::cur->gdim)

static inline unsigned emu_lane() { return emu::cur->tid.x & 31; }

static inline void __syncthreads() { emu::block_barrier(0, (int)emu::cur->bdim.x); }
static inline int __syncthreads_or(int p) { return emu::block_barrier_or(p); }
static inline void emu_named_barrier(int id, int nthreads) { emu::block_barrier(id, nthreads); }

static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    uint64_t all[32]; unsigned am;
    emu::warp_exchange(mask, 0, all, &am);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    uint64_t all[32]; unsigned am;
    emu::warp_exchange(mask, pred ? 1 : 0, all, &am);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (((mask >> i) & 1) && all[i]) r |= 1u << i;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline unsigned __activemask() { return 0xffffffffu; }

template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shfl");
    uint64_t all[32]; unsigned am; uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu::warp_exchange(mask, raw, all, &am);
    int lane = (int)emu_lane();
    int base = lane & ~(width - 1);
    int s = base + (src & (width - 1));
    T out; memcpy(&out, &all[s], sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    uint64_t all[32]; unsigned am; uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu::warp_exchange(mask, raw, all, &am);
    int lane = (int)emu_lane();
    int base = lane & ~(width - 1);
    int s = lane - (int)delta;
    if (s < base) s = lane;
    T out; memcpy(&out, &all[s], sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    uint64_t all[32]; unsigned am; uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu::warp_exchange(mask, raw, all, &am);
    int lane = (int)emu_lane();
    int base = lane & ~(width - 1);
    int s = lane + (int)delta;
    if (s >= base + width) s = lane;
    T out; memcpy(&out, &all[s], sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lm, int width = 32) {
    uint64_t all[32]; unsigned am; uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu::warp_exchange(mask, raw, all, &am);
    int lane = (int)emu_lane();
    int s = lane ^ lm;
    (void)width;
    T out; memcpy(&out, &all[s], sizeof(T));
    return out;
}
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v) {
    uint64_t all[32]; unsigned am; uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu::warp_exchange(mask, raw, all, &am);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (((mask >> i) & 1) && all[i] == raw) r |= 1u << i;
    return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    uint64_t all[32]; unsigned am;
    emu::warp_exchange(mask, v, all, &am);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((mask >> i) & 1) r += (unsigned)all[i];
    return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    uint64_t all[32]; unsigned am;
    emu::warp_exchange(mask, v, all, &am);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (((mask >> i) & 1) && (unsigned)all[i] > r) r = (unsigned)all[i];
    return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    uint64_t all[32]; unsigned am;
    emu::warp_exchange(mask, v, all, &am);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; i++) if (((mask >> i) & 1) && (unsigned)all[i] < r) r = (unsigned)all[i];
    return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    uint64_t all[32]; unsigned am;
    emu::warp_exchange(mask, v, all, &am);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((mask >> i) & 1) r |= (unsigned)all[i];
    return r;
}

// scalar intrinsics
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((v >> i) & 1) r |= 1u << (31 - i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (sh & 31)) >> 32);
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 7;
        r |= (unsigned)((v >> (8 * sel)) & 0xff) << (8 * i);
    }
    return r;
}
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
template <typename T> static inline T __ldg(const T *p) { return *p; }
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
