// TEST INFRASTRUCTURE ONLY -- a minimal host emulation of the CUDA execution model, enough to compile
// tamm_b200/csrc/tamm_kernels.cu (kernels AND the C ABI that launches them) with g++ and run it without a GPU:
//
//     g++ -std=c++20 -pthread -DTAMM_EMULATE -include tests/cuda_emu/cuda_emu.h -x c++ tamm_b200/csrc/tamm_kernels.cu ...
//
// One OS thread per CUDA thread, CTAs of a grid one after the other:
//   threadIdx / blockIdx / blockDim / gridDim   thread-local values
//   __shared__                                  `static` (one CTA is resident at a time); dynamic shared memory per CTA
//   __syncthreads / __syncwarp / __shfl_*_sync  std::barrier per CTA / per warp (+ an exchange buffer per warp)
//   mbarrier + cp.async.bulk (TMA)              a 64-bit word {pending bytes, phase}; the copy is a memcpy by the issuer
//   <<<grid, block, dyn, stream>>>              cuda_emu::launch (synchronous; streams are ignored)
//   cudaMemsetAsync, cudaFuncSetAttribute, ...  trivial host versions
// What this checks that the arithmetic-only simulator (tests/hostsim) cannot: launch geometry, lane / profile / chunk
// indexing, the staging protocol of the layer tiles, the scan, the reductions, every global-memory address the kernels
// form (the arrays are ordinary host allocations: an out-of-range access is caught by the guard pages the tests put around
// them or by AddressSanitizer; a shared-memory access that no barrier orders is a data race between two OS threads, which
// ThreadSanitizer reports) -- the roles compute-sanitizer's memcheck and racecheck would play on the device.  What it cannot
// check: timing, the weaker memory model of a real GPU, and nvcc's own code generation.
// Nothing in the product (tamm_b200/, lib/) includes or loads this.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <tuple>
#include <vector>

#ifndef TAMM_EMULATE
#define TAMM_EMULATE 1
#endif

// ---- qualifiers -------------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types, dim3 ----------------------------------------------------------------------------------------------
struct alignas(16) double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- runtime API ------------------------------------------------------------------------------------------------------
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr int cudaFuncAttributeMaxDynamicSharedMemorySize = 8;
struct cudaDeviceProp { char name[64]; int major, minor, multiProcessorCount; };
inline const char* cudaGetErrorString(cudaError_t) { return "cuda_emu: no error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    std::snprintf(p->name, sizeof(p->name), "cuda_emu (host threads)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

namespace cuda_emu {

constexpr size_t MAX_DYN_SMEM = 232448;   // 227 KB, the sm_100 limit per CTA: a launch asking for more is refused like on the device

struct Warp {
    std::barrier<> bar;
    alignas(16) unsigned char buf[2][32][16];  // two sets of one exchange slot per lane (up to 16 bytes), used alternately
    explicit Warp(int n) : bar(n) {}
};
struct Cta {
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    unsigned char* dyn = nullptr;
    explicit Cta(int n) : bar(n) {}
};
struct Ctx {
    dim3 tid, bid, bdim, gdim;
    Cta* cta = nullptr;
    Warp* warp = nullptr;
    int lane = 0;
    unsigned shfl_count = 0;                   // shuffles this thread has executed (warp-uniform): selects the slot set
    int sync_count = 0;                        // __syncthreads calls of this thread in this launch (fault injection below)
};
// Fault injection for the negative control of the race check (tests/test_cuda_emu.py): with CUDA_EMU_SKIP_SYNCTHREADS=k every
// thread skips its k-th __syncthreads of every launch -- consistently, so nothing deadlocks, but the accesses that barrier
// ordered now race, which ThreadSanitizer must report.
inline int skip_sync() {
    static const int k = [] { const char* e = std::getenv("CUDA_EMU_SKIP_SYNCTHREADS"); return e ? std::atoi(e) : 0; }();
    return k;
}
inline thread_local Ctx ctx;

inline void* dyn_smem() { return ctx.cta->dyn; }

// Worker threads are kept from CTA to CTA and from launch to launch (creating 384 threads per CTA costs more than running a
// small CTA): worker i sleeps on its own generation word and is woken only when a CTA has a thread i.
class Pool {
    struct Slot {
        std::atomic<unsigned> go{0};
        std::thread th;
    };
    std::vector<std::unique_ptr<Slot>> slots;
    std::atomic<int> remaining{0};
    const void* task = nullptr;
    void (*call)(const void*, int) = nullptr;
    std::atomic<bool> stop{false};

    void worker(Slot* s, int id) {
        unsigned seen = 0;
        for (;;) {
            s->go.wait(seen, std::memory_order_acquire);
            seen = s->go.load(std::memory_order_acquire);
            if (stop.load(std::memory_order_acquire)) return;
            call(task, id);
            if (remaining.fetch_sub(1, std::memory_order_acq_rel) == 1) remaining.notify_all();
        }
    }

public:
    template <class F>
    void run(int nt, const F& f) {          // f(t) for t in [0, nt), each on its own thread, all alive at the same time
        while ((int)slots.size() < nt) {
            slots.emplace_back(new Slot);
            Slot* s = slots.back().get();
            const int id = (int)slots.size() - 1;
            s->th = std::thread([this, s, id] { worker(s, id); });
        }
        task = &f;
        call = [](const void* p, int t) { (*static_cast<const F*>(p))(t); };
        remaining.store(nt, std::memory_order_release);
        for (int t = 0; t < nt; ++t) {
            slots[t]->go.fetch_add(1, std::memory_order_release);
            slots[t]->go.notify_one();
        }
        for (int r; (r = remaining.load(std::memory_order_acquire)) != 0;) remaining.wait(r, std::memory_order_acquire);
    }
    ~Pool() {
        stop.store(true, std::memory_order_release);
        for (auto& s : slots) {
            s->go.fetch_add(1, std::memory_order_release);
            s->go.notify_one();
            s->th.join();
        }
    }
};
inline Pool& pool() {
    static Pool* p = new Pool;      // never destroyed: the workers may outlive static destruction order at process exit
    return *p;
}

template <class... KArgs>
struct Launch {
    void (*kernel)(KArgs...);
    dim3 grid, block;
    size_t dyn;
    template <class... A>
    void operator()(A&&... a) const {
        std::tuple<KArgs...> args(static_cast<KArgs>(a)...);          // converted once, like a kernel's parameter buffer
        const int nt = (int)(block.x * block.y * block.z);
        if (nt < 1 || nt > 1024 || nt % 32 != 0 || dyn > MAX_DYN_SMEM || grid.y > 65535u || grid.z > 65535u) {
            std::fprintf(stderr, "cuda_emu: invalid launch configuration (block %d threads, %zu bytes of dynamic shared memory, grid %u x %u x %u)\n",
                         nt, dyn, grid.x, grid.y, grid.z);
            std::abort();
        }
        std::unique_ptr<unsigned char[]> mem(new unsigned char[dyn + 256]);
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    Cta cta(nt);
                    for (int w = 0; w < nt / 32; ++w) cta.warps.emplace_back(new Warp(32));
                    // poisoned dynamic shared memory for every CTA (the device's is uninitialised too)
                    std::memset(mem.get(), 0xA5, dyn + 256);
                    cta.dyn = mem.get() + (128 - (reinterpret_cast<uintptr_t>(mem.get()) & 127)) % 128;
                    pool().run(nt, [&](int t) {
                        Ctx& c = ctx;
                        c.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        c.bid = dim3(bx, by, bz);
                        c.bdim = block; c.gdim = grid;
                        c.cta = &cta; c.warp = cta.warps[t / 32].get(); c.lane = t % 32; c.shfl_count = 0; c.sync_count = 0;
                        std::apply(kernel, args);
                        // a thread that has returned no longer takes part in barriers (CUDA semantics)
                        c.warp->bar.arrive_and_drop();
                        cta.bar.arrive_and_drop();
                    });
                }
    }
};
template <class... KArgs>
inline Launch<KArgs...> launch(dim3 g, dim3 b, size_t dyn, void* /*stream*/, void (*k)(KArgs...)) { return Launch<KArgs...>{k, g, b, dyn}; }

template <class T>
inline T shfl(T v, int src) {
    static_assert(sizeof(T) <= 16, "shuffle payload");
    // ONE barrier per shuffle: consecutive shuffles alternate between two slot sets, and a lane can only write set s again
    // (two shuffles later) after every lane has passed the barrier of the shuffle in between, i.e. has finished reading s
    Warp& w = *ctx.warp;
    auto& slots = w.buf[ctx.shfl_count++ & 1u];
    std::memcpy(slots[ctx.lane], &v, sizeof(T));
    w.bar.arrive_and_wait();
    T r;
    std::memcpy(&r, slots[(src >= 0 && src < 32) ? src : ctx.lane], sizeof(T));
    return r;
}

}  // namespace cuda_emu

#define threadIdx (cuda_emu::ctx.tid)
#define blockIdx (cuda_emu::ctx.bid)
#define blockDim (cuda_emu::ctx.bdim)
#define gridDim (cuda_emu::ctx.gdim)
#define TAMM_KERNEL(grid, block, dyn, stream, ...) cuda_emu::launch((grid), (block), (dyn), (stream), &__VA_ARGS__)
#define TAMM_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(cuda_emu::dyn_smem())

// ---- synchronisation, shuffles, cache-hinted accesses, atomics ----------------------------------------------------------------
inline void __syncthreads() {
    cuda_emu::Ctx& c = cuda_emu::ctx;
    if (++c.sync_count == cuda_emu::skip_sync()) return;
    c.cta->bar.arrive_and_wait();
}
inline void __syncwarp(unsigned = 0xffffffffu) { cuda_emu::ctx.warp->bar.arrive_and_wait(); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return cuda_emu::shfl(v, cuda_emu::ctx.lane ^ m); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) { return cuda_emu::shfl(v, cuda_emu::ctx.lane >= d ? cuda_emu::ctx.lane - d : cuda_emu::ctx.lane); }
template <class T> inline T __shfl_sync(unsigned, T v, int s) { return cuda_emu::shfl(v, s & 31); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
using std::max;
using std::min;

// ---- mbarrier + bulk copy: the barrier word holds {pending transaction bytes (high 32 bits), phase counter (low 32 bits)} ------
// Usage in the kernels: one elected thread arms the barrier with the byte count of a tile (arrive.expect_tx, arrival count 1) and
// issues the copies; everybody waits on the phase parity.
inline void mbar_init(uint64_t* bar, uint32_t /*count*/) { std::atomic_ref<uint64_t>(*bar).store(0, std::memory_order_release); }
inline void mbar_fence_init() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    std::atomic_ref<uint64_t>(*bar).fetch_add((uint64_t)bytes << 32, std::memory_order_acq_rel);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    std::atomic_ref<uint64_t> a(*bar);
    while (((uint32_t)a.load(std::memory_order_acquire) & 1u) == (parity & 1u)) std::this_thread::yield();
}
inline void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    // cp.async.bulk needs 16-byte aligned addresses and a size that is a multiple of 16
    if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15u) {
        std::fprintf(stderr, "cuda_emu: misaligned bulk copy (dst %p, src %p, %u bytes)\n", dst, src, bytes);
        std::abort();
    }
    std::memcpy(dst, src, bytes);
    std::atomic_ref<uint64_t> a(*bar);
    const uint64_t now = a.fetch_sub((uint64_t)bytes << 32, std::memory_order_acq_rel) - ((uint64_t)bytes << 32);
    if ((now >> 32) == 0) a.fetch_add(1, std::memory_order_acq_rel);      // all bytes of this phase have landed: next phase
}
inline void multimem_st(uint4* p, uint4 v) { *p = v; }
inline void multimem_st(uint2* p, uint2 v) { *p = v; }
