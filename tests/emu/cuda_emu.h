// cuda_emu.h -- test infrastructure: run a CUDA kernel body as plain C++ on the CPU, one OS thread per CUDA thread of a
// CTA (CTAs run one after another), __syncthreads = std::barrier.  Lets the CPU-only test tier execute the real kernel
// source (index arithmetic, shared-memory ring, barriers), not a restatement of it.
#pragma once
#include <barrier>
#include <thread>
#include <vector>
#include <mutex>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <memory>

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
inline std::barrier<> *emu_barrier = nullptr;
inline std::mutex emu_mutex;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__
inline void __syncthreads() { emu_barrier->arrive_and_wait(); }
template <class T> inline T __ldg(const T *p) { return *p; }
using std::min; using std::max;

// block-wide max then one "atomic" update: under emulation every thread just takes the lock
inline void block_atomic_max(double v, double *out) { { std::lock_guard<std::mutex> g(emu_mutex); if (v > *out) *out = v; } __syncthreads(); }

// kernels without __syncthreads: the threads of a CTA can simply run one after another on the calling thread
template <class Kernel, class Args>
void emu_launch_seq(Kernel k, dim3 grid, dim3 block, const Args &a)
{
    blockDim = block; gridDim = grid;
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned tz = 0; tz < block.z; ++tz) for (unsigned ty = 0; ty < block.y; ++ty) for (unsigned tx = 0; tx < block.x; ++tx) {
            threadIdx = dim3(tx, ty, tz); blockIdx = dim3(bx, by, bz);
            k(a);
        }
}

template <class Kernel, class Args>
void emu_launch(Kernel k, dim3 grid, int nthreads, const Args &a)
{
    blockDim = dim3(nthreads); gridDim = grid;
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(nthreads);
        emu_barrier = &bar;
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() { threadIdx = dim3(t); blockIdx = dim3(bx, by, bz); k(a); });
        for (auto &x : th) x.join();
    }
}

// ---- 2-D thread blocks with dynamic shared memory and warp shuffles (the plane-marching Godunov kernels) ----
// A warp = 32 consecutive linear thread ids; a shuffle is a rendezvous of the warp's threads on a per-warp barrier.
inline void *emu_smem = nullptr;
inline std::vector<std::barrier<> *> emu_warp_bar;
inline std::vector<double> emu_shfl_buf;
inline thread_local int emu_tid = 0;
inline double emu_shfl(double v, int delta)
{
    const int lane = emu_tid & 31, base = emu_tid - lane, src = lane + delta;
    emu_shfl_buf[emu_tid] = v;
    emu_warp_bar[emu_tid >> 5]->arrive_and_wait();
    const double r = (src < 0 || src > 31) ? v : emu_shfl_buf[base + src];
    emu_warp_bar[emu_tid >> 5]->arrive_and_wait();
    return r;
}
inline double __shfl_up_sync(unsigned, double v, int d) { return emu_shfl(v, -d); }
inline double __shfl_down_sync(unsigned, double v, int d) { return emu_shfl(v, d); }

template <class Kernel, class Args>
void emu_launch2(Kernel k, dim3 grid, dim3 block, size_t smem_bytes, const Args &a)
{
    const int nthreads = block.x * block.y * block.z;
    blockDim = block; gridDim = grid;
    std::vector<char> sm(smem_bytes + 64);
    emu_smem = sm.data();
    emu_shfl_buf.assign(nthreads, 0.0);
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(nthreads);
        emu_barrier = &bar;
        std::vector<std::unique_ptr<std::barrier<>>> wb;
        emu_warp_bar.clear();
        for (int w = 0; w < (nthreads + 31) / 32; ++w) { wb.emplace_back(new std::barrier<>(std::min(32, nthreads - 32 * w))); emu_warp_bar.push_back(wb.back().get()); }
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                emu_tid = t;
                threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y)); blockIdx = dim3(bx, by, bz);
                k(a);
            });
        for (auto &x : th) x.join();
    }
}
