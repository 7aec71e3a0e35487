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
