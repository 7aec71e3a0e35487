// vdn_common.cuh -- shared device/host definitions for the B200 VARDEN hot path (sm_100a, FP64).
#pragma once
#ifndef VDN_EMU
#include <cuda_runtime.h>
#endif
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>
#include "../../include/vdn.h"

// ---- BC codes (FBoxLib bc_module; define_bc_tower.f90:158-340) ----
enum : int {
    BC_PERIODIC = -1, BC_INTERIOR = 0, BC_INLET = 11, BC_OUTLET = 12, BC_SYMMETRY = 13,
    BC_SLIP_WALL = 14, BC_NO_SLIP_WALL = 15,
    BC_REFLECT_ODD = 20, BC_REFLECT_EVEN = 21, BC_FOEXTRAP = 22, BC_EXT_DIR = 23, BC_HOEXTRAP = 24
};
enum : int { ELL_PER = -1, ELL_INT = 0, ELL_DIR = 1, ELL_NEU = 2 };

struct VdnError : std::runtime_error { using std::runtime_error::runtime_error; };

#define VDN_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    char b_[512]; snprintf(b_, sizeof b_, "%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    throw VdnError(b_); } } while (0)
#define VDN_REQUIRE(cond, msg) do { if (!(cond)) { char b_[512]; snprintf(b_, sizeof b_, "%s:%d %s", __FILE__, __LINE__, msg); throw VdnError(b_); } } while (0)

// A strided view of one field over the rank's region.  p points at LOCAL cell/face (0,0,0), comp 0,
// so ghost cells are reached with negative indices.  Local index = global index - region_lo.
struct View {
    double *p;
    // 32-bit strides and offsets: the stage kernels are issue-bound and 64-bit index arithmetic costs 3-4 instructions per
    // multiply; alloc_field refuses fields of 2^31 elements (16 GB) or more
    int sy, sz, cs;
    __host__ __device__ __forceinline__ double &operator()(int i, int j, int k, int c = 0) const {
        return p[i + sy * j + sz * k + cs * c];
    }
    __host__ __device__ __forceinline__ View comp(int c) const { View v = *this; v.p += (long)cs * (long)c; return v; }
    // stride along direction d
    __host__ __device__ __forceinline__ int st(int d) const { return d == 0 ? 1 : (d == 1 ? sy : sz); }
};

constexpr int VDN_MAXCUT = 16;

// Geometry + BC tables handed to kernels by value.
struct Geo {
    int dim;
    int n[3];               // region cells (n[2] == 1 in 2-D)
    int pbc[3][2];          // physical BC on the REGION faces: domain code, PERIODIC, or INTERIOR (rank boundary)
    double h[3];
    // reference boxes inside the region (tensor-product cuts, local coords) -> per-box eps
    int nb[3];
    int cut[3][VDN_MAXCUT + 1];
    __host__ __device__ __forceinline__ int box1(int d, int i) const {
        int c = i < 0 ? 0 : (i > n[d] - 1 ? n[d] - 1 : i);
        int b = 0;
        for (int q = 1; q < nb[d]; ++q) b += (c >= cut[d][q]);
        return b;
    }
    __host__ __device__ __forceinline__ int box(int i, int j, int k) const {
        if (nb[0] * nb[1] * nb[2] == 1) return 0;
        return box1(0, i) + nb[0] * (box1(1, j) + nb[1] * box1(2, k));
    }
};

// device field (host-side descriptor)
struct DField {
    double *base = nullptr;
    int ng = 0, nc = 0, fdir = -1;
    int ext[3] = {1, 1, 1};
    size_t bytes = 0;
    long sy = 0, sz = 0, cs = 0;
    int ngd[3] = {0, 0, 0};     // ghost width per direction (0 in the unused 3rd direction of 2-D)
    bool in_heap = false;       // storage lives in the symmetric heap of the peer-memory transport (not cudaFree'd on its own)
    View view() const {
        View v; v.sy = (int)sy; v.sz = (int)sz; v.cs = (int)cs;
        v.p = base + ngd[0] + sy * ngd[1] + sz * ngd[2];
        return v;
    }
    long ncell() const { return (long)ext[0] * ext[1] * ext[2]; }
};

#ifdef __CUDACC__
// block-wide max of non-negative values, then ONE atomicMax per block on the bit pattern (non-negative doubles order
// like their bit patterns).  Every thread of the block must call it (no early returns before it).
__device__ __forceinline__ void block_atomic_max(double v, double *out)
{
    __shared__ double sm_max_[32];
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int tid = threadIdx.z * blockDim.y * blockDim.x + threadIdx.y * blockDim.x + threadIdx.x;
    const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
    if ((tid & 31) == 0) sm_max_[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        v = tid < nw ? sm_max_[tid] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (tid == 0 && v > *(volatile double *)out)
            atomicMax((unsigned long long *)out, (unsigned long long)__double_as_longlong(v));
    }
}
#endif

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

struct Range { int lo[3], hi[3]; };   // inclusive local index range
static inline dim3 grid3(const Range &r, dim3 b)
{
    return dim3(cdiv(r.hi[0] - r.lo[0] + 1, b.x), cdiv(r.hi[1] - r.lo[1] + 1, b.y), cdiv(r.hi[2] - r.lo[2] + 1, b.z));
}
static inline Range mk_range(int l0, int h0, int l1, int h1, int l2, int h2)
{
    Range r; r.lo[0] = l0; r.hi[0] = h0; r.lo[1] = l1; r.hi[1] = h1; r.lo[2] = l2; r.hi[2] = h2; return r;
}
