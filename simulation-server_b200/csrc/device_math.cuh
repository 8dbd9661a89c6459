// float3 helpers with the numeric contract of the reference's utilities/math.cuh:11-104
// (IEEE sqrt/div, normalize() maps NaN to the zero vector; the device printf of the reference is dropped).
#pragma once
#include <cuda_runtime.h>

namespace bcs {

__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 xyz(const float4& v) { return make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 operator*(float a, float3 v) { return make_float3(a * v.x, a * v.y, a * v.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator/(float3 v, float a) { return make_float3(v.x / a, v.y / a, v.z / a); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross(float3 u, float3 v)
{
    return make_float3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
__device__ __forceinline__ float length_squared(float3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
__device__ __forceinline__ float length(float3 v) { return sqrtf(length_squared(v)); }
__device__ __forceinline__ float3 normalize(float3 v)
{
    float3 vn = v / sqrtf(dot(v, v));
    if (isnan(vn.x) || isnan(vn.y) || isnan(vn.z)) return make_float3(0.f, 0.f, 0.f);
    return vn;
}

// calculateIdForCell (grids/uniform_grid.cu:24-36) with the `max`/`min` macros of :20-21 written out:
//   max(0, q)    -> (0 > q ? 0 : q)        min(len, m) -> (len > m ? m : len)
// The division must stay a true IEEE division (cell sizes such as 25 are not powers of two).
// exact quotient (p - mn) / cs: a power-of-two cell size (the reference's 2-unit particle grid) divides exactly as a
// product with its reciprocal, anything else (e.g. the 25-unit triangle grid) needs the IEEE division
__device__ __forceinline__ float cell_quotient(float d, int cs)
{
    return (cs & (cs - 1)) == 0 ? d * (1.0f / (float)cs) : __fdiv_rn(d, (float)cs);
}

__device__ __forceinline__ int axis_cell(float p, float mn, float len, int cs)
{
    float q = cell_quotient(p - mn, cs);
    float m = (0.f > q) ? 0.f : q;
    float r = (len > m) ? m : len;
    return (int)r;
}

// the unclamped per-axis index the collision kernels use to trim the 27-cell stencil
// (particle_collisions.cuh:117-119)
__device__ __forceinline__ int axis_cell_raw(float p, float mn, int cs) { return (int)cell_quotient(p - mn, cs); }

// Philox4x32-10 (counter-based respawn RNG; replaces the cuRAND XORWOW states the reference shares
// between the threads of a cell, vein_end.cu:103-105)
__device__ __forceinline__ void philox4x32_10(unsigned c[4], unsigned k0, unsigned k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        unsigned n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// (0,1] like curand_uniform
__device__ __forceinline__ float u01(unsigned x) { return (float)x * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }

// number of items of the rank's enumeration: owned blood cells x maxP (padding included) + ghosts
__device__ __forceinline__ int item_total(const ActiveItems& A) { return A.cellPrefix[A.types->n] * A.maxP + *A.ghostCount; }

// item -> particle id (or -1 for padding / past the end); flag: 1 owned, 2 ghost
__device__ __forceinline__ int active_item(const ActiveItems A, int item, int& flag)
{
    const TypesDev* ty = A.types;
    const int nTypes = ty->n;
    const int ownedItems = A.cellPrefix[nTypes] * A.maxP;
    if (item < ownedItems) {
        const int w = item / A.maxP, k = item - w * A.maxP;
        int t = 0;
        while (t + 1 < nTypes && w >= A.cellPrefix[t + 1]) ++t;
        const int P = ty->t[t].P;
        if (k >= P) return -1;
        const int c = A.cells[ty->t[t].cStart + (w - A.cellPrefix[t])];
        flag = 1;
        return ty->t[t].pStart + (c - ty->t[t].cStart) * P + k;
    }
    const int gi = item - ownedItems;
    if (gi < *A.ghostCount) {
        flag = 2;
        return A.ghostList[gi];
    }
    return -1;
}

}  // namespace bcs
