// Fixed-point accumulator bins in shared memory (96 bits: three 32-bit limbs per accumulator), shared by the tile kernels.
#pragma once
#include "common.cuh"

namespace sfb {

// ---- fixed-point bins in shared memory ---------------------------------------------------------------------------
// b[0..2] += (x0, x1, x2) as one 96-bit integer: every limb add is a native 32-bit ATOMS.ADD that returns the old
// value, a wrap of a limb is carried into the next one by the thread that caused it -- exactly once per wrap, so
// the 96-bit total is exact whatever the interleaving of the threads.
__device__ __forceinline__ void fx_smem_add(uint32_t* b, Fx f) {
    uint32_t carry = 0;
    if (f.x0) {
        const uint32_t old = atomicAdd(b, f.x0);
        carry = (uint32_t)(old + f.x0 < old);
    }
    uint32_t t1 = f.x1 + carry;
    uint32_t c1 = (uint32_t)(t1 < carry);          // x1 = 2^32 - 1 and a carry: t1 wraps to 0
    if (t1) {
        const uint32_t old = atomicAdd(b + 1, t1);
        c1 |= (uint32_t)(old + t1 < old);
    }
    const uint32_t t2 = f.x2 + c1;
    if (t2) atomicAdd(b + 2, t2);
}
// one accumulator out to the global planes (or to the fp64 array in the plain mode)
__device__ __forceinline__ void fx_flush(double* p, i64 det_stride, const uint32_t* b) {
    const uint32_t x0 = b[0], x1 = b[1], x2 = b[2];
    if ((x0 | x1 | x2) == 0) return;
    if (det_stride) {
        Fx f;
        f.x0 = x0; f.x1 = x1; f.x2 = x2;
        fx_red(p, det_stride, f);
    } else {
        atomicAdd(p, (double)x2 + (double)x1 * 0x1p-32 + (double)x0 * 0x1p-64);
    }
}
// The node loop: cov[i] * weight into the bins b + 3 i of the walk's slots.  Records that cover their node (aligned ==
// node length) add the weight itself, split once; with weight 1 (`unit`) that is the integer 1.
__device__ __forceinline__ void bins_add_record(uint32_t* b, const sfb_reads& r, int v, bool unit, double weight, const Fx& whole) {
    if (v >= 0 && (v & 0xffff) == (v >> 16)) {
        if (unit) atomicAdd(b + 2, 1u); else fx_smem_add(b, whole);
        return;
    }
    const double cov = v < 0 ? r.exc_cov[-(i64)v - 1] : (double)(v & 0xffff) / (double)(v >> 16);
    fx_smem_add(b, fx_split(unit ? cov : cov * weight));
}
__device__ __forceinline__ void bins_add_records(uint32_t* b, const sfb_reads& r, i64 w0, int L, bool unit, double weight) {
    Fx whole;
    whole.x0 = whole.x1 = 0u;
    whole.x2 = 1u;
    if (!unit) whole = fx_split(weight);
    // the packed records four at a time: the loads of a batch are in flight together (behind the shared-memory atomics
    // of the previous record they would queue up one by one)
    for (int i0 = 0; i0 < L; i0 += 4) {
        int v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = i0 + u < L ? r.walk_alen[w0 + i0 + u] : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u < L) bins_add_record(b + 3 * (i0 + u), r, v[u], unit, weight, whole);
    }
}
}  // namespace sfb
