// Ray ordering for the wavefront (RTX_ORDER=1): the trace kernel takes the rays of a pass in the order of WHERE THEY START
// and WHICH WAY THEY GO instead of slot order.
//
// Why. The trace kernel is issue bound at 8 of 32 lanes: the rays of a warp leave the node loop after very different trip
// counts. tools/time_trace.py / tools/time_trace_bins.py measured the fixed-ray kernel on launch-sized sets (256 Ki) of
// third-generation rays in different orders: sorted by (direction octant, record the ray starts on) it runs 36 % faster on
// scene 9, 20 % on the Cornell box, 28 % on scene 1; with 4096 bins and random order inside a bin still 24-27 %. Records are
// laid out in BVH leaf order (flatten.cpp), so the record index IS a space-filling position, with no scale to choose.
//
// How (one counting sort per pass, spread over kernels that exist anyway):
//   shade kernel  every slot that holds a ray for the next trace pass computes key = octant * groups + (record >> shift)
//                 (camera rays: one bin of their own, in arrival order, which keeps a tile's rays together) and takes
//                 rank = atomicAdd(hist[key], 1). The LAST CTA of the launch to finish turns hist into exclusive offsets
//                 (and clears it for the next pass).
//   order kernel  order[offsets[key] + rank] = slot                                        (4 bytes per ray, one tiny launch)
//   trace kernel  thread p traces the ray of slot order[p]; p >= offsets[n_bins] (= rays in flight) exits at once, so empty
//                 slots cost nothing.
// The path state never moves: only the trace kernel's accesses to the pool are permuted (7 gathered doubles, 3 scattered
// stores per ray, all in L2). Results do not depend on the order: every draw is keyed by (pixel, sample, bounce).
#pragma once
#include <stdint.h>

namespace rtx {

constexpr uint32_t kOrderDead = 0xffffffffu;

struct OrderArgs {
    uint32_t* hist;      // [n_bins_padded] zero between passes
    uint32_t* offsets;   // [n_bins_padded + 1] exclusive prefix sums of hist; offsets[n_bins_padded] = rays in flight
    uint2* key_rank;     // [n_slots] (key, rank inside the bin) of the slot's ray, key = kOrderDead: no ray
    uint32_t* order;     // [n_slots] slots in key order
    unsigned int* done;  // CTAs of the shade launch that have finished (back to 0 when the last one has scanned)
    int32_t n_bins_padded;  // 8 * groups + 1, rounded up to a multiple of 4 * the shade CTA
    int32_t groups, shift;  // record index >> shift < groups
};

__device__ __forceinline__ uint32_t order_key(const OrderArgs& o, int32_t src_rec, double dx, double dy, double dz) {
    if (src_rec < 0) return 8u * (uint32_t)o.groups;  // a camera ray
    const uint32_t oct = (dx > 0.0 ? 1u : 0u) | (dy > 0.0 ? 2u : 0u) | (dz > 0.0 ? 4u : 0u);
    uint32_t g = (uint32_t)src_rec >> o.shift;
    if (g >= (uint32_t)o.groups) g = (uint32_t)o.groups - 1u;
    return oct * (uint32_t)o.groups + g;
}

// Called by every thread of every CTA of the shade launch, after its atomic on hist has returned. The CTA that finishes
// last scans the histogram: thread t owns n_bins_padded / kThreads consecutive bins.
template <int kThreads>
__device__ __forceinline__ void order_scan_by_last_cta(const OrderArgs& o, int tid) {
    __shared__ uint32_t s_warp[kThreads / 32];
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(o.done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int per = o.n_bins_padded / kThreads;  // a multiple of 4
    const uint4* src = reinterpret_cast<const uint4*>(o.hist + tid * per);
    uint32_t sum = 0;
    for (int k = 0; k < per / 4; ++k) {
        const uint4 v = __ldcg(src + k);
        sum += v.x + v.y + v.z + v.w;
    }
    // exclusive scan of the per-thread sums over the CTA
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t incl = sum;
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = incl - sum;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    uint4* hist4 = reinterpret_cast<uint4*>(o.hist + tid * per);
    uint32_t* dst = o.offsets + tid * per;
    for (int k = 0; k < per / 4; ++k) {
        const uint4 v = __ldcg(src + k);
        dst[4 * k + 0] = base; base += v.x;
        dst[4 * k + 1] = base; base += v.y;
        dst[4 * k + 2] = base; base += v.z;
        dst[4 * k + 3] = base; base += v.w;
        hist4[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == kThreads - 1) o.offsets[o.n_bins_padded] = base;
    if (tid == 0) *o.done = 0u;
}

__global__ void __launch_bounds__(256) wf_order_kernel(OrderArgs o, int n_slots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const uint2 kr = o.key_rank[i];
    if (kr.x != kOrderDead) o.order[__ldg(o.offsets + kr.x) + kr.y] = (uint32_t)i;
}

}  // namespace rtx
