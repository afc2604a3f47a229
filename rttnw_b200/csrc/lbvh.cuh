// lbvh.cuh — device-side BVH construction (replaces BvhTree::new, hittable.rs:260-321,
// whose topology is not part of the contract — only the closest-hit answers are).
//
// The host builder (flatten.cpp, binned SAH) gives the better trees and is the default; this one is
// for scenes whose primitive count makes a host build the bottleneck: Morton codes of the box
// centroids, one radix sort (cub), Karras' parallel hierarchy ("Maximizing parallelism in the
// construction of BVHs, octrees and k-d trees", HPG 2012), a bottom-up fit with one atomic per
// node, and an emit pass that writes the traversal layout of device_types.h (both child boxes in
// the parent, 64-byte nodes). Leaves hold one record each and reference the records in their
// original order, so nothing else of the scene moves.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>

#include "device_types.h"

namespace rtx {

struct LbvhScratch {
    uint32_t *keys, *keys_alt;      // Morton codes
    int32_t *vals, *vals_alt;       // item index
    int32_t *left, *right;          // internal node i: children (>= 0 internal, < 0: ~sorted leaf position)
    int32_t *parent_of_internal;    // [n - 1]
    int32_t *parent_of_leaf;        // [n]
    unsigned int* visits;           // [n - 1] arrival counter of the fit pass
    float* ibox;                    // [6 (n - 1)] boxes of the internal nodes
    int32_t* depth;                 // [n - 1] inner nodes on the longest path below (and including) the node
    void* cub_temp;
    size_t cub_temp_bytes;
};

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {  // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// boxes: 6 floats per item (lo.xyz, hi.xyz). Centroids are normalised to [lo, lo + extent].
__global__ void lbvh_morton_kernel(int n, const float* __restrict__ boxes, float3 lo, float3 inv_extent, uint32_t* __restrict__ keys,
                                   int32_t* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = boxes + 6 * (size_t)i;
    float cx = (0.5f * (b[0] + b[3]) - lo.x) * inv_extent.x;
    float cy = (0.5f * (b[1] + b[4]) - lo.y) * inv_extent.y;
    float cz = (0.5f * (b[2] + b[5]) - lo.z) * inv_extent.z;
    uint32_t x = (uint32_t)fminf(fmaxf(cx * 1024.f, 0.f), 1023.f);
    uint32_t y = (uint32_t)fminf(fmaxf(cy * 1024.f, 0.f), 1023.f);
    uint32_t z = (uint32_t)fminf(fmaxf(cz * 1024.f, 0.f), 1023.f);
    keys[i] = (expand_bits10(x) << 2) | (expand_bits10(y) << 1) | expand_bits10(z);
    vals[i] = i;
}

// length of the common prefix of the keys at sorted positions i and j (ties broken by position), -1 outside
__device__ __forceinline__ int lbvh_delta(const uint32_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = keys[i], b = keys[j];
    if (a == b) return 32 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clz(a ^ b);
}

__global__ void lbvh_hierarchy_kernel(int n, const uint32_t* __restrict__ keys, int32_t* __restrict__ left, int32_t* __restrict__ right,
                                      int32_t* __restrict__ parent_of_internal, int32_t* __restrict__ parent_of_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    // direction of the range, its other end (Karras 2012, Fig. 4)
    int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    // split position
    int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int32_t lc, rc;
    if (lo == gamma) { lc = ~gamma; parent_of_leaf[gamma] = i; } else { lc = gamma; parent_of_internal[gamma] = i; }
    if (hi == gamma + 1) { rc = ~(gamma + 1); parent_of_leaf[gamma + 1] = i; } else { rc = gamma + 1; parent_of_internal[gamma + 1] = i; }
    left[i] = lc;
    right[i] = rc;
    if (i == 0) parent_of_internal[0] = -1;
}

// One thread per leaf climbs towards the root; the second thread to arrive at a node owns it.
__global__ void lbvh_fit_kernel(int n, const float* __restrict__ boxes, const int32_t* __restrict__ vals, const int32_t* __restrict__ left,
                                const int32_t* __restrict__ right, const int32_t* __restrict__ parent_of_internal,
                                const int32_t* __restrict__ parent_of_leaf, unsigned int* visits, float* ibox, int32_t* depth) {
    int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int node = parent_of_leaf[leaf];
    while (node >= 0) {
        if (atomicAdd(&visits[node], 1u) == 0u) return;  // the sibling subtree is not done yet
        __threadfence();
        float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        int dmax = 0;
        const int32_t kids[2] = {left[node], right[node]};
        for (int k = 0; k < 2; ++k) {
            const volatile float* b;
            if (kids[k] < 0) {
                b = boxes + 6 * (size_t)vals[~kids[k]];
            } else {
                b = ibox + 6 * (size_t)kids[k];
                dmax = max(dmax, ((volatile int32_t*)depth)[kids[k]]);
            }
            for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], b[a]); hi[a] = fmaxf(hi[a], b[3 + a]); }
        }
        float* o = ibox + 6 * (size_t)node;
        for (int a = 0; a < 3; ++a) { o[a] = lo[a]; o[3 + a] = hi[a]; }
        depth[node] = dmax + 1;
        __threadfence();
        node = parent_of_internal[node];
    }
}

// Traversal layout: node i holds the boxes of its two children. Leaf codes reference the record
// first_record + item, one record per leaf.
__global__ void lbvh_emit_kernel(int n, const float* __restrict__ boxes, const int32_t* __restrict__ vals, const int32_t* __restrict__ left,
                                 const int32_t* __restrict__ right, const float* __restrict__ ibox, int32_t node_base, int32_t first_record,
                                 BvhNode* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    BvhNode nd;
    const int32_t kids[2] = {left[i], right[i]};
    int32_t code[2];
    float b[2][6];
    for (int k = 0; k < 2; ++k) {
        const float* src;
        if (kids[k] < 0) {
            int item = vals[~kids[k]];
            src = boxes + 6 * (size_t)item;
            code[k] = ~(((first_record + item) << 4) | 1);
        } else {
            src = ibox + 6 * (size_t)kids[k];
            code[k] = node_base + kids[k];
        }
        for (int a = 0; a < 6; ++a) b[k][a] = src[a];
    }
    nd.c0x[0] = b[0][0]; nd.c0x[1] = b[0][3]; nd.c0y[0] = b[0][1]; nd.c0y[1] = b[0][4]; nd.c0z[0] = b[0][2]; nd.c0z[1] = b[0][5];
    nd.c1x[0] = b[1][0]; nd.c1x[1] = b[1][3]; nd.c1y[0] = b[1][1]; nd.c1y[1] = b[1][4]; nd.c1z[0] = b[1][2]; nd.c1z[1] = b[1][5];
    nd.child0 = code[0];
    nd.child1 = code[1];
    nd._pad[0] = nd._pad[1] = 0;
    out[i] = nd;
}

// Builds the n - 1 nodes of a BVH over n >= 2 boxes into d_nodes (node indices node_base ..); returns
// the depth of the tree (inner nodes on the longest root-to-leaf path) through *depth_out.
// d_boxes: 6 n floats on the device; lo / extent: bounds of the box centroids.
inline cudaError_t lbvh_build(cudaStream_t stream, int n, const float* d_boxes, const float lo[3], const float extent[3], int32_t first_record,
                              int32_t node_base, BvhNode* d_nodes, int* depth_out) {
    LbvhScratch s{};
    const size_t un = (size_t)n;
    cudaError_t e = cudaSuccess;
    uint8_t* block = nullptr;
    // one allocation for everything but the cub temp
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_keys = take(un * 4), o_keys2 = take(un * 4), o_vals = take(un * 4), o_vals2 = take(un * 4), o_left = take(un * 4),
           o_right = take(un * 4), o_pi = take(un * 4), o_pl = take(un * 4), o_visits = take(un * 4), o_ibox = take(un * 24),
           o_depth = take(un * 4);
    cub::DeviceRadixSort::SortPairs(nullptr, s.cub_temp_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, n, 0, 30, stream);
    size_t o_cub = take(s.cub_temp_bytes);
    if ((e = cudaMalloc(&block, off)) != cudaSuccess) return e;
    s.keys = (uint32_t*)(block + o_keys); s.keys_alt = (uint32_t*)(block + o_keys2);
    s.vals = (int32_t*)(block + o_vals); s.vals_alt = (int32_t*)(block + o_vals2);
    s.left = (int32_t*)(block + o_left); s.right = (int32_t*)(block + o_right);
    s.parent_of_internal = (int32_t*)(block + o_pi); s.parent_of_leaf = (int32_t*)(block + o_pl);
    s.visits = (unsigned int*)(block + o_visits); s.ibox = (float*)(block + o_ibox); s.depth = (int32_t*)(block + o_depth);
    s.cub_temp = block + o_cub;
    const int tb = 256;
    const int gn = (n + tb - 1) / tb;
    float3 flo = make_float3(lo[0], lo[1], lo[2]);
    float3 inv = make_float3(extent[0] > 0 ? 1.f / extent[0] : 0.f, extent[1] > 0 ? 1.f / extent[1] : 0.f, extent[2] > 0 ? 1.f / extent[2] : 0.f);
    lbvh_morton_kernel<<<gn, tb, 0, stream>>>(n, d_boxes, flo, inv, s.keys, s.vals);
    e = cub::DeviceRadixSort::SortPairs(s.cub_temp, s.cub_temp_bytes, s.keys, s.keys_alt, s.vals, s.vals_alt, n, 0, 30, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s.visits, 0, un * 4, stream);
    if (e == cudaSuccess) {
        lbvh_hierarchy_kernel<<<gn, tb, 0, stream>>>(n, s.keys_alt, s.left, s.right, s.parent_of_internal, s.parent_of_leaf);
        lbvh_fit_kernel<<<gn, tb, 0, stream>>>(n, d_boxes, s.vals_alt, s.left, s.right, s.parent_of_internal, s.parent_of_leaf, s.visits, s.ibox,
                                               s.depth);
        lbvh_emit_kernel<<<gn, tb, 0, stream>>>(n, d_boxes, s.vals_alt, s.left, s.right, s.ibox, node_base, first_record, d_nodes);
        e = cudaGetLastError();
    }
    int32_t depth = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&depth, s.depth, sizeof(depth), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(block);
    if (depth_out) *depth_out = depth;
    return e;
}

}  // namespace rtx
