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
#include <cub/device/device_scan.cuh>

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


// ---------------------------------------------------------------------------
// PLOC — parallel locally-ordered clustering (Meister & Bittner, "Parallel Locally-Ordered Clustering for Bounding
// Volume Hierarchy Construction", TVCG 2018), the builder for scenes too large for the host and too incoherent for
// an LBVH. Bottom-up agglomeration along the Morton curve: the clusters (at first the leaves, in Morton order) each
// look r positions to the left and right for the neighbour whose union with them has the smallest surface area;
// pairs that chose each other merge into an inner node that takes the place of the left partner; the array is
// compacted in order; repeat until one cluster is left. The Morton curve only has to bring candidates close — the
// merges are decided by the surface areas of real boxes, which is where an LBVH (splits decided by code bits) loses.
//
// Ties: a cluster prefers its "partner" position i ^ 1, then the smaller position. Every pass then merges at least
// one pair (take the smallest union area d of the pass: if some partner pair has it, both prefer each other; if none
// has, the smallest position a on any d-edge and the smallest b among a's d-neighbours choose each other), and
// identical boxes (all unions equal) pair up 0-1, 2-3, ... instead of merging once per pass.
// Node numbers are handed out downwards from n - 2, so that the last merge — the root — is node 0 like Karras' root.
// ---------------------------------------------------------------------------
constexpr int kPlocBlock = 256;
constexpr int kPlocMaxRadius = 32;

__global__ void ploc_init_kernel(int n, const float* __restrict__ boxes, const int32_t* __restrict__ vals, int32_t* __restrict__ cid,
                                 float* __restrict__ cbox) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = boxes + 6 * (size_t)vals[i];
    float* o = cbox + 6 * (size_t)i;
    for (int a = 0; a < 6; ++a) o[a] = b[a];
    cid[i] = ~i;
}

// nn[i] = the position in [i - r, i + r] whose box, united with box i, has the smallest area (-1 when m == 1).
// `pair_up`: skip the search and name the partner position (the guard against a pass count that runs away).
__global__ void __launch_bounds__(kPlocBlock) ploc_nn_kernel(int m, int r, bool pair_up, const float* __restrict__ cbox, int32_t* __restrict__ nn) {
    __shared__ float s_box[(kPlocBlock + 2 * kPlocMaxRadius) * 6];
    const int first = (int)blockIdx.x * kPlocBlock - r;  // position of s_box[0]
    const int span = kPlocBlock + 2 * r;
    for (int k = threadIdx.x; k < span * 6; k += kPlocBlock) {
        const int pos = first + k / 6;
        s_box[k] = (pos >= 0 && pos < m) ? cbox[6 * (size_t)pos + (k % 6)] : 0.f;
    }
    __syncthreads();
    const int i = blockIdx.x * kPlocBlock + threadIdx.x;
    if (i >= m) return;
    if (pair_up) {
        nn[i] = (i ^ 1) < m ? (i ^ 1) : -1;
        return;
    }
    const float* me = s_box + (threadIdx.x + r) * 6;
    const float lx = me[0], ly = me[1], lz = me[2], hx = me[3], hy = me[4], hz = me[5];
    float best = 3.4e38f;
    int best_j = -1;
    const int j0 = max(i - r, 0), j1 = min(i + r, m - 1);
    for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        const float* o = s_box + (j - first) * 6;
        const float dx = fmaxf(hx, o[3]) - fminf(lx, o[0]), dy = fmaxf(hy, o[4]) - fminf(ly, o[1]), dz = fmaxf(hz, o[5]) - fminf(lz, o[2]);
        const float area = dx * dy + dy * dz + dz * dx;
        if (area < best || (area == best && j == (i ^ 1))) { best = area; best_j = j; }
    }
    nn[i] = best_j;
}

// Mutual pairs merge: the left partner makes the node and keeps the place, the right partner's place is dropped.
__global__ void ploc_merge_kernel(int m, int n, int32_t* __restrict__ cid, float* __restrict__ cbox, const int32_t* __restrict__ nn,
                                  unsigned int* __restrict__ n_made, int32_t* __restrict__ left, int32_t* __restrict__ right,
                                  float* __restrict__ ibox, int32_t* __restrict__ depth, uint32_t* __restrict__ keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = nn[i];
    uint32_t k = 1u;
    if (j >= 0 && nn[j] == i) {
        if (i < j) {
            const int32_t node = n - 2 - (int32_t)atomicAdd(n_made, 1u);
            const int32_t a = cid[i], b = cid[j];
            float* bi = cbox + 6 * (size_t)i;
            const float* bj = cbox + 6 * (size_t)j;
            float u[6];
            for (int x = 0; x < 3; ++x) { u[x] = fminf(bi[x], bj[x]); u[3 + x] = fmaxf(bi[3 + x], bj[3 + x]); }
            float* o = ibox + 6 * (size_t)node;
            for (int x = 0; x < 6; ++x) { o[x] = u[x]; bi[x] = u[x]; }
            left[node] = a;
            right[node] = b;
            depth[node] = 1 + max(a >= 0 ? depth[a] : 0, b >= 0 ? depth[b] : 0);
            cid[i] = node;
        } else {
            k = 0u;
        }
    }
    keep[i] = k;
}

__global__ void ploc_compact_kernel(int m, const int32_t* __restrict__ cid, const float* __restrict__ cbox, const uint32_t* __restrict__ keep,
                                    const uint32_t* __restrict__ pos, int32_t* __restrict__ cid_out, float* __restrict__ cbox_out,
                                    int32_t* __restrict__ m_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (keep[i]) {
        const uint32_t p = pos[i];
        cid_out[p] = cid[i];
        const float* b = cbox + 6 * (size_t)i;
        float* o = cbox_out + 6 * (size_t)p;
        for (int a = 0; a < 6; ++a) o[a] = b[a];
    }
    if (i == m - 1) *m_out = (int32_t)(pos[i] + keep[i]);
}

// Same contract as lbvh_build. radius: neighbours searched on each side (1 .. kPlocMaxRadius); passes_out: how many
// passes the clustering took.
inline cudaError_t ploc_build(cudaStream_t stream, int n, const float* d_boxes, const float lo[3], const float extent[3], int32_t first_record,
                              int32_t node_base, BvhNode* d_nodes, int* depth_out, int radius, int* passes_out) {
    const size_t un = (size_t)n;
    cudaError_t e = cudaSuccess;
    uint8_t* block = nullptr;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_keys = take(un * 4), o_keys2 = take(un * 4), o_vals = take(un * 4), o_vals2 = take(un * 4), o_left = take(un * 4),
                 o_right = take(un * 4), o_ibox = take(un * 24), o_depth = take(un * 4), o_cid = take(un * 4), o_cid2 = take(un * 4),
                 o_cbox = take(un * 24), o_cbox2 = take(un * 24), o_nn = take(un * 4), o_keep = take(un * 4), o_pos = take(un * 4),
                 o_scalars = take(64);
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, n, 0,
                                    30, stream);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, stream);
    const size_t o_cub = take(sort_bytes > scan_bytes ? sort_bytes : scan_bytes);
    if ((e = cudaMalloc(&block, off)) != cudaSuccess) return e;
    uint32_t *keys = (uint32_t*)(block + o_keys), *keys_alt = (uint32_t*)(block + o_keys2);
    int32_t *vals = (int32_t*)(block + o_vals), *vals_alt = (int32_t*)(block + o_vals2);
    int32_t *left = (int32_t*)(block + o_left), *right = (int32_t*)(block + o_right), *depth = (int32_t*)(block + o_depth);
    float* ibox = (float*)(block + o_ibox);
    int32_t* cid[2] = {(int32_t*)(block + o_cid), (int32_t*)(block + o_cid2)};
    float* cbox[2] = {(float*)(block + o_cbox), (float*)(block + o_cbox2)};
    int32_t* nn = (int32_t*)(block + o_nn);
    uint32_t *keep = (uint32_t*)(block + o_keep), *pos = (uint32_t*)(block + o_pos);
    unsigned int* n_made = (unsigned int*)(block + o_scalars);
    int32_t* d_m = (int32_t*)(block + o_scalars + 16);
    void* cub_temp = block + o_cub;
    size_t cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    const int tb = 256;
    const int gn = (n + tb - 1) / tb;
    const int r = radius < 1 ? 1 : (radius > kPlocMaxRadius ? kPlocMaxRadius : radius);
    float3 flo = make_float3(lo[0], lo[1], lo[2]);
    float3 inv = make_float3(extent[0] > 0 ? 1.f / extent[0] : 0.f, extent[1] > 0 ? 1.f / extent[1] : 0.f, extent[2] > 0 ? 1.f / extent[2] : 0.f);
    lbvh_morton_kernel<<<gn, tb, 0, stream>>>(n, d_boxes, flo, inv, keys, vals);
    e = cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, keys, keys_alt, vals, vals_alt, n, 0, 30, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(block + o_scalars, 0, 64, stream);
    int passes = 0;
    if (e == cudaSuccess) {
        ploc_init_kernel<<<gn, tb, 0, stream>>>(n, d_boxes, vals_alt, cid[0], cbox[0]);
        int m = n, cur = 0;
        // ~40 % of the clusters merge per pass on ordinary input; past this many passes the rest is paired up blindly
        const int patience = 96;
        while (m > 1 && e == cudaSuccess) {
            const int gm = (m + kPlocBlock - 1) / kPlocBlock;
            ploc_nn_kernel<<<gm, kPlocBlock, 0, stream>>>(m, r, passes >= patience, cbox[cur], nn);
            ploc_merge_kernel<<<gm, kPlocBlock, 0, stream>>>(m, n, cid[cur], cbox[cur], nn, n_made, left, right, ibox, depth, keep);
            cub_bytes = scan_bytes;
            e = cub::DeviceScan::ExclusiveSum(cub_temp, cub_bytes, keep, pos, m, stream);
            if (e != cudaSuccess) break;
            ploc_compact_kernel<<<gm, kPlocBlock, 0, stream>>>(m, cid[cur], cbox[cur], keep, pos, cid[cur ^ 1], cbox[cur ^ 1], d_m);
            int32_t m_new = 0;
            e = cudaMemcpyAsync(&m_new, d_m, sizeof(m_new), cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) break;
            if (m_new >= m || m_new < 1) { e = cudaErrorUnknown; break; }  // (cannot happen: every pass merges a pair)
            m = m_new;
            cur ^= 1;
            ++passes;
        }
        if (e == cudaSuccess) {
            lbvh_emit_kernel<<<gn, tb, 0, stream>>>(n, d_boxes, vals_alt, left, right, ibox, node_base, first_record, d_nodes);
            e = cudaGetLastError();
        }
    }
    int32_t d = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&d, depth, sizeof(d), cudaMemcpyDeviceToHost, stream);  // node 0 is the root
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(block);
    if (depth_out) *depth_out = d;
    if (passes_out) *passes_out = passes;
    return e;
}

}  // namespace rtx
