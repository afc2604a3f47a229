// trace3.cuh — K2b, "sorted" form: the closest surface hit of every ray in flight, with the rays of one SM regrouped
// by the KIND OF WORK they need next, every few steps, through shared memory.
//
// Why. In the first two forms a thread keeps one ray from its first node to its last leaf, so a warp's instruction
// count is set by its slowest lane in every phase: 8-9 of 32 lanes did useful work per instruction (ncu, profiles/),
// and the persistent voted form (trace2.cuh) recovered only what its bookkeeping cost. What a ray needs next is one of
// three things — fp32 node steps, the f64 tests of a leaf, or a new ray — and which one is known at the end of every
// step. So here the traversal state of a ray lives in SHARED memory (slab parameters, t_max, current node, stack: 120
// bytes), not in a thread's registers, and one CTA per SM owns as many rays as it has threads. Work proceeds in rounds:
// at the start of a round every ray is on exactly one of three lists (node / leaf / free), the lists are cut into
// chunks of 32, and each warp takes a chunk — all of its lanes doing the same kind of work by construction: leaf chunks
// run the f64 primitive tests (ray re-read from the pool), node chunks take up to `burst` node steps, free chunks fetch
// new rays from the CTA's share of the pool — and every lane then appends its ray to next round's list for whatever it
// needs next (one shared-memory atomic per warp and list). Leaf chunks (f64, global loads) and node chunks (fp32,
// shared memory) run side by side on different warps, so the two kinds of latency cover each other; two barriers
// separate the rounds. The world BVH is staged in shared memory as in trace2.cuh.
//
// What it measured (scene 9, one B200): 25 of 32 lanes active per instruction, as designed — and 238-342 M
// samples/s against 600 M. Moving a ray's 14 words of state in and out of shared memory, three list appends and the
// round's bookkeeping cost ~260 thread instructions per ray and round, as much as four node steps, so the warp
// instructions per ray only fell by 10 % (8.06 G against 8.94 G per 17 spp) while thread instructions tripled; and
// rounds are as long as their slowest chunk (an f64 leaf chunk behind two dependent global loads), during which most
// warps wait at the barrier: 0.72 instructions per cycle and SM against 2.3. Regrouping through shared memory pays
// only if a round does much more than four steps per ray, and then the lanes diverge again inside the round. Kept as
// a measured negative result, opt-in (RTX_TRACE=3).
//
// Replaces the recursion of BvhTree::hit (hittable.rs:355-368) over Bound::hit (bound.rs:13-32) and the List::hit
// scan of the leaves (hittable.rs:153-163); same answers as the other forms
// (tests/test_gpu_parity.py::test_other_kernel_forms_trace_the_same_rays).
#pragma once
#include "trace2.cuh"

namespace rtx {

struct Trace3Cfg {
    int32_t stage_first;  // first node of the world BVH in SceneView::nodes (its root)
    int32_t n_stage;      // nodes staged in shared memory: [stage_first, stage_first + n_stage)
    int32_t cap;          // capacity of the staged copy, in nodes (>= n_stage)
    int32_t burst;        // node steps a ray takes per round, at most
};

constexpr int kT3StateFloats = 10;  // idx idy idz anx any anz afx afy afz tmax
constexpr int kT3StateInts = 4;     // cur, top, depth, slot
__host__ __device__ constexpr size_t trace3_smem_bytes(int threads, int cap) {
    return (size_t)cap * kNodeStride + (size_t)threads * (4 * kT3StateFloats + 4 * kT3StateInts + 4 * (1 + kShortStack) + 2 * 6) + 64;
}

// Appends `id` to a list for every lane with `pred`: one shared atomic per warp.
__device__ __forceinline__ void t3_push(bool pred, uint16_t* list, int* count, int id, int lane, uint32_t lt) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) list[base + __popc(m & lt)] = (uint16_t)id;
}

template <bool kCount, int kThreads, bool kAllStaged>
__global__ void __launch_bounds__(kThreads, 1) wf_trace3_kernel(SceneView sc, PathPool pool, int n_slots, Trace3Cfg cfg,
                                                                 unsigned long long* ray_count, Counters* counters) {
    extern __shared__ __align__(16) unsigned char t3_smem[];
    constexpr int R = kThreads;  // rays in flight per CTA
    constexpr int kWarps = kThreads / 32;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    // layout: staged nodes | float state [10][R] | int state [4][R] | stack [1 + kShortStack][R] (row 0 = pad) |
    //         lists node[2][R], leaf[2][R], free[2][R] (uint16) | counters
    const uint32_t s_nodes = (uint32_t)__cvta_generic_to_shared(t3_smem);
    float* st_f = reinterpret_cast<float*>(t3_smem + (size_t)cfg.cap * kNodeStride);
    int32_t* st_i = reinterpret_cast<int32_t*>(st_f + kT3StateFloats * R);
    int32_t* stk = st_i + kT3StateInts * R;
    uint16_t* l_node = reinterpret_cast<uint16_t*>(stk + (1 + kShortStack) * R);
    uint16_t* l_leaf = l_node + 2 * R;
    uint16_t* l_free = l_leaf + 2 * R;
    int* ctr = reinterpret_cast<int*>(l_free + 2 * R);  // [0,1] nodes  [2,3] leaves  [4,5] free  [6] next piece of the share
    const uint32_t s_stk = (uint32_t)__cvta_generic_to_shared(stk);
    Tally<kCount> tally;

    // ---- stage the top of the world BVH, child references made relative to its first node ----
    {
        const uint4* g = reinterpret_cast<const uint4*>(sc.nodes + cfg.stage_first);
        const int n_chunks = 4 * cfg.n_stage;
        for (int k = tid; k < n_chunks; k += kThreads) {
            uint4 v = __ldg(g + k);
            const int c = k & 3, n = k >> 2;
            if (c == 3) {
                if ((int32_t)v.x >= 0) v.x -= (uint32_t)cfg.stage_first;
                if ((int32_t)v.y >= 0) v.y -= (uint32_t)cfg.stage_first;
            }
            *reinterpret_cast<uint4*>(t3_smem + (size_t)n * kNodeStride + (size_t)c * 16) = v;
        }
    }
    l_free[tid] = (uint16_t)tid;  // every state slot is free
    if (tid < 8) ctr[tid] = tid == 4 ? R : 0;
    const int G = (int)gridDim.x, bid = (int)blockIdx.x;
    const BvhNode* g_nodes = sc.nodes + cfg.stage_first;
    const int32_t root = sc.world_root - cfg.stage_first;
    const float tmin_f = __int_as_float(0x3a83126e);  // the largest float below 0.001 (main.rs:36's t_min, rounded down)
    TraceCfg ncfg{cfg.stage_first, cfg.n_stage, cfg.cap, 0, 0, 0};
    unsigned int my_rays = 0;
    int b = 0;  // which buffer of each list this round reads

    while (true) {
        __syncthreads();  // every append of the previous round has landed
        const int nN = ctr[b], nL = ctr[2 + b], nF = ctr[4 + b], base = ctr[6];
        // this CTA's share of the pool: every G-th piece of 32 slots; local index j -> slot ((j / 32) G + bid) 32 + j % 32
        const bool dry = ((((base >> 5) * G + bid) << 5)) >= n_slots;
        if (nN == 0 && nL == 0 && dry) break;
        const int nb = b ^ 1;
        if (tid == 0) {
            ctr[nb] = 0;
            ctr[2 + nb] = 0;
            ctr[4 + nb] = 0;
        }
        __syncthreads();  // next round's lists are empty; everybody has read this round's counts
        const int cL = (nL + 31) >> 5, cN = (nN + 31) >> 5, cF = dry ? 0 : (nF + 31) >> 5;
        for (int c = warp; c < cL + cN + cF; c += kWarps) {
            int id = 0, kind = 0;  // what the lane's ray needs next: 1 node steps, 2 a leaf, 3 nothing (finished / no ray)
            if (c < cL) {
                // ================= a chunk of leaves: the f64 tests of List::hit over the leaf's records =================
                const int e = (c << 5) + lane;
                if (e < nL) {
                    id = l_leaf[b * R + e];
                    const int32_t code = st_i[0 * R + id];
                    int32_t depth = st_i[2 * R + id];
                    const int slot = st_i[3 * R + id];
                    const int32_t v = ~code;
                    const int32_t first = v >> 4, count = v & 15;
                    const RayD ray{mk(pool.ox[slot], pool.oy[slot], pool.oz[slot]), mk(pool.dx[slot], pool.dy[slot], pool.dz[slot]), pool.time[slot]};
                    double best_t = pool.best_t[slot];
                    for (int32_t k = 0; k < count; ++k) {
                        const int4 h = __ldg(reinterpret_cast<const int4*>(sc.records + first + k));
                        double t;
                        int32_t hit_rec;
                        if (test_geometry(sc, first + k, h, ray, 0.001, best_t, t, hit_rec, tally)) {
                            best_t = t;
                            pool.best_t[slot] = t;  // closer than anything so far (the medium candidate included)
                            pool.best_rec[slot] = hit_rec;
                            pool.best_chain[slot] = h.w;
                            st_f[9 * R + id] = __double2float_ru(t);
                        }
                    }
                    // pop: the deferred node (or leaf, or the sentinel) is next
                    const int32_t cur = st_i[1 * R + id];
                    st_i[0 * R + id] = cur;
                    st_i[1 * R + id] = stk[depth * R + id];  // row `depth` is the entry below the top (row 0: pad)
                    st_i[2 * R + id] = depth - 1;
                    kind = cur >= 0 ? 1 : (cur == kSentinel ? 3 : 2);
                }
            } else if (c < cL + cN) {
                // ================= a chunk of rays at inner nodes: up to `burst` fp32 steps each =================
                const int e = ((c - cL) << 5) + lane;
                if (e < nN) {
                    id = l_node[b * R + e];
                    SlabRay s;
                    s.idx = st_f[0 * R + id]; s.idy = st_f[1 * R + id]; s.idz = st_f[2 * R + id];
                    s.anx = st_f[3 * R + id]; s.any = st_f[4 * R + id]; s.anz = st_f[5 * R + id];
                    s.afx = st_f[6 * R + id]; s.afy = st_f[7 * R + id]; s.afz = st_f[8 * R + id];
                    const float tmax_f = st_f[9 * R + id];
                    int32_t cur = st_i[0 * R + id], top = st_i[1 * R + id];
                    const int32_t depth0 = st_i[2 * R + id];
                    uint32_t sp = s_stk + (uint32_t)((depth0 + 1) * R + id) * 4;  // first free entry of this ray's column
                    int left = cfg.burst;
                    while (cur >= 0 && left > 0) {
                        --left;
                        cur = node_step2<R, kAllStaged>(s_nodes, ncfg, g_nodes, cur, s, tmin_f, tmax_f, top, sp, tally);
                    }
                    st_i[0 * R + id] = cur;
                    st_i[1 * R + id] = top;
                    st_i[2 * R + id] = (int32_t)((sp - s_stk) >> 2) / R - 1;
                    kind = cur >= 0 ? 1 : (cur == kSentinel ? 3 : 2);
                }
            } else {
                // ================= a chunk of free state slots: new rays from the CTA's share of the pool =================
                const int e = ((c - cL - cN) << 5) + lane;
                if (e < nF) {
                    id = l_free[b * R + e];
                    const int j = base + e;
                    const int slot = ((((j >> 5) * G + bid) << 5) | (j & 31));
                    kind = 3;
                    if (slot < n_slots && pool.bounce[slot] >= 0) {
                        const d3 o = mk(pool.ox[slot], pool.oy[slot], pool.oz[slot]), d = mk(pool.dx[slot], pool.dy[slot], pool.dz[slot]);
                        SlabRay s;
                        make_slab(o, d, s);
                        st_f[0 * R + id] = s.idx; st_f[1 * R + id] = s.idy; st_f[2 * R + id] = s.idz;
                        st_f[3 * R + id] = s.anx; st_f[4 * R + id] = s.any; st_f[5 * R + id] = s.anz;
                        st_f[6 * R + id] = s.afx; st_f[7 * R + id] = s.afy; st_f[8 * R + id] = s.afz;
                        st_f[9 * R + id] = __double2float_ru(pool.best_t[slot]);
                        st_i[0 * R + id] = root;
                        st_i[1 * R + id] = kSentinel;
                        st_i[2 * R + id] = 0;
                        st_i[3 * R + id] = slot;
                        ++my_rays;
                        kind = 1;
                    }
                }
            }
            t3_push(kind == 1, l_node + nb * R, ctr + nb, id, lane, lt);
            t3_push(kind == 2, l_leaf + nb * R, ctr + 2 + nb, id, lane, lt);
            t3_push(kind == 3, l_free + nb * R, ctr + 4 + nb, id, lane, lt);
        }
        if (!dry && tid == 0) ctr[6] = base + nF;  // (once the share is dry the free list is not looked at any more)
        b = nb;
    }
    if (ray_count) {
        for (int off = 16; off > 0; off >>= 1) my_rays += __shfl_xor_sync(FULL, my_rays, off);
        if (lane == 0 && my_rays) atomicAdd(ray_count, (unsigned long long)my_rays);
    }
    if constexpr (kCount) {
        uint32_t vals[3] = {tally.n_node, tally.n_sphere, tally.n_rect};
#pragma unroll
        for (int k = 0; k < 3; ++k)
            for (int off = 16; off > 0; off >>= 1) vals[k] += __shfl_xor_sync(FULL, vals[k], off);
        if (lane == 0) {
            atomicAdd(&counters->node_visits, (unsigned long long)vals[0]);
            atomicAdd(&counters->box_tests, 2ull * vals[0]);
            atomicAdd(&counters->sphere_tests, (unsigned long long)vals[1]);
            atomicAdd(&counters->rect_tests, (unsigned long long)vals[2]);
        }
    }
}

}  // namespace rtx
