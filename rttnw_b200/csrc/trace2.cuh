// trace2.cuh — K2b, second form: the closest surface hit of every ray in flight, with the world BVH and the
// traversal stacks in SHARED memory and the warp as a persistent scheduling unit.
//
// Why (profiles/r1_wf_trace_kernel_details.txt + the raw counters of the same capture): the first form of this
// kernel (wf_trace_kernel, kernels.cuh) is bound by the L1TEX data pipe, not by issue slots —
// l1tex__data_pipe_lsu_wavefronts 64 % of peak while active, long_scoreboard the top stall — because a warp whose
// lanes sit in different nodes pays one L1 wavefront PER LANE for each of the four LDG.128 of a node and for each
// local-memory stack access (2.2-2.5 of 32 bytes used per sector). That is also why none of the round-1 schedules
// that raised the lanes per instruction gained anything: the wavefront count per ray is the same under all of them.
// Here the node array (85 KB for the book-2 final scene) is staged once per CTA into shared memory as four planes of
// 16-byte chunks (chunk c of node i at plane c, index i: the bank group is i mod 8 for every chunk, so lanes in
// different nodes spread over the banks without a swizzle), the stack is a per-thread column of a shared array
// (entry k of thread t at [k][t]: conflict-free by construction) with its top kept in a register, so the pop that
// follows a miss is off the critical path, and nothing on the node loop touches L1.
//
// With L1 out of the way the binding resource is issue slots at whatever lane utilisation the schedule reaches, so
// the warp runs a small state machine instead of a plain while-while loop: every lane is in one of three states —
// at an inner node, at a leaf, or idle (ray finished / none yet) — and each round the warp runs ONE kind of work,
// chosen by vote with two thresholds: leaves are tested once `t_leaf` lanes wait at one (or nobody has a node left),
// idle lanes are refilled from the CTA's share of the pool once `t_refill` of them wait (or nothing else is left to
// do), otherwise every lane that has a node takes up to `burst` node steps. t_leaf = t_refill = 33 is the plain
// while-while loop with whole-warp refills.
//
// Replaces the recursion of BvhTree::hit (hittable.rs:355-368) over Bound::hit (bound.rs:13-32); same answers as
// wf_trace_kernel (tests/test_gpu_parity.py runs every render test under both).
#pragma once
#include "kernels.cuh"

namespace rtx {

struct TraceCfg {
    int32_t stage_first;  // first node of the world BVH in SceneView::nodes (its root)
    int32_t n_stage;      // nodes staged in shared memory: [stage_first, stage_first + n_stage)
    int32_t cap;          // plane stride of the staged copy, in nodes (>= n_stage)
    int32_t t_leaf, t_refill, burst;
};

constexpr int kShortStack = 16;  // shared-memory stack entries per thread (+ the top in a register): BVHs up to depth 16

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// One inner node from the staged copy (or, past the staged range, from global memory): both child boxes against
// the ray; returns the next node and defers the farther child when both are hit. Node references are relative to
// the first staged node. `top` is the top of the stack (in a register), `sp` points at the first free shared entry
// of this thread's column (stride kThreads entries).
template <int kThreads, bool kAllStaged, bool kCount>
__device__ __forceinline__ int32_t node_step2(const uint4* __restrict__ s_nodes, const TraceCfg& cfg, const BvhNode* __restrict__ g_nodes,
                                              int32_t cur, const SlabRay& s, float tmin_f, float tmax_f, int32_t& top, int32_t*& sp,
                                              Tally<kCount>& tally) {
    float4 q0, q1, q2;
    int4 meta;
    if (kAllStaged || cur < cfg.n_stage) {
        const uint4* p = s_nodes + cur;
        uint4 a = p[0], b = p[cfg.cap], c = p[2 * cfg.cap], d = p[3 * cfg.cap];
        q0 = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
        q1 = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w));
        q2 = make_float4(__uint_as_float(c.x), __uint_as_float(c.y), __uint_as_float(c.z), __uint_as_float(c.w));
        meta = make_int4((int)d.x, (int)d.y, 0, 0);
    } else {
        const float4* np = reinterpret_cast<const float4*>(g_nodes + cur);
        q0 = __ldg(np); q1 = __ldg(np + 1); q2 = __ldg(np + 2);
        meta = __ldg(reinterpret_cast<const int4*>(np + 3));
        if (meta.x >= 0) meta.x -= cfg.stage_first;
        if (meta.y >= 0) meta.y -= cfg.stage_first;
    }
    tally.node();
    float n0, n1;
    const bool h0 = slab(s, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, tmin_f, tmax_f, n0);
    const bool h1 = slab(s, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, tmin_f, tmax_f, n1);
    if (h0 && h1) {
        const bool swap = n1 < n0;
        *sp = top;
        sp += kThreads;
        top = swap ? meta.x : meta.y;
        return swap ? meta.y : meta.x;
    }
    if (h0) return meta.x;
    if (h1) return meta.y;
    const int32_t r = top;
    sp -= kThreads;
    top = *sp;  // (below the first entry lies the pad row: a finished ray pops it once, nobody looks at the value)
    return r;
}

template <bool kCount, int kThreads, bool kAllStaged>
__global__ void __launch_bounds__(kThreads, 1) wf_trace2_kernel(SceneView sc, PathPool pool, int n_slots, TraceCfg cfg,
                                                                 unsigned long long* ray_count, Counters* counters) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    uint4* s_nodes = reinterpret_cast<uint4*>(tr_smem);
    int32_t* s_stack = reinterpret_cast<int32_t*>(s_nodes + 4 * (size_t)cfg.cap);  // [1 + kShortStack][kThreads], row 0 = pad
    int* s_next = reinterpret_cast<int*>(s_stack + (1 + kShortStack) * kThreads);
    Tally<kCount> tally;

    // ---- stage the top of the world BVH: four planes of 16-byte chunks, child references made relative ----
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
            s_nodes[(size_t)c * cfg.cap + n] = v;
        }
    }
    // this CTA's share of the pool: every gridDim.x-th piece of 32 slots (neighbouring slots hold rays of similar cost —
    // they were filled together — so contiguous shares left some SMs idle for a third of the launch), handed out to
    // its warps a few lanes at a time through a shared counter: local index j -> slot ((j / 32) * G + b) * 32 + j % 32
    const int G = (int)gridDim.x, b = (int)blockIdx.x;
    if (tid == 0) *s_next = 0;
    __syncthreads();

    const BvhNode* g_nodes = sc.nodes + cfg.stage_first;
    const int32_t root = sc.world_root - cfg.stage_first;
    int32_t* const sp0 = s_stack + kThreads + tid;  // first entry of this thread's column
    const float tmin_f = __int_as_float(0x3a83126e);  // the largest float below 0.001 (main.rs:36's t_min, rounded down)
    const uint32_t lt = lanemask_lt();

    // lane state
    int32_t cur = kSentinel;  // >= 0: inner node; < 0: leaf code; kSentinel: idle
    int32_t top = kSentinel;
    int32_t* sp = sp0;
    int slot = -1;
    RayD ray{mk(0, 0, 0), mk(0, 0, 1), 0.0};
    SlabRay sr{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    Best best{0.0, -1, 0};
    float tmax_f = 0.f;
    bool dry = (b << 5) >= n_slots;  // warp-uniform: the CTA's share has been handed out
    unsigned int my_rays = 0;

    while (true) {
        const int nn = __popc(__ballot_sync(FULL, cur >= 0));
        const int nl = __popc(__ballot_sync(FULL, cur < 0 && cur != kSentinel));
        const int ni = 32 - nn - nl;
        if (nl > 0 && (nl >= cfg.t_leaf || nn == 0)) {
            // ================= leaf phase: f64 primitive tests =================
            if (cur < 0 && cur != kSentinel) {
                const int32_t v = ~cur;
                const int32_t first = v >> 4, count = v & 15;
                cur = top;
                sp -= kThreads;
                top = *sp;
                for (int32_t i = 0; i < count; ++i) {
                    const int4 h = __ldg(reinterpret_cast<const int4*>(sc.records + first + i));
                    double t;
                    int32_t hit_rec;
                    if (test_geometry(sc, first + i, h, ray, 0.001, best.t, t, hit_rec, tally)) {
                        best.t = t;
                        best.rec = hit_rec;
                        best.chain = h.w;
                        tmax_f = __double2float_ru(t);
                    }
                }
            }
        } else if (!dry && ni > 0 && (ni >= cfg.t_refill || nn == 0)) {
            // ================= refill: write the finished rays back, take new slots =================
            const bool idle = cur == kSentinel;
            if (idle && slot >= 0 && best.rec >= 0) {  // closer than the medium candidate (if any) the shade kernel left there
                pool.best_t[slot] = best.t;
                pool.best_rec[slot] = best.rec;
                pool.best_chain[slot] = best.chain;
            }
            if (idle) slot = -1;
            const unsigned m = __ballot_sync(FULL, idle);
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(s_next, __popc(m));
            base = __shfl_sync(FULL, base, leader);
            const int j = base + __popc(m & lt);
            const int my = ((((j >> 5) * G + b) << 5) | (j & 31));
            if (idle && my < n_slots && pool.bounce[my] >= 0) {
                slot = my;
                ray.o = mk(pool.ox[my], pool.oy[my], pool.oz[my]);
                ray.d = mk(pool.dx[my], pool.dy[my], pool.dz[my]);
                ray.time = pool.time[my];
                best.t = pool.best_t[my];
                best.rec = -1;
                best.chain = 0;
                make_slab(ray.o, ray.d, sr);
                tmax_f = __double2float_ru(best.t);
                cur = root;
                top = kSentinel;
                sp = sp0;
                ++my_rays;
            }
            if (((((base + __popc(m)) >> 5) * G + b) << 5) >= n_slots) dry = true;
        } else if (nn > 0) {
            // ================= node phase: up to `burst` steps for every lane that has a node =================
#pragma unroll 1
            for (int it = 0; it < cfg.burst; ++it) {
                if (cur >= 0) cur = node_step2<kThreads, kAllStaged>(s_nodes, cfg, g_nodes, cur, sr, tmin_f, tmax_f, top, sp, tally);
                if (!__any_sync(FULL, cur >= 0)) break;
            }
        } else {
            break;  // dry, and no ray in flight
        }
    }
    if (slot >= 0 && best.rec >= 0) {  // rays that finished after the share ran dry
        pool.best_t[slot] = best.t;
        pool.best_rec[slot] = best.rec;
        pool.best_chain[slot] = best.chain;
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
