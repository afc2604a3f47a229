// trace2.cuh — K2b, second form: the closest surface hit of every ray in flight, with the world BVH and the
// traversal stacks in SHARED memory and the warp as a persistent scheduling unit.
//
// Why it was built (profiles/r1_wf_trace_kernel_details.txt + the raw counters of the same capture): in the first form of
// this kernel (wf_trace_kernel, kernels.cuh) the L1TEX data pipe is the busiest unit (l1tex__data_pipe_lsu_wavefronts
// 64 % of peak while active) and long_scoreboard the top stall — a warp whose lanes sit in different nodes pays one L1
// wavefront PER LANE for each of the four LDG.128 of a node and for each local-memory stack access (2.2-2.5 of 32
// bytes used per sector) — and only 8.5 of 32 lanes are active per instruction. Here the node array (85 KB for the
// book-2 final scene, 106 KB with the padding that spreads it over the banks) is staged once per CTA into shared
// memory, the stack is a per-thread column of a shared array (entry k of thread t at [k][t]: conflict-free by
// construction) with its top in a register, so the pop that follows a miss is off the critical path, and nothing in the
// node loop touches L1; and the warp is a persistent scheduling unit that runs a small state machine instead of a plain
// while-while loop: every lane is at an inner node, at a leaf, or idle (ray finished / none yet), and each round the
// warp runs ONE kind of work, chosen by vote with two thresholds — leaves are tested once `t_leaf` lanes wait at one
// (or nobody has a node left), idle lanes are refilled from the CTA's share of the pool once `t_refill` of them wait (or
// nothing else is left to do), otherwise every lane that has a node takes up to `burst` node steps. t_leaf = t_refill
// = 33 is the plain while-while loop with whole-warp refills.
//
// What it measured (scene 9, one B200; DESIGN.md §4 has the table): L1 data-pipe load fell from 64 % to 31 % and the
// local-memory traffic is gone, but a warp still issues once every 9.3 cycles (9.4 before): the kernel is bound by the
// latency of its dependent chain (fetch, 12 FMAs, min / max, compare, branch) at the 6-7 warps per scheduler its
// registers allow, not by L1. The votes (16 / 8 / 8) cut its warp instructions by 21 % against whole-warp refills and
// by 8 % against the first form; alone it is 12 % faster than the first form (16.0 against 18.3 ms of trace time per
// 17 spp). In production it is 5 % SLOWER (570 against 600 M samples/s): one CTA of 896 threads x 72 registers holds the
// whole register file of its SM, so the other partition's shade kernel cannot share the SM with it the way it does with
// the first form's 128-thread CTAs, and that overlap is worth more than the trace kernel gained. Opt-in (RTX_TRACE=2).
//
// Replaces the recursion of BvhTree::hit (hittable.rs:355-368) over Bound::hit (bound.rs:13-32); same answers as
// wf_trace_kernel (tests/test_gpu_parity.py::test_other_kernel_forms_trace_the_same_rays).
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

constexpr uint32_t kNodeStride = 80;  // bytes between staged nodes: 64 of data + 16 of padding, so that chunk c of node i starts in
                                      // 16-byte bank group (5 i + c) mod 8 — lanes in different nodes spread over the banks

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int2 lds64i(uint32_t addr) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ int32_t lds32i(uint32_t addr) {
    int32_t v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32i(uint32_t addr, int32_t v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v)); }

// One inner node from the staged copy (or, past the staged range, from global memory): both child boxes against
// the ray; returns the next node and defers the farther child when both are hit. Node references are relative to
// the first staged node. `top` is the top of the stack (in a register), `sp` the shared-window address of the first
// free entry of this thread's column (stride kThreads entries). Shared memory is addressed through 32-bit window
// addresses: with generic pointers ptxas rebuilt the window base (S2UR + two uniform ops) in every trip of the loop.
template <int kThreads, bool kAllStaged, bool kCount>
__device__ __forceinline__ int32_t node_step2(uint32_t s_nodes, const TraceCfg& cfg, const BvhNode* __restrict__ g_nodes,
                                              int32_t cur, const SlabRay& s, float tmin_f, float tmax_f, int32_t& top, uint32_t& sp,
                                              Tally<kCount>& tally) {
    float4 q0, q1, q2;
    int2 meta;
    if (kAllStaged || cur < cfg.n_stage) {
        const uint32_t p = s_nodes + (uint32_t)cur * kNodeStride;
        q0 = lds128(p);
        q1 = lds128(p + 16);
        q2 = lds128(p + 32);
        meta = lds64i(p + 48);
    } else {
        const float4* np = reinterpret_cast<const float4*>(g_nodes + cur);
        q0 = __ldg(np); q1 = __ldg(np + 1); q2 = __ldg(np + 2);
        const int4 mt = __ldg(reinterpret_cast<const int4*>(np + 3));
        meta = make_int2(mt.x >= 0 ? mt.x - cfg.stage_first : mt.x, mt.y >= 0 ? mt.y - cfg.stage_first : mt.y);
    }
    tally.node();
    float n0, n1;
    const bool h0 = slab(s, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, tmin_f, tmax_f, n0);
    const bool h1 = slab(s, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, tmin_f, tmax_f, n1);
    if (h0 && h1) {
        const bool swap = n1 < n0;
        sts32i(sp, top);
        sp += kThreads * 4;
        top = swap ? meta.x : meta.y;
        return swap ? meta.y : meta.x;
    }
    if (h0) return meta.x;
    if (h1) return meta.y;
    const int32_t r = top;
    sp -= kThreads * 4;
    top = lds32i(sp);  // (below the first entry lies the pad row: a finished ray pops it once, nobody looks at the value)
    return r;
}

template <bool kCount, int kThreads, bool kAllStaged>
__global__ void __launch_bounds__(kThreads, 1) wf_trace2_kernel(SceneView sc, PathPool pool, int n_slots, TraceCfg cfg,
                                                                 unsigned long long* ray_count, Counters* counters) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    // layout: staged nodes (kNodeStride bytes each) | stack [1 + kShortStack][kThreads] int32, row 0 = pad | counter
    const uint32_t s_nodes = (uint32_t)__cvta_generic_to_shared(tr_smem);
    const uint32_t s_stack = s_nodes + (uint32_t)cfg.cap * kNodeStride;
    int* s_next = reinterpret_cast<int*>(tr_smem + (size_t)cfg.cap * kNodeStride + (size_t)(1 + kShortStack) * kThreads * 4);
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
            *reinterpret_cast<uint4*>(tr_smem + (size_t)n * kNodeStride + (size_t)c * 16) = v;
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
    const uint32_t sp0 = s_stack + (uint32_t)(kThreads + tid) * 4;  // first entry of this thread's column
    const float tmin_f = __int_as_float(0x3a83126e);  // the largest float below 0.001 (main.rs:36's t_min, rounded down)
    const uint32_t lt = lanemask_lt();

    // lane state
    int32_t cur = kSentinel;  // >= 0: inner node; < 0: leaf code; kSentinel: idle
    int32_t top = kSentinel;
    uint32_t sp = sp0;
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
                sp -= kThreads * 4;
                top = lds32i(sp);
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
            int left = cfg.burst;
            while (cur >= 0 && left > 0) {  // (a per-lane loop: a vote per step cost 7 instructions and 13 % of the stall samples)
                --left;
                cur = node_step2<kThreads, kAllStaged>(s_nodes, cfg, g_nodes, cur, sr, tmin_f, tmax_f, top, sp, tally);
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

namespace rtx {

// ---------------------------------------------------------------------------
// K2b, form "1c": the launch shape of the first form (one slot per thread, 128-thread CTAs, nothing persistent, so the
// other partition's shade CTAs interleave with these on every SM) with the two things the ncu source view of the first
// form shows its warps waiting for taken off the critical path: the stack is a shared-memory column with its top in a
// register (the pop that follows a miss was 11.6 % of all stall samples as a local-memory load; local stack traffic
// used 2.2-2.5 of the 32 bytes of each sector it moved), and a node is two 32-byte loads instead of four 16-byte ones
// (one L1 wavefront per lane per load instruction: the L1 data pipe was the busiest unit at 64 %). Measured: 596 M
// samples/s against 604 M for the first form on scene 9, 746 against 752 on scene 7 — neither the local-memory stack
// nor the number of node loads is what the first form waits for. Opt-in (RTX_TRACE=4).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

template <bool kCount>
__global__ void __launch_bounds__(kTraceBlock, WF_TRACE_MINB*(128 / kTraceBlock)) wf_trace1c_kernel(SceneView sc, PathPool pool, int n_slots,
                                                                                                    unsigned long long* ray_count, Counters* counters) {
    __shared__ int32_t s_stack[(1 + kShortStack) * kTraceBlock];  // row 0 = pad (a finished ray pops it once)
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    const int i = blockIdx.x * kTraceBlock + tid;
    Tally<kCount> tally;
    const bool act = i < n_slots && pool.bounce[i] >= 0;
    if (act) {
        const RayD ray{mk(pool.ox[i], pool.oy[i], pool.oz[i]), mk(pool.dx[i], pool.dy[i], pool.dz[i]), pool.time[i]};
        Best best{pool.best_t[i], -1, 0};
        SlabRay s;
        make_slab(ray.o, ray.d, s);
        const float tmin_f = __int_as_float(0x3a83126e);  // the largest float below 0.001
        float tmax_f = __double2float_ru(best.t);
        int32_t top = kSentinel;
        uint32_t sp = (uint32_t)__cvta_generic_to_shared(s_stack) + (uint32_t)(kTraceBlock + tid) * 4;
        int32_t cur = sc.world_root;
        while (true) {
            while (cur >= 0) {
                const float4* np = reinterpret_cast<const float4*>(sc.nodes + cur);
                float4 q0, q1, q2, q3;
                ldg256(np, q0, q1);
                ldg256(np + 2, q2, q3);
                tally.node();
                float n0, n1;
                const bool h0 = slab(s, q0.x, q0.y, q0.z, q0.w, q2.x, q2.y, tmin_f, tmax_f, n0);
                const bool h1 = slab(s, q1.x, q1.y, q1.z, q1.w, q2.z, q2.w, tmin_f, tmax_f, n1);
                const int32_t c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
                if (h0 && h1) {
                    const bool swap = n1 < n0;
                    sts32i(sp, top);
                    sp += kTraceBlock * 4;
                    top = swap ? c0 : c1;
                    cur = swap ? c1 : c0;
                } else if (h0) {
                    cur = c0;
                } else if (h1) {
                    cur = c1;
                } else {
                    cur = top;
                    sp -= kTraceBlock * 4;
                    top = lds32i(sp);
                }
            }
            if (cur == kSentinel) break;
            const int32_t v = ~cur;
            const int32_t first = v >> 4, count = v & 15;
            cur = top;
            sp -= kTraceBlock * 4;
            top = lds32i(sp);
            for (int32_t k = 0; k < count; ++k) {
                const int4 h = __ldg(reinterpret_cast<const int4*>(sc.records + first + k));
                double t;
                int32_t hit_rec;
                if (test_geometry(sc, first + k, h, ray, 0.001, best.t, t, hit_rec, tally)) {
                    best.t = t;
                    best.rec = hit_rec;
                    best.chain = h.w;
                    tmax_f = __double2float_ru(t);
                }
            }
        }
        if (best.rec >= 0) {  // closer than the medium candidate (if any) the shade kernel left there
            pool.best_t[i] = best.t;
            pool.best_rec[i] = best.rec;
            pool.best_chain[i] = best.chain;
        }
    }
    if (ray_count) {
        const unsigned am = __ballot_sync(FULL, act);
        if (lane == 0 && am != 0) atomicAdd(ray_count, (unsigned long long)__popc(am));
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
