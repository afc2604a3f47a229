// tools/sah_cost.cpp — surface-area-heuristic cost of the world BVH the host flattener builds for a built-in scene
// (expected node steps and primitive tests of a random ray that hits the root box). A quick, GPU-free proxy for
// builder experiments:   g++ -O2 -std=c++17 -Iinclude tools/sah_cost.cpp rttnw_b200/csrc/flatten.cpp \
//                            rttnw_b200/csrc/scenes.cpp rttnw_b200/csrc/png_io.cpp -lz -o /tmp/sah_cost && /tmp/sah_cost 9
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../rttnw_b200/csrc/flatten.hpp"
#include "../rttnw_b200/csrc/scene_api.hpp"

using namespace rtx;

static double half_area(const float lo[3], const float hi[3]) {
    double d[3] = {(double)hi[0] - lo[0], (double)hi[1] - lo[1], (double)hi[2] - lo[2]};
    if (d[0] < 0 || d[1] < 0 || d[2] < 0) return 0.0;
    return d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
}

struct Acc { double node_area = 0, leaf_area = 0, prim_area = 0; int nodes = 0, leaves = 0, prims = 0, depth = 0; double by_depth[64] = {0}; };

static void walk(const FlatScene& fs, int32_t ref, double area, int depth, Acc& a) {
    if (depth > a.depth) a.depth = depth;
    if (ref < 0) {
        int count = (~ref) & 15;
        a.leaf_area += area; a.prim_area += area * count; a.leaves++; a.prims += count;
        return;
    }
    const BvhNode& n = fs.nodes[(size_t)ref];
    a.node_area += area; a.nodes++; a.by_depth[depth < 64 ? depth : 63] += area;
    float lo0[3] = {n.c0x[0], n.c0y[0], n.c0z[0]}, hi0[3] = {n.c0x[1], n.c0y[1], n.c0z[1]};
    float lo1[3] = {n.c1x[0], n.c1y[0], n.c1z[0]}, hi1[3] = {n.c1x[1], n.c1y[1], n.c1z[1]};
    walk(fs, n.child0, half_area(lo0, hi0), depth + 1, a);
    walk(fs, n.child1, half_area(lo1, hi1), depth + 1, a);
}

// the 4-wide copy: node = entries ni, ni + 1; a step tests four boxes
static void walk_wide(const FlatScene& fs, int32_t ni, double area, int depth, Acc& a) {
    if (depth > a.depth) a.depth = depth;
    a.node_area += area; a.nodes++;
    for (int j = 0; j < 4; ++j) {
        const BvhNode& n = fs.nodes[(size_t)ni + (size_t)(j >> 1)];
        const int32_t ref = (j & 1) == 0 ? n.child0 : n.child1;
        float lo[3] = {(j & 1) == 0 ? n.c0x[0] : n.c1x[0], (j & 1) == 0 ? n.c0y[0] : n.c1y[0], (j & 1) == 0 ? n.c0z[0] : n.c1z[0]};
        float hi[3] = {(j & 1) == 0 ? n.c0x[1] : n.c1x[1], (j & 1) == 0 ? n.c0y[1] : n.c1y[1], (j & 1) == 0 ? n.c0z[1] : n.c1z[1]};
        const double ar = half_area(lo, hi);
        if (ref >= 0) walk_wide(fs, ref, ar, depth + 1, a);
        else if (ref != kEmptyChild) { a.leaf_area += ar; a.prim_area += ar * ((~ref) & 15); a.leaves++; a.prims += (~ref) & 15; }
    }
}

int main(int argc, char** argv) {
    int number = argc > 1 ? atoi(argv[1]) : 9;
    rttnw::SceneBuilder builder;
    if (!rttnw::builtin_scene(number, 0, "assets/earth.png", builder)) { fprintf(stderr, "no scene %d\n", number); return 1; }
    FlatScene fs;
    std::string err;
    if (!flatten_scene(builder.desc, fs, err, false, true)) { fprintf(stderr, "flatten: %s\n", err.c_str()); return 1; }
    const BvhNode& r = fs.nodes[(size_t)fs.world_root];
    float lo[3], hi[3];
    lo[0] = std::min(r.c0x[0], r.c1x[0]); hi[0] = std::max(r.c0x[1], r.c1x[1]);
    lo[1] = std::min(r.c0y[0], r.c1y[0]); hi[1] = std::max(r.c0y[1], r.c1y[1]);
    lo[2] = std::min(r.c0z[0], r.c1z[0]); hi[2] = std::max(r.c0z[1], r.c1z[1]);
    double root = half_area(lo, hi);
    Acc a;
    walk(fs, fs.world_root, root, 1, a);
    printf("scene %d: %d nodes, %d leaves, %d leaf records, depth %d | expected per random ray through the root box: "
           "%.3f node steps, %.3f leaf visits, %.3f primitive tests\n",
           number, a.nodes, a.leaves, a.prims, a.depth, a.node_area / root, a.leaf_area / root, a.prim_area / root);
    if (fs.wide_root >= 0) {
        Acc w;
        walk_wide(fs, fs.wide_root, root, 1, w);
        printf("  4-wide copy: %d nodes, %d leaves, depth %d | %.3f node steps (4 boxes each), %.3f leaf visits, %.3f primitive tests\n",
               w.nodes, w.leaves, w.depth, w.node_area / root, w.leaf_area / root, w.prim_area / root);
    }
    if (argc > 2) {
        printf("  node steps by depth:");
        for (int d = 1; d <= a.depth; ++d) printf(" %.2f", a.by_depth[d] / root);
        printf("\n");
    }
    return 0;
}
