// Host check (nvcc -std=c++17 -I.. tools/tile_order_check.cu): tile_of_order is a bijection with order_of_tile as its inverse, ragged grids and > 2^24 tiles included.
#include <cstdio>
#include <vector>
#include <algorithm>
using std::min;
#include "../rttnw_b200/csrc/kernels.cuh"
int main() {
    int dims[][2] = {{100, 200}, {1, 1}, {8, 16}, {9, 17}, {7, 15}, {75, 150}, {13, 33}, {1024, 2048}, {4099, 8191}};
    for (auto& d : dims) {
        int tx_n = d[0], ty_n = d[1];
        float inv = 1.0f / (float)(tx_n * rtx::kBlockH);
        std::vector<char> seen((size_t)tx_n * ty_n, 0);
        for (unsigned k = 0; k < (unsigned)(tx_n * ty_n); ++k) {
            int tx, ty;
            rtx::tile_of_order(k, tx_n, ty_n, inv, tx, ty);
            if (tx < 0 || tx >= tx_n || ty < 0 || ty >= ty_n || seen[(size_t)ty * tx_n + tx] || rtx::order_of_tile(tx, ty, tx_n, ty_n) != k) {
                printf("FAIL %dx%d k=%u -> %d,%d\n", tx_n, ty_n, k, tx, ty);
                return 1;
            }
            seen[(size_t)ty * tx_n + tx] = 1;
        }
    }
    printf("tile order is a bijection on every grid tried\n");
    return 0;
}
