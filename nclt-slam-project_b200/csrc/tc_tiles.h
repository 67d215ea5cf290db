// Tile table of the block-scaled fp4 operand images (tc_hamming.cu): plain host C++, shared with the diagnostics
// library so that tests/test_tc_tiles_cpu.py can check it without a GPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace nclt_tc4 {
constexpr int TILE_ROWS = 240;                  // rows (TMEM accumulator columns) of a full tile
constexpr int SEG4 = 48;                        // keyframes are padded to whole 48-row segments; a tile is <= 5 segments
constexpr int GROUP_MIN_ROWS = 16 * TILE_ROWS;    // a tile group closes at the first keyframe boundary after this many rows

// Image row space: every keyframe's rows are padded to a multiple of SEG4 (zero rows with a zero bias row; an empty
// keyframe owns one all-padding segment), and tiles are 240 consecutive image rows ACROSS keyframe boundaries, so a
// keyframe can only end at a segment boundary of a tile.  The host precomputes where: bit s of `endmask` says that a
// keyframe ends with segment s; keyframes inside a tile are consecutive, starting at kf0.
struct LibTile4 {
    uint32_t img_off256;   // byte offset / 256 into the library image
    uint16_t n;            // rows in the tile image (multiple of 48, <= 240)
    uint16_t endmask;      // bit s: a keyframe ends at tile column 48 * (s + 1)
    int kf0;               // keyframe of column 0
    int prow0;             // image row of column 0
};
// Tile table of an fp4 image: 240-row tiles over the IMAGE row space (every keyframe padded to whole 48-row segments; an
// empty keyframe owns one all-padding segment), across keyframe boundaries.  A tile group closes (its last tile is short)
// at the first keyframe boundary after GROUP_MIN_ROWS image rows; work splits start at group boundaries, so the image does
// not depend on the batch size.  counts == nullptr: every keyframe has `stride` rows (the frames of a batch as a library).
struct Tiles4 {
    std::vector<LibTile4> tiles;
    std::vector<int> pstart, grp_tile;
    size_t off256 = 0;
};
inline void build_tiles4(const int* counts, int n_kf, int row_bytes, Tiles4& o, int stride = 0) {
    o.pstart.assign(n_kf + 1, 0);
    for (int k = 0; k < n_kf; ++k)
        o.pstart[k + 1] = o.pstart[k] + std::max(SEG4, ((counts ? counts[k] : stride) + SEG4 - 1) / SEG4 * SEG4);
    const std::vector<int>& pstart = o.pstart;
    int k = 0;
    while (k < n_kf) {
        o.grp_tile.push_back((int)o.tiles.size());
        const int k_a = k, row_a = pstart[k];
        int row_b = row_a;
        while (k < n_kf && row_b - row_a < GROUP_MIN_ROWS) row_b = pstart[++k];
        int kf = k_a;                                          // keyframe of the tile's column 0
        for (int r = row_a; r < row_b; r += TILE_ROWS) {
            const int n = std::min(TILE_ROWS, row_b - r);        // a multiple of 48
            while (pstart[kf + 1] <= r) ++kf;
            uint16_t endmask = 0;
            for (int q = kf; q < k && pstart[q + 1] <= r + n; ++q) endmask |= (uint16_t)(1u << ((pstart[q + 1] - r) / SEG4 - 1));
            o.tiles.push_back(LibTile4{(uint32_t)o.off256, (uint16_t)n, endmask, kf, r});
            o.off256 += (size_t)n * row_bytes / 256;           // 48 rows x 160 (192) bytes = 30 (36) x 256
        }
    }
    o.grp_tile.push_back((int)o.tiles.size());
}

}  // namespace nclt_tc4
