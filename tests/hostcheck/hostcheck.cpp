// Host build of the FPB_HD device routines, for the no-GPU test suite only
// (tests/test_hostcheck_*.py).  Never loaded by the product package.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "hd_geometry.h"

extern "C" {

// largest-contour convex-hull fill of an 8-bit mask (non-zero = set), sequential driver
// mirroring k_seg_geometry.  out_hull: H*W bytes {0,255}; bbox = x,y,w,h; returns number
// of hull vertices (0 when the mask is empty); hull vertices are written to hullxy (x0,y0,x1,y1..).
int hc_largest_hull_fill(const uint8_t* mask, int H, int W, uint8_t* out_hull, int* bbox,
                         long long* best_area2, int* hullxy, int hull_cap) {
    const int wpr = (W + 31) / 32;
    std::vector<uint32_t> bits((size_t)wpr * H, 0u);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (mask[(size_t)y * W + x]) bits[(size_t)y * wpr + (x >> 5)] |= 1u << (x & 31);
    long long best = -1; int bx = -1, by = -1;
    std::vector<uint32_t> vis((size_t)wpr * H, 0u);          // same raster-order walk with "visited" marks as k_seg_main
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (!fpb_bit(bits.data(), wpr, W, H, x, y)) continue;
            if (fpb_bit(bits.data(), wpr, W, H, x - 1, y) || fpb_bit(bits.data(), wpr, W, H, x - 1, y - 1) ||
                fpb_bit(bits.data(), wpr, W, H, x, y - 1) || fpb_bit(bits.data(), wpr, W, H, x + 1, y - 1)) continue;
            if ((vis[(size_t)y * wpr + (x >> 5)] >> (x & 31)) & 1u) continue;
            long long a2 = 0;
            fpb_trace_border(bits.data(), wpr, W, H, x, y, &a2, nullptr, nullptr, 8 * W * H + 16, vis.data());
            if (a2 < 0) a2 = -a2;
            if (a2 > best) { best = a2; bx = x; by = y; }
        }
    memset(out_hull, 0, (size_t)H * W);
    *best_area2 = best;
    if (best < 0) return 0;
    std::vector<int> rowmin(H, 1 << 30), rowmax(H, -1);
    long long a2 = 0;
    fpb_trace_border(bits.data(), wpr, W, H, bx, by, &a2, rowmin.data(), rowmax.data(), 8 * W * H + 16);
    std::vector<int> hx(2 * H + 4), hy(2 * H + 4), tx(2 * H + 4), ty(2 * H + 4);
    const int n = fpb_hull_from_rows(rowmin.data(), rowmax.data(), 0, H - 1, hx.data(), hy.data(), tx.data(), ty.data());
    int minx = 1 << 30, maxx = -1, miny = 1 << 30, maxy = -1;
    for (int i = 0; i < n; ++i) {
        if (hx[i] < minx) minx = hx[i]; if (hx[i] > maxx) maxx = hx[i];
        if (hy[i] < miny) miny = hy[i]; if (hy[i] > maxy) maxy = hy[i];
        if (i < hull_cap) { hullxy[2 * i] = hx[i]; hullxy[2 * i + 1] = hy[i]; }
    }
    bbox[0] = minx; bbox[1] = miny; bbox[2] = maxx - minx + 1; bbox[3] = maxy - miny + 1;
    for (int y = 0; y < H; ++y) {
        int xa, xb;
        if (fpb_fill_span(hx.data(), hy.data(), n, y, W, &xa, &xb))
            for (int x = xa; x <= xb; ++x) out_hull[(size_t)y * W + x] = 255;
    }
    for (int i = 0; i < n; ++i) {
        const int j = (i == 0) ? n - 1 : i - 1;
        FpbLine L = fpb_line_begin(hx[j], hy[j], hx[i], hy[i]);
        for (int k = 0; k < L.count; ++k) {
            if ((unsigned)L.x < (unsigned)W && (unsigned)L.y < (unsigned)H) out_hull[(size_t)L.y * W + L.x] = 255;
            fpb_line_next(L);
        }
    }
    return n;
}

}  // extern "C"

#include "hd_scalar.h"
extern "C" {
// K1: stretch LUT from a 256-bin histogram
void hc_stretch_lut(const unsigned* hist, int npix, uint8_t* lut) {
    unsigned cum[256]; unsigned run = 0;
    for (int i = 0; i < 256; ++i) { run += hist[i]; cum[i] = run; }
    const float lo = fpb_percentile_u8_unit(cum, npix, 0.5f), hi = fpb_percentile_u8_unit(cum, npix, 99.5f);
    for (int v = 0; v < 256; ++v) lut[v] = fpb_stretch_value(v, lo, hi);
}
int hc_otsu_u8(const unsigned* hist, int npix) { return fpb_otsu_u8(hist, npix); }
float hc_patch_otsu(const unsigned* ih) {
    float counts[256], centers[256], tmp[512];
    return fpb_patch_otsu(ih, counts, centers, tmp);
}
}

extern "C" int hc_gauss_weights(double sigma, double* w) { return fpb_gauss_weights_fill(sigma, w, 64); }

// word-parallel Zhang-Suen deletion test (hd_scalar.h, used by k_thin_extract when the built-in table is installed):
// out[c] bit 0 / bit 1 = neighbourhood code c (NW=1 N=2 NE=4 E=8 SE=16 S=32 SW=64 W=128) is deleted in pass 1 / 2
extern "C" void hc_zs_codes(uint8_t* out) {
    for (int base = 0; base < 256; base += 32) {
        uint32_t pl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < 32; ++j) for (int k = 0; k < 8; ++k) if (((base + j) >> k) & 1) pl[k] |= 1u << j;
        // planes in code-bit order: NW N NE E SE S SW W
        const uint32_t m1 = fpb_zs_delete_mask(pl[0], pl[1], pl[2], pl[3], pl[4], pl[5], pl[6], pl[7], 1);
        const uint32_t m2 = fpb_zs_delete_mask(pl[0], pl[1], pl[2], pl[3], pl[4], pl[5], pl[6], pl[7], 2);
        for (int j = 0; j < 32; ++j) out[base + j] = (uint8_t)(((m1 >> j) & 1u) | (((m2 >> j) & 1u) << 1));
    }
}
