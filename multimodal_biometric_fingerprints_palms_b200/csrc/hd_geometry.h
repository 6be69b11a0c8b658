// Sequential geometry of the segmentation stage (K3): outer-border following with
// shoelace area, convex hull from per-row extremes, OpenCV-exact filled-polygon spans
// and 8-connected line strokes.
//
// Written as FPB_HD (= __host__ __device__) functions: the CUDA kernel
// `k_seg_geometry` (segment.cu) runs them on a bit-packed mask held in shared memory;
// `tests/hostcheck/hostcheck.cpp` compiles the very same code with g++ so the no-GPU
// test suite can compare it with OpenCV (findContours / contourArea / convexHull /
// drawContours / boundingRect) on random masks.  It is not a CPU execution path of the
// product: nothing in the Python package loads the host build.
//
// Reference behaviour replaced: /root/reference/src/preprocessing/fingerprint_preprocess.py:112-129
//   contours = cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)
//   largest  = max(contours, key=cv2.contourArea); hull = cv2.convexHull(largest)
//   cv2.drawContours(mask_hull, [hull], -1, 255, -1); x,y,w,h = cv2.boundingRect(hull)
#pragma once
#include <stdint.h>

#ifndef FPB_HD
#ifdef __CUDACC__
#define FPB_HD __host__ __device__ __forceinline__
#else
#define FPB_HD static inline
#endif
#endif

// ---- bit-packed binary image: row y occupies words [y*wpr, (y+1)*wpr), bit x&31 of word x>>5
FPB_HD int fpb_bit(const uint32_t* bits, int wpr, int W, int H, int x, int y) {
    if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) return 0;
    return (bits[y * wpr + (x >> 5)] >> (x & 31)) & 1u;
}

// 8-neighbourhood, clockwise on screen (y grows downwards) starting at West.
//            0:W  1:NW  2:N  3:NE  4:E  5:SE  6:S  7:SW
FPB_HD int fpb_dir_dx(int d) { return (d == 0 || d == 1 || d == 7) ? -1 : ((d == 2 || d == 6) ? 0 : 1); }
FPB_HD int fpb_dir_dy(int d) { return (d == 1 || d == 2 || d == 3) ? -1 : ((d == 0 || d == 4) ? 0 : 1); }
#define FPB_DIR_DX(d) fpb_dir_dx(d)
#define FPB_DIR_DY(d) fpb_dir_dy(d)

#ifdef __CUDA_ARCH__
#define FPB_FFS(x) __ffs((int)(x))
#define FPB_CLZ(x) __clz((int)(x))
#else
#define FPB_FFS(x) __builtin_ffs((int)(x))
#define FPB_CLZ(x) ((x) ? __builtin_clz((unsigned)(x)) : 32)
#endif

// 3 horizontally adjacent pixels (x-1, x, x+1) of row y as bits 0..2 (0 outside the image; the words carry no
// set bits beyond column W-1)
FPB_HD unsigned fpb_row3(const uint32_t* bits, int wpr, int H, int x, int y) {
    if ((unsigned)y >= (unsigned)H) return 0u;
    const uint32_t* row = bits + y * wpr;
    const int k = x >> 5, j = x & 31;
    const uint32_t cur = row[k];
    unsigned long long s = (unsigned long long)cur << 1;
    if (j == 0 && k > 0) s |= (unsigned long long)(row[k - 1] >> 31);
    if (j == 31 && k + 1 < wpr) s |= (unsigned long long)(row[k + 1] & 1u) << 33;
    return (unsigned)(s >> j) & 7u;
}

// the 8 neighbours of (x,y) as a ring byte: bit d = neighbour in direction d (0:W 1:NW 2:N 3:NE 4:E 5:SE 6:S 7:SW)
FPB_HD unsigned fpb_ring8(const uint32_t* bits, int wpr, int H, int x, int y) {
    const unsigned t = fpb_row3(bits, wpr, H, x, y - 1), m = fpb_row3(bits, wpr, H, x, y), b = fpb_row3(bits, wpr, H, x, y + 1);
    return (m & 1u) | ((t & 1u) << 1) | (((t >> 1) & 1u) << 2) | (((t >> 2) & 1u) << 3) | (((m >> 2) & 1u) << 4) |
           (((b >> 2) & 1u) << 5) | (((b >> 1) & 1u) << 6) | ((b & 1u) << 7);
}

// Follow the border that starts at set pixel (sx,sy) whose West neighbour is 0
// (Suzuki & Abe 1985, steps 3.1-3.5 - the procedure behind cv2.findContours).  Accumulates
// twice the signed shoelace area of the closed pixel-centre polygon (what
// cv2.contourArea returns, doubled, before fabs).  When rowmin/rowmax are non-null the
// per-row extreme x of the visited pixels are recorded (rowmin must be pre-filled with a
// large value, rowmax with -1).  Returns the number of border steps.
// Each step reads the 3x3 neighbourhood once as a ring byte and finds the next direction with one clz.
// When `visited` is non-null every border pixel is marked in that bit image (same layout as `bits`), so that a
// caller walking the candidates in raster order can skip start pixels that lie on an already followed border.
FPB_HD int fpb_trace_border(const uint32_t* bits, int wpr, int W, int H, int sx, int sy,
                            long long* area2, int* rowmin, int* rowmax, int max_steps, uint32_t* visited = nullptr) {
    (void)W;
    long long acc = 0;
    int steps = 0;
    if (rowmin) { if (sx < rowmin[sy]) rowmin[sy] = sx; if (sx > rowmax[sy]) rowmax[sy] = sx; }
    if (visited) visited[sy * wpr + (sx >> 5)] |= 1u << (sx & 31);
    // 3.1: from West, clockwise (increasing direction index), first set neighbour
    const unsigned ring0 = fpb_ring8(bits, wpr, H, sx, sy);
    if (!ring0) { *area2 = 0; return 0; }                 // isolated pixel
    const int d1 = FPB_FFS(ring0) - 1;
    const int x1 = sx + FPB_DIR_DX(d1), y1 = sy + FPB_DIR_DY(d1);
    // 3.2: (i2,j2) <- (i1,j1), (i3,j3) <- (i,j)
    int cx = sx, cy = sy;           // current pixel (i3,j3)
    int from = d1;                  // direction from the current pixel to (i2,j2)
    unsigned ring = ring0;
    for (;;) {
        // 3.3: counter-clockwise (decreasing index) starting just after `from`, wrapping back to `from` itself:
        // rot bit i = ring bit (from + i) & 7, so the search order from-1, from-2, .., from is rot bit 7, 6, .., 0
        const unsigned rot = (((ring << 8) | ring) >> from) & 255u;
        const int i = 31 - FPB_CLZ(rot);
        const int dn = (from + i) & 7;
        const int nx = cx + FPB_DIR_DX(dn), ny = cy + FPB_DIR_DY(dn);
        acc += (long long)cx * ny - (long long)nx * cy;
        ++steps;
        // 3.5: back at the start, about to repeat the first move?
        if (nx == sx && ny == sy && cx == x1 && cy == y1) break;
        from = (dn + 4) & 7;        // direction from the new pixel back to the old one
        cx = nx; cy = ny;
        if (rowmin) { if (cx < rowmin[cy]) rowmin[cy] = cx; if (cx > rowmax[cy]) rowmax[cy] = cx; }
        if (visited) visited[cy * wpr + (cx >> 5)] |= 1u << (cx & 31);
        if (steps >= max_steps) break;   // safety net, never reached on valid input
        ring = fpb_ring8(bits, wpr, H, cx, cy);
    }
    *area2 = acc;
    return steps;
}

// ---- convex hull (strict vertices only) of the points (rowmin[y],y),(rowmax[y],y), y0<=y<=y1,
// rows with rowmax[y] < 0 are empty.  Output: closed polygon, no repeated first point.
// hx/hy need room for 2*(y1-y0+1)+2 entries.
FPB_HD long long fpb_cross(int ox, int oy, int ax, int ay, int bx, int by) {
    return (long long)(ax - ox) * (by - oy) - (long long)(ay - oy) * (bx - ox);
}

// one monotone chain: dir = +1 walks the rows upwards taking (rowmin, rowmax) per row, dir = -1 downwards taking
// (rowmax, rowmin).  Returns the number of points written to ox/oy.
FPB_HD int fpb_hull_chain(const int* rowmin, const int* rowmax, int y0, int y1, int dir, int* ox, int* oy) {
    int n = 0;
    for (int y = dir > 0 ? y0 : y1; dir > 0 ? y <= y1 : y >= y0; y += dir) {
        if (rowmax[y] < 0) continue;
        for (int s = 0; s < 2; ++s) {
            if (s == 1 && rowmax[y] == rowmin[y]) break;
            const int x = ((s == 1) == (dir > 0)) ? rowmax[y] : rowmin[y];
            while (n >= 2 && fpb_cross(ox[n - 2], oy[n - 2], ox[n - 1], oy[n - 1], x, y) <= 0) --n;
            ox[n] = x; oy[n] = y; ++n;
        }
    }
    return n;
}

// joins the two chains (lower in hx/hy with nl points, upper in tx/ty with nu points) into hx/hy
FPB_HD int fpb_hull_join(int* hx, int* hy, int nl, const int* tx, const int* ty, int nu) {
    if (nl == 0) return 0;
    if (nl == 1) return 1;
    int n = nl - 1;                      // lower chain without its last point
    for (int i = 0; i + 1 < nu; ++i) { hx[n] = tx[i]; hy[n] = ty[i]; ++n; }
    return n;
}

FPB_HD int fpb_hull_from_rows(const int* rowmin, const int* rowmax, int y0, int y1,
                              int* hx, int* hy, int* tx, int* ty) {
    // Andrew's monotone chain over the points sorted by (y, x); `<= 0` pops keep strict
    // vertices only (cv2.convexHull drops collinear points too).  hx/hy receive the hull,
    // tx/ty are scratch; all four need 2*(y1-y0+1)+2 entries.
    const int nl = fpb_hull_chain(rowmin, rowmax, y0, y1, +1, hx, hy);
    const int nu = fpb_hull_chain(rowmin, rowmax, y0, y1, -1, tx, ty);
    return fpb_hull_join(hx, hy, nl, tx, ty, nu);
}

// ---- OpenCV FillEdgeCollection span of scanline y for a convex polygon (drawing.cpp):
// edges carry x in 16.16 fixed point advanced by a truncated per-row slope from their
// upper vertex; active for y0 <= y < y1; the span is [ceil(xl), floor(xr)].
// Returns 0 when the scanline has no span.
FPB_HD long long fpb_cdiv(long long a, long long b) {   // C truncating division, explicit
    long long q = (a < 0 ? -a : a) / (b < 0 ? -b : b);
    return ((a < 0) == (b < 0)) ? q : -q;
}

FPB_HD int fpb_fill_span(const int* hx, const int* hy, int n, int y, int W, int* xa, int* xb) {
    long long lo = 0, hi = 0;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        const int j = (i == 0) ? n - 1 : i - 1;
        int ax = hx[j], ay = hy[j], bx = hx[i], by = hy[i];
        if (ay == by) continue;
        const long long dx = fpb_cdiv((long long)(bx - ax) * 65536, (long long)(by - ay));
        long long x; int ey0, ey1;
        if (ay < by) { ey0 = ay; ey1 = by; x = (long long)ax * 65536; }
        else         { ey0 = by; ey1 = ay; x = (long long)bx * 65536; }
        if (y < ey0 || y >= ey1) continue;
        x += (long long)(y - ey0) * dx;
        if (cnt == 0) { lo = hi = x; } else { if (x < lo) lo = x; if (x > hi) hi = x; }
        ++cnt;
    }
    if (cnt < 2) return 0;
    int a = (int)((lo + 65535) >> 16), b = (int)(hi >> 16);
    if (!(a < W && b >= 0)) return 0;
    if (a < 0) a = 0;
    if (b >= W) b = W - 1;
    if (b < a) return 0;
    *xa = a; *xb = b;
    return 1;
}

// ---- cv::LineIterator (connectivity 8, left-to-right normalised) as used by cv::line /
// the edge strokes of a filled contour.  count = max(|dx|,|dy|)+1 pixels; a minor-axis move
// accompanies an increment iff err < 0 before it.
struct FpbLine { int x, y, sy, steep, dmaj, dmin, err, count; };

FPB_HD FpbLine fpb_line_begin(int x1, int y1, int x2, int y2) {
    int dx = x2 - x1, dy = y2 - y1;
    if (dx < 0) { int t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; dx = -dx; dy = -dy; }
    FpbLine L;
    L.x = x1; L.y = y1;
    L.sy = dy < 0 ? -1 : 1;
    if (dy < 0) dy = -dy;
    L.steep = dy > dx;
    L.dmaj = L.steep ? dy : dx;
    L.dmin = L.steep ? dx : dy;
    L.err = L.dmaj - 2 * L.dmin;
    L.count = L.dmaj + 1;
    return L;
}

FPB_HD void fpb_line_next(FpbLine& L) {
    const int neg = L.err < 0;
    L.err += -2 * L.dmin + (neg ? 2 * L.dmaj : 0);
    if (L.steep) { L.y += L.sy; L.x += neg; }
    else         { L.x += 1;    L.y += neg ? L.sy : 0; }
}
