// Border following for the device instance-info path.
//
// The reference asks OpenCV for `cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0]`
// of every instance's bounding-box crop (loader/postproc.py:27-31; tiatoolbox
// HoVerNet.get_instance_info does the same, infer/wsi.py:150). OpenCV's raster scan (Suzuki &
// Abe border following) links every new border at the HEAD of its parent's child list, so
// element [0] of the flattened tree is the outer border of the top-level 8-connected component
// found LAST by the scan; CHAIN_APPROX_SIMPLE keeps a border pixel iff the chain direction
// changes there. Both rules are restated here from the published algorithm: one scan per
// instance over its own box with the marks of the work image kept in an int32 plane (unique
// border numbers, so OpenCV's 7-bit number recycling and its disambiguation step vanish).
//
// Compiled twice: by nvcc for the kernels in instinfo.cu and by g++ for the CPU unit test that
// diffs it against the OpenCV of this image (tests/native/contour_host.cpp) — the test build is
// a checker of this file, not a product path.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CC_HD __host__ __device__ __forceinline__
#else
#define CC_HD static inline
#endif

struct cc_view {
  const int32_t* lab;  // [H, W] instance labels
  int32_t* mark;       // [H*up, W*up] zero-initialised; an instance writes only its own pixels
  int W;               // row stride of `lab`
  int up;              // nearest-neighbour upsampling factor of the label image (infer/tile.py:196)
  int r0, c0, h, w;    // bounding box of the instance in upsampled coordinates
  int32_t id;
};

struct cc_border {
  int y, x;  // start pixel of the chosen border (box coordinates), -1 if the instance is empty
  int npts;  // points CHAIN_APPROX_SIMPLE keeps
};

CC_HD bool cc_fg(const cc_view& v, int y, int x) {
  if ((unsigned)y >= (unsigned)v.h || (unsigned)x >= (unsigned)v.w) return false;
  const int Y = v.r0 + y, X = v.c0 + x;
  return v.lab[(size_t)(Y / v.up) * v.W + X / v.up] == v.id;
}

CC_HD int32_t* cc_mark(const cc_view& v, int y, int x) {
  return v.mark + (size_t)(v.r0 + y) * ((size_t)v.W * v.up) + (v.c0 + x);
}

// Value of OpenCV's work image: 0 background (and the zero frame around the crop), 1 untouched
// foreground, otherwise +-(border number).
CC_HD int32_t cc_pix(const cc_view& v, int y, int x) {
  if (!cc_fg(v, y, x)) return 0;
  const int32_t m = *cc_mark(v, y, x);
  return m ? m : 1;
}

// Follows one border from its start pixel (the search order of OpenCV's contour fetch: first
// neighbour clockwise from west / east, then counter-clockwise from the direction of arrival).
// MARK: write +-nbd into the mark plane (negative = a zero pixel lies to the east). out_xy
// (nullable): receives the kept points as (x, y) in box coordinates.
template <bool MARK>
CC_HD int cc_trace(const cc_view& v, int y0, int x0, int is_hole, int32_t nbd, int32_t* out_xy,
                   int add_x, int add_y) {
  // OpenCV's eight chain directions, counter-clockwise from east (y grows downwards):
  // dx = {1,1,0,-1,-1,-1,0,1}, dy = {0,-1,-1,-1,0,1,1,1}, as nibble tables (no local-memory array)
#define CC_DX(s) ((int)((0x21000122u >> (4 * (s))) & 15u) - 1)
#define CC_DY(s) ((int)((0x22210001u >> (4 * (s))) & 15u) - 1)
  int s_end = is_hole ? 0 : 4, s = s_end;
  int y1, x1;
  do {
    s = (s - 1) & 7;
    y1 = y0 + CC_DY(s);
    x1 = x0 + CC_DX(s);
  } while (!cc_fg(v, y1, x1) && s != s_end);
  if (s == s_end) {  // isolated pixel
    if (MARK) *cc_mark(v, y0, x0) = -nbd;
    if (out_xy) {
      out_xy[0] = x0 + add_x;
      out_xy[1] = y0 + add_y;
    }
    return 1;
  }
  int y3 = y0, x3 = x0, y4 = y0, x4 = x0, prev_s = s ^ 4, n = 0;
  for (;;) {
    s_end = s;
    while (s < 15) {
      ++s;
      y4 = y3 + CC_DY(s & 7);
      x4 = x3 + CC_DX(s & 7);
      if (cc_fg(v, y4, x4)) break;
    }
    s &= 7;
    if (MARK) {
      int32_t* m = cc_mark(v, y3, x3);
      if ((unsigned)(s - 1) < (unsigned)s_end)
        *m = -nbd;
      else if (*m == 0)
        *m = nbd;
    }
    if (s != prev_s) {
      if (out_xy) {
        out_xy[2 * n] = x3 + add_x;
        out_xy[2 * n + 1] = y3 + add_y;
      }
      ++n;
      prev_s = s;
    }
    if (y4 == y0 && x4 == x0 && y3 == y1 && x3 == x1) break;
    y3 = y4;
    x3 = x4;
    s = (s + 4) & 7;
  }
  return n;
}

// The raster scan: finds every outer / hole border in OpenCV's order, tracks for outer borders
// whether their parent is the frame, and returns the LAST such border (= contours[0]).
// Mark encoding: |mark| = 4 * number + 2 * is_hole + parent_is_frame.
CC_HD cc_border cc_scan(const cc_view& v) {
  cc_border res;
  res.y = res.x = -1;
  res.npts = 0;
  int32_t number = 0;
  for (int y = 0; y < v.h; ++y) {
    int32_t lnbd = 0, prev = 0;
    for (int x = 0; x <= v.w; ++x) {  // x == w is the zero frame column
      int32_t p = cc_pix(v, y, x);
      if (p == prev) continue;
      int is_hole = 0;
      bool start = true;
      if (!(prev == 0 && p == 1)) {
        if (p != 0 || prev < 1) {
          start = false;
        } else {
          if (prev & -2) lnbd = prev;
          is_hole = 1;
        }
      }
      if (start) {
        int top = 0;
        if (!is_hole) {
          if (lnbd == 0) {
            top = 1;
          } else {
            const int32_t a = lnbd < 0 ? -lnbd : lnbd;
            top = (a & 2) ? 0 : (a & 1);
          }
        }
        ++number;
        const int32_t nbd = number * 4 + is_hole * 2 + top;
        const int oy = y, ox = x - is_hole;
        const int n = cc_trace<true>(v, oy, ox, is_hole, nbd, nullptr, 0, 0);
        if (!is_hole && top) {
          res.y = oy;
          res.x = ox;
          res.npts = n;
        }
        lnbd = cc_pix(v, oy, ox);
        p = cc_pix(v, y, x);
      }
      prev = p;
      if (prev & -2) lnbd = prev;
    }
  }
  return res;
}

// ---- the same scan, 32 pixels of a row at a time (one warp per instance on the device) --------
// A chunk is classified with three ballots (border starts, hole starts, marked pixels); every
// other decision is warp-uniform and re-reads the (L1-resident) work image, so the control flow
// below is the code the CPU unit test exercises with a 32-iteration loop in place of the ballots.
struct cc_masks {
  unsigned start, hole, marked;
};

#ifdef __CUDA_ARCH__
#define CC_CLZ(m) __clz((int)(m))
#define CC_FFS(m) __ffs((int)(m))
__device__ __forceinline__ cc_masks cc_chunk(const cc_view& v, int y, int x, int32_t prev,
                                             int lane) {
  const int xi = x + lane;
  const bool valid = xi <= v.w;
  const int32_t p = valid ? cc_pix(v, y, xi) : 0;
  int32_t pp = __shfl_up_sync(0xffffffffu, p, 1);
  if (lane == 0) pp = prev;
  const bool outer = valid && pp == 0 && p == 1;
  const bool hole = valid && p == 0 && pp >= 1;
  cc_masks m;
  m.start = __ballot_sync(0xffffffffu, outer || hole);
  m.hole = __ballot_sync(0xffffffffu, hole);
  m.marked = __ballot_sync(0xffffffffu, valid && (p & -2));
  return m;
}
#else
#define CC_CLZ(m) __builtin_clz((unsigned)(m))
#define CC_FFS(m) __builtin_ffs((int)(m))
static inline cc_masks cc_chunk(const cc_view& v, int y, int x, int32_t prev, int) {
  cc_masks m = {0u, 0u, 0u};
  int32_t pp = prev;
  for (int lane = 0; lane < 32; ++lane) {
    const int xi = x + lane;
    const bool valid = xi <= v.w;
    const int32_t p = valid ? cc_pix(v, y, xi) : 0;
    const bool outer = valid && pp == 0 && p == 1;
    const bool hole = valid && p == 0 && pp >= 1;
    if (outer || hole) m.start |= 1u << lane;
    if (hole) m.hole |= 1u << lane;
    if (valid && (p & -2)) m.marked |= 1u << lane;
    pp = p;
  }
  return m;
}
#endif

CC_HD cc_border cc_scan_warp(const cc_view& v, int lane) {
  cc_border res;
  res.y = res.x = -1;
  res.npts = 0;
  int32_t number = 0;
  for (int y = 0; y < v.h; ++y) {
    int32_t lnbd = 0, prev = 0;
    int x = 0;
    while (x <= v.w) {
      const cc_masks m = cc_chunk(v, y, x, prev, lane);
      if (!m.start) {
        if (m.marked) lnbd = cc_pix(v, y, x + 31 - CC_CLZ(m.marked));
        prev = cc_pix(v, y, x + 31);
        x += 32;
        continue;
      }
      const int k = CC_FFS(m.start) - 1;
      const unsigned below = m.marked & ((1u << k) - 1u);
      if (below) lnbd = cc_pix(v, y, x + 31 - CC_CLZ(below));
      const int is_hole = (m.hole >> k) & 1;
      int top = 0;
      if (!is_hole) {
        if (lnbd == 0) {
          top = 1;
        } else {
          const int32_t a = lnbd < 0 ? -lnbd : lnbd;
          top = (a & 2) ? 0 : (a & 1);
        }
      }
      ++number;
      const int32_t nbd = number * 4 + is_hole * 2 + top;
      const int oy = y, ox = x + k - is_hole;
      int n = 0;
      if (lane == 0) n = cc_trace<true>(v, oy, ox, is_hole, nbd, nullptr, 0, 0);
#ifdef __CUDA_ARCH__
      __syncwarp();
      n = __shfl_sync(0xffffffffu, n, 0);
#endif
      if (!is_hole && top) {
        res.y = oy;
        res.x = ox;
        res.npts = n;
      }
      lnbd = cc_pix(v, oy, ox);
      prev = cc_pix(v, y, x + k);
      if (prev & -2) lnbd = prev;
      x += k + 1;
    }
  }
  return res;
}
