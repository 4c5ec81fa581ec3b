// CPU build of cerberus_b200/csrc/contour_core.h for tests/test_contour_host.py: the same
// border-following code the instance-info kernels run, callable from ctypes so that it can be
// diffed against the OpenCV of this image on thousands of masks without a GPU. Test-only.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../cerberus_b200/csrc/contour_core.h"

extern "C" int contour0_host(const int32_t* lab, int H, int W, int32_t id, int up, int32_t* out_xy,
                             int cap, int32_t* box4, int chunked) {
  int r0 = H * up, r1 = -1, c0 = W * up, c1 = -1;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      if (lab[(size_t)y * W + x] == id) {
        if (y * up < r0) r0 = y * up;
        if (y * up + up - 1 > r1) r1 = y * up + up - 1;
        if (x * up < c0) c0 = x * up;
        if (x * up + up - 1 > c1) c1 = x * up + up - 1;
      }
  if (r1 < 0) return -1;
  std::vector<int32_t> mark((size_t)H * up * W * up, 0);
  cc_view v{lab, mark.data(), W, up, r0, c0, r1 - r0 + 1, c1 - c0 + 1, id};
  box4[0] = r0; box4[1] = c0; box4[2] = r1 + 1; box4[3] = c1 + 1;
  cc_border b = chunked ? cc_scan_warp(v, 0) : cc_scan(v);
  if (b.npts > cap) return -2;
  if (b.y < 0) return 0;
  int n = cc_trace<false>(v, b.y, b.x, 0, 0, out_xy, 0, 0);
  return n == b.npts ? n : -3;
}

// Whole-table CPU twin of cerb_inst_info's OUTPUT FORMAT for tests/test_instinfo_host.py: the
// contours come from the shared header above, the per-instance sums from plain loops. It lets the
// Python side of the instance tables (cerberus_b200/instinfo.py: row selection, dtypes, offsets,
// type rule decoding) be checked against the OpenCV loop without a GPU. Returns the number of
// instances; arrays are sized by the caller (max label + 1 rows, cap_xy points).
extern "C" int inst_table_host(const int32_t* lab, int H, int W, const float* type, int up,
                               int32_t* ids, int32_t* box, int64_t* mom, int32_t* typ,
                               int64_t* off, int32_t* xy, int64_t cap_xy, int32_t* any_bg) {
  int32_t mx = 0;
  *any_bg = 0;
  for (size_t p = 0; p < (size_t)H * W; ++p) {
    if (lab[p] > mx) mx = lab[p];
    if (lab[p] == 0) *any_bg = 1;
  }
  std::vector<int32_t> mark((size_t)H * up * W * up, 0);
  int n = 0;
  int64_t npts = 0;
  off[0] = 0;
  for (int32_t id = 1; id <= mx; ++id) {
    int r0 = H, r1 = -1, c0 = W, c1 = -1;
    long long cnt = 0, sx = 0, sy = 0;
    long long hist[64] = {0};
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        if (lab[(size_t)y * W + x] == id) {
          if (y < r0) r0 = y;
          if (y > r1) r1 = y;
          if (x < c0) c0 = x;
          if (x > c1) c1 = x;
          ++cnt; sx += x; sy += y;
          if (type) hist[(int)(type[(size_t)y * W + x] * 4.f)]++;
        }
    if (!cnt) continue;
    const long long u = up, cu = cnt * u * u;
    const long long sX = u * u * u * sx + u * u * (u - 1) / 2 * cnt, sY = u * u * u * sy + u * u * (u - 1) / 2 * cnt;
    ids[n] = id;
    box[n * 4 + 0] = r0 * up; box[n * 4 + 1] = c0 * up; box[n * 4 + 2] = (r1 + 1) * up; box[n * 4 + 3] = (c1 + 1) * up;
    mom[n * 3 + 0] = cu; mom[n * 3 + 1] = sX - (long long)c0 * up * cu; mom[n * 3 + 2] = sY - (long long)r0 * up * cu;
    int best = -1; long long best_n = 0;
    if (type) {
      for (int t = 0; t < 64; ++t) if (hist[t] > best_n) { best = t; best_n = hist[t]; }
      if (best == 0) {
        int second = -1; long long second_n = 0;
        for (int t = 1; t < 64; ++t) if (hist[t] > second_n) { second = t; second_n = hist[t]; }
        if (second > 0) { best = second; best_n = second_n; }
      }
    }
    typ[n * 2 + 0] = best; typ[n * 2 + 1] = (int32_t)(best_n * up * up);
    cc_view v{lab, mark.data(), W, up, r0 * up, c0 * up, (r1 - r0 + 1) * up, (c1 - c0 + 1) * up, id};
    cc_border b = cc_scan_warp(v, 0);
    if (b.y >= 0) {
      if (npts + b.npts > cap_xy) return -2;
      cc_trace<false>(v, b.y, b.x, 0, 0, xy + npts * 2, v.c0, v.r0);
      npts += b.npts;
    }
    ++n;
    off[n] = npts;
  }
  return n;
}
