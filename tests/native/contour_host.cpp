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
