/* ORACLE — test infrastructure only; never linked or called by the product path.
 *
 * Plain-C CPU restatement of the reference's instance post-processing
 * (loader/postproc.py:268-407, PostProcInstErodedContourMap) and of the third-party
 * arithmetic it calls:
 *   scipy.ndimage.label (4-connectivity, raster-first-pixel numbering)   -> orc_label4
 *   scipy.ndimage.binary_fill_holes                                      -> orc_fill_holes
 *   cv2.erode / cv2.dilate / cv2.getStructuringElement(MORPH_ELLIPSE)    -> orc_erode_cross,
 *                                                                           orc_dilate, orc_ellipse
 *   skimage.morphology.remove_small_objects (scikit-image 0.19.2, environment.yml:20,
 *     un-vendored)                                                       -> orc_remove_small
 *   skimage.segmentation.watershed (same pin; _watershed_cy.pyx + heap_general.pxi,
 *     restated from the published algorithm)       -> orc_watershed
 * Parity pin: scipy and OpenCV are present in the build image, so tests/test_oracle_postproc.py
 * checks the first three families against them directly, and tests/golden/postproc_*.npz hold
 * outputs of the UNMODIFIED reference post_process code (run with these restated skimage
 * functions injected; scikit-image itself is absent -> the watershed heap order is "parity
 * unpinned" against skimage proper and says so in DESIGN.md).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ labelling */
/* Returns the number of components. lab: 1.. in order of first pixel (raster). */
int orc_label4(const uint8_t* fg, int H, int W, int32_t* lab) {
  const int hw = H * W;
  int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)hw);
  int count = 0;
  memset(lab, 0, sizeof(int32_t) * (size_t)hw);
  for (int p = 0; p < hw; ++p) {
    if (!fg[p] || lab[p]) continue;
    ++count;
    int sp = 0;
    stack[sp++] = p;
    lab[p] = count;
    while (sp) {
      const int q = stack[--sp];
      const int x = q % W, y = q / W;
      if (y > 0 && fg[q - W] && !lab[q - W]) { lab[q - W] = count; stack[sp++] = q - W; }
      if (x > 0 && fg[q - 1] && !lab[q - 1]) { lab[q - 1] = count; stack[sp++] = q - 1; }
      if (x < W - 1 && fg[q + 1] && !lab[q + 1]) { lab[q + 1] = count; stack[sp++] = q + 1; }
      if (y < H - 1 && fg[q + W] && !lab[q + W]) { lab[q + W] = count; stack[sp++] = q + W; }
    }
  }
  free(stack);
  return count;
}

/* skimage remove_small_objects on a label image: zero labels whose pixel count < min_size. */
void orc_remove_small(int32_t* lab, int hw, int min_size) {
  int32_t maxl = 0;
  for (int p = 0; p < hw; ++p) if (lab[p] > maxl) maxl = lab[p];
  int32_t* cnt = (int32_t*)calloc((size_t)maxl + 1, sizeof(int32_t));
  for (int p = 0; p < hw; ++p) cnt[lab[p]]++;
  for (int p = 0; p < hw; ++p) if (lab[p] && cnt[lab[p]] < min_size) lab[p] = 0;
  free(cnt);
}

/* scipy.ndimage.binary_fill_holes: background not 4-connected to the border becomes fg. */
void orc_fill_holes(const uint8_t* fg, int H, int W, uint8_t* out) {
  const int hw = H * W;
  uint8_t* outside = (uint8_t*)calloc((size_t)hw, 1);
  int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)hw);
  int sp = 0;
  for (int p = 0; p < hw; ++p) {
    const int x = p % W, y = p / W;
    if ((x == 0 || y == 0 || x == W - 1 || y == H - 1) && !fg[p] && !outside[p]) {
      outside[p] = 1;
      stack[sp++] = p;
    }
  }
  while (sp) {
    const int q = stack[--sp];
    const int x = q % W, y = q / W;
    if (y > 0 && !fg[q - W] && !outside[q - W]) { outside[q - W] = 1; stack[sp++] = q - W; }
    if (x > 0 && !fg[q - 1] && !outside[q - 1]) { outside[q - 1] = 1; stack[sp++] = q - 1; }
    if (x < W - 1 && !fg[q + 1] && !outside[q + 1]) { outside[q + 1] = 1; stack[sp++] = q + 1; }
    if (y < H - 1 && !fg[q + W] && !outside[q + W]) { outside[q + W] = 1; stack[sp++] = q + W; }
  }
  for (int p = 0; p < hw; ++p) out[p] = !outside[p];
  free(stack);
  free(outside);
}

/* ------------------------------------------------------------------ morphology */
/* cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k)) -> k*k bytes. */
void orc_ellipse(int k, uint8_t* elem) {
  const int r = k / 2, c = k / 2;
  const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
  memset(elem, 0, (size_t)k * k);
  for (int i = 0; i < k; ++i) {
    int j1 = 0, j2 = 0;
    const int dy = i - r;
    if (abs(dy) <= r) {
      const int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));
      j1 = c - dx > 0 ? c - dx : 0;
      j2 = c + dx + 1 < k ? c + dx + 1 : k;
    }
    for (int j = j1; j < j2; ++j) elem[i * k + j] = 1;
  }
}

/* cv2.erode(img, ELLIPSE 3x3 = cross): out-of-image neighbours do not constrain. */
void orc_erode_cross(const uint8_t* in, int H, int W, uint8_t* out) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int p = y * W + x;
      uint8_t v = in[p];
      if (x > 0) v &= in[p - 1];
      if (x < W - 1) v &= in[p + 1];
      if (y > 0) v &= in[p - W];
      if (y < H - 1) v &= in[p + W];
      out[p] = v;
    }
}

/* cv2.dilate(img, elem k x k), default anchor (k/2, k/2), outside pixels ignored:
 * dst(y,x) = max over elem(ky,kx) != 0 of src(y + ky - a, x + kx - a). */
void orc_dilate(const uint8_t* in, int H, int W, const uint8_t* elem, int k, uint8_t* out) {
  const int a = k / 2;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      uint8_t v = 0;
      for (int ky = 0; ky < k && !v; ++ky) {
        const int sy = y + ky - a;
        if (sy < 0 || sy >= H) continue;
        for (int kx = 0; kx < k; ++kx) {
          if (!elem[ky * k + kx]) continue;
          const int sx = x + kx - a;
          if (sx < 0 || sx >= W) continue;
          if (in[sy * W + sx]) { v = 1; break; }
        }
      }
      out[y * W + x] = v;
    }
}

/* ------------------------------------------------------------------ watershed */
typedef struct { double value; int32_t age; int64_t index; int64_t source; } HeapItem;
typedef struct { HeapItem* d; int64_t n, cap; } Heap;

static int smaller(const HeapItem* a, const HeapItem* b) {
  if (a->value != b->value) return a->value < b->value;
  return a->age < b->age;
}
static void heappush(Heap* h, const HeapItem* e) {
  if (h->n == h->cap) { h->cap *= 2; h->d = (HeapItem*)realloc(h->d, sizeof(HeapItem) * (size_t)h->cap); }
  int64_t child = h->n;
  h->d[child] = *e;
  h->n++;
  while (child > 0) {
    const int64_t parent = (child + 1) / 2 - 1;
    if (smaller(&h->d[child], &h->d[parent])) {
      HeapItem t = h->d[child]; h->d[child] = h->d[parent]; h->d[parent] = t;
      child = parent;
    } else break;
  }
}
static void heappop(Heap* h, HeapItem* dst) {
  *dst = h->d[0];
  h->n--;
  if (h->n == 0) return;
  { HeapItem t = h->d[0]; h->d[0] = h->d[h->n]; h->d[h->n] = t; }
  int64_t i = 0, smallest = 0;
  for (;;) {
    const int64_t l = i * 2 + 1, r = i * 2 + 2;
    if (l < h->n) {
      if (smaller(&h->d[l], &h->d[i])) smallest = l;
      if (r < h->n && smaller(&h->d[r], &h->d[smallest])) smallest = r;
    } else break;
    if (smallest == i) break;
    { HeapItem t = h->d[i]; h->d[i] = h->d[smallest]; h->d[smallest] = t; }
    i = smallest;
  }
}

/* skimage.segmentation.watershed(image, markers, mask=mask), connectivity 1, compactness 0,
 * watershed_line False. image: float64 [H,W]; markers: int32; mask: bytes. out: int32 [H,W]. */
void orc_watershed(const double* image, const int32_t* markers, const uint8_t* mask, int H, int W,
                   int32_t* out) {
  const int PW = W + 2, PH = H + 2;
  const int64_t phw = (int64_t)PW * PH;
  double* img = (double*)calloc((size_t)phw, sizeof(double));
  uint8_t* msk = (uint8_t*)calloc((size_t)phw, 1);
  int32_t* o = (int32_t*)calloc((size_t)phw, sizeof(int32_t));
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int64_t q = (int64_t)(y + 1) * PW + x + 1;
      const int p = y * W + x;
      img[q] = image[p];
      msk[q] = mask[p] != 0;
      o[q] = mask[p] ? markers[p] : 0; /* markers * mask */
    }
  Heap h; h.n = 0; h.cap = 1024; h.d = (HeapItem*)malloc(sizeof(HeapItem) * (size_t)h.cap);
  HeapItem e, ne;
  for (int64_t q = 0; q < phw; ++q) {
    if (!o[q]) continue;
    e.value = img[q]; e.age = 0; e.index = q; e.source = q;
    heappush(&h, &e);
  }
  const int64_t nb[4] = {-PW, -1, 1, PW};
  int32_t age = 1;
  while (h.n > 0) {
    heappop(&h, &e);
    for (int i = 0; i < 4; ++i) {
      const int64_t j = e.index + nb[i];
      if (!msk[j]) continue;
      if (o[j]) continue;
      age += 1;
      ne.value = img[j]; ne.age = age; ne.index = j; ne.source = e.source;
      o[j] = o[e.index];
      heappush(&h, &ne);
    }
  }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) out[y * W + x] = o[(int64_t)(y + 1) * PW + x + 1];
  free(h.d); free(o); free(msk); free(img);
}

/* ------------------------------------------------------------------ pipelines */
/* loader/postproc.py:352-381. fg: float32 [H,W,2] with pixel stride `cs` floats (inner at
 * fg[0], contour at fg[1]). Returns 1 and fills out (int32) when the mask is non-empty, else
 * returns 0 and zero-fills (the reference then returns float64 zeros). */
int orc_proc_nuclei(const float* fg, int cs, int H, int W, int32_t* out) {
  const int hw = H * W;
  uint8_t* msk = (uint8_t*)malloc((size_t)hw);
  uint8_t* tmp = (uint8_t*)malloc((size_t)hw);
  int32_t* lab = (int32_t*)malloc(sizeof(int32_t) * (size_t)hw);
  int32_t* mrk = (int32_t*)malloc(sizeof(int32_t) * (size_t)hw);
  double* img = (double*)malloc(sizeof(double) * (size_t)hw);
  long sum = 0;
  for (int p = 0; p < hw; ++p) {
    const float inner = fg[(size_t)p * cs], cnt = fg[(size_t)p * cs + 1];
    const float raw = inner + cnt; /* float32 add, as numpy */
    msk[p] = raw > 0.5f;
    sum += msk[p];
    img[p] = -(double)inner;
  }
  memset(out, 0, sizeof(int32_t) * (size_t)hw);
  if (sum == 0) { free(msk); free(tmp); free(lab); free(mrk); free(img); return 0; }
  orc_erode_cross(msk, H, W, tmp);
  orc_label4(tmp, H, W, lab);
  orc_remove_small(lab, hw, 8);
  for (int p = 0; p < hw; ++p) msk[p] = lab[p] > 0;
  for (int p = 0; p < hw; ++p) tmp[p] = fg[(size_t)p * cs] > 0.5f;
  orc_label4(tmp, H, W, mrk);
  orc_remove_small(mrk, hw, 4);
  for (int p = 0; p < hw; ++p) tmp[p] = mrk[p] != 0;
  uint8_t* filled = (uint8_t*)malloc((size_t)hw);
  orc_fill_holes(tmp, H, W, filled);
  orc_label4(filled, H, W, mrk);
  orc_watershed(img, mrk, msk, H, W, out);
  free(filled); free(msk); free(tmp); free(lab); free(mrk); free(img);
  return 1;
}

/* Shared tail of loader/postproc.py:147-265 (PostProcInstErodedMap) and :270-350
 * (PostProcInstErodedContourMap gland / lumen): `b` is the thresholded foreground. */
static void orc_instances_dilate_fill(uint8_t* b, int H, int W, int k, int min_size, int32_t* out) {
  const int hw = H * W;
  uint8_t* elem = (uint8_t*)malloc((size_t)(k > 0 ? k * k : 1));
  orc_ellipse(k, elem);
  int32_t* lab = (int32_t*)malloc(sizeof(int32_t) * (size_t)hw);
  orc_label4(b, H, W, lab);           /* remove_small_objects on a bool array labels it first */
  orc_remove_small(lab, hw, min_size);
  for (int p = 0; p < hw; ++p) b[p] = lab[p] != 0;
  const int n = orc_label4(b, H, W, lab);
  memset(out, 0, sizeof(int32_t) * (size_t)hw);
  uint8_t* crop = (uint8_t*)malloc((size_t)hw);
  uint8_t* dil = (uint8_t*)malloc((size_t)hw);
  uint8_t* fil = (uint8_t*)malloc((size_t)hw);
  /* Reference quirk (loader/postproc.py:291,332): `np.unique(inst_lab).tolist()[1:]` drops
   * the smallest value assuming it is the background 0; with no background pixel at all the
   * first instance is dropped instead. */
  int has_bg = 0;
  for (int p = 0; p < hw; ++p) if (!lab[p]) { has_bg = 1; break; }
  for (int id = has_bg ? 1 : 2; id <= n; ++id) {
    int y1 = H, y2 = -1, x1 = W, x2 = -1;
    for (int p = 0; p < hw; ++p)
      if (lab[p] == id) {
        const int x = p % W, y = p / W;
        if (y < y1) y1 = y;
        if (y > y2) y2 = y;
        if (x < x1) x1 = x;
        if (x > x2) x2 = x;
      }
    y2 += 1; x2 += 1; /* misc/utils.py:82-91 */
    const int pad = k * 2;
    y1 = y1 - pad >= 0 ? y1 - pad : y1;
    x1 = x1 - pad >= 0 ? x1 - pad : x1;
    x2 = x2 + pad <= W - 1 ? x2 + pad : x2;
    y2 = y2 + pad <= H - 1 ? y2 + pad : y2;
    const int ch = y2 - y1, cw = x2 - x1;
    for (int y = 0; y < ch; ++y)
      for (int x = 0; x < cw; ++x) crop[y * cw + x] = lab[(y1 + y) * W + x1 + x] == id;
    orc_dilate(crop, ch, cw, elem, k, dil);
    orc_fill_holes(dil, ch, cw, fil);
    for (int y = 0; y < ch; ++y)
      for (int x = 0; x < cw; ++x)
        if (fil[y * cw + x]) out[(y1 + y) * W + x1 + x] = id;
  }
  free(crop); free(dil); free(fil); free(lab); free(elem);
}

/* loader/postproc.py:270-309 (tissue 0, gland) / :312-350 (tissue 1, lumen). out: int32
 * (the reference holds the same integers in a float64 map). */
void orc_proc_gland_lumen(const float* fg, int cs, int H, int W, int tissue, double ds, int32_t* out) {
  const int hw = H * W;
  const int ksize_ = tissue == 0 ? 11 : 3;
  const int k = (int)((ksize_ - 1) * ds);
  const int min_size = (int)((tissue == 0 ? 1000 : 150) * (ds * ds));
  const float thr = tissue == 0 ? 0.55f : 0.5f;
  uint8_t* b = (uint8_t*)malloc((size_t)hw);
  for (int p = 0; p < hw; ++p) {
    const float inner = fg[(size_t)p * cs];
    const float cnt = fg[(size_t)p * cs + 1] > 0.5f ? 1.0f : 0.0f;
    const float d = inner - cnt;
    b[p] = d > thr;
  }
  orc_instances_dilate_fill(b, H, W, k, min_size, out);
  free(b);
}

/* loader/postproc.py:147-265 PostProcInstErodedMap: tissue 0 gland (:149-176: ksize 11, min 1500),
 * 1 lumen (:179-206: 3, 150), 2 nuclei (:209-236: 3, 8); foreground = the single channel > 0.5. */
void orc_proc_eroded_map(const float* fg, int cs, int H, int W, int tissue, int32_t* out) {
  const int hw = H * W;
  const int k = tissue == 0 ? 11 : 3;
  const int min_size = tissue == 0 ? 1500 : tissue == 1 ? 150 : 8;
  uint8_t* b = (uint8_t*)malloc((size_t)hw);
  for (int p = 0; p < hw; ++p) b[p] = fg[(size_t)p * cs] > 0.5f;
  orc_instances_dilate_fill(b, H, W, k, min_size, out);
  free(b);
}
