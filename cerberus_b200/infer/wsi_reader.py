"""Array-backed slide reader for the WSI path.

The reference opens slides with tiatoolbox `WSIReader.open` (OpenSlide / JP2 / TIFF backends,
infer/wsi.py:522-531); none of those libraries is in this image. What the hot path needs from a
reader is small: the slide dimensions at the processing resolution, the scan resolution, and
zero-padded reads of pixel windows - so slides are accepted as arrays:

  *.npy                    uint8 [H, W, 3] (opened memory-mapped)
  *.png / *.jpg / *.tif    decoded with OpenCV (BGR -> RGB)

A sidecar `<file>.json` may carry {"mpp": <microns per pixel>} (default: the processing
resolution, i.e. no resampling). Resampling between scan and processing resolution is NOT
implemented: a slide whose mpp differs from --wsi_proc_mag is rejected. Pyramidal vendor formats
(.svs, .ndpi, .mrxs, ...) raise: convert them to an array at the processing resolution first.
"""
import json
import os

import numpy as np

ARRAY_EXTS = (".npy", ".png", ".jpg", ".jpeg", ".tif", ".tiff", ".bmp")


class ArraySlide:
    def __init__(self, img, mpp):
        img = np.asarray(img)
        if img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
            raise ValueError("slide array must be uint8 [H, W, 3] (got %r %s)" % (img.shape, img.dtype))
        self.img = img
        self.mpp = float(mpp)

    @classmethod
    def open(cls, path, default_mpp):
        ext = os.path.splitext(path)[1].lower()
        if ext not in ARRAY_EXTS:
            raise NotImplementedError(
                "cerberus_b200 reads array-backed slides only (%s); %r needs OpenSlide / tiatoolbox, "
                "which this build does not ship. Export the slide at the processing resolution to "
                ".npy / .tif first." % (", ".join(ARRAY_EXTS), ext))
        mpp = default_mpp
        side = path + ".json"
        if os.path.exists(side):
            with open(side) as f:
                mpp = float(json.load(f).get("mpp", default_mpp))
        if ext == ".npy":
            img = np.load(path, mmap_mode="r")
        else:
            import cv2
            img = cv2.imread(path, cv2.IMREAD_COLOR)
            if img is None:
                raise IOError("cannot decode %s" % path)
            img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        return cls(img, mpp)

    def slide_dimensions(self, mpp):
        """(width, height) at `mpp` microns per pixel."""
        if abs(mpp - self.mpp) > 1e-6 * max(mpp, self.mpp):
            raise NotImplementedError(
                "slide scanned at %.4f mpp but processing resolution is %.4f mpp: resampling is "
                "not implemented in this build" % (self.mpp, mpp))
        return np.array([self.img.shape[1], self.img.shape[0]])

    def read_bounds(self, bounds):
        """Zero-padded read of [x0, y0, x1, y1] (tiatoolbox read_bounds(..., coord_space=
        "resolution", pad_constant_values=0), used by WSIStreamDataset: infer/wsi.py:936-942)."""
        x0, y0, x1, y1 = [int(v) for v in bounds]
        H, W = self.img.shape[:2]
        out = np.zeros((y1 - y0, x1 - x0, 3), dtype=np.uint8)
        sx0, sy0, sx1, sy1 = max(x0, 0), max(y0, 0), min(x1, W), min(y1, H)
        if sx1 > sx0 and sy1 > sy0:
            out[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = self.img[sy0:sy1, sx0:sx1]
        return out

    def thumbnail(self, power=1.25):
        """Approximate slide_thumbnail(resolution=1.25, units="power"): objective power is taken as
        10 / mpp (0.25 mpp = 40x), area-averaged down."""
        import cv2
        f = power / (10.0 / self.mpp)
        return cv2.resize(np.asarray(self.img), (0, 0), fx=f, fy=f, interpolation=cv2.INTER_AREA)
