"""CPU restatement of cv::ORB::compute on provided level-0 keypoints (first half).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Reference call site: DisparityUtil.cpp:107 (`ORB::create()`,
all defaults) and :127-134 (`f2d->compute(img, keypoints, descriptors)` with a KeyPoint at every pixel).  The
arithmetic lives in OpenCV features2d `orb.cpp` (un-vendored dependency, 4.5.x in the reference, 4.13 in this image):

  detectAndCompute(image, noArray(), keypoints, descriptors, useProvidedKeypoints=true)
    1. non-gray input -> cvtColor(BGR2GRAY): (B*3735 + G*19235 + R*9798 + 2^14) >> 15
    2. level 0 of the pyramid = the image with a reflect-101 border (keypoints with octave 0 use no other level)
    3. KeyPointsFilter::runByImageBorder(kp, size, edgeThreshold=31): keep 31 <= cvRound(x) < w-31, same for y
    4. GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) -- on the pyramid SUB-matrix, which OpenCV does not send
       to its fixed-point 8-bit kernel but to the floating-point one (IPP in the stock packages): the result is the
       Gaussian-weighted sum rounded to nearest.  Restated here in double precision (kernel = getGaussianKernel(7, 2)).
       Pinned against cv2: 0 differing descriptor bits on the committed fixtures; the float variants one can write
       (float32/float64 accumulation, either tap order) differ from each other in ~1e-6 of the pixels (sums that
       land within float rounding of x.5), so an exact match of those pixels with IPP's internal order is not claimed.
    5. computeOrbDescriptors: angle (degrees, as given: -1 for a default KeyPoint) -> radians in float,
       a = (float)cos, b = (float)sin; for each of the 512 pattern points  x = px*a - py*b, y = px*b + py*a  in float
       (products and sum rounded separately), sample blurred[cy + cvRound(y), cx + cvRound(x)] around
       (cx, cy) = (cvRound(pt.x), cvRound(pt.y)); bit k of byte i = sample[16i + 2k] < sample[16i + 2k + 1].
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATTERN = np.load(os.path.join(_HERE, "orb_pattern.npy")).astype(np.int32)   # (512, 2): x, y
EDGE_THRESHOLD = 31


def gray_from_bgr(img):
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def gaussian_kernel_7_2():
    """cv::getGaussianKernel(7, 2.0, CV_64F): exp(-x^2 / (2 sigma^2)) normalised to sum 1."""
    x = np.arange(7, dtype=np.float64) - 3.0
    k = np.exp(-(x * x) / (2.0 * 2.0 * 2.0))
    return k / k.sum()


def gaussian_blur_7(gray):
    k = gaussian_kernel_7_2()
    h, w = gray.shape
    xp = np.pad(gray.astype(np.float64), 3, mode="reflect")      # numpy 'reflect' == BORDER_REFLECT_101
    rows = np.zeros((h + 6, w), np.float64)
    for i in range(7):
        rows += xp[:, i:i + w] * k[i]
    out = np.zeros((h, w), np.float64)
    for i in range(7):
        out += rows[i:i + h, :] * k[i]
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def orb_compute(image, pts, angles):
    """-> (kept int32[n], desc uint8[n, 32]).  image: HxW (gray) or HxWx3 (BGR) uint8; pts float32 (N, 2) = (x, y);
    angles float32 (N,) in degrees."""
    image = np.asarray(image)
    gray = gray_from_bgr(image) if image.ndim == 3 else image
    h, w = gray.shape
    pts = np.asarray(pts, np.float32).reshape(-1, 2)
    ang = np.asarray(angles, np.float32).reshape(-1)
    cx = np.rint(pts[:, 0]).astype(np.int64)                   # cvRound: half to even
    cy = np.rint(pts[:, 1]).astype(np.int64)
    keep = (cx >= EDGE_THRESHOLD) & (cx < w - EDGE_THRESHOLD) & (cy >= EDGE_THRESHOLD) & (cy < h - EDGE_THRESHOLD)
    kept = np.nonzero(keep)[0].astype(np.int32)
    if len(kept) == 0:
        return kept, np.zeros((0, 32), np.uint8)
    blurred = gaussian_blur_7(gray)
    f = np.float32
    rad = ang[kept] * f(np.pi / f(180.0))                       # float * (float)(CV_PI/180.f)
    a = np.cos(rad.astype(np.float64)).astype(f)[:, None]
    b = np.sin(rad.astype(np.float64)).astype(f)[:, None]
    px = PATTERN[:, 0].astype(f)[None, :]
    py = PATTERN[:, 1].astype(f)[None, :]
    x = (px * a).astype(f) - (py * b).astype(f)
    y = (px * b).astype(f) + (py * a).astype(f)
    ix = np.rint(x).astype(np.int64)
    iy = np.rint(y).astype(np.int64)
    vals = blurred[cy[kept][:, None] + iy, cx[kept][:, None] + ix].astype(np.int32)
    bits = (vals[:, 0::2] < vals[:, 1::2]).astype(np.uint8)      # (n, 256)
    desc = np.packbits(bits.reshape(len(kept), 32, 8)[:, :, ::-1], axis=2).reshape(len(kept), 32)
    return kept, desc
