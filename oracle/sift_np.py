"""CPU restatement of cv::SIFT (detect + compute) — TEST INFRASTRUCTURE, not a product path.

Scope: SURVEY.md 8(f) rank 3, `SfM::extractFeatures` (reference SfM.cpp:577-597):
    featureDetector->detect(image, keypoints); descriptorExtractor->compute(image, keypoints, descriptors)
with the detector `cv::SIFT::create(featureLimit, 3, 0.09)` that PhotogrammetrieCli.cpp:342-357 configures
(feature-limit defaults to 10000: retainBest keeps the strongest responses).  The algorithm lives
in a third-party dependency that is absent from /root/reference: OpenCV (the reference links the system OpenCV 4.x;
the build container carries cv2 4.13.0, which is what this restatement is pinned against —
tests/golden/make_golden_sift.py writes cv2's keypoints / descriptors for the reference's own image
images/insel/1.jpg and for a synthetic image, tests/test_sift_oracle.py compares).  What is restated is the published
algorithm of modules/features2d/src/sift.dispatch.cpp + sift.simd.hpp (float pipeline, SIFT_FIXPT_SCALE = 1):

    createInitialImage     2x INTER_LINEAR upsampling, blur with sqrt(sigma^2 - 4 * 0.5^2)
    buildGaussianPyramid   nOctaveLayers + 3 images per octave, incremental sigmas, octave base = image [nOctaveLayers] of
                           the previous octave, every second pixel
    buildDoGPyramid        differences of neighbouring levels
    findScaleSpaceExtrema  26-neighbour extrema above floor(0.5 * contrastThreshold / nOctaveLayers * 255),
                           adjustLocalExtrema (<= 5 Newton steps, contrast and edge tests), calcOrientationHist
                           (36 bins, [1 4 6 4 1] / 16 smoothing, peaks >= 0.8 max, parabolic bin interpolation)
    removeDuplicatedSorted sort by (x, y, -size, angle, -response, -octave), drop equal (x, y, size, angle)
    retainBest             nfeatures > 0: keep responses >= the nfeatures-th largest
    calcSIFTDescriptor     4 x 4 x 8 histogram, trilinear votes, 0.2 clipping, * 512, saturate to u8

Everything is computed in float32 in OpenCV's operation order where that is observable; the remaining differences
to cv2 are last-bit effects of FMA contraction / SIMD summation inside GaussianBlur, exp and atan2, which can flip a
borderline threshold decision: parity with cv2 is therefore stated as a tolerance (see the test), not bit-exact.
"""
import numpy as np

f32 = np.float32

SIFT_DESCR_WIDTH = 4
SIFT_DESCR_HIST_BINS = 8
SIFT_INIT_SIGMA = f32(0.5)
SIFT_IMG_BORDER = 5
SIFT_MAX_INTERP_STEPS = 5
SIFT_ORI_HIST_BINS = 36
SIFT_ORI_SIG_FCTR = f32(1.5)
SIFT_ORI_RADIUS = f32(4.5)
SIFT_ORI_PEAK_RATIO = f32(0.8)
SIFT_DESCR_SCL_FCTR = f32(3.0)
SIFT_DESCR_MAG_THR = f32(0.2)
SIFT_INT_DESCR_FCTR = f32(512.0)
FLT_EPSILON = f32(1.1920929e-07)

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4")])          # cv::KeyPoint without class_id (always -1 here)


# ---------------------------------------------------------------------------------------------- hal restatements
_P1 = f32(0.9997878412794807) * f32(180 / np.pi)
_P3 = f32(-0.3258083974640975) * f32(180 / np.pi)
_P5 = f32(0.1555786518463281) * f32(180 / np.pi)
_P7 = f32(-0.04432655554792128) * f32(180 / np.pi)


def fast_atan2_deg(y, x):
    """cv::hal::fastAtan2(..., angleInDegrees = true): odd 7th-order polynomial on the smaller / larger ratio."""
    y = np.asarray(y, f32)
    x = np.asarray(x, f32)
    ax, ay = np.abs(x), np.abs(y)
    big = ax >= ay
    c = (np.where(big, ay, ax) / (np.where(big, ax, ay) + f32(2.220446049250313e-16))).astype(f32)
    c2 = c * c
    a = (((_P7 * c2 + _P5) * c2 + _P3) * c2 + _P1) * c
    a = np.where(big, a, f32(90) - a)
    a = np.where(x < 0, f32(180) - a, a)
    a = np.where(y < 0, f32(360) - a, a)
    return a.astype(f32)


def cv_round(v):
    """cvRound: round half to even (lrint / cvtss2si)."""
    return int(np.rint(v))


def bgr_to_gray(img):
    """cvtColor(COLOR_BGR2GRAY) on 8-bit data, what cv::SIFT applies to a colour image first (color_rgb.simd.hpp RGB2Gray<uchar>:
    BT.601 weights in 15-bit fixed point)."""
    b, g, r = (np.asarray(img)[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


# ---------------------------------------------------------------------------------------------- image pyramid
def gaussian_kernel(sigma):
    """cv::GaussianBlur(src, dst, Size(), sigma) on CV_32F: ksize = cvRound(sigma * 4 * 2 + 1) | 1, normalised exp kernel."""
    ksize = cv_round(float(sigma) * 8 + 1) | 1
    r = (ksize - 1) // 2
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-(x * x) / (2.0 * float(sigma) * float(sigma)))
    return (k / k.sum()).astype(f32)


def _reflect101(n, r):
    """cv::borderInterpolate(p, n, BORDER_REFLECT_101) for p in [-r, n + r)."""
    idx = np.arange(-r, n + r)
    if n == 1:
        return np.zeros_like(idx)
    while True:
        bad = (idx < 0) | (idx >= n)
        if not bad.any():
            return idx
        idx = np.where(idx < 0, -idx, np.where(idx >= n, 2 * (n - 1) - idx, idx))


def gaussian_blur(img, sigma):
    """Separable blur, BORDER_REFLECT_101, rows first then columns, float32 accumulation."""
    k = gaussian_kernel(sigma)
    r = len(k) // 2
    h, w = img.shape
    src = img[:, _reflect101(w, r)]
    tmp = np.zeros((h, w), f32)
    for t in range(len(k)):
        tmp += k[t] * src[:, t:t + w]
    src = tmp[_reflect101(h, r), :]
    out = np.zeros((h, w), f32)
    for t in range(len(k)):
        out += k[t] * src[t:t + h, :]
    return out


def resize_linear_2x(img):
    """cv::resize(src, dst, Size(2w, 2h), 0, 0, INTER_LINEAR) on CV_32F: sample at (d + 0.5) / 2 - 0.5, clamped."""
    def taps(n):
        d = np.arange(2 * n)
        s = (d + 0.5) * 0.5 - 0.5
        i0 = np.floor(s).astype(np.int64)
        fr = (s - i0).astype(f32)
        lo = i0 < 0
        i0[lo] = 0
        fr[lo] = 0
        hi = i0 >= n - 1
        i0[hi] = n - 1
        fr[hi] = 0
        return i0, np.minimum(i0 + 1, n - 1), fr
    h, w = img.shape
    x0, x1, fx = taps(w)
    y0, y1, fy = taps(h)
    rows = img[:, x0] * (f32(1) - fx)[None, :] + img[:, x1] * fx[None, :]
    return (rows[y0, :] * (f32(1) - fy)[:, None] + rows[y1, :] * fy[:, None]).astype(f32)


def create_initial_image(gray, sigma=1.6, double_size=True):
    """sift.dispatch.cpp createInitialImage (enable_precise_upscale = false)."""
    g = np.asarray(gray).astype(f32)
    sigma = f32(sigma)
    if double_size:
        sig_diff = np.sqrt(max(sigma * sigma - SIFT_INIT_SIGMA * SIFT_INIT_SIGMA * f32(4), f32(0.01)), dtype=f32)
        return gaussian_blur(resize_linear_2x(g), sig_diff)
    sig_diff = np.sqrt(max(sigma * sigma - SIFT_INIT_SIGMA * SIFT_INIT_SIGMA, f32(0.01)), dtype=f32)
    return gaussian_blur(g, sig_diff)


def n_octaves_for(base_shape, first_octave=-1):
    return cv_round(np.log(float(min(base_shape))) / np.log(2.0) - 2) - first_octave


def level_sigmas(n_layers=3, sigma=1.6):
    sig = [float(sigma)]
    k = 2.0 ** (1.0 / n_layers)
    for i in range(1, n_layers + 3):
        prev = k ** (i - 1) * sigma
        total = prev * k
        sig.append(np.sqrt(total * total - prev * prev))
    return sig


def build_gaussian_pyramid(base, n_octaves, n_layers=3, sigma=1.6):
    sig = level_sigmas(n_layers, sigma)
    pyr = []
    for o in range(n_octaves):
        for i in range(n_layers + 3):
            if o == 0 and i == 0:
                pyr.append(base)
            elif i == 0:
                src = pyr[(o - 1) * (n_layers + 3) + n_layers]
                pyr.append(np.ascontiguousarray(src[0:2 * (src.shape[0] // 2):2, 0:2 * (src.shape[1] // 2):2]))   # INTER_NEAREST, half size
            else:
                pyr.append(gaussian_blur(pyr[-1], sig[i]))
    return pyr


def build_dog_pyramid(gpyr, n_octaves, n_layers=3):
    dog = []
    for o in range(n_octaves):
        for i in range(n_layers + 2):
            a = gpyr[o * (n_layers + 3) + i]
            dog.append((gpyr[o * (n_layers + 3) + i + 1] - a).astype(f32))
    return dog


# ---------------------------------------------------------------------------------------------- keypoints
def _solve3(H, b):
    """Matx33f::solve(b, DECOMP_LU) = closed-form Cramer rule in float32; singular -> zeros."""
    a = H
    d = (a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[0][1] * (a[1][0] * a[2][2] - a[2][0] * a[1][2]) +
         a[0][2] * (a[1][0] * a[2][1] - a[2][0] * a[1][1]))
    if d == 0:
        return f32(0), f32(0), f32(0)
    d = f32(1) / d
    x0 = d * (b[0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (b[1] * a[2][2] - a[1][2] * b[2]) +
              a[0][2] * (b[1] * a[2][1] - a[1][1] * b[2]))
    x1 = d * (a[0][0] * (b[1] * a[2][2] - a[1][2] * b[2]) - b[0] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
              a[0][2] * (a[1][0] * b[2] - b[1] * a[2][0]))
    x2 = d * (a[0][0] * (a[1][1] * b[2] - b[1] * a[2][1]) - a[0][1] * (a[1][0] * b[2] - b[1] * a[2][0]) +
              b[0] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]))
    return x0, x1, x2


def adjust_local_extrema(dog, octv, layer, r, c, n_layers, contrast_threshold, edge_threshold, sigma):
    """sift.simd.hpp adjustLocalExtrema; returns None or (keypoint fields, layer, r, c)."""
    img_scale = f32(1.0) / f32(255)
    deriv_scale = img_scale * f32(0.5)
    second_deriv_scale = img_scale
    cross_deriv_scale = img_scale * f32(0.25)
    contrast_threshold = f32(contrast_threshold)
    edge_threshold = f32(edge_threshold)
    xi = xr = xc = f32(0)
    i = 0
    while i < SIFT_MAX_INTERP_STEPS:
        idx = octv * (n_layers + 2) + layer
        img, prv, nxt = dog[idx], dog[idx - 1], dog[idx + 1]
        dD = ((img[r, c + 1] - img[r, c - 1]) * deriv_scale, (img[r + 1, c] - img[r - 1, c]) * deriv_scale,
              (nxt[r, c] - prv[r, c]) * deriv_scale)
        v2 = img[r, c] * f32(2)
        dxx = (img[r, c + 1] + img[r, c - 1] - v2) * second_deriv_scale
        dyy = (img[r + 1, c] + img[r - 1, c] - v2) * second_deriv_scale
        dss = (nxt[r, c] + prv[r, c] - v2) * second_deriv_scale
        dxy = (img[r + 1, c + 1] - img[r + 1, c - 1] - img[r - 1, c + 1] + img[r - 1, c - 1]) * cross_deriv_scale
        dxs = (nxt[r, c + 1] - nxt[r, c - 1] - prv[r, c + 1] + prv[r, c - 1]) * cross_deriv_scale
        dys = (nxt[r + 1, c] - nxt[r - 1, c] - prv[r + 1, c] + prv[r - 1, c]) * cross_deriv_scale
        X = _solve3(((dxx, dxy, dxs), (dxy, dyy, dys), (dxs, dys, dss)), dD)
        xi, xr, xc = -X[2], -X[1], -X[0]
        if abs(xi) < 0.5 and abs(xr) < 0.5 and abs(xc) < 0.5:
            break
        lim = f32(2147483647 // 3)
        if abs(xi) > lim or abs(xr) > lim or abs(xc) > lim or not (np.isfinite(xi) and np.isfinite(xr) and np.isfinite(xc)):
            return None
        c += cv_round(xc)
        r += cv_round(xr)
        layer += cv_round(xi)
        if (layer < 1 or layer > n_layers or c < SIFT_IMG_BORDER or c >= img.shape[1] - SIFT_IMG_BORDER or
                r < SIFT_IMG_BORDER or r >= img.shape[0] - SIFT_IMG_BORDER):
            return None
        i += 1
    if i >= SIFT_MAX_INTERP_STEPS:
        return None
    idx = octv * (n_layers + 2) + layer
    img, prv, nxt = dog[idx], dog[idx - 1], dog[idx + 1]
    dD = ((img[r, c + 1] - img[r, c - 1]) * deriv_scale, (img[r + 1, c] - img[r - 1, c]) * deriv_scale,
          (nxt[r, c] - prv[r, c]) * deriv_scale)
    t = dD[0] * xc + dD[1] * xr + dD[2] * xi
    contr = img[r, c] * img_scale + t * f32(0.5)
    if abs(contr) * f32(n_layers) < contrast_threshold:
        return None
    v2 = img[r, c] * f32(2)
    dxx = (img[r, c + 1] + img[r, c - 1] - v2) * second_deriv_scale
    dyy = (img[r + 1, c] + img[r - 1, c] - v2) * second_deriv_scale
    dxy = (img[r + 1, c + 1] - img[r + 1, c - 1] - img[r - 1, c + 1] + img[r - 1, c - 1]) * cross_deriv_scale
    tr = dxx + dyy
    det = dxx * dyy - dxy * dxy
    if det <= 0 or tr * tr * edge_threshold >= (edge_threshold + f32(1)) * (edge_threshold + f32(1)) * det:
        return None
    scale = f32(1 << octv)
    kx = (f32(c) + xc) * scale
    ky = (f32(r) + xr) * scale
    octave = octv + (layer << 8) + (cv_round((float(xi) + 0.5) * 255) << 16)
    size = f32(sigma) * np.power(f32(2), (f32(layer) + xi) / f32(n_layers), dtype=f32) * scale * f32(2)
    return (kx, ky, f32(size), abs(contr), octave), layer, r, c


def orientation_hist(img, px, py, radius, sigma, n=SIFT_ORI_HIST_BINS):
    """sift.simd.hpp calcOrientationHist -> (smoothed histogram, its maximum)."""
    h, w = img.shape
    expf_scale = f32(-1.0) / (f32(2.0) * sigma * sigma)
    ii, jj = np.meshgrid(np.arange(-radius, radius + 1), np.arange(-radius, radius + 1), indexing="ij")
    y, x = py + ii, px + jj
    ok = (y > 0) & (y < h - 1) & (x > 0) & (x < w - 1)
    y, x, ii, jj = y[ok], x[ok], ii[ok], jj[ok]              # row-major order = OpenCV's k order
    dx = img[y, x + 1] - img[y, x - 1]
    dy = img[y - 1, x] - img[y + 1, x]
    wgt = np.exp(((ii * ii + jj * jj).astype(f32) * expf_scale).astype(f32)).astype(f32)
    ori = fast_atan2_deg(dy, dx)
    mag = np.sqrt(dx * dx + dy * dy).astype(f32)
    bins = np.rint(f32(n / 360.0) * ori).astype(np.int64)
    bins = np.where(bins >= n, bins - n, bins)
    bins = np.where(bins < 0, bins + n, bins)
    temphist = np.zeros(n, f32)
    np.add.at(temphist, bins, (wgt * mag).astype(f32))       # sequential float32 accumulation in k order
    t = np.concatenate([temphist[-2:], temphist, temphist[:2]])
    hist = ((t[0:n] + t[4:n + 4]) * f32(1 / 16) + (t[1:n + 1] + t[3:n + 3]) * f32(4 / 16) + t[2:n + 2] * f32(6 / 16)).astype(f32)
    return hist, hist.max() if len(hist) else f32(0)


def find_scale_space_extrema(gpyr, dog, n_octaves, n_layers=3, contrast_threshold=0.04, edge_threshold=10.0, sigma=1.6):
    """sift.dispatch.cpp findScaleSpaceExtrema + sift.simd.hpp findScaleSpaceExtremaT::process (all orientations)."""
    threshold = int(np.floor(0.5 * contrast_threshold / n_layers * 255))
    n = SIFT_ORI_HIST_BINS
    out = []
    B = SIFT_IMG_BORDER
    for o in range(n_octaves):
        for i in range(1, n_layers + 1):
            idx = o * (n_layers + 2) + i
            prv, cur, nxt = dog[idx - 1], dog[idx], dog[idx + 1]
            rows, cols = cur.shape
            if rows <= 2 * B or cols <= 2 * B:
                continue
            val = cur[B:rows - B, B:cols - B]
            nmax = np.full(val.shape, -np.inf, f32)
            nmin = np.full(val.shape, np.inf, f32)
            for im in (prv, cur, nxt):
                for dr in (-1, 0, 1):
                    for dc in (-1, 0, 1):
                        if im is cur and dr == 0 and dc == 0:
                            continue
                        nb = im[B + dr:rows - B + dr, B + dc:cols - B + dc]
                        np.maximum(nmax, nb, out=nmax)
                        np.minimum(nmin, nb, out=nmin)
            cand = (np.abs(val) > threshold) & (((val > 0) & (val >= nmax)) | ((val < 0) & (val <= nmin)))
            rr, cc = np.nonzero(cand)                           # row-major: the order of the serial OpenCV loop
            for r, c in zip(rr + B, cc + B):
                res = adjust_local_extrema(dog, o, i, int(r), int(c), n_layers, contrast_threshold, edge_threshold, sigma)
                if res is None:
                    continue
                (kx, ky, size, resp, octave), layer, r1, c1 = res
                scl_octv = size * f32(0.5) / f32(1 << o)
                hist, omax = orientation_hist(gpyr[o * (n_layers + 3) + layer], c1, r1, cv_round(SIFT_ORI_RADIUS * scl_octv),
                                              SIFT_ORI_SIG_FCTR * scl_octv, n)
                mag_thr = f32(omax * SIFT_ORI_PEAK_RATIO)
                for j in range(n):
                    l = j - 1 if j > 0 else n - 1
                    r2 = j + 1 if j < n - 1 else 0
                    if hist[j] > hist[l] and hist[j] > hist[r2] and hist[j] >= mag_thr:
                        b = f32(j) + f32(0.5) * (hist[l] - hist[r2]) / (hist[l] - f32(2) * hist[j] + hist[r2])
                        b = f32(n) + b if b < 0 else (b - f32(n) if b >= n else b)
                        angle = f32(360) - f32(f32(360.0 / n) * b)
                        if abs(angle - f32(360)) < FLT_EPSILON:
                            angle = f32(0)
                        out.append((kx, ky, size, angle, resp, octave))
    return np.array(out, dtype=KEYPOINT_DTYPE)


def remove_duplicated_sorted(kps):
    """KeyPointsFilter::removeDuplicatedSorted (keypoint.cpp): sort, then keep the first of equal (x, y, size, angle)."""
    if len(kps) < 2:
        return kps
    order = np.lexsort((-kps["octave"], -kps["response"], kps["angle"], -kps["size"], kps["y"], kps["x"]))
    k = kps[order]
    same = (k["x"][1:] == k["x"][:-1]) & (k["y"][1:] == k["y"][:-1]) & (k["size"][1:] == k["size"][:-1]) & \
           (k["angle"][1:] == k["angle"][:-1])
    return k[np.concatenate([[True], ~same])]


def retain_best(kps, n_points):
    """KeyPointsFilter::retainBest (keypoint.cpp): keep every keypoint whose response is >= the n_points-th largest response
    (ties at the boundary stay).  OpenCV leaves the survivors in the order std::nth_element / std::partition produce; the
    SET is what is defined, and it is returned here in the order of the input list."""
    if n_points <= 0 or len(kps) <= n_points:
        return kps
    boundary = np.sort(kps["response"])[::-1][n_points - 1]
    return kps[kps["response"] >= boundary]


def detect(gray, n_layers=3, contrast_threshold=0.04, edge_threshold=10.0, sigma=1.6, return_pyramid=False, nfeatures=0):
    """cv::SIFT::detect = detectAndCompute without descriptors (sift.dispatch.cpp SIFT_Impl::detectAndCompute, firstOctave = -1)."""
    base = create_initial_image(gray, sigma, True)
    n_oct = n_octaves_for(base.shape, -1)
    gpyr = build_gaussian_pyramid(base, n_oct, n_layers, sigma)
    dog = build_dog_pyramid(gpyr, n_oct, n_layers)
    kps = find_scale_space_extrema(gpyr, dog, n_oct, n_layers, contrast_threshold, edge_threshold, sigma)
    kps = remove_duplicated_sorted(kps)
    kps = retain_best(kps, nfeatures)
    # firstOctave < 0: back to the coordinates of the input image
    oc = kps["octave"]
    kps["octave"] = (oc & ~255) | ((oc - 1) & 255)
    kps["x"] *= f32(0.5)
    kps["y"] *= f32(0.5)
    kps["size"] *= f32(0.5)
    return (kps, gpyr) if return_pyramid else kps


def unpack_octave(octave_field):
    octave = octave_field & 255
    layer = (octave_field >> 8) & 255
    if octave >= 128:
        octave |= -128
    scale = f32(1.0) / f32(1 << octave) if octave >= 0 else f32(1 << -octave)
    return octave, layer, scale


def sift_descriptor(img, ptx, pty, ori, scl, d=SIFT_DESCR_WIDTH, n=SIFT_DESCR_HIST_BINS):
    """sift.simd.hpp calcSIFTDescriptor -> 128 integer-valued float32 (the CV_32F descriptor row of cv2 >= 4.x)."""
    rows, cols = img.shape
    px, py = cv_round(ptx), cv_round(pty)
    cos_t = np.cos(ori * f32(np.pi / 180), dtype=f32)
    sin_t = np.sin(ori * f32(np.pi / 180), dtype=f32)
    bins_per_rad = f32(n / 360.0)
    exp_scale = f32(-1.0) / f32(d * d * 0.5)
    hist_width = SIFT_DESCR_SCL_FCTR * scl
    radius = cv_round(hist_width * f32(1.4142135623730951) * f32(d + 1) * f32(0.5))
    radius = min(radius, int(np.sqrt(float(cols) * cols + float(rows) * rows)))
    cos_t = cos_t / hist_width
    sin_t = sin_t / hist_width
    ii, jj = np.meshgrid(np.arange(-radius, radius + 1), np.arange(-radius, radius + 1), indexing="ij")
    fi, fj = ii.astype(f32), jj.astype(f32)
    c_rot = fj * cos_t - fi * sin_t
    r_rot = fj * sin_t + fi * cos_t
    rbin = r_rot + f32(d // 2) - f32(0.5)
    cbin = c_rot + f32(d // 2) - f32(0.5)
    r, c = py + ii, px + jj
    ok = (rbin > -1) & (rbin < d) & (cbin > -1) & (cbin < d) & (r > 0) & (r < rows - 1) & (c > 0) & (c < cols - 1)
    r, c, rbin, cbin, c_rot, r_rot = r[ok], c[ok], rbin[ok], cbin[ok], c_rot[ok], r_rot[ok]
    dx = img[r, c + 1] - img[r, c - 1]
    dy = img[r - 1, c] - img[r + 1, c]
    w = np.exp(((c_rot * c_rot + r_rot * r_rot) * exp_scale).astype(f32)).astype(f32)
    orient = fast_atan2_deg(dy, dx)
    mag = np.sqrt(dx * dx + dy * dy).astype(f32) * w
    obin = (orient - ori) * bins_per_rad
    r0 = np.floor(rbin).astype(np.int64)
    c0 = np.floor(cbin).astype(np.int64)
    o0 = np.floor(obin).astype(np.int64)
    rbin = rbin - r0.astype(f32)
    cbin = cbin - c0.astype(f32)
    obin = obin - o0.astype(f32)
    o0 = np.where(o0 < 0, o0 + n, o0)
    o0 = np.where(o0 >= n, o0 - n, o0)
    v_r1 = mag * rbin
    v_r0 = mag - v_r1
    v_rc11 = v_r1 * cbin
    v_rc10 = v_r1 - v_rc11
    v_rc01 = v_r0 * cbin
    v_rc00 = v_r0 - v_rc01
    v111 = v_rc11 * obin
    v110 = v_rc11 - v111
    v101 = v_rc10 * obin
    v100 = v_rc10 - v101
    v011 = v_rc01 * obin
    v010 = v_rc01 - v011
    v001 = v_rc00 * obin
    v000 = v_rc00 - v001
    hist = np.zeros((d + 2) * (d + 2) * (n + 2), f32)
    idx = ((r0 + 1) * (d + 2) + c0 + 1) * (n + 2) + o0
    # the eight votes of sample k are added before sample k + 1 is looked at: interleave to keep that order
    offs = np.array([0, 1, n + 2, n + 3, (d + 2) * (n + 2), (d + 2) * (n + 2) + 1, (d + 3) * (n + 2), (d + 3) * (n + 2) + 1])
    votes = np.stack([v000, v001, v010, v011, v100, v101, v110, v111], axis=1).astype(f32)
    np.add.at(hist, (idx[:, None] + offs[None, :]).ravel(), votes.ravel())
    h3 = hist.reshape(d + 2, d + 2, n + 2)
    h3[:, :, 0] += h3[:, :, n]
    h3[:, :, 1] += h3[:, :, n + 1]
    raw = h3[1:d + 1, 1:d + 1, :n].reshape(-1).copy()
    nrm2 = f32(0)
    for v in raw:
        nrm2 += v * v
    thr = np.sqrt(nrm2, dtype=f32) * SIFT_DESCR_MAG_THR
    raw = np.minimum(raw, thr)
    nrm2 = f32(0)
    for v in raw:
        nrm2 += v * v
    scale = SIFT_INT_DESCR_FCTR / max(np.sqrt(nrm2, dtype=f32), FLT_EPSILON)
    return np.clip(np.rint(raw * scale), 0, 255).astype(f32)            # saturate_cast<uchar>: round half to even, clamp


def compute(gray, kps, n_layers=3, sigma=1.6, gpyr=None):
    """cv::SIFT::compute = detectAndCompute(useProvidedKeypoints = true): pyramid from the octave range of the keypoints,
    then calcDescriptors.  (Feature2D::compute's border / size filters never remove a SIFT keypoint of `detect`.)"""
    if len(kps) == 0:
        return np.zeros((0, 128), f32)
    octs = [unpack_octave(int(o))[0] for o in kps["octave"]]
    first_octave = min(min(octs), 0)
    assert first_octave >= -1
    n_oct = max(octs) - first_octave + 1
    if gpyr is None:
        base = create_initial_image(gray, sigma, first_octave < 0)
        gpyr = build_gaussian_pyramid(base, n_oct, n_layers, sigma)
    out = np.zeros((len(kps), 128), f32)
    for i, kp in enumerate(kps):
        octave, layer, scale = unpack_octave(int(kp["octave"]))
        size = kp["size"] * scale
        img = gpyr[(octave - first_octave) * (n_layers + 3) + layer]
        angle = f32(360) - kp["angle"]
        if abs(angle - f32(360)) < FLT_EPSILON:
            angle = f32(0)
        out[i] = sift_descriptor(img, kp["x"] * scale, kp["y"] * scale, angle, size * f32(0.5))
    return out


def detect_and_compute(gray, n_layers=3, contrast_threshold=0.04, edge_threshold=10.0, sigma=1.6, nfeatures=0):
    """detect() then compute() as SfM::extractFeatures calls them (SfM.cpp:584-588)."""
    kps = detect(gray, n_layers, contrast_threshold, edge_threshold, sigma, nfeatures=nfeatures)
    return kps, compute(gray, kps, n_layers, sigma)
