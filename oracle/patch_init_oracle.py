"""CPU restatement (numpy, float64) of the per-patch 2-dof candidate cost of the reference's Optuna initialiser -- TEST
INFRASTRUCTURE ONLY (only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import anything under oracle/).

The reference evaluates, for every patch of a pyramid level and every TPE trial, ONE translation candidate on the events
cropped to the patch, on its numpy path (src/solver/patch_contrast_pyramid.py:320-415):

  crop + origin shift        src/utils/event_utils.py:50-88  (x0 <= x < x1, y0 <= y < y1; x is the row axis)
  theta *= t_max - t_min     src/solver/patch_contrast_pyramid.py:366-371 (objective_initial, normalize_t_in_batch)
  2-dof warp, 'middle'       src/warp.py:201-259 (reference time = t_min + 0.5 period; dt /= max(dt) - min(dt)),
                             src/warp.py:483-522 (x' = x + dt theta_0, y' = y + dt theta_1)
  bilinear vote, numpy       src/event_image_converter.py:257-312 (floor(x + 1e-8), per-corner in-image masks, padding offset)
  scipy gaussian_filter      src/event_image_converter.py:122-124 -- third-party (scipy 1.x `ndimage.gaussian_filter`): separable,
                             radius int(4 sigma + 0.5), weights exp(-x^2 / (2 sigma^2)) normalised, axis 0 then axis 1,
                             boundary 'reflect' (d c b a | a b c d | d c b a)
  cv2.Sobel(ksize=3) / 8     src/costs/gradient_magnitude.py:78-95 -- third-party (OpenCV 4.x): [1 2 1]^T x [-1 0 1] and its
                             transpose, boundary BORDER_REFLECT_101 (g f e d c b | a b c d e f g h | g f e d c b a)
  cost                       src/costs/normalized_gradient_magnitude.py:81-94 with direction 'minimize', omit_boundary False:
                             mean(|grad IWE(events)|^2) / mean(|grad IWE(warped)|^2); NaN -> 0.0 (pyramid.py:374-375, :411-415)

Pinned by tests/golden/reference_patch_init.npz (made by tests/golden/make_golden_patch_init.py from the UNMODIFIED reference
method `calculate_cost_for_small_patch` / `objective_initial`, i.e. through the real scipy and cv2) in tests/test_patch_init.py.
"""
from __future__ import annotations

import numpy as np


def crop_to_patch(events: np.ndarray, x_min: int, x_max: int, y_min: int, y_max: int) -> np.ndarray:
    """Events of one patch in patch-local coordinates (event_utils.py:50-88)."""
    e = np.asarray(events, dtype=np.float64)
    keep = (x_min <= e[:, 0]) & (e[:, 0] < x_max) & (y_min <= e[:, 1]) & (e[:, 1] < y_max)
    return e[keep] - np.array([x_min, y_min, 0.0, 0.0])


def warp_2dof_middle(events: np.ndarray, theta: np.ndarray, normalize_t: bool = True) -> np.ndarray:
    t = events[:, 2]
    t_ref = t.min() + (t.max() - t.min()) * 0.5
    dt = t - t_ref
    if normalize_t:
        dt = dt / (dt.max() - dt.min())
    return np.stack([events[:, 0] + dt * theta[0], events[:, 1] + dt * theta[1], dt, events[:, 3]], axis=1)


def bilinear_vote(events: np.ndarray, image_size, pad=(0, 0)) -> np.ndarray:
    ph, pw = pad
    h, w = image_size[0] + 2 * ph, image_size[1] + 2 * pw
    img = np.zeros(h * w, dtype=np.float64)
    fl = np.floor(events[:, :2] + 1e-8)
    fr = events[:, :2] - fl
    col, row = fl[:, 1] + pw, fl[:, 0] + ph
    for d_row, d_col, wt in ((0, 0, (1 - fr[:, 0]) * (1 - fr[:, 1])), (1, 0, fr[:, 0] * (1 - fr[:, 1])),
                             (0, 1, (1 - fr[:, 0]) * fr[:, 1]), (1, 1, fr[:, 0] * fr[:, 1])):
        r, c = row + d_row, col + d_col
        ok = (0 <= c) & (c < w) & (0 <= r) & (r < h)
        np.add.at(img, (c[ok] + r[ok] * w).astype(np.int64), wt[ok])
    return img.reshape(h, w)


def _reflect(i: np.ndarray, n: int) -> np.ndarray:
    """scipy 'reflect': the mirror axis sits on the pixel EDGE (-1 -> 0, n -> n-1); repeated for very small n."""
    i = np.mod(i, 2 * n)
    return np.where(i >= n, 2 * n - 1 - i, i)


def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    """OpenCV BORDER_REFLECT_101: the mirror axis sits on the border PIXEL (-1 -> 1, n -> n-2)."""
    if n == 1:
        return np.zeros_like(i)
    i = np.mod(i, 2 * n - 2)
    return np.where(i >= n, 2 * n - 2 - i, i)


def gaussian_weights(sigma: float) -> np.ndarray:
    radius = int(4.0 * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    w = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return w / w.sum()


def gaussian_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    if sigma <= 0:
        return img
    w = gaussian_weights(sigma)
    r = len(w) // 2
    out = img
    for axis in (0, 1):
        n = out.shape[axis]
        acc = np.zeros_like(out)
        for k in range(-r, r + 1):
            acc = acc + w[k + r] * np.take(out, _reflect(np.arange(n) + k, n), axis=axis)
        out = acc
    return out


def sobel_energy(img: np.ndarray) -> float:
    """mean(gx^2 + gy^2) of the Sobel / 8 gradients over the WHOLE image (omit_boundary False)."""
    h, w = img.shape
    rows = [_reflect101(np.arange(h) + d, h) for d in (-1, 0, 1)]
    cols = [_reflect101(np.arange(w) + d, w) for d in (-1, 0, 1)]
    at = lambda a, b: img[np.ix_(rows[a + 1], cols[b + 1])]  # noqa: E731
    d_col = (at(-1, 1) + 2 * at(0, 1) + at(1, 1) - at(-1, -1) - 2 * at(0, -1) - at(1, -1)) / 8.0  # cv2.Sobel(dx=1, dy=0): along the width
    d_row = (at(1, -1) + 2 * at(1, 0) + at(1, 1) - at(-1, -1) - 2 * at(-1, 0) - at(-1, 1)) / 8.0  # cv2.Sobel(dx=0, dy=1): along the height
    return float(np.mean(d_col ** 2 + d_row ** 2))


def small_patch_cost(events: np.ndarray, motion: np.ndarray, image_size, pad=(0, 0), sigma: float = 1.0, normalize_t: bool = True) -> float:
    """`calculate_cost_for_small_patch(events, motion, "2d-translation")` (pyramid.py:379-415) for patch-local events."""
    warped = warp_2dof_middle(events, np.asarray(motion, dtype=np.float64), normalize_t)
    e_orig = sobel_energy(gaussian_blur(bilinear_vote(events, image_size, pad), sigma))
    e_warp = sobel_energy(gaussian_blur(bilinear_vote(warped, image_size, pad), sigma))
    with np.errstate(divide="ignore", invalid="ignore"):
        loss = np.float64(-e_orig) / np.float64(-e_warp)
    return 0.0 if np.isnan(loss) else float(loss)


def candidate_loss(events: np.ndarray, trans, image_size, pad=(0, 0), sigma: float = 1.0, normalize_t: bool = True) -> float:
    """`objective_initial` (pyramid.py:364-377) for one sampled (trans_x, trans_y): the candidate is scaled by the patch's
    time span when the solver normalises t in the batch."""
    theta = np.array([trans[0], trans[1]], dtype=np.float64)
    if normalize_t:
        theta = theta * (events[:, 2].max() - events[:, 2].min())
    return small_patch_cost(events, theta, image_size, pad, sigma, normalize_t)
