"""Image comparison metrics of the parity checks (tests/ and bench.py's `parity` block).

relMSE as rendering papers use it, the fraction of pixels that agree within a tolerance, and LDR-FLIP
(Andersson, Nilsson, Akenine-Moller, Oskarsson, Astrom, Fairchild: "FLIP: A Difference Evaluator for Alternating
Images", HPG 2020), implemented from the paper: colour pipeline (contrast-sensitivity filtering in YCxCz, Hunt
adjustment, HyAB distance, error redistribution) and feature pipeline (edge / point detectors on luminance),
combined as dE = dEc ^ (1 - dEf).  Inputs to flip() are LINEAR radiance images; they are brought to display
space the way the reference presents them (tone mapping 1 - exp(-c), toneMapping.comp:21, then the sRGB OETF).
numpy only; no test-infrastructure or CUDA dependency.
"""
from __future__ import annotations

import numpy as np

# ---- generic -------------------------------------------------------------------------------------


def rel_mse(img, ref, eps=1e-2):
    """Mean over pixels of |img - ref|^2 / (ref^2 + eps) on RGB."""
    img, ref = np.asarray(img, np.float64)[..., :3], np.asarray(ref, np.float64)[..., :3]
    return float(np.mean(np.sum((img - ref) ** 2, -1) / (np.sum(ref**2, -1) + eps)))


def close_fraction(img, ref, tol=1e-4):
    """Fraction of pixels whose RGB agrees within tol * max(1, |ref|)."""
    d = np.abs(np.asarray(img)[..., :3] - np.asarray(ref)[..., :3]).max(-1)
    return float((d <= tol * np.maximum(1.0, np.abs(np.asarray(ref)[..., :3]).max(-1))).mean())


def tonemap_srgb(linear):
    """1 - exp(-c) (toneMapping.comp:21) followed by the sRGB OETF, in [0, 1]."""
    c = 1 - np.exp(-np.maximum(np.asarray(linear, np.float64), 0))
    return np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(np.maximum(c, 1e-12), 1 / 2.4) - 0.055)


# ---- LDR-FLIP ------------------------------------------------------------------------------------

_QC, _PC, _PT = 0.7, 0.4, 0.95  # colour pipeline
_QF, _W_FEATURE = 0.5, 0.082  # feature pipeline: exponent, detector width in degrees
_REF_WHITE = np.array([0.950428545, 1.0, 1.088900371])  # D65 in XYZ
DEFAULT_PPD = 0.7 * 3840 / 0.7 * np.pi / 180  # 0.7 m from a 0.7 m wide 3840-pixel display: 67.02 pixels per degree


def _srgb_to_linear(c):
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


_RGB2XYZ = np.array([[10135552 / 24577794, 8788810 / 24577794, 4435075 / 24577794],
                     [2613072 / 12288897, 8788810 / 12288897, 887015 / 12288897],
                     [1425312 / 73733382, 8788810 / 73733382, 70074185 / 73733382]])
_XYZ2RGB = np.linalg.inv(_RGB2XYZ)


def _xyz_to_ycxcz(xyz):
    n = xyz / _REF_WHITE
    return np.stack([116 * n[..., 1] - 16, 500 * (n[..., 0] - n[..., 1]), 200 * (n[..., 1] - n[..., 2])], -1)


def _ycxcz_to_xyz(c):
    y = (c[..., 0] + 16) / 116
    return np.stack([y + c[..., 1] / 500, y, y - c[..., 2] / 200], -1) * _REF_WHITE


def _xyz_to_lab(xyz):
    n = xyz / _REF_WHITE
    d = 6 / 29
    f = np.where(n > d**3, np.cbrt(np.maximum(n, 0)), n / (3 * d * d) + 4 / 29)
    return np.stack([116 * f[..., 1] - 16, 500 * (f[..., 0] - f[..., 1]), 200 * (f[..., 1] - f[..., 2])], -1)


def _hunt(lab):
    return np.stack([lab[..., 0], 0.01 * lab[..., 0] * lab[..., 1], 0.01 * lab[..., 0] * lab[..., 2]], -1)


def _hyab(a, b):
    d = a - b
    return np.abs(d[..., 0]) + np.sqrt(d[..., 1] ** 2 + d[..., 2] ** 2)


def _conv1(img, k, axis):
    """Correlation with a 1-D kernel along `axis` of an (H, W) image, edges replicated (the paper's choice)."""
    r = len(k) // 2
    pad = [(0, 0), (0, 0)]
    pad[axis] = (r, r)
    p = np.pad(img, pad, mode="edge")
    out = np.zeros_like(img, dtype=np.float64)
    n = img.shape[axis]
    for d in range(len(k)):
        if k[d] != 0.0:
            out += k[d] * (p[d : d + n, :] if axis == 0 else p[:, d : d + n])
    return out


def _sep(img, kx, ky):
    """Separable 2-D filter kx(x) * ky(y)."""
    return _conv1(_conv1(img, kx, 1), ky, 0)


def _csf_filters(ppd):
    """Spatial-domain contrast sensitivity filters of the achromatic, red-green and blue-yellow channels as sums
    of separable Gaussians a * sqrt(pi / b) * exp(-pi^2 d^2 / b), each channel normalised to unit sum:
    a list of [(weight, 1-D kernel), ...] per channel."""
    params = {"A": [(1.0, 0.0047)], "RG": [(1.0, 0.0053)], "BY": [(34.1, 0.04), (13.5, 0.025)]}
    radius = int(np.ceil(3 * np.sqrt(0.04 / (2 * np.pi**2)) * ppd))
    x = np.arange(-radius, radius + 1) / ppd
    out = []
    for c in ("A", "RG", "BY"):
        terms = [(a * np.sqrt(np.pi / b), np.exp(-np.pi**2 * x**2 / b)) for a, b in params[c]]
        total = sum(w * k.sum() ** 2 for w, k in terms)
        out.append([(w / total, k) for w, k in terms])
    return out


def _feature_filters(ppd):
    """1-D factors of the edge (first derivative of a Gaussian) and point (second derivative) detectors, their positive
    and negative lobes normalised to +1 / -1 like the paper's 2-D kernels."""
    sd = 0.5 * _W_FEATURE * ppd
    radius = int(np.ceil(3 * sd))
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    g = np.exp(-(x**2) / (2 * sd * sd))
    g /= g.sum()

    def norm(k):
        return np.where(k > 0, k / k[k > 0].sum(), k / -k[k < 0].sum())

    return g, norm(-x * np.exp(-(x**2) / (2 * sd * sd))), norm((x**2 / (sd * sd) - 1) * np.exp(-(x**2) / (2 * sd * sd)))


def flip_map_ldr(test_srgb, ref_srgb, ppd=DEFAULT_PPD):
    """Per-pixel LDR-FLIP error in [0, 1] of two sRGB-encoded images in [0, 1] (H, W, 3)."""
    test, ref = np.clip(np.asarray(test_srgb, np.float64)[..., :3], 0, 1), np.clip(np.asarray(ref_srgb, np.float64)[..., :3], 0, 1)
    opp = [_xyz_to_ycxcz(_srgb_to_linear(x) @ _RGB2XYZ.T) for x in (ref, test)]
    # ---- colour pipeline
    filters = _csf_filters(ppd)
    lab = []
    for o in opp:
        f = np.stack([sum(w * _sep(o[..., c], k, k) for w, k in filters[c]) for c in range(3)], -1)
        rgb = np.clip(_ycxcz_to_xyz(f) @ _XYZ2RGB.T, 0, 1)
        lab.append(_hunt(_xyz_to_lab(rgb @ _RGB2XYZ.T)))
    de_c = _hyab(lab[0], lab[1]) ** _QC
    green = _hunt(_xyz_to_lab(np.array([0.0, 1.0, 0.0]) @ _RGB2XYZ.T))
    blue = _hunt(_xyz_to_lab(np.array([0.0, 0.0, 1.0]) @ _RGB2XYZ.T))
    cmax = _hyab(green, blue) ** _QC
    limit = _PC * cmax
    de_c = np.where(de_c < limit, _PT / limit * de_c, _PT + (de_c - limit) / (cmax - limit) * (1 - _PT))
    # ---- feature pipeline (on achromatic Y normalised to [0, 1])
    g, d1, d2 = _feature_filters(ppd)
    feat = []
    for o in opp:
        y = (o[..., 0] + 16) / 116
        edge = np.hypot(_sep(y, d1, g), _sep(y, g, d1))
        point = np.hypot(_sep(y, d2, g), _sep(y, g, d2))
        feat.append((edge, point))
    de_f = (np.maximum(np.abs(feat[0][0] - feat[1][0]), np.abs(feat[0][1] - feat[1][1])) / np.sqrt(2)) ** _QF
    return de_c ** (1 - de_f)


def flip(img_linear, ref_linear, ppd=DEFAULT_PPD):
    """Mean LDR-FLIP error of two LINEAR radiance images, displayed like the reference displays them
    (tone mapping 1 - exp(-c), sRGB encoding)."""
    return float(np.mean(flip_map_ldr(tonemap_srgb(img_linear), tonemap_srgb(ref_linear), ppd)))
