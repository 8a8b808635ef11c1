"""Image comparison metrics used by the parity tests."""
import numpy as np


def rel_mse(img, ref, eps=1e-2):
    """Mean over pixels of |img - ref|^2 / (ref^2 + eps) on RGB (the usual relMSE of rendering papers)."""
    img, ref = np.asarray(img, np.float64)[..., :3], np.asarray(ref, np.float64)[..., :3]
    return float(np.mean(np.sum((img - ref) ** 2, -1) / (np.sum(ref**2, -1) + eps)))


def tonemap_srgb(linear):
    """1 - exp(-c) (toneMapping.comp:21) followed by the sRGB OETF, in [0, 1]."""
    c = 1 - np.exp(-np.maximum(np.asarray(linear, np.float64), 0))
    return np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(np.maximum(c, 1e-12), 1 / 2.4) - 0.055)


def flip_lite(img, ref):
    """A simplified FLIP-style perceptual error in [0, 1]: mean over pixels of the Hunt-adjusted
    colour difference of the tone-mapped images after a 3x3 spatial filter, in a YCxCz-like
    opponent space.  (The full FLIP pipeline's contrast-sensitivity filters and feature maps are
    omitted; this bound is therefore stated as 'FLIP-lite' wherever it is quoted.)"""
    a, b = tonemap_srgb(img)[..., :3], tonemap_srgb(ref)[..., :3]

    def blur(x):
        p = np.pad(x, ((1, 1), (1, 1), (0, 0)), mode="edge")
        return sum(p[i : i + x.shape[0], j : j + x.shape[1]] for i in range(3) for j in range(3)) / 9.0

    def opponent(x):
        y = 0.2126 * x[..., 0] + 0.7152 * x[..., 1] + 0.0722 * x[..., 2]
        return np.stack([y, x[..., 0] - y, x[..., 2] - y], -1)

    d = opponent(blur(a)) - opponent(blur(b))
    return float(np.mean(np.sqrt(np.sum(d**2, -1))))


def close_fraction(img, ref, tol=1e-4):
    """Fraction of pixels whose RGB agrees within tol * max(1, |ref|)."""
    d = np.abs(np.asarray(img)[..., :3] - np.asarray(ref)[..., :3]).max(-1)
    return float((d <= tol * np.maximum(1.0, np.abs(np.asarray(ref)[..., :3]).max(-1))).mean())
