"""Post-process + output chain of the oracle (oracle/pt_oracle_post.cpp) against independent closed forms.

The reference has no test or golden image for postprocess.comp / bloom*.comp / composition.comp /
toneMapping.comp (parity unpinned); what can be pinned is the arithmetic each stage
states: half-precision stores, the soft-knee prefilter, kernel weights that sum to one, the tone
curve and the sRGB transfer function."""
import numpy as np


def srgb8(linear):
    l = np.clip(np.asarray(linear, np.float64), 0.0, 1.0)
    e = np.where(l <= 0.0031308, 12.92 * l, 1.055 * l ** (1 / 2.4) - 0.055)
    return np.rint(e * 255).astype(np.uint8)


def half(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def test_round_half_matches_numpy(oracle_mod):
    rs = np.random.default_rng(7)
    bits = rs.integers(0, 2 ** 32, 200000, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    x = x[np.isfinite(x)]
    special = np.array([0.0, -0.0, 1.0, 65504.0, 65519.99, 65520.0, 1e9, -1e9, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.0001,
                        1 + 2.0 ** -11, 1 + 3 * 2.0 ** -11, 6.1e-5, 6.0e-8, np.inf, -np.inf], np.float32)
    x = np.concatenate([x, special, rs.uniform(-70000, 70000, 100000).astype(np.float32),
                        (rs.uniform(0, 1, 100000) ** 8).astype(np.float32)])
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).astype(np.float32)
    got = oracle_mod.round_half(x)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.isnan(oracle_mod.round_half(np.array([np.nan], np.float32))[0])


def test_constant_image_below_threshold(oracle_mod):
    """No bloom (nothing above threshold - knee = 0.5): out = srgb8(half(1 - exp(-half(c / n * exposure))))."""
    for c, n, exposure in ((0.5, 4, 1.0), (0.18, 16, 2.0), (0.45, 1, 1.0), (0.0, 3, 1.0)):
        acc = np.zeros((24, 40, 4), np.float32)
        acc[..., :3] = np.float32(c) * n
        acc[..., 3] = 1
        out = oracle_mod.postprocess(acc, n, exposure=exposure)
        col = half(np.float32(c) * n / np.float32(n) * np.float32(exposure))
        want = srgb8(half(1.0 - np.exp(-np.float64(col))))
        assert out.shape == (24, 40, 4) and (out[..., 3] == 255).all()
        assert (out[..., :3] == want).all(), (c, out[0, 0], want)
        hdr = oracle_mod.postprocess(acc, n, exposure=exposure, hdr=True)
        assert (hdr[..., :3] == col).all() and (hdr[..., 3] == 1).all()


def test_nan_and_inf_markers(oracle_mod):
    """postprocess.comp:25-28: NaN -> (5000, 0, 0), Inf -> (0, 5000, 0); intensity 0 keeps bloom out."""
    acc = np.full((16, 16, 4), 0.25, np.float32)
    acc[3, 5, 1] = np.nan
    acc[9, 2, 0] = np.inf
    acc[12, 12, 2] = -np.inf
    hdr = oracle_mod.postprocess(acc, 1, bloom_intensity=0.0, hdr=True)
    assert tuple(hdr[3, 5, :3]) == (5000.0, 0.0, 0.0)
    assert tuple(hdr[9, 2, :3]) == (0.0, 5000.0, 0.0) and tuple(hdr[12, 12, :3]) == (0.0, 5000.0, 0.0)
    assert (hdr[0, 0, :3] == 0.25).all()
    sdr = oracle_mod.postprocess(acc, 1, bloom_intensity=0.0)
    assert tuple(sdr[3, 5, :3]) == (255, 0, 0) and tuple(sdr[9, 2, :3]) == (0, 255, 0)


def test_bloom_of_a_constant_image(oracle_mod):
    """Down- and upsample kernels both sum to one, so on a constant image every level holds the
    prefiltered colour b and the upsample chain adds one b per level: bloom[0] = maxMip * b."""
    w, h = 64, 48
    levels = int(np.floor(np.log2(max(w, h)))) + 1
    max_mip = min(levels - 3, 12)
    c, threshold, intensity = np.float32(3.0), np.float32(1.0), np.float32(0.5)
    acc = np.zeros((h, w, 4), np.float32)
    acc[..., :3] = c
    hdr = oracle_mod.postprocess(acc, 1, bloom_threshold=threshold, bloom_intensity=intensity, hdr=True)
    br = c
    rq = np.clip(br - (threshold - 0.5), 0.0, 1.0)
    rq = 0.5 * rq * rq
    b = half(c * (max(rq, br - threshold) / max(br, 1e-4)))
    want = intensity * np.float32(0.1) * (max_mip * b) + c
    assert np.allclose(hdr[..., :3], want, rtol=2e-3), (hdr[0, 0], want)
    # interior and border agree: clamp-to-edge sampling of a constant is the constant
    assert np.ptp(hdr[..., 0]) <= 2e-3 * want


def test_bloom_spreads_and_conserves_energy(oracle_mod):
    """One bright pixel on black: the composite keeps the pixel, adds a halo that decays with
    distance and stays non-negative; with intensity 0 the halo vanishes."""
    acc = np.zeros((128, 128, 4), np.float32)
    acc[64, 64, :3] = 400.0
    base = oracle_mod.postprocess(acc, 1, bloom_intensity=0.0, hdr=True)
    assert base[64, 64, 0] == 400.0 and base[..., :3].sum() == 1200.0
    img = oracle_mod.postprocess(acc, 1, bloom_intensity=1.0, hdr=True)
    halo = img[..., 0] - base[..., 0]
    assert (halo >= 0).all() and halo[64, 64] > 0
    assert halo[64, 70] > halo[64, 90] > 0 and halo[70, 64] > halo[90, 64] > 0
    # symmetric kernel on a symmetric input (pixel centre is not the image centre: allow fp16 noise)
    assert np.isclose(halo[64, 60], halo[60, 64], rtol=5e-2)


def test_small_frames_skip_bloom(oracle_mod):
    """Frames under 8 pixels have no bloom mips to run (the reference's unsigned min(levels - 3, 12) would
    wrap): only the prefiltered level-0 colour b = c * (c - threshold) / c = 3 is composited."""
    acc = np.full((4, 6, 4), 8.0, np.float32)
    hdr = oracle_mod.postprocess(acc, 2, bloom_intensity=1.0, hdr=True)
    assert (hdr[..., :3] == half(np.float32(1.0) * np.float32(0.1) * np.float32(3.0) + np.float32(4.0))).all()


def test_output_file_writers(tmp_path):
    """The Python mirror's .hdr (RGBE) and .tga writers round-trip what they are given."""
    import importlib
    import os

    core = importlib.import_module("path-tracing_b200.core")
    rs = np.random.default_rng(2)
    img = (rs.random((7, 9, 4)) ** 3 * 40).astype(np.float32)
    img[0, 0, :3] = 0
    path = os.path.join(tmp_path, "a.hdr")
    core.write_hdr(path, img)
    data = open(path, "rb").read()
    head, body = data.split(b"\n\n", 1)
    assert head.startswith(b"#?RADIANCE") and b"32-bit_rle_rgbe" in head
    dims, raw = body.split(b"\n", 1)
    assert dims == b"-Y 7 +X 9"
    rgbe = np.frombuffer(raw, np.uint8).reshape(7, 9, 4).astype(np.float64)
    dec = rgbe[..., :3] * np.exp2(rgbe[..., 3:4] - 136.0)
    dec[rgbe[..., 3] == 0] = 0
    quantum = img[..., :3].max(-1, keepdims=True) / 128.0  # 8-bit mantissas under the exponent of the largest component
    assert (np.abs(dec - img[..., :3]) <= quantum + 1e-6).all() and (dec <= img[..., :3] + 1e-6).all()  # truncation
    rgba8 = rs.integers(0, 256, (5, 6, 4), dtype=np.uint8)
    path = os.path.join(tmp_path, "a.tga")
    core.write_tga(path, rgba8)
    data = open(path, "rb").read()
    assert data[2] == 2 and int.from_bytes(data[12:14], "little") == 6 and int.from_bytes(data[14:16], "little") == 5 and data[16] == 32
    assert np.array_equal(np.frombuffer(data[18:], np.uint8).reshape(5, 6, 4)[..., [2, 1, 0, 3]], rgba8)
