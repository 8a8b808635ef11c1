"""Independent numpy decoders for BC1 / BC3 / BC5 blocks (test reference for the oracle's and the core's decoders)."""
import numpy as np

BC1, BC3, BC5 = 2, 3, 4


def _alpha(blocks):
    """(N, 8) uint8 -> (N, 16) uint8."""
    a0, a1 = blocks[:, 0].astype(np.int64), blocks[:, 1].astype(np.int64)
    pal = np.zeros((len(blocks), 8), np.int64)
    pal[:, 0], pal[:, 1] = a0, a1
    for i in range(1, 7):
        pal[:, 1 + i] = ((7 - i) * a0 + i * a1 + 3) // 7
    five = np.zeros_like(pal)
    five[:, 0], five[:, 1] = a0, a1
    for i in range(1, 5):
        five[:, 1 + i] = ((5 - i) * a0 + i * a1 + 2) // 5
    five[:, 6], five[:, 7] = 0, 255
    pal = np.where((a0 > a1)[:, None], pal, five)
    bits = np.zeros(len(blocks), np.uint64)
    for k in range(6):
        bits |= blocks[:, 2 + k].astype(np.uint64) << np.uint64(8 * k)
    idx = np.stack([(bits >> np.uint64(3 * t)) & np.uint64(7) for t in range(16)], 1).astype(np.int64)
    return np.take_along_axis(pal, idx, 1).astype(np.uint8)


def _color(blocks, punch_through):
    """(N, 8) uint8 -> (N, 16, 4) uint8."""
    c0 = blocks[:, 0].astype(np.int64) | (blocks[:, 1].astype(np.int64) << 8)
    c1 = blocks[:, 2].astype(np.int64) | (blocks[:, 3].astype(np.int64) << 8)

    def expand(c):
        r, g, b = c >> 11, (c >> 5) & 63, c & 31
        return np.stack([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (b << 3) | (b >> 2), np.full_like(c, 255)], 1)

    p0, p1 = expand(c0), expand(c1)
    four = (c0 > c1) | (not punch_through)
    p2 = np.where(four[:, None], (2 * p0 + p1 + 1) // 3, (p0 + p1 + 1) // 2)
    p3 = np.where(four[:, None], (p0 + 2 * p1 + 1) // 3, 0)
    p2[:, 3] = 255
    p3[:, 3] = np.where(four, 255, 0)
    pal = np.stack([p0, p1, p2, p3], 1)  # (N, 4, 4)
    idx = (blocks[:, 4].astype(np.int64) | (blocks[:, 5].astype(np.int64) << 8) | (blocks[:, 6].astype(np.int64) << 16)
           | (blocks[:, 7].astype(np.int64) << 24))
    sel = np.stack([(idx >> (2 * t)) & 3 for t in range(16)], 1)  # (N, 16)
    return np.take_along_axis(pal, sel[:, :, None].repeat(4, 2), 1).astype(np.uint8)


def level_bytes(fmt, w, h):
    return ((w + 3) // 4) * ((h + 3) // 4) * (8 if fmt == BC1 else 16)


def decode_level(fmt, data, w, h):
    """Flat uint8 block data of one level -> (h, w, 4) uint8."""
    bw, bh = (w + 3) // 4, (h + 3) // 4
    bs = 8 if fmt == BC1 else 16
    blocks = np.frombuffer(bytes(data[: bw * bh * bs]), np.uint8).reshape(bw * bh, bs)
    if fmt == BC1:
        tex = _color(blocks, True)
    elif fmt == BC3:
        tex = _color(blocks[:, 8:], False)
        tex[:, :, 3] = _alpha(blocks[:, :8])
    else:
        tex = np.zeros((len(blocks), 16, 4), np.uint8)
        tex[:, :, 0], tex[:, :, 1], tex[:, :, 3] = _alpha(blocks[:, :8]), _alpha(blocks[:, 8:]), 255
    img = tex.reshape(bh, bw, 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(bh * 4, bw * 4, 4)
    return img[:h, :w]


def encode_bc1_opaque(rgb):
    """A plain min/max-endpoint BC1 encoder for test textures: (h, w, 3) uint8, h and w multiples of 4."""
    h, w = rgb.shape[:2]
    b = rgb.reshape(h // 4, 4, w // 4, 4, 3).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 3).astype(np.int64)
    # endpoints: the two texels of the block that are farthest apart
    d2 = ((b[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1).reshape(len(b), 256)
    far = d2.argmax(1)
    hi, lo = b[np.arange(len(b)), far // 16], b[np.arange(len(b)), far % 16]

    def to565(c):
        return ((c[:, 0] >> 3) << 11) | ((c[:, 1] >> 2) << 5) | (c[:, 2] >> 3)

    c0, c1 = to565(hi), to565(lo)
    swap = c0 < c1
    c0, c1 = np.where(swap, c1, c0), np.where(swap, c0, c1)
    out = np.zeros((len(b), 8), np.uint8)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = c0 & 255, c0 >> 8, c1 & 255, c1 >> 8
    # palette of each block from the decoder itself (indices 0..3 in the first four texels)
    probe = out.copy()
    probe[:, 4:] = [0b11100100, 0, 0, 0]
    pal = _color(probe, True)[:, :4, :3].astype(np.int64)
    d = ((b[:, :, None, :] - pal[:, None, :, :]) ** 2).sum(-1)  # (N, 16, 4)
    sel = d.argmin(-1)
    sel = np.where((c0 == c1)[:, None], 0, sel)  # equal endpoints select the 3-colour mode: index 3 is transparent
    idx = np.zeros(len(b), np.int64)
    for t in range(16):
        idx |= sel[:, t] << (2 * t)
    for k in range(4):
        out[:, 4 + k] = (idx >> (8 * k)) & 255
    return out.reshape(-1)


def _blocks(img):
    h, w = img.shape[:2]
    c = img.shape[2] if img.ndim == 3 else 1
    return img.reshape(h // 4, 4, w // 4, 4, c).transpose(0, 2, 1, 3, 4).reshape(-1, 16, c)


def encode_alpha(values):
    """(N, 16) uint8 -> (N, 8) uint8 single-channel blocks (max / min endpoints, nearest palette entry)."""
    v = values.astype(np.int64)
    out = np.zeros((len(v), 8), np.uint8)
    out[:, 0], out[:, 1] = v.max(1), v.min(1)
    probe = out.copy()
    bits = 0
    for t in range(8):
        bits |= t << (3 * t)
    probe[:, 2:] = np.frombuffer(int(bits).to_bytes(6, "little"), np.uint8)
    pal = _alpha(probe)[:, :8].astype(np.int64)
    sel = np.abs(v[:, :, None] - pal[:, None, :]).argmin(-1)
    idx = np.zeros(len(v), np.uint64)
    for t in range(16):
        idx |= sel[:, t].astype(np.uint64) << np.uint64(3 * t)
    for k in range(6):
        out[:, 2 + k] = ((idx >> np.uint64(8 * k)) & np.uint64(255)).astype(np.uint8)
    return out


def encode(fmt, rgba):
    """(h, w, 4) uint8 (h, w multiples of 4) -> flat block data.  BC1 ignores alpha, BC5 keeps r and g."""
    if fmt == BC1:
        return encode_bc1_opaque(rgba[..., :3])
    b = _blocks(rgba)
    if fmt == BC5:
        return np.concatenate([encode_alpha(b[:, :, 0]), encode_alpha(b[:, :, 1])], 1).reshape(-1)
    color = encode_bc1_opaque(rgba[..., :3]).reshape(-1, 8)
    return np.concatenate([encode_alpha(b[:, :, 3]), color], 1).reshape(-1)


def encode_chain(fmt, rgba, levels):
    """Block data of `levels` mip levels (2x2 box filter), level 0 first; extents stay multiples of 4."""
    out, img = [], rgba.astype(np.float64)
    for _ in range(levels):
        out.append(encode(fmt, np.rint(img).astype(np.uint8)))
        h, w = img.shape[:2]
        img = img.reshape(h // 2, 2, w // 2, 2, 4).mean((1, 3))
    return np.concatenate(out)
