"""Block-compressed textures in the oracle: decoders against an independent numpy restatement and against
hand-computed blocks."""
import copy
import importlib

import numpy as np
import pytest

import bc_ref
import conftest

sc = conftest.pkg.scene


def _scene_with(default_scene, tex):
    s = copy.copy(default_scene)
    s.textures = list(default_scene.textures) + [tex]
    return s, sc.SCENE_TEXTURE_OFFSET + len(s.textures) - 1


def test_hand_computed_blocks(default_scene, oracle_mod):
    # BC1, c0 = 0xF800 (red) > c1 = 0x001F (blue): palette red, blue, (2r+b)/3, (r+2b)/3; indices 0,1,2,3 repeated
    blk = bytes([0x00, 0xF8, 0x1F, 0x00, 0b11100100] * 1 + [0b11100100] * 3)
    tex = sc.Texture(np.frombuffer(blk, np.uint8), srgb=True, bc_format=sc.TEXTURE_BC1, bc_extent=(4, 4), levels=1)
    s, slot = _scene_with(default_scene, tex)
    o = oracle_mod.OracleScene(s)
    assert o.texture_info(slot) == (4, 4, 1)
    lvl = o.texture_level(slot, 0)
    assert [tuple(p) for p in lvl[0]] == [(255, 0, 0, 255), (0, 0, 255, 255), (170, 0, 85, 255), (85, 0, 170, 255)]
    # c0 <= c1: three-colour mode with transparent black
    blk = bytes([0x1F, 0x00, 0x00, 0xF8, 0b11100100, 0, 0, 0])
    tex = sc.Texture(np.frombuffer(blk, np.uint8), bc_format=sc.TEXTURE_BC1, bc_extent=(4, 4))
    s, slot = _scene_with(default_scene, tex)
    lvl = oracle_mod.OracleScene(s).texture_level(slot, 0)
    assert [tuple(p) for p in lvl[0]] == [(0, 0, 255, 255), (255, 0, 0, 255), (128, 0, 128, 255), (0, 0, 0, 0)]
    # BC5 red block a0 = 255 > a1 = 0: 255, 0, then 6/7 .. 1/7 of 255 rounded
    bits = 0
    for t in range(16):
        bits |= (t % 8) << (3 * t)
    red = [255, 0] + list(bits.to_bytes(6, "little"))
    blk = bytes(red + [7, 7] + [0] * 6)
    tex = sc.Texture(np.frombuffer(blk, np.uint8), bc_format=sc.TEXTURE_BC5, bc_extent=(4, 4))
    s, slot = _scene_with(default_scene, tex)
    lvl = oracle_mod.OracleScene(s).texture_level(slot, 0)
    assert list(lvl[0, :, 0]) == [255, 0, 219, 182] and list(lvl[1, :, 0]) == [146, 109, 73, 36]
    assert (lvl[..., 1] == 7).all() and (lvl[..., 2] == 0).all() and (lvl[..., 3] == 255).all()


@pytest.mark.parametrize("fmt", [bc_ref.BC1, bc_ref.BC3, bc_ref.BC5])
@pytest.mark.parametrize("extent", [(64, 32), (20, 12), (7, 5)])
def test_random_blocks_match_numpy_decoder(default_scene, oracle_mod, fmt, extent):
    """Random bytes are valid blocks and reach every mode; three stored mip levels, odd extents."""
    w, h = extent
    rs = np.random.default_rng(fmt * 100 + w)
    sizes = [(max(1, w >> l), max(1, h >> l)) for l in range(3)]
    data = rs.integers(0, 256, sum(bc_ref.level_bytes(fmt, *e) for e in sizes), dtype=np.uint8)
    tex = sc.Texture(data, srgb=(fmt == bc_ref.BC1), bc_format=fmt, bc_extent=(w, h), levels=3)
    s, slot = _scene_with(default_scene, tex)
    o = oracle_mod.OracleScene(s)
    assert o.texture_info(slot) == (w, h, 3)
    off = 0
    for l, (lw, lh) in enumerate(sizes):
        want = bc_ref.decode_level(fmt, data[off:], lw, lh)
        off += bc_ref.level_bytes(fmt, lw, lh)
        assert np.array_equal(o.texture_level(slot, l), want), (fmt, l)


def test_encoder_round_trip():
    """The test encoder + reference decoder reproduce a two-colour image exactly."""
    img = np.zeros((8, 8, 3), np.uint8)
    img[:, :4] = (255, 0, 0)
    img[:, 4:] = (0, 0, 255)
    img[2, 1] = (0, 0, 255)
    dec = bc_ref.decode_level(bc_ref.BC1, bc_ref.encode_bc1_opaque(img), 8, 8)
    assert np.array_equal(dec[..., :3], img) and (dec[..., 3] == 255).all()
