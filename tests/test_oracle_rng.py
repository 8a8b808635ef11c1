"""Pins the oracle's RNG bit-exactly to the known answers derived from the integer spec at
Path-Tracing/Shaders/common.glsl:133-165."""
import numpy as np

MODE_RNG = 11


def _rng(oracle_mod, px, py, width, frame):
    inp = np.array([[px, py, width, frame]], np.uint32).view(np.float32)
    out = oracle_mod.test_shading(MODE_RNG, inp)[0]
    bits = out.view(np.uint32)
    return int(bits[0]), [int(b) for b in bits[1:5]], [float(f) for f in out[5:9]]


def test_init_and_stream_kats(oracle_mod):
    seed, st, fl = _rng(oracle_mod, 1, 0, 512, 0)
    assert seed == 0x124EA49D
    assert st == [0x1D719993, 0xE63E38F2, 0x052D6422, 0x9C876E36]
    assert np.allclose(fl, [0.115014553, 0.899386883, 0.020223856, 0.611441493], rtol=0, atol=1e-7)

    seed, st, fl = _rng(oracle_mod, 0, 1, 512, 0)
    assert seed == 0x9E003B6E and st == [0xB4DB4CD8, 0x75446D78, 0xE58930AD, 0xB002DD03]
    assert abs(fl[0] - 0.706471205) < 1e-7

    seed, st, fl = _rng(oracle_mod, 255, 255, 512, 0)
    assert seed == 0x9CF7ABFA and st[0] == 0x589338FE and abs(fl[0] - 0.34599638) < 1e-7

    seed, st, fl = _rng(oracle_mod, 0, 0, 512, 1)
    assert seed == 0xC0738807 and st == [0x9F15277E, 0x44A5AAE3, 0xCECFF1FF, 0x2EF13967]
    assert abs(fl[0] - 0.62141645) < 1e-7

    seed, st, fl = _rng(oracle_mod, 1919, 1079, 1920, 7)
    assert seed == 0xD81BF370 and st[0] == 0x68C1A90A and abs(fl[0] - 0.40920496) < 1e-7


def test_zero_seed_quirk(oracle_mod):
    """Q2: pixel (0,0) of frame 0 seeds 0 and xorshift never leaves 0."""
    seed, st, fl = _rng(oracle_mod, 0, 0, 1920, 0)
    assert seed == 0 and st == [0, 0, 0, 0] and fl == [0.0, 0.0, 0.0, 0.0]


def test_floats_in_unit_interval(oracle_mod):
    rs = np.random.default_rng(1)
    inp = rs.integers(0, 4096, (2000, 4)).astype(np.uint32)
    inp[:, 2] = 4096
    out = oracle_mod.test_shading(MODE_RNG, inp.view(np.float32))
    f = out[:, 5:9]
    assert (f >= 0).all() and (f < 1).all()
