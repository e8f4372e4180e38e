"""CPU: the numpy reader / writer of the checkpoint file (layout of csrc/state.cu) round-trips and rejects broken files."""
import numpy as np
import pytest

from coupledwateranimation_b200 import checkpoint as ck


def test_header_layout_matches_the_c_struct():
    assert ck.HEADER.itemsize == 256
    off = {n: ck.HEADER.fields[n][1] for n in ck.HEADER.names}
    assert (off["frame"], off["n_particles"], off["read_index"], off["constants"], off["boundary"], off["wave"], off["sim"]) == (16, 24, 48, 80, 96, 128, 160)
    assert ck.PARTICLE.itemsize == 64


@pytest.mark.parametrize("ch", [1, 4])
def test_roundtrip(tmp_path, ch):
    rng = np.random.default_rng(3)
    p = np.zeros(37, ck.PARTICLE)
    for f in ck.PARTICLE.names:
        p[f] = rng.standard_normal((37, 4)).astype(np.float32)
    p["pos"][5, 0] = np.nan
    shape = (6, 10) if ch == 1 else (6, 10, 4)
    ims = [rng.standard_normal(shape).astype(np.float32) for _ in range(3)]
    path = str(tmp_path / "a.ckpt")
    ck.write(path, 123456789012, p, ims, read_index=(2, 0), write_index=1, unit=(1, 2, 0), tex_unit0=2, wave_variant=1,
             sim=(0.005, 4000.0, 2e-5, -9806.65, 0.3, 0.01, 25.0, 2.0 / 7.0, 0.0, 0.125, 0.0, 0.0))
    r = ck.read(path)
    assert r["frame"] == 123456789012
    assert np.array_equal(r["particles"].view(np.uint8), p.view(np.uint8))
    assert all(np.array_equal(a, b) for a, b in zip(r["images"], ims))
    h = r["header"]
    assert list(h["read_index"]) == [2, 0] and int(h["write_index"]) == 1 and list(h["unit"]) == [1, 2, 0] and int(h["tex_unit0"]) == 2
    assert int(h["wave_ch"]) == ch and int(h["wave_variant"]) == 1 and h["sim"][9] == np.float32(0.125)
    assert h["constants"].tolist() == [np.float32(0.02), 2.0, 3000.0, 1000.0]          # the reference's ConstantsUniform defaults


def test_rejects_broken_files(tmp_path):
    path = str(tmp_path / "b.ckpt")
    ck.write(path, 1, np.zeros(4, ck.PARTICLE), [np.zeros((3, 3), np.float32)] * 3)
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[:-8])
    with pytest.raises(ValueError, match="truncated"):
        ck.read(path)
    open(path, "wb").write(b"X" + raw[1:])
    with pytest.raises(ValueError, match="not a version-1 checkpoint"):
        ck.read(path)
