"""Reader / writer of the checkpoint file cwa_checkpoint_save produces (layout: csrc/state.cu).  Pure numpy: tools and tests can
inspect, build or edit a checkpoint without a GPU."""
from __future__ import annotations

import numpy as np

MAGIC = b"CWACKPT1"
HEADER = np.dtype([
    ("magic", "S8"), ("version", "<u4"), ("header_bytes", "<u4"), ("frame", "<u8"),
    ("n_particles", "<u4"), ("particle_bytes", "<u4"), ("wave_w", "<u4"), ("wave_h", "<u4"), ("wave_ch", "<u4"), ("wave_variant", "<u4"),
    ("read_index", "<i4", 2), ("write_index", "<i4"), ("unit", "<i4", 3), ("tex_unit0", "<i4"), ("evolve", "<i4"),
    ("constants", "<f4", 4), ("boundary", "<f4", 8), ("wave", "<f4", 8), ("sim", "<f4", 12),
    ("pad", "V48"),
])
assert HEADER.itemsize == 256
PARTICLE = np.dtype([("pos", "<f4", 4), ("vel", "<f4", 4), ("force", "<f4", 4), ("extras", "<f4", 4)])


def read(path: str) -> dict:
    with open(path, "rb") as f:
        h = np.frombuffer(f.read(256), HEADER)[0]
        if h["magic"] != MAGIC or h["version"] != 1 or h["header_bytes"] != 256:
            raise ValueError(f"{path} is not a version-1 checkpoint")
        n, w, hh, ch = int(h["n_particles"]), int(h["wave_w"]), int(h["wave_h"]), int(h["wave_ch"])
        raw_p = f.read(n * 64)
        raw_i = [f.read(w * hh * ch * 4) for _ in range(3)]
        if len(raw_p) != n * 64 or any(len(r) != w * hh * ch * 4 for r in raw_i):
            raise ValueError(f"{path} is truncated")
        particles = np.frombuffer(raw_p, PARTICLE).copy()
        shape = (hh, w) if ch == 1 else (hh, w, ch)
        images = [np.frombuffer(r, "<f4").reshape(shape).copy() for r in raw_i]
    return {"header": h, "frame": int(h["frame"]), "particles": particles, "images": images}


def write(path: str, frame: int, particles: np.ndarray, images, *, read_index=(0, 1), write_index=2, unit=(0, 1, 2), tex_unit0=-1,
          evolve=1, wave_variant=0, constants=(0.02, 2.0, 3000.0, 1000.0), boundary=(0.48, 1.0, 0.48, 500.0, 0.0, -0.02, 0.0, 50.0),
          wave=(0.01, 0.985, 0.001, 1.0, 2.0, 0.35, -1.0, 0.0), sim=(0.005, 4000.0, 0.00005, -9806.65, 0.3, 0.01, 25.0, 2.0, 0.0, 0.0, 0.0, 0.0)):
    assert particles.dtype == PARTICLE and len(images) == 3
    im0 = np.asarray(images[0], "<f4")
    hh, w = im0.shape[0], im0.shape[1]
    ch = 1 if im0.ndim == 2 else im0.shape[2]
    h = np.zeros(1, HEADER)
    h["magic"], h["version"], h["header_bytes"], h["frame"] = MAGIC, 1, 256, frame
    h["n_particles"], h["particle_bytes"], h["wave_w"], h["wave_h"], h["wave_ch"], h["wave_variant"] = particles.size, 64, w, hh, ch, wave_variant
    h["read_index"], h["write_index"], h["unit"], h["tex_unit0"], h["evolve"] = read_index, write_index, unit, tex_unit0, evolve
    h["constants"], h["boundary"], h["wave"], h["sim"] = constants, boundary, wave, sim
    with open(path, "wb") as f:
        f.write(h.tobytes())
        f.write(np.ascontiguousarray(particles).tobytes())
        for im in images:
            f.write(np.ascontiguousarray(im, "<f4").tobytes())
