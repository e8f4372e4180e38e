"""Shared helpers for the parity tests."""
import numpy as np

RTOL_FP32 = 1e-4   # north-star: single-step rho / pres / force / height within 1e-4 relative


def rel_err(a, b, scale=None):
    """|a-b| / (|b| + scale).  `scale` defaults to the RMS of b: sums that cancel (interior forces of
    a lattice) are compared against the magnitude of the field, not against their own near-zero value."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if scale is None:
        fin = np.isfinite(b)
        scale = float(np.sqrt(np.mean(b[fin] ** 2))) if fin.any() else 0.0
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(a - b) / (np.abs(b) + scale + 1e-300)
    return e


def assert_close(a, b, rtol=RTOL_FP32, scale=None, what=""):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert (na == nb).all(), f"{what}: NaN masks differ ({int(na.sum())} vs {int(nb.sum())})"
    ok = ~nb
    ia, ib = np.isinf(a) & ok, np.isinf(b) & ok
    assert (ia == ib).all() and (a[ia] == b[ib]).all(), f"{what}: inf entries differ"
    ok &= ~ib
    if not ok.any():
        return 0.0
    e = rel_err(a[ok], b[ok], scale)
    worst = float(e.max())
    assert worst <= rtol, f"{what}: max rel err {worst:.3e} > {rtol:.1e} at flat index {int(np.argmax(e))}"
    return worst


def jittered_block(O, nx, ny, nz, prm, seed=1234, jitter=0.1, vel=0.0):
    """make_cube lattice with +-jitter*spacing uniform noise (default_rng(1234))."""
    p = O.make_cube(nx, ny, nz, prm)
    rng = np.random.default_rng(seed)
    spacing = prm.smoothing_coeff * 0.85 * prm.particle_radius
    p["pos"][:, :3] += rng.uniform(-jitter, jitter, (p.size, 3)).astype(np.float32) * np.float32(spacing)
    if vel > 0:
        p["vel"][:, :3] = rng.uniform(-vel, vel, (p.size, 3)).astype(np.float32)
    return p


def random_field(h, w, seed=7, ch=1, amp=0.5):
    rng = np.random.default_rng(seed)
    shape = (h, w) if ch == 1 else (h, w, ch)
    return (amp * rng.standard_normal(shape)).astype(np.float32)


def smooth_field(h, w, ch=1, amp=0.05):
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    f = amp * (np.sin(x * 0.37) * np.cos(y * 0.23) + 0.5 * np.sin((x + y) * 0.11)).astype(np.float32)
    if ch == 1:
        return np.ascontiguousarray(f)
    out = np.zeros((h, w, ch), np.float32)
    out[..., 0] = f
    return out
