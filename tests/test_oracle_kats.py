"""Known-answer tests derived BY HAND from the shader text (closed forms evaluated in float64 here, nothing below calls the oracle to
produce an expected value): the floating-point half of the oracle has no reference fixture to pin it (the GLSL cannot run in this
image), so these are its independent check -- and, in the `gpu` variants, the CUDA path's.

Scene: two particles closer than h = 0.01, one above CREST_THRESHOLD and one below, over a LINEAR height field
T[j][i] = a*i + b*j + c0.  GL_LINEAR sampling of a linear field is exact away from the border: texture(s, t) = a*(s*W - 0.5) +
b*(t*H - 0.5) + c0; CLAMP_TO_EDGE makes the `uv + 1` taps of WaveNormal read column W-1 / row H-1.

  rho_pres_comp.glsl:60-80   density (self included), EOS pressure, wave coupling
  force_comp.glsl:68-139     spiky pressure, viscosity, crest rule, torque on the stored force, wave drag / normal force, gravity
  integrate_comp.glsl:62-84  symplectic Euler, damping, foam rule, surface clamp
"""
import numpy as np
import pytest

PI = 3.141592741                     # rho_pres_comp.glsl:8
R = 0.005
H_S = 2.0 * R                        # smoothing_length = smoothing_coeff * PARTICLE_RADIUS = 0.01
MASS, RHO0, VISC = 0.02, 1000.0, 3000.0
GAS, DT, G_Y, DAMPING = 4000.0, 0.00005, -9806.65, 0.3
W = H = 32
A, B, C0 = 0.002, -0.003, 0.014      # height field T[j][i] = A*i + B*j + C0 (about 0.008 under the particles: between their heights)


def field():
    i, j = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    return (A * i + B * j + C0).astype(np.float32)


def tex(s, t):
    """exact GL_LINEAR + CLAMP_TO_EDGE sample of the linear field"""
    u = min(max(s * W - 0.5, 0.0), W - 1.0)
    v = min(max(t * H - 0.5, 0.0), H - 1.0)
    return A * u + B * v + C0


def particles(cwa_dtype):
    p = np.zeros(2, cwa_dtype)
    p["pos"][0] = (0.100, 0.005, 0.100, 1.0)            # below the crest threshold 0.01
    p["pos"][1] = (0.103, 0.012, 0.102, 1.0)            # above it; |delta| = 7.87e-3 < h
    p["vel"][0] = (5.0, -30.0, 10.0, 0.0)               # fast: the foam rule fires, and it ends the step below the surface
    p["vel"][1] = (-1.5, 0.75, 0.25, 0.0)
    p["force"][0] = (1000.0, -2000.0, 3000.0, 7.0)      # last frame's force: feeds the torque term (and is scaled by the crest rule)
    p["force"][1] = (-500.0, 250.0, 125.0, 3.0)
    p["extras"][:] = (RHO0, 0.0, 500.0, 50.0)
    return p


def expected_rho_pres(p):
    pos = p["pos"][:, :3].astype(np.float64)
    r = np.linalg.norm(pos[0] - pos[1])
    w = lambda rr: MASS * 315.0 * (H_S * H_S - rr * rr) ** 3 / (64.0 * PI * H_S ** 9)      # :62-67
    out = []
    for i in range(2):
        rho = w(0.0) + w(r)                                                                # self included
        pres = max(GAS * (rho - RHO0), 0.0)                                                # :70
        height = tex(2.0 * pos[i, 0], 2.0 * pos[i, 2])                                     # :72-73
        wave_force = height * rho                                                          # :74
        pres += wave_force                                                                 # :75
        rho += wave_force / (GAS * R)                                                      # :76
        out.append((max(RHO0, rho), pres))                                                 # :79-80
    return np.array(out), r


def expected_force(p, rho_pres):
    pos = p["pos"][:, :3].astype(np.float64); vel = p["vel"][:, :3].astype(np.float64); fprev = p["force"].astype(np.float64)
    spiky = -20.0 / (PI * H_S ** 6)
    lap = -spiky
    out = []
    for i in range(2):
        j = 1 - i
        d = pos[i] - pos[j]
        r = np.linalg.norm(d)
        pres = -MASS * (rho_pres[i, 1] + rho_pres[j, 1]) / (2.0 * rho_pres[j, 0]) * spiky * (H_S - r) ** 2 * d / r      # :85
        visc = MASS * (vel[j] - vel[i]) / rho_pres[j, 0] * lap * (H_S - r)                                            # :86
        f_mem = fprev[i].copy()
        if pos[i, 1] > 0.01:                                                                                         # :91-95
            f_mem /= 0.25
            visc *= 0.5
        visc *= VISC                                                                                                 # :97
        s, t = 2.0 * pos[i, 0], 2.0 * pos[i, 2]
        height = tex(s, t)
        torque = 0.25 * np.cross(pos[i], f_mem[:3])                                                                  # :103-104
        hx, hy = tex(s + 0.01, t), tex(s, t + 0.01)                                                                  # WaveVelocity :117-128
        wv = np.array([(hx - height) / DT, (hy - height) / DT, (hx - hy) / 0.01])
        drag = -0.25 * (vel[i] - wv)                                                                                 # :107-108
        normal = np.array([tex(s + 1.0, t) - height, tex(s, t + 1.0) - height, -1.0])                                # WaveNormal :131-139
        wave = -height * normal * 0.5                                                                                # :110
        grav = rho_pres[i, 0] * np.array([0.0, G_Y, 0.0])                                                            # :113
        out.append(np.concatenate([pres + visc + grav + torque + drag + wave, [f_mem[3]]]))                          # :114 (force.w: only the crest rule touches it)
    return np.array(out)


def expected_integrate(p, rho_pres, force):
    pos = p["pos"][:, :3].astype(np.float64); vel = p["vel"][:, :3].astype(np.float64)
    out = []
    for i in range(2):
        acc = force[i, :3] / rho_pres[i, 0]                                    # :62
        nv = vel[i] + DT * acc                                                 # :63
        npos = pos[i] + DT * nv                                                # :64
        nv = nv * (1.0 - DAMPING * DT)                                         # :66
        f, rho, prs = force[i].copy(), rho_pres[i, 0], rho_pres[i, 1]
        if np.linalg.norm(nv) > 25.0:                                          # :69-76
            f *= 0.5; rho *= 0.1; prs *= 0.25; nv = nv * 0.1
        th = tex(2.0 * npos[0], 2.0 * npos[2])                                 # :79
        if npos[1] < th:
            npos[1] = th - R                                                   # :80-83
        out.append((npos, nv, f, rho, prs))                                    # (the box is far away: CheckBoundary does nothing here)
    return out


def check(got, what, tol=2e-5):
    exp_rp, r = expected_rho_pres(what["p0"])
    assert 0.0 < r < H_S
    if "rho" in got:
        for i in range(2):
            assert abs(got["rho"][i] - exp_rp[i, 0]) <= tol * exp_rp[i, 0], ("rho", i, got["rho"][i], exp_rp[i, 0])
            assert abs(got["pres"][i] - exp_rp[i, 1]) <= tol * exp_rp[i, 1], ("pressure", i, got["pres"][i], exp_rp[i, 1])
    exp_f = expected_force(what["p0"], exp_rp)
    if "force" in got:
        for i in range(2):
            scale = np.linalg.norm(exp_f[i, :3])
            assert np.abs(got["force"][i, :3] - exp_f[i, :3]).max() <= tol * scale, ("force", i, got["force"][i], exp_f[i])
            assert got["force"][i, 3] == exp_f[i, 3], "force.w: x4 above the crest, untouched below"
    if "pos" in got:
        exp_i = expected_integrate(what["p0"], exp_rp, exp_f)
        for i in range(2):
            npos, nv, f, rho, prs = exp_i[i]
            assert np.abs(got["pos"][i] - npos).max() <= tol * 0.1, ("pos", i, got["pos"][i], npos)
            assert np.abs(got["vel"][i] - nv).max() <= tol * np.linalg.norm(nv), ("vel", i, got["vel"][i], nv)
            assert abs(got["rho_after"][i] - rho) <= tol * rho and abs(got["pres_after"][i] - prs) <= tol * prs
            assert np.abs(got["force_after"][i] - f).max() <= tol * np.linalg.norm(f[:3])


def test_hand_derived_values_exercise_every_branch():
    """The scene is built so that every rule of the three shaders fires at least once."""
    dt = np.dtype([("pos", "<f4", 4), ("vel", "<f4", 4), ("force", "<f4", 4), ("extras", "<f4", 4)])
    p = particles(dt)
    rp, r = expected_rho_pres(p)
    f = expected_force(p, rp)
    it = expected_integrate(p, rp, f)
    assert rp[0, 0] > RHO0 and rp[0, 1] > 0                                   # EOS pressure active
    assert f[1, 3] == 12.0 and f[0, 3] == 7.0                                 # crest rule on particle 1 only
    assert it[0][3] == pytest.approx(rp[0, 0] * 0.1) and it[1][3] == pytest.approx(rp[1, 0])       # foam rule: particle 0 only
    th0, th1 = tex(2.0 * it[0][0][0], 2.0 * it[0][0][2]), tex(2.0 * it[1][0][0], 2.0 * it[1][0][2])
    assert it[0][0][1] == pytest.approx(th0 - R) and it[1][0][1] > th1                           # surface clamp: particle 0 only
    assert abs(f[0, 0]) > 1e6 and abs(tex(0.2, 0.2)) > 1e-3                                      # pair forces and wave terms are not degenerate


def test_oracle_against_hand_derived_values(oracle):
    prm = oracle.default_params3()
    p0 = particles(oracle.PARTICLE) if hasattr(oracle, "PARTICLE") else particles(oracle.make_cube(1, 1, 1, prm).dtype)
    t = field()
    q = p0.copy()
    oracle.sph3_rho_pres(q, prm, t)
    got = {"rho": q["extras"][:, 0].astype(np.float64), "pres": q["extras"][:, 1].astype(np.float64)}
    oracle.sph3_force(q, prm, t)
    got["force"] = q["force"].astype(np.float64)
    oracle.sph3_integrate(q, prm, t)
    got.update(pos=q["pos"][:, :3].astype(np.float64), vel=q["vel"][:, :3].astype(np.float64), rho_after=q["extras"][:, 0].astype(np.float64),
               pres_after=q["extras"][:, 1].astype(np.float64), force_after=q["force"].astype(np.float64))
    check(got, {"p0": p0})


def test_oracle_sampler_is_exact_on_a_linear_field(oracle):
    """GL_LINEAR of a linear field reproduces it; CLAMP_TO_EDGE outside: the sampler both code paths implement by hand."""
    t = field()
    rng = np.random.default_rng(5)
    for s, tt in rng.uniform(-0.3, 1.6, (200, 2)):
        assert abs(oracle.tex_bilinear(t, float(s), float(tt)) - tex(s, tt)) <= 2e-6 * max(1.0, abs(tex(s, tt)))


@pytest.mark.gpu
@pytest.mark.parametrize("grid", [False, True])
def test_cuda_against_hand_derived_values(cwa, ctx, oracle, grid):
    """The CUDA passes against the same closed forms -- all-pairs as shipped and on a uniform grid (row-mask kernels)."""
    prm = oracle.default_params3()
    ctx.set_params_from_oracle(prm)
    p0 = particles(cwa.PARTICLE)
    g = cwa.UniformGrid(ctx, 3, (0.0, -0.02, 0.0), (0.26, 0.18, 0.26), (25, 19, 25), 2) if grid else None
    sph = cwa.Sph(ctx, 2, g, particles=p0)
    wave = cwa.StencilImage2DTripleBuffered(ctx, W, H, 1, cwa.WAVE_COUPLED)
    t = field()
    for im in range(3):
        wave.write_image(im, t)
    sph.bind_wave(wave, 0)
    sph.rho_pres()
    q = sph.download()
    got = {"rho": q["extras"][:, 0].astype(np.float64), "pres": q["extras"][:, 1].astype(np.float64)}
    sph.force()
    q = sph.download()
    got["force"] = q["force"].astype(np.float64)
    sph.integrate()
    q = sph.download()
    got.update(pos=q["pos"][:, :3].astype(np.float64), vel=q["vel"][:, :3].astype(np.float64), rho_after=q["extras"][:, 0].astype(np.float64),
               pres_after=q["extras"][:, 1].astype(np.float64), force_after=q["force"].astype(np.float64))
    check(got, {"p0": p0})
