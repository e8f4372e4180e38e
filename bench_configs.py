"""BASELINE.json configs 1-3 for bench.py (`--config D | C2 | C3`); config 4 (C4) and 5 (C5, --gpus N) live in bench.py itself.

  D   the shipped scene (CoupledWaterAnimation/Main.cpp:28-35,184-204): 20 480 particles, all-pairs passes, 64^2 RGBA wave,
      AS_SHIPPED coupling.  Six launches of a few microseconds each: the frame runs as ONE CUDA graph launch (cwa_coupled_step).
  C2  SphWave2D Koschier 2-D SPH on the uniform grid + prefix scan (SphWave2D/StencilBuffer.cpp:138-179), 65 536 particles,
      128 x 32 cells, 2 substeps per frame.
  C3  Wave2DSimp 4096^2 triple-buffered stencil (Wave2DSimp/Wave2D_cs.glsl:78-103), scalar (R32F) and RGBA32F-compatible.

Every function returns one JSON-ready dict with the keys of bench.py's line: value (state resident in HBM), e2e (through the C ABI with
pinned HOST buffers copied in and out every step), roofline of the dominant kernel, cpu_baseline (the oracle on the host cores, bounded
sample) and gpu_launches.  `brief=True` skips the CPU leg and shortens the runs (used for the "other_configs" rows of the default line).
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # SMs x FP32 lanes x 2 flop x max SM clock: 74.4 TFLOP/s (non-tensor FP32; not in MEASURED_PEAKS.json)


def _timed(ctx, fn, k):
    ctx.synchronize()
    l0 = ctx.launch_count
    ctx.timer_begin()
    fn(k)
    ms = ctx.timer_end()
    return ms, ctx.launch_count - l0


def _kernels(prof, k, algo, peak):
    out = []
    tot = sum(v[0] for v in prof.values()) or 1.0
    for name, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        b = algo.get(name, 0)
        avg = ms / cnt
        gbs = b / (avg * 1e-3) / 1e9 if avg > 0 else 0.0
        out.append({"kernel": name, "launches": cnt, "avg_us": avg * 1e3, "share": ms / tot, "algo_bytes": b, "achieved_gbs": gbs, "frac": gbs / peak})
    return out


def _pin(torch, nbytes):
    return torch.empty(nbytes, dtype=torch.uint8).pin_memory()


# ---------------------------------------------------------------------------------------------------------------------------------
def run_D(cwa, ctx, torch, K, W, peak, peak_src, brief=False, oracle=None):
    n, nx, ny, nz, wv = 20480, 64, 5, 64, 64
    sph = cwa.Sph(ctx, n, None)
    sph.init_cube(nx, ny, nz)
    wave = cwa.StencilImage2DTripleBuffered(ctx, wv, wv, 4, cwa.WAVE_COUPLED)
    cpl = cwa.COUPLING_AS_SHIPPED
    sph.coupled_step(wave, W, cpl)
    ms, launches = _timed(ctx, lambda k: sph.coupled_step(wave, k, cpl), K)
    ctx.set_tuning(graph=0)
    ms_nograph, _ = _timed(ctx, lambda k: sph.coupled_step(wave, k, cpl), K)
    ctx.profile_begin()
    sph.coupled_step(wave, K, cpl)
    prof = ctx.profile_end()
    ctx.set_tuning(graph=1)
    # all-pairs passes: n^2 pair tests each; SURVEY 8d: ~12 flop per density test, ~15 per force test (+ the accepted pairs, a few per target)
    flops = {"density": 12.0 * n * n, "force": 15.0 * n * n}
    kern = _kernels(prof, K, {"integrate": 96 * n, "wave_evolve": 48 * wv * wv}, peak)
    dom = kern[0]
    tf = flops.get(dom["kernel"], 0.0) / (dom["avg_us"] * 1e-6) / 1e12
    roofline = {"kernel": dom["kernel"], "bound": "fp32 (non-tensor; no pass is a dense contraction)", "achieved": tf, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": tf / FP32_PEAK_TFLOPS, "traffic": None, "peak_source": "148 SMs x 128 lanes x 2 x 1.965 GHz (nominal; MEASURED_PEAKS.json holds no FP32 figure)",
                "note": f"{n}^2 pair tests per pass at ~{flops.get(dom['kernel'], 0) / n / n:.0f} flop each (SURVEY 8d) -- EFFECTIVE rate: candidate tiles whose bounding box is farther than h from a CTA's targets are skipped, exactly (tuning allpairs_cull); the frame is 7 launches, replayed as one CUDA graph"}
    # end to end: particles + the two wave levels the stencil reads go up, one frame, particles + the new level come back
    lib, h = ctx.lib, ctx.h
    pp = _pin(torch, n * 64); hp = pp.numpy().view(cwa.PARTICLE)
    pw = [_pin(torch, wv * wv * 16) for _ in range(2)]; hw = [w_.numpy().view(np.float32).reshape(wv, wv, 4) for w_ in pw]
    hp[:] = sph.download(); hw[0][:] = wave.read_role(0); hw[1][:] = wave.read_role(1)

    def e2e_step():
        cwa.check(lib.cwa_buffer_sub_data(h, sph.buffer.h, 0, hp.nbytes, C.c_void_p(hp.ctypes.data)))
        cwa.check(lib.cwa_wave_write_image(h, wave.h, wave.role_image(0), C.c_void_p(hw[0].ctypes.data)))
        cwa.check(lib.cwa_wave_write_image(h, wave.h, wave.role_image(1), C.c_void_p(hw[1].ctypes.data)))
        cwa.check(lib.cwa_coupled_step(h, sph.h, wave.h, 1, cpl))
        cwa.check(lib.cwa_buffer_read_async(h, sph.buffer.h, 0, hp.nbytes, C.c_void_p(hp.ctypes.data)))
        hw[0], hw[1] = hw[1], hw[0]
        cwa.check(lib.cwa_wave_read_image_async(h, wave.h, wave.role_image(0), C.c_void_p(hw[0].ctypes.data)))
        ctx.synchronize()

    for _ in range(3):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    cpu = None
    if not brief and oracle is not None:
        prm = oracle.default_params3()
        oc = oracle.Coupled(n, wv, wv, 4, prm, oracle.COUPLING_AS_SHIPPED)
        oc.particles[:] = oracle.make_cube(nx, ny, nz, prm)
        oc.step(2)
        nc = 40
        t0 = time.perf_counter(); oc.step(nc); dt = time.perf_counter() - t0
        oc.close()
        cores = oracle.lib().orc_num_threads()
        cpu = {"value": n * nc / dt, "unit": "particle-updates/s", "cores": cores, "kind": "port", "ms_per_step": dt / nc * 1e3,
               "sample": f"{nc} full frames of the shipped scene after 2 warm-up frames, CPU restatement of the reference GLSL (OpenMP, {cores} threads)"}
    line = {"metric": "particle_updates_per_sec", "value": n * K / (ms * 1e-3), "unit": "particle-updates/s", "ms_per_step": ms / K, "steps_per_sec": K / (ms * 1e-3),
            "ms_per_step_without_graph": ms_nograph / K, "dtype": "f32",
            "config": {"workload": f"D: the shipped scene, {n} particles (64x5x64 lattice) all-pairs + {wv}^2 RGBA32F wave, coupling AS_SHIPPED (Main.cpp:28-35,184-204)",
                       "l2": "working set 1.5 MB: L2-resident by nature (the reference's own configuration; compute-bound all-pairs loops)",
                       "timing": "cudaEvent on the context stream around K frames, each frame one CUDA graph launch"},
            "e2e": {"value": n * K / e2e_s, "unit": "particle-updates/s", "h2d_bytes_per_step": n * 64 + 2 * wv * wv * 16, "d2h_bytes_per_step": n * 64 + wv * wv * 16,
                    "ms_per_step": e2e_s / K * 1e3, "how": "C ABI, pinned host buffers: particles + two wave levels up, one coupled frame, particles + the new level back, synchronised every step"},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_kernels": kern, "cpu_baseline": cpu}
    sph.destroy(); wave.destroy()
    return line


# ---------------------------------------------------------------------------------------------------------------------------------
def run_C2(cwa, ctx, torch, K, W, peak, peak_src, brief=False, oracle=None):
    n, ext = 65536, ((0.0, 0.0), (38.4, 9.6), (128, 32))
    grid = cwa.UniformGrid(ctx, 2, *ext, n)
    s = cwa.SphUgrid(ctx, n, grid, cwa.SPH2_WAVE, substeps=2)
    s.set_uniforms(init_width=512, view_width=38.4)
    s.Reinit()
    s.Compute(W)
    ms, launches = _timed(ctx, lambda k: s.Compute(k), K)
    ctx.profile_begin()
    s.Compute(K)
    prof = ctx.profile_end()
    cells = 128 * 32
    # per substep: grid build (hash 20 B + insert 16 B per particle, scan 8 B per cell), two reorders (48 B in + 48 B out + 8), density / forces read+write one record
    algo = {"clear(memset)": 4 * cells, "grid_hash_count": 20 * n, "scan_lookback": 8 * cells, "grid_insert": 16 * n, "grid_cell_order": 16 * n,
            "reorder": 104 * n, "density": 96 * n, "force": 96 * n}
    kern = _kernels(prof, K, algo, peak)
    dom = kern[0]
    roofline = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": None,
                "peak_source": peak_src, "note": "65 536 particles (3 MB of records) are L2-resident and every kernel lasts a few microseconds: the frame is launch- and latency-bound, "
                                                 "the GB/s figure is algorithmic bytes over the kernel's CUDA-event time"}
    lib, h = ctx.lib, ctx.h
    pp = _pin(torch, n * 48); hp = pp.numpy().view(cwa.PARTICLE2D)
    hp[:] = s.download()

    def e2e_step():
        cwa.check(lib.cwa_sph2_write(h, s.h, C.c_void_p(hp.ctypes.data)))
        cwa.check(lib.cwa_sph2_compute(h, s.h, 1))
        cwa.check(lib.cwa_sph2_read(h, s.h, C.c_void_p(hp.ctypes.data)))

    for _ in range(3):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    ctx.synchronize()
    e2e_s = time.perf_counter() - t0
    cpu = None
    if not brief and oracle is not None:
        prm = oracle.default_params2(1)
        prm.init_width = 512
        prm.view_width = 38.4
        b0 = oracle.sph2_init(n, prm); b1 = np.zeros_like(b0)
        g = oracle.grid2(*ext)
        r, _ = oracle.sph2_step(b0, b1, 0, 2, prm, None, g)
        nc = 40
        t0 = time.perf_counter()
        for _ in range(nc):
            r, _ = oracle.sph2_step(b0, b1, r, 2, prm, None, g)
        dt = time.perf_counter() - t0
        cores = oracle.lib().orc_num_threads()
        cpu = {"value": n * nc / dt, "unit": "particle-updates/s", "cores": cores, "kind": "port", "ms_per_step": dt / nc * 1e3,
               "sample": f"{nc} frames (2 substeps each) after 1 warm-up frame, CPU restatement of SphWaveKoschier2D_grid_cs + uniform_grid_sph_cs + prefix_sum_cs ({cores} threads)"}
    line = {"metric": "particle_updates_per_sec", "value": n * K / (ms * 1e-3), "unit": "particle-updates/s (one update = one frame = 2 substeps)", "ms_per_step": ms / K,
            "steps_per_sec": K / (ms * 1e-3), "dtype": "f32",
            "config": {"workload": f"C2: SphWave2D Koschier 2-D SPH (wave variant) on the uniform grid + prefix scan, {n} particles (512x128 lattice), 128x32 cells of 0.3, 2 substeps per frame",
                       "l2": "working set 6 MB: L2-resident by nature of the configuration", "timing": "cudaEvent on the context stream around K frames"},
            "e2e": {"value": n * K / e2e_s, "unit": "particle-updates/s", "h2d_bytes_per_step": n * 48, "d2h_bytes_per_step": n * 48, "ms_per_step": e2e_s / K * 1e3,
                    "how": "C ABI, pinned host buffer: particle records up, one frame, records back (blocking read)"},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_kernels": kern, "cpu_baseline": cpu}
    s.destroy(); grid.destroy()
    return line


# ---------------------------------------------------------------------------------------------------------------------------------
def run_C3(cwa, ctx, torch, K, W, peak, peak_src, brief=False, oracle=None, size=4096):
    rows = {}
    for ch in (1, 4):
        wave = cwa.StencilImage2DTripleBuffered(ctx, size, size, ch, cwa.WAVE_SIMP)
        wave.Compute(W)
        ms, launches = _timed(ctx, lambda k: wave.Compute(k), K)
        rows[ch] = (ms, launches)
        if ch == 1:
            ctx.profile_begin()
            wave.Compute(K)
            prof = ctx.profile_end()
            # end to end: the two levels the stencil reads go up, one step, the new level comes back
            lib, h = ctx.lib, ctx.h
            pw = [_pin(torch, size * size * 4) for _ in range(2)]
            hw = [w_.numpy().view(np.float32).reshape(size, size) for w_ in pw]
            hw[0][:] = wave.read_role(0); hw[1][:] = wave.read_role(1)

            def e2e_step():
                cwa.check(lib.cwa_wave_write_image(h, wave.h, wave.role_image(0), C.c_void_p(hw[0].ctypes.data)))
                cwa.check(lib.cwa_wave_write_image(h, wave.h, wave.role_image(1), C.c_void_p(hw[1].ctypes.data)))
                cwa.check(lib.cwa_wave_compute(h, wave.h, 1))
                hw[0], hw[1] = hw[1], hw[0]
                cwa.check(lib.cwa_wave_read_image_async(h, wave.h, wave.role_image(0), C.c_void_p(hw[0].ctypes.data)))
                ctx.synchronize()

            for _ in range(2):
                e2e_step()
            ne = max(3, min(K, 20))
            t0 = time.perf_counter()
            for _ in range(ne):
                e2e_step()
            e2e_s = (time.perf_counter() - t0) / ne
        wave.destroy()
    cells = size * size
    ms1, launches = rows[1]
    kern = _kernels(prof, K, {"wave_evolve": 12 * cells}, peak)
    dom = kern[0]
    roofline = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": None, "peak_source": peak_src,
                "note": "12 B per cell: u(t-1) read once (TMA halo tiles), u(t-2) read, u(t) written; three 64 MB levels = 192 MB working set > 126 MB L2"}
    cpu = None
    if not brief and oracle is not None:
        u0 = oracle.wave_init(size, size, 1, oracle.WAVE_SIMP); u1 = u0.copy()
        nc = 10
        t0 = time.perf_counter()
        for _ in range(nc):
            u0, u1 = oracle.wave_evolve(u0, u1, oracle.WAVE_SIMP, 0.01, 0.9995, 0.001), u0
        dt = time.perf_counter() - t0
        cores = oracle.lib().orc_num_threads()
        cpu = {"value": cells * nc / dt, "unit": "cell-updates/s", "cores": cores, "kind": "port", "ms_per_step": dt / nc * 1e3,
               "sample": f"{nc} steps of the {size}^2 field, CPU restatement of Wave2D_cs.glsl ({cores} threads)"}
    line = {"metric": "wave_cell_updates_per_sec", "value": cells * K / (ms1 * 1e-3), "unit": "cell-updates/s", "ms_per_step": ms1 / K, "steps_per_sec": K / (ms1 * 1e-3), "dtype": "f32",
            "rgba32f_compat": {"value": cells * K / (rows[4][0] * 1e-3), "ms_per_step": rows[4][0] / K, "note": "the as-shipped 4-channel layout, 48 B per cell"},
            "config": {"workload": f"C3: Wave2DSimp {size}^2 triple-buffered wave stencil (Wave2D_cs.glsl:78-103), scalar R32F; RGBA32F-compatible row beside it",
                       "l2": "three 64 MB levels (192 MB) > 126 MB L2: inputs larger than L2, no flush needed", "timing": "cudaEvent on the context stream around K steps"},
            "e2e": {"value": cells / e2e_s, "unit": "cell-updates/s", "h2d_bytes_per_step": 2 * cells * 4, "d2h_bytes_per_step": cells * 4, "ms_per_step": e2e_s * 1e3,
                    "how": "C ABI, pinned host buffers: two levels up, one step, the new level back, synchronised every step"},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_kernels": kern, "cpu_baseline": cpu}
    return line


RUNNERS = {"D": run_D, "C2": run_C2, "C3": run_C3}


def brief_rows(cwa, ctx, torch, peak, peak_src):
    """Short runs of configs 1-3 for the `other_configs` key of the default (C4) line: value, e2e and the dominant kernel's fraction."""
    out = {}
    for name, (k, w) in (("D", (200, 20)), ("C2", (100, 10)), ("C3", (50, 5))):
        try:
            r = RUNNERS[name](cwa, ctx, torch, k, w, peak, peak_src, brief=True)
            out[name] = {"workload": r["config"]["workload"], "metric": r["metric"], "value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"],
                         "e2e_value": r["e2e"]["value"], "e2e_ms_per_step": r["e2e"]["ms_per_step"], "gpu_launches": r["gpu_launches"], "steps": k,
                         "roofline": {key: r["roofline"][key] for key in ("kernel", "bound", "achieved", "peak", "unit", "frac")}}
            if "rgba32f_compat" in r:
                out[name]["rgba32f_compat"] = r["rgba32f_compat"]
            if "ms_per_step_without_graph" in r:
                out[name]["ms_per_step_without_graph"] = r["ms_per_step_without_graph"]
        except Exception as e:                      # a side row must never take the headline line down
            out[name] = {"error": str(e)[:200]}
    return out


def cpu_reference(config: str, oracle, steps: int):
    """The reference arm of configs 1-3 (`bench.py --impl reference --config X`): the CPU restatement of the reference on the host cores,
    bounded to `steps` steps.  Returns (metric, unit, value, ms_per_step, cores, sample, workload)."""
    cores = oracle.lib().orc_num_threads()
    if config == "D":
        n, wv = 20480, 64
        prm = oracle.default_params3()
        oc = oracle.Coupled(n, wv, wv, 4, prm, oracle.COUPLING_AS_SHIPPED)
        oc.particles[:] = oracle.make_cube(64, 5, 64, prm)
        oc.step(2)
        t0 = time.perf_counter(); oc.step(steps); dt = time.perf_counter() - t0
        oc.close()
        return ("particle_updates_per_sec", "particle-updates/s", n * steps / dt, dt / steps * 1e3, cores,
                f"{steps} full frames of the shipped scene after 2 warm-up frames, OpenMP on {cores} host threads",
                f"D: the shipped scene, {n} particles (64x5x64 lattice) all-pairs + {wv}^2 RGBA32F wave, coupling AS_SHIPPED (Main.cpp:28-35,184-204)")
    if config == "C2":
        n, ext = 65536, ((0.0, 0.0), (38.4, 9.6), (128, 32))
        prm = oracle.default_params2(1)
        prm.init_width = 512
        prm.view_width = 38.4
        b0 = oracle.sph2_init(n, prm); b1 = np.zeros_like(b0)
        g = oracle.grid2(*ext)
        r, _ = oracle.sph2_step(b0, b1, 0, 2, prm, None, g)
        t0 = time.perf_counter()
        for _ in range(steps):
            r, _ = oracle.sph2_step(b0, b1, r, 2, prm, None, g)
        dt = time.perf_counter() - t0
        return ("particle_updates_per_sec", "particle-updates/s (one update = one frame = 2 substeps)", n * steps / dt, dt / steps * 1e3, cores,
                f"{steps} frames (2 substeps each) after 1 warm-up frame, {cores} host threads",
                f"C2: SphWave2D Koschier 2-D SPH (wave variant) on the uniform grid + prefix scan, {n} particles (512x128 lattice), 128x32 cells of 0.3, 2 substeps per frame")
    if config == "C3":
        size = 4096
        u0 = oracle.wave_init(size, size, 1, oracle.WAVE_SIMP); u1 = u0.copy()
        t0 = time.perf_counter()
        for _ in range(steps):
            u0, u1 = oracle.wave_evolve(u0, u1, oracle.WAVE_SIMP, 0.01, 0.9995, 0.001), u0
        dt = time.perf_counter() - t0
        return ("wave_cell_updates_per_sec", "cell-updates/s", size * size * steps / dt, dt / steps * 1e3, cores,
                f"{steps} steps of the {size}^2 field, {cores} host threads",
                f"C3: Wave2DSimp {size}^2 triple-buffered wave stencil (Wave2D_cs.glsl:78-103), scalar R32F; RGBA32F-compatible row beside it")
    raise ValueError(config)
