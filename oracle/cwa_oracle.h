/*
 * cwa_oracle.h -- CPU restatement ("oracle") of the CoupledWaterAnimation simulation step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (coupledwateranimation_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or as the
 * timed CPU baseline.
 *
 * Every function cites the reference file:line it restates (paths relative to the reference
 * repository root).  The reference's device code is GLSL and cannot be executed in this image
 * (no GL stack), so:
 *   - scan:  pinned by the reference's own known-answer tests (SphWave2D/ParallelScan.cpp:124-155,
 *            UniformGrid2D/ParallelScan.cpp:117);
 *   - grid:  pinned against the reference's CPU twin UniformGrid2D::Build compiled from its own
 *            source into oracle/_ref/ (see oracle/Makefile, oracle/ref_grid_shim.cpp);
 *   - SPH passes, wave stencil, bilinear sampler:  PARITY UNPINNED -- the reference holds no test,
 *            golden file or recorded output for them; this restatement is the pin.
 *
 * Canonical floating-point choices (shared with the CUDA kernels, documented in DESIGN.md):
 *   - length(v)   = sqrtf(fmaf(v.z,v.z, fmaf(v.y,v.y, v.x*v.x)))          (3-D)
 *                   sqrtf(fmaf(v.y,v.y, v.x*v.x))                          (2-D)
 *     so that the neighbour acceptance test r < h is bit-identical on CPU and GPU;
 *   - pow(x,2), pow(x,3), pow(h,6), pow(h,9) = repeated FP32 multiplication;
 *   - everything else is evaluated in the GLSL's written association order, FP32, no contraction
 *     (compile with -ffp-contract=off);
 *   - float -> cell index conversion clamps in the float domain first (NaN -> cell 0).
 */
#ifndef CWA_ORACLE_H
#define CWA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- a-0: particle structs ------------------------------------------------------------- */
/* CoupledWaterAnimation/Main.cpp:167-173, rho_pres_comp.glsl:14-20 (std430, 64 B) */
typedef struct { float pos[4], vel[4], force[4], extras[4]; } orc_particle3;
/* SphWave2D/Main.cpp:35-40, SphWaveKoschier2D_grid_cs.glsl:39-44 (48 B) */
typedef struct { float pos[4], vel[4], acc[4]; } orc_particle2;

/* ---- a-0': parameter blocks ------------------------------------------------------------ */
/* ConstantsUniform / BoundaryUniform / WaveUniforms, CoupledWaterAnimation/Main.cpp:184-204.
 * The trailing fields are shader constants in the reference (rho_pres_comp.glsl:5,41,43;
 * force_comp.glsl:7-9,52-55; integrate_comp.glsl:8,51) that the north-star turns into
 * parameters; defaults reproduce the reference exactly. */
typedef struct {
    /* ConstantsUniform (binding 1) */
    float mass, smoothing_coeff, visc, resting_rho;
    /* BoundaryUniform (binding 2) */
    float upper[4], lower[4];
    /* WaveUniforms (binding 3) */
    float attributes[4], mesh_ws_pos[4];
    /* shader constants */
    float particle_radius;   /* PARTICLE_RADIUS 0.005 */
    float gas_const;         /* GAS_CONST 4000 */
    float dt;                /* 5e-5 */
    float gravity_y;         /* -9806.65 */
    float damping;           /* DAMPING 0.3 */
    float crest_threshold;   /* CREST_THRESHOLD 0.01 */
    float foam_speed;        /* 25.0 */
    float uv_scale;          /* 2.0: coord = 2*pos.xz */
    float uv_scale_z;        /* extension: t = uv_scale_z * pos.z; 0 = uv_scale (the reference) */
    float torque_coeff;      /* extension: force_comp.glsl:103 literal 0.25; 0 = 0.25 (the reference) */
} orc_params3;

void orc_params3_default(orc_params3* p);

/* ---- texture (A.3) ---------------------------------------------------------------------- */
typedef struct { const float* data; int w, h, ch; } orc_tex;   /* data==NULL: unbound -> 0 */
float orc_tex_bilinear(const orc_tex* t, float s, float tt);

/* ---- a-2: scan --------------------------------------------------------------------------- */
int  orc_scan_blelloch(int* x, int n);                 /* in place; -1 if n not a power of 2 */
void orc_scan_exclusive(const int* in, int* out, int n);

/* ---- a-1, a-3: grid ---------------------------------------------------------------------- */
typedef struct { float min[2], max[2]; int ncells[2]; float cell[2]; } orc_grid2;
typedef struct { float min[4], max[4]; int ncells[4]; float cell[4]; } orc_grid3;
void orc_grid2_init(orc_grid2* g, const float mn[2], const float mx[2], const int n[2]);
void orc_grid3_init(orc_grid3* g, const float mn[3], const float mx[3], const int n[3]);
int  orc_grid2_cell_index(const orc_grid2* g, float x, float y);   /* linear index */
int  orc_grid3_cell_index(const orc_grid3* g, float x, float y, float z);
/* pos: pointer to first particle's pos.x; stride in floats.  cell_of[i] = -1 if not inserted */
void orc_grid2_build(const orc_grid2* g, const float* pos, int stride, int n,
                     int* cell_of, int* counter, int* offset, int* index_list);
void orc_grid3_build(const orc_grid3* g, const float* pos, int stride, int n,
                     int* cell_of, int* counter, int* offset, int* index_list);

/* ---- a-7, a-8: wave ---------------------------------------------------------------------- */
enum { ORC_WAVE_COUPLED = 0 /* wave_comp.glsl */, ORC_WAVE_SIMP = 1 /* Wave2D_cs.glsl */ };
void orc_wave_init(float* out, int w, int h, int ch, int variant, float type);
void orc_wave_evolve(const float* u0, const float* u1, float* out, int w, int h, int ch,
                     int variant, float lambda, float atten, float beta, float type);

/* ---- a-4, a-5, a-6: 3-D SPH passes (in place on the particle array, like the shaders) --- */
/* grid == NULL -> all-pairs loops exactly as shipped; else neighbour search through the grid
 * lists (counter/offset/index_list from orc_grid3_build on the same positions). */
void orc_sph3_rho_pres(orc_particle3* p, int n, const orc_params3* prm, const orc_tex* tex,
                       const orc_grid3* grid, const int* counter, const int* offset,
                       const int* index_list);
void orc_sph3_force(orc_particle3* p, int n, const orc_params3* prm, const orc_tex* tex,
                    const orc_grid3* grid, const int* counter, const int* offset,
                    const int* index_list);
void orc_sph3_integrate(orc_particle3* p, int n, const orc_params3* prm, const orc_tex* tex);
/* neighbour count (r < h, self included) -- for the "neighbour set identical" assertions */
void orc_sph3_neighbour_count(const orc_particle3* p, int n, float h, const orc_grid3* grid,
                              const int* counter, const int* offset, const int* index_list,
                              int* out_count);

/* ---- a-10: scene init --------------------------------------------------------------------- */
/* make_cube/init_particles, Main.cpp:735-776, generalised to nx*ny*nz (shipped: 64,5,64) */
void orc_make_cube(orc_particle3* p, int nx, int ny, int nz, const orc_params3* prm);

/* ---- a-8, a-9: coupled driver -------------------------------------------------------------- */
enum { ORC_COUPLING_AS_SHIPPED = 0, ORC_COUPLING_LATEST = 1 };
typedef struct orc_coupled orc_coupled;
orc_coupled* orc_coupled_create(int n, int wave_w, int wave_h, int wave_ch,
                                const orc_params3* prm, int coupling,
                                int use_grid, const float gmin[3], const float gmax[3],
                                const int gn[3]);
void orc_coupled_destroy(orc_coupled* c);
orc_particle3* orc_coupled_particles(orc_coupled* c);
/* role 0 = newest (u^t), 1 = previous, 2 = next write target */
float* orc_coupled_wave(orc_coupled* c, int role);
void orc_coupled_wave_reinit(orc_coupled* c);            /* StencilImage2DTripleBuffered::Reinit */
void orc_coupled_step(orc_coupled* c, int nframes);      /* idle() + display() bind, per frame */
int  orc_coupled_sampled_image(const orc_coupled* c);    /* physical image bound to tex unit 0, -1 none */
void orc_coupled_set_params(orc_coupled* c, const orc_params3* prm);

/* ---- a-4b, a-5b: 2-D Koschier SPH on the grid ---------------------------------------------- */
enum { ORC_SPH2_KOSCHIER = 0 /* SphKoschier2D_grid_cs */, ORC_SPH2_WAVE = 1 /* SphWaveKoschier2D_grid_cs */ };
typedef struct {
    int   variant;
    float time;        /* uniform time (orbiting circle in variant 0) */
    float bottom;      /* uniform bottom = 0.3 (variant 1) */
    float psi;         /* uniform PSI; <0 -> default REST_DENS/(1.5*k2) */
    int   init_width;  /* const WIDTH of the init lattice: 32 (variant 0) / 128 (variant 1) */
    float view_width;  /* const VIEW_WIDTH = 9.6; a parameter so config C2 can widen the tank */
} orc_params2;
void orc_params2_default(orc_params2* p, int variant);
/* 1-D wave texture for variant 1: RGBA32F, w texels (h=1); data==NULL -> unbound (0,0,0,1) */
void orc_sph2_init(orc_particle2* out, int n, const orc_params2* prm);
void orc_sph2_density(const orc_particle2* in, orc_particle2* out, int n, const orc_params2* prm,
                      const orc_tex* wave1d, const orc_grid2* g, const int* counter,
                      const int* offset, const int* index_list);
void orc_sph2_forces(const orc_particle2* in, orc_particle2* out, int n, const orc_params2* prm,
                     const orc_tex* wave1d, const orc_grid2* g, const int* counter,
                     const int* offset, const int* index_list);
/* SphUgrid::Compute: substeps x (grid build on read buffer, mode 1, swap, mode 2, swap).
 * buf0/buf1 are the two ping-pong buffers; returns index (0/1) of the read buffer afterwards. */
int orc_sph2_step(orc_particle2* buf0, orc_particle2* buf1, int read_index, int n, int substeps,
                  const orc_params2* prm, const orc_tex* wave1d, const orc_grid2* g,
                  int* counter, int* offset, int* index_list, int* cell_of);

/* ---- the 1-D wave substrates of the 2-D app -----------------------------------
 * SphWave2D/Shallow1D_cs.glsl (two-phase Lax-Wendroff shallow water, RGBA32F texel = (h, uh, hm, uhm), double buffered) and
 * SphWave2D/Wave1D_cs.glsl (damped 1-D wave equation, texel = (u, v, a, -), triple buffered), driven by ImageStencil
 * (SphWave2D/StencilImage2D.cpp:67-164).  One call = ONE dispatch of the shader: `out` holds the previous contents of the output
 * image (texels the shader does not store stay as they are); imageLoad outside the image returns 0 (GL robust access). */
enum { ORC_BC_REFLECT = 0, ORC_BC_FREE = 1, ORC_BC_FIXED = 2 };          /* Shallow1D_cs.glsl:84-87: const int BC = FREE */
typedef struct {
    float lambda;        /* location 2: 0.001 (shallow) / 0.01 (wave) */
    float dx_or_atten;   /* location 3: dx = 0.1 (shallow) / atten = 0.9995 (wave) */
    float beta;          /* location 4: 0.001 */
    float boundary[2];   /* location 5: vec2(0) */
    int   bc;            /* shader const BC, promoted; FREE as shipped */
} orc_stencil1d_params;
void orc_stencil1d_params_default(orc_stencil1d_params* p, int shader /*0 shallow, 1 wave*/);
/* uMode: 0 InitWave, 1 Splash, 2 ITERATE0, 3 ITERATE1 (Shallow1D_cs.glsl:11-15,57-76) */
void orc_shallow1d_dispatch(const float* in, float* out, int w, int mode, const orc_stencil1d_params* p);
/* uMode: 0 / 1 InitWave, 2 ITERATE; in0 = t-1 (unit 0), in1 = t-2 (unit 1) (Wave1D_cs.glsl:11-14,57-71) */
void orc_wave1d_dispatch(const float* in0, const float* in1, float* out, int w, int mode, const orc_stencil1d_params* p);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
