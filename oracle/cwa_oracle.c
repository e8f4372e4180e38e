/*
 * cwa_oracle.c -- CPU restatement of the CoupledWaterAnimation simulation step.
 * TEST INFRASTRUCTURE ONLY (see cwa_oracle.h).  Build: gcc -O3 -march=native -fopenmp
 * -ffp-contract=off -shared -fPIC (oracle/Makefile).
 *
 * Parity status: scan pinned by the reference KATs; grid pinned by the reference CPU twin
 * (oracle/_ref); SPH passes, wave stencil and sampler: PARITY UNPINNED (no reference test,
 * golden vector or runnable GLSL exists -- SURVEY.md F10, 8c).
 */
#include "cwa_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.141592741f /* rho_pres_comp.glsl:8, force_comp.glsl:12 */

/* ------------------------------------------------------------------------------------------ */
/* 1-D wave substrates (SphWave2D/Shallow1D_cs.glsl, Wave1D_cs.glsl)               */
/* ------------------------------------------------------------------------------------------ */
void orc_stencil1d_params_default(orc_stencil1d_params* p, int shader)
{
    p->lambda = shader == 0 ? 0.001f : 0.01f;          /* Shallow1D_cs.glsl:23 / Wave1D_cs.glsl:19 */
    p->dx_or_atten = shader == 0 ? 0.1f : 0.9995f;     /* :24 / :20 */
    p->beta = 0.001f;                                  /* :25 / :21 */
    p->boundary[0] = p->boundary[1] = 0.0f;            /* :26 / :22 */
    p->bc = ORC_BC_FREE;                               /* const int BC = FREE */
}

/* imageLoad with out-of-range coordinates returns zeros */
static inline void load4(const float* img, int w, int x, float v[4])
{
    if (x < 0 || x >= w) { v[0] = v[1] = v[2] = v[3] = 0.0f; return; }
    memcpy(v, img + 4 * (size_t)x, 16);
}
static inline void store4(float* img, int x, const float v[4]) { memcpy(img + 4 * (size_t)x, v, 16); }
static inline float signf_glsl(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
/* exp() of the shaders: evaluated in double and rounded once, so CPU and GPU agree to the last bit
 * (GLSL leaves the precision of exp implementation-defined) */
static inline float expf_canon(float x) { return (float)exp((double)x); }

/* EnforceBC of both shaders (Shallow1D_cs.glsl:169-235, Wave1D_cs.glsl:132-196).  `negate_all`: Wave1D's ReflectBC stores -c,
 * Shallow1D's reflects the velocity channel only.  Stores happen in program order: own texel first, boundary copies after. */
static void enforce_bc(float* out, int w, int coord, const float c[4], const orc_stencil1d_params* p, int negate_all)
{
    const float boundary_scale = 0.1f;
    if (p->bc == ORC_BC_FIXED) {                                         /* FixedBC */
        if (coord == 0)     { const float l[4] = { p->boundary[0] * boundary_scale, 0, 0, 0 }; store4(out, coord, l); return; }
        if (coord == w - 1) { const float r[4] = { p->boundary[1] * boundary_scale, 0, 0, 0 }; store4(out, coord, r); return; }
        store4(out, coord, c);
        return;
    }
    if (coord == 0 || coord == w - 1) return;                            /* neighbours write the boundary values */
    store4(out, coord, c);
    float b[4] = { c[0], c[1], c[2], c[3] };
    if (p->bc == ORC_BC_REFLECT) {
        if (negate_all) { for (int k = 0; k < 4; k++) b[k] = -b[k]; } else b[1] = -b[1];
    }
    if (coord == 1) store4(out, 0, b);
    if (coord == w - 2) store4(out, w - 1, b);
}

void orc_shallow1d_dispatch(const float* in, float* out, int w, int mode, const orc_stencil1d_params* p)
{
    const float VIEW_HEIGHT = 2.0f * 4.8f;
    const float G = 9.8f;
    const float lambda = p->lambda, dx = p->dx_or_atten;
    /* Two sweeps reproduce the store order of a lock-step GPU: within one invocation the stores happen in program order, and an
     * invocation's LATER store wins over a neighbour's EARLIER one to the same texel (ITERATE0's unconditional final store of the
     * half-step values, :160, lands after every EnforceBC copy). */
    for (int coord = 0; coord < w; coord++) {
        if (mode == 0) {                                                 /* InitWave :120-134 */
            float v[4] = { 0, 0, 0, 0 };
            const float x = (float)coord / (float)(w - 1);
            const float xc = x - 0.5f;
            const float h_free = VIEW_HEIGHT * 0.5f;
            const float h = (VIEW_HEIGHT * 0.1f) * expf_canon(-xc * xc / 0.005f);
            v[0] = h_free + h;
            v[1] = 0.5f * fabsf(h) * signf_glsl(xc);
            store4(out, coord, v);
        } else if (mode == 1) {                                          /* Splash :103-118 */
            float v[4]; load4(in, w, coord, v);
            const float x = (float)coord / (float)(w - 1);
            const float xc = x - 0.5f;
            const float h = (-VIEW_HEIGHT * 0.1f) * expf_canon(-xc * xc / 0.005f);
            v[0] += h;
            v[1] += 0.2f * fabsf(h) * signf_glsl(xc);
            store4(out, coord, v);
        } else if (mode == 2 || mode == 3) {                             /* IterateWave :136-167 */
            float c[4], e[4], ww[4];
            load4(in, w, coord, c); load4(in, w, coord + 1, e); load4(in, w, coord - 1, ww);
            float r[4] = { c[0], c[1], c[2], c[3] };
            if (mode == 2) {
                const float hm = (c[0] + e[0]) / 2.0f - lambda / 2.0f * (e[1] - c[1]) / dx;
                const float uhm = (c[1] + e[1]) / 2.0f
                                - lambda / 2.0f * (e[1] * e[1] / e[0] + 0.5f * G * e[0] * e[0] - c[1] * c[1] / c[0] - 0.5f * G * c[0] * c[0]) / dx;
                r[2] = hm; r[3] = uhm;
                enforce_bc(out, w, coord, r, p, 0);
            } else {
                const float h = c[0] - lambda * (c[3] - ww[3]) / dx;
                const float uh = c[1] - lambda * (c[3] * c[3] / c[2] + 0.5f * G * c[2] * c[2] - ww[3] * ww[3] / ww[2] - 0.5f * G * ww[2] * ww[2]) / dx;
                r[0] = h; r[1] = uh;
                enforce_bc(out, w, coord, r, p, 0);
            }
        }
    }
    if (mode == 2) {                                                     /* :160 "no BC needed for half values": every invocation, last */
        for (int coord = 0; coord < w; coord++) {
            float c[4], e[4];
            load4(in, w, coord, c); load4(in, w, coord + 1, e);
            float r[4] = { c[0], c[1], 0, 0 };
            r[2] = (c[0] + e[0]) / 2.0f - lambda / 2.0f * (e[1] - c[1]) / dx;
            r[3] = (c[1] + e[1]) / 2.0f
                 - lambda / 2.0f * (e[1] * e[1] / e[0] + 0.5f * G * e[0] * e[0] - c[1] * c[1] / c[0] - 0.5f * G * c[0] * c[0]) / dx;
            store4(out, coord, r);
        }
    }
}

void orc_wave1d_dispatch(const float* in0, const float* in1, float* out, int w, int mode, const orc_stencil1d_params* p)
{
    const float lambda = p->lambda, atten = p->dx_or_atten, beta = p->beta;
    const float boundary_scale = 0.1f;
    if (mode == 0 || mode == 1) {                                        /* InitWave :98-121 */
        /* sweep 1: every invocation's first store (the initial profile) ... */
        for (int coord = 0; coord < w; coord++) {
            float v[4] = { 0, 0, 0, 0 };
            const int cen = (int)(0.25f * (float)w) + 1 * mode;
            const int x = cen - coord;
            float d0 = 0.0f;
            if (p->bc == ORC_BC_FIXED) {
                const float a = p->boundary[0] * boundary_scale, b = p->boundary[1] * boundary_scale, t = (float)coord / (float)(w - 1);
                d0 = a * (1.0f - t) + b * t;                              /* mix(a, b, t) */
            }
            v[0] = d0 + 0.1f * expf_canon((float)(-x * x) / 5000.0f);
            store4(out, coord, v);
        }
        if (mode == 1) {                                                 /* ... sweep 2: MODE_INIT_1 then steps once from the INPUT images */
            for (int coord = 0; coord < w; coord++) {
                float c1[4], c0[4], e0[4], w0[4], r[4];
                load4(in1, w, coord, c1); load4(in0, w, coord, c0); load4(in0, w, coord + 1, e0); load4(in0, w, coord - 1, w0);
                for (int k = 0; k < 4; k++) r[k] = c0[k] - (0.5f * lambda) * (e0[k] - 2.0f * c0[k] + w0[k]);
                const float rx = r[0];
                r[1] = (rx - c1[0]) / 2.0f;                               /* ComputeVelocityAcceleration :73-79 */
                r[2] = (rx - 2.0f * c0[0] + c1[0]);
                enforce_bc(out, w, coord, r, p, 1);
            }
        }
        return;
    }
    if (mode != 2) return;
    const float kc = 2.0f - 2.0f * lambda - beta, k1 = 1.0f - beta;
    for (int coord = 0; coord < w; coord++) {                            /* IterateWave :123-130 */
        float c1[4], c0[4], e0[4], w0[4], r[4];
        load4(in1, w, coord, c1); load4(in0, w, coord, c0); load4(in0, w, coord + 1, e0); load4(in0, w, coord - 1, w0);
        for (int k = 0; k < 4; k++) r[k] = atten * (kc * c0[k] + lambda * (e0[k] + w0[k]) - k1 * c1[k]);
        const float rx = r[0];
        r[1] = (rx - c1[0]) / 2.0f;
        r[2] = (rx - 2.0f * c0[0] + c1[0]);
        enforce_bc(out, w, coord, r, p, 1);
    }
}


int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* small helpers: GLSL built-ins restated                                                      */
/* ------------------------------------------------------------------------------------------ */
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* GLSL smoothstep(e0,e1,x): t = clamp((x-e0)/(e1-e0),0,1); t*t*(3-2t).  Reversed edges are
 * evaluated with the same formula. */
static inline float smoothstepf(float e0, float e1, float x)
{
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

static inline float signf(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

/* canonical length (see header) */
static inline float length3(float x, float y, float z)
{
    return sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
}
static inline float length2(float x, float y) { return sqrtf(fmaf(y, y, x * x)); }

/* float -> cell coordinate with the clamp done in the float domain (NaN -> 0).
 * ivec(floor(q)) then clamp(cell, 0, n-1): uniform_grid_sph_cs.glsl:144-145 */
static inline int cell_coord(float q, int n)
{
    float f = floorf(q);
    if (!(f >= 0.0f)) return 0;
    if (f > (float)(n - 1)) return n - 1;
    return (int)f;
}

/* ------------------------------------------------------------------------------------------ */
/* parameters                                                                                  */
/* ------------------------------------------------------------------------------------------ */
/* texture t scale: uv_scale_z when set, else the reference's single factor (2.0*pos.xz) */
static float orc_uvz(const orc_params3* p) { return p->uv_scale_z != 0.0f ? p->uv_scale_z : p->uv_scale; }

void orc_params3_default(orc_params3* p)
{
    /* CoupledWaterAnimation/Main.cpp:184-204 */
    p->mass = 0.02f; p->smoothing_coeff = 2.0f; p->visc = 3000.0f; p->resting_rho = 1000.0f;
    p->upper[0] = 0.48f; p->upper[1] = 1.0f; p->upper[2] = 0.48f; p->upper[3] = 500.0f;
    p->lower[0] = 0.0f; p->lower[1] = -0.02f; p->lower[2] = 0.0f; p->lower[3] = 50.0f;
    p->attributes[0] = 0.01f; p->attributes[1] = 0.985f; p->attributes[2] = 0.001f; p->attributes[3] = 1.0f;
    p->mesh_ws_pos[0] = 2.0f; p->mesh_ws_pos[1] = 0.35f; p->mesh_ws_pos[2] = -1.0f; p->mesh_ws_pos[3] = 0.0f;
    /* shader constants: rho_pres_comp.glsl:5,41,43; force_comp.glsl:7,52,55; integrate_comp.glsl:8,51,69 */
    p->particle_radius = 0.005f; p->gas_const = 4000.0f; p->dt = 0.00005f; p->gravity_y = -9806.65f;
    p->damping = 0.3f; p->crest_threshold = 0.01f; p->foam_speed = 25.0f; p->uv_scale = 2.0f; p->uv_scale_z = 0.0f; p->torque_coeff = 0.0f;
}

/* ------------------------------------------------------------------------------------------ */
/* A.3 bilinear sampler: texture() on GL_LINEAR / GL_CLAMP_TO_EDGE / LOD 0, .r channel          */
/* StencilImage2DTripleBuffered.cpp:25-26 (filter/wrap), rho_pres_comp.glsl:73                  */
/* ------------------------------------------------------------------------------------------ */
static inline int tex_index(float f, int n)
{
    /* clamp an (already floored) float texel coordinate to [0,n-1]; NaN -> 0 */
    if (!(f >= 0.0f)) return 0;
    if (f > (float)(n - 1)) return n - 1;
    return (int)f;
}

float orc_tex_bilinear(const orc_tex* t, float s, float tt)
{
    if (!t || !t->data) return 0.0f; /* unbound texture samples (0,0,0,1): Appendix B frame 1 */
    const int W = t->w, H = t->h, C = t->ch;
    float u = s * (float)W - 0.5f;
    float v = tt * (float)H - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int i0 = tex_index(fu, W), i1 = tex_index(fu + 1.0f, W);
    int j0 = tex_index(fv, H), j1 = tex_index(fv + 1.0f, H);
    float t00 = t->data[((size_t)j0 * W + i0) * C];
    float t10 = t->data[((size_t)j0 * W + i1) * C];
    float t01 = t->data[((size_t)j1 * W + i0) * C];
    float t11 = t->data[((size_t)j1 * W + i1) * C];
    /* lerp(x,y,a) = x + a*(y-x) */
    float r0 = t00 + a * (t10 - t00);
    float r1 = t01 + a * (t11 - t01);
    return r0 + b * (r1 - r0);
}

/* 1-D RGBA sampler for the 2-D app (sampler1D wave_tex, SphWaveKoschier2D_grid_cs.glsl:283).
 * Unbound -> (0,0,0,1). */
static void tex1d_linear(const orc_tex* t, float s, float out[4])
{
    if (!t || !t->data) { out[0] = out[1] = out[2] = 0.0f; out[3] = 1.0f; return; }
    const int W = t->w, C = t->ch;
    float u = s * (float)W - 0.5f;
    float fu = floorf(u);
    float a = u - fu;
    int i0 = tex_index(fu, W), i1 = tex_index(fu + 1.0f, W);
    for (int c = 0; c < 4; c++) {
        float x0 = (c < C) ? t->data[(size_t)i0 * C + c] : (c == 3 ? 1.0f : 0.0f);
        float x1 = (c < C) ? t->data[(size_t)i1 * C + c] : (c == 3 ? 1.0f : 0.0f);
        out[c] = x0 + a * (x1 - x0);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a-2 scan                                                                                    */
/* ------------------------------------------------------------------------------------------ */
/* One dispatch of prefix_sum_cs.glsl:18-43 over all gid (every "thread" sees the state left by
 * the previous dispatch; within a dispatch the touched indices are disjoint). */
static void prefix_sum_dispatch(int* x, int n, int phase, int stride, int nthreads)
{
    for (int gid = 0; gid < nthreads; gid++) {
        int ka = (stride / 2 - 1) + gid * stride;
        int kb = ka + stride / 2;
        if (kb >= n) continue;
        if (phase == 0) {
            x[kb] = x[ka] + x[kb];
        } else {
            if (stride == n) x[n - 1] = 0;
            int t = x[ka];
            x[ka] = x[kb];
            x[kb] = t + x[kb];
        }
    }
}

/* ParallelScan::Compute, SphWave2D/ParallelScan.cpp:43-95 (in place on the copy). */
int orc_scan_blelloch(int* x, int num)
{
    if (num < 2 || (num & (num - 1)) != 0) return -1; /* ParallelScan.cpp:15-16 assert */
    int n = num / 2, pass = 0;
    for (;;) { /* upsweep :57-72 */
        int stride = 2 << pass;
        int groups = (n + 1023) / 1024;
        prefix_sum_dispatch(x, num, 0, stride, groups * 1024);
        if (n == 1) break;
        n = n / 2; pass = pass + 1;
    }
    for (;;) { /* downsweep :77-91 */
        int stride = 2 << pass;
        int groups = (n + 1023) / 1024;
        prefix_sum_dispatch(x, num, 1, stride, groups * 1024);
        if (n == num / 2) break;
        n = n * 2; pass = pass - 1;
    }
    return 0;
}

/* "truth" loop of ParallelScanTest, ParallelScan.cpp:132-137; UniformGrid2D.cpp:60-69 */
void orc_scan_exclusive(const int* in, int* out, int n)
{
    int sum = 0;
    for (int i = 0; i < n; i++) { out[i] = sum; sum += in[i]; }
}

/* ------------------------------------------------------------------------------------------ */
/* a-1, a-3 grid                                                                               */
/* ------------------------------------------------------------------------------------------ */
/* UniformGridSph2D ctor, SphWave2D/UniformGridGpu2D.cpp:156-161 */
void orc_grid2_init(orc_grid2* g, const float mn[2], const float mx[2], const int n[2])
{
    for (int a = 0; a < 2; a++) {
        g->min[a] = mn[a]; g->max[a] = mx[a]; g->ncells[a] = n[a];
        g->cell[a] = (mx[a] - mn[a]) / (float)n[a];
    }
}

/* Ugrid3D ctor, UniformGrid2D/UniformGridParticles3D.cpp:92-97 */
void orc_grid3_init(orc_grid3* g, const float mn[3], const float mx[3], const int n[3])
{
    for (int a = 0; a < 3; a++) {
        g->min[a] = mn[a]; g->max[a] = mx[a]; g->ncells[a] = n[a];
        g->cell[a] = (mx[a] - mn[a]) / (float)n[a];
    }
    g->min[3] = g->max[3] = 0.0f; g->ncells[3] = 1; g->cell[3] = 0.0f;
}

/* ComputeCellIndex + Index, uniform_grid_sph_cs.glsl:141-152 */
static inline void grid2_cell(const orc_grid2* g, float x, float y, int* ci, int* cj)
{
    *ci = cell_coord((x - g->min[0]) / g->cell[0], g->ncells[0]);
    *cj = cell_coord((y - g->min[1]) / g->cell[1], g->ncells[1]);
}
int orc_grid2_cell_index(const orc_grid2* g, float x, float y)
{
    int i, j; grid2_cell(g, x, y, &i, &j);
    return i * g->ncells[1] + j;
}

/* ComputeCellIndex + Index, ugrid_particles_cs.glsl:97-108.  NOTE the (sic) index formula
 * (i*Ny + j)*Nx + k -- restated as written. */
static inline void grid3_cell(const orc_grid3* g, float x, float y, float z, int* ci, int* cj, int* ck)
{
    *ci = cell_coord((x - g->min[0]) / g->cell[0], g->ncells[0]);
    *cj = cell_coord((y - g->min[1]) / g->cell[1], g->ncells[1]);
    *ck = cell_coord((z - g->min[2]) / g->cell[2], g->ncells[2]);
}
static inline int grid3_index(const orc_grid3* g, int i, int j, int k)
{
    return (i * g->ncells[1] + j) * g->ncells[0] + k;
}
int orc_grid3_cell_index(const orc_grid3* g, float x, float y, float z)
{
    int i, j, k; grid3_cell(g, x, y, z, &i, &j, &k);
    return grid3_index(g, i, j, k);
}
static inline int grid3_num_cells(const orc_grid3* g)
{
    /* Ugrid3D allocates Nx*Ny*Nz counters (UniformGridParticles3D.cpp:102) */
    return g->ncells[0] * g->ncells[1] * g->ncells[2];
}

/* point_in_aabb, uniform_grid_sph_cs.glsl:19-23: strict on both axes */
static inline int point_in_aabb2(const orc_grid2* g, float x, float y)
{
    return (x > g->min[0] && y > g->min[1] && x < g->max[0] && y < g->max[1]);
}

/* UniformGridSph2D::CollisionQuery (= build), UniformGridGpu2D.cpp:220-258 with
 * uniform_grid_sph_cs.glsl ComputeCounts :112-125 / InsertPoint :154-165.
 * The GPU insert order inside a cell is nondeterministic (F7); the canonical order restated here
 * is ascending particle id, which is what the CPU twin UniformGrid2D::Build produces. */
void orc_grid2_build(const orc_grid2* g, const float* pos, int stride, int n,
                     int* cell_of, int* counter, int* offset, int* index_list)
{
    const int C = g->ncells[0] * g->ncells[1];
    memset(counter, 0, sizeof(int) * (size_t)C);           /* ClearCounter */
    memset(offset, 0, sizeof(int) * (size_t)C);            /* ClearOffset  */
    for (int i = 0; i < n; i++) {                          /* COMPUTE_COUNTS */
        float x = pos[(size_t)i * stride], y = pos[(size_t)i * stride + 1];
        if (!point_in_aabb2(g, x, y)) { if (cell_of) cell_of[i] = -1; continue; }
        int ix = orc_grid2_cell_index(g, x, y);
        if (cell_of) cell_of[i] = ix;
        counter[ix]++;
    }
    orc_scan_exclusive(counter, offset, C);                /* ParallelScan::Compute */
    memset(counter, 0, sizeof(int) * (size_t)C);           /* ClearCounter */
    for (int i = 0; i < n; i++) {                          /* INSERT_BOXES */
        float x = pos[(size_t)i * stride], y = pos[(size_t)i * stride + 1];
        if (!point_in_aabb2(g, x, y)) continue;
        int ix = orc_grid2_cell_index(g, x, y);
        int count = counter[ix]++;
        index_list[offset[ix] + count] = i;
    }
}

/* UgridParticles3D::BuildGrid, UniformGridParticles3D.cpp:170-211 with ugrid_particles_cs.glsl
 * ComputeCounts :90-95 / InsertParticle :110-117 (no in-extent test). */
void orc_grid3_build(const orc_grid3* g, const float* pos, int stride, int n,
                     int* cell_of, int* counter, int* offset, int* index_list)
{
    /* Canonical choice for an undefined case: ivec3(floor(NaN)) is undefined in GLSL, so a particle
     * whose position holds a NaN lands in an implementation-defined cell.  Such a particle can never
     * pass a distance test again (every r is NaN), so it is simply NOT inserted (cell_of = -1). */
    const int C = grid3_num_cells(g);
    memset(counter, 0, sizeof(int) * (size_t)C);
    memset(offset, 0, sizeof(int) * (size_t)C);
    for (int i = 0; i < n; i++) {
        const float* p = pos + (size_t)i * stride;
        if (p[0] != p[0] || p[1] != p[1] || p[2] != p[2]) { if (cell_of) cell_of[i] = -1; continue; }
        int ix = orc_grid3_cell_index(g, p[0], p[1], p[2]);
        if (cell_of) cell_of[i] = ix;
        counter[ix]++;
    }
    orc_scan_exclusive(counter, offset, C);
    memset(counter, 0, sizeof(int) * (size_t)C);
    for (int i = 0; i < n; i++) {
        const float* p = pos + (size_t)i * stride;
        if (p[0] != p[0] || p[1] != p[1] || p[2] != p[2]) continue;
        int ix = orc_grid3_cell_index(g, p[0], p[1], p[2]);
        int count = counter[ix]++;
        index_list[offset[ix] + count] = i;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a-7 wave                                                                                    */
/* ------------------------------------------------------------------------------------------ */
static inline float dist2i(int x, int y, int cx, int cy)
{
    /* distance(coord, cen) on ivec2 -> vec2: length of the difference */
    float dx = (float)x - (float)cx, dy = (float)y - (float)cy;
    return sqrtf(dx * dx + dy * dy);
}

/* InitWave: wave_comp.glsl:82-140 (variant COUPLED) / Wave2D_cs.glsl:66-76 (variant SIMP).
 * Only .x is written non-zero (F9); ch = 1 stores the scalar field, ch = 4 the RGBA image. */
void orc_wave_init(float* out, int w, int h, int ch, int variant, float type)
{
    int cen0[2], cen1[2], cen2[2] = {0, 0};
    int use_third = 0;
    float peak, e0;
    if (variant == ORC_WAVE_SIMP) {
        cen0[0] = (int)(0.25f * (float)w); cen0[1] = (int)(0.25f * (float)h);
        cen1[0] = (int)(0.75f * (float)w); cen1[1] = (int)(0.75f * (float)h);
        peak = 0.5f; e0 = 3.0f;
    } else {
        e0 = 5.0f;
        if (type == 1.0f) {            /* splash :91-98 */
            cen0[0] = (int)(0.25f * (float)w); cen0[1] = (int)(0.25f * (float)h);
            cen1[0] = (int)(0.75f * (float)w); cen1[1] = (int)(0.75f * (float)h);
            peak = 0.5f;
        } else if (type == 0.0f) {     /* wave :99-107 */
            cen0[0] = (int)(0.25f * (float)w); cen0[1] = h;
            cen2[0] = (int)(0.5f * (float)w);  cen2[1] = h;
            cen1[0] = (int)(0.75f * (float)w); cen1[1] = h;
            peak = 1.0f; use_third = 1;
        } else {                       /* boat wake :108-116 */
            cen0[0] = (int)(0.5f * (float)w); cen0[1] = (int)(0.1f * (float)h);
            cen1[0] = cen0[0]; cen1[1] = cen0[1];
            cen2[0] = cen0[0]; cen2[1] = cen0[1];
            peak = 0.1f; use_third = 1;
        }
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            float d = fminf(dist2i(x, y, cen0[0], cen0[1]), dist2i(x, y, cen1[0], cen1[1]));
            if (use_third) d = fminf(d, dist2i(x, y, cen2[0], cen2[1]));
            float v = peak * smoothstepf(e0, 0.0f, d);
            float* o = out + ((size_t)y * w + x) * ch;
            o[0] = v;
            for (int c = 1; c < ch; c++) o[c] = 0.0f;
        }
    }
}

/* EvolveWave + get_clamp: wave_comp.glsl:171-205 / Wave2D_cs.glsl:78-103.
 * u0 = wave at t-1 (image unit 0), u1 = wave at t-2 (unit 1), out = unit 2. */
void orc_wave_evolve(const float* u0, const float* u1, float* out, int w, int h, int ch,
                     int variant, float lambda, float atten, float beta, float type)
{
    const float kc = 2.0f - 4.0f * lambda - beta; /* (2.0f - 4.0f*a[0] - a[2]) */
    const float k1 = 1.0f - beta;
    const int mid_x = w / 2; /* CoordOnLine :159-169 */
    const int wake = (variant == ORC_WAVE_COUPLED) && (type > 0.0f && type < 1.0f);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const int yn = clampi(y + 1, 0, h - 1), ys = clampi(y - 1, 0, h - 1);
        for (int x = 0; x < w; x++) {
            const int xe = clampi(x + 1, 0, w - 1), xw = clampi(x - 1, 0, w - 1);
            for (int c = 0; c < ch; c++) {
                float c1 = u1[((size_t)y * w + x) * ch + c];
                float c0 = u0[((size_t)y * w + x) * ch + c];
                float n0 = u0[((size_t)yn * w + x) * ch + c];
                float s0 = u0[((size_t)ys * w + x) * ch + c];
                float e0 = u0[((size_t)y * w + xe) * ch + c];
                float w0 = u0[((size_t)y * w + xw) * ch + c];
                float v = kc * c0 + lambda * (n0 + s0 + e0 + w0) - k1 * c1;
                v = v * atten; /* w *= attributes[1]  /  w = atten*w */
                if (c == 0 && wake && v > 0.0001f && x == mid_x) v += 0.001f; /* :177-184 */
                out[((size_t)y * w + x) * ch + c] = v;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a-4 .. a-6: 3-D SPH                                                                          */
/* ------------------------------------------------------------------------------------------ */
static inline float pow2f(float x) { return x * x; }
static inline float pow3f(float x) { return (x * x) * x; }
static inline float pow6f(float x) { float x2 = x * x; return (x2 * x2) * x2; }
static inline float pow9f(float x) { float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4; return x8 * x; }

typedef struct { int i0, i1, j0, j1, k0, k1; } cell_range3;

/* cells overlapped by [pos-h, pos+h] via the clamped ComputeCellIndex -- the reference idiom of
 * SphKoschier2D_grid_cs.glsl:384-391 applied to the 3-D point grid. */
static inline cell_range3 query_range3(const orc_grid3* g, const float* pos, float h)
{
    cell_range3 r;
    grid3_cell(g, pos[0] - h, pos[1] - h, pos[2] - h, &r.i0, &r.j0, &r.k0);
    grid3_cell(g, pos[0] + h, pos[1] + h, pos[2] + h, &r.i1, &r.j1, &r.k1);
    return r;
}

/* rho_pres_comp.glsl:49-81 */
void orc_sph3_rho_pres(orc_particle3* p, int n, const orc_params3* prm, const orc_tex* tex,
                       const orc_grid3* grid, const int* counter, const int* offset,
                       const int* index_list)
{
    const float h = prm->smoothing_coeff * prm->particle_radius; /* :54 */
    const float mass = prm->mass;
    const float h9 = pow9f(h);
    float* rho_out = (float*)malloc(sizeof(float) * (size_t)n);
    float* prs_out = (float*)malloc(sizeof(float) * (size_t)n);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; i++) {
        const float* pi = p[i].pos;
        float rho = 0.0f;
#define ORC_RHO_PAIR(J)                                                                         \
        {                                                                                       \
            const float* pj = p[(J)].pos;                                                       \
            float r = length3(pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]);                     \
            if (r < h)                                                                          \
                rho += mass * 315.0f * pow3f(h * h - r * r) / (64.0f * ORC_PI * h9); /* :66 */  \
        }
        if (!grid) {
            for (int j = 0; j < n; j++) ORC_RHO_PAIR(j)
        } else {
            cell_range3 q = query_range3(grid, pi, h);
            for (int ci = q.i0; ci <= q.i1; ci++)
                for (int cj = q.j0; cj <= q.j1; cj++)
                    for (int ck = q.k0; ck <= q.k1; ck++) {
                        int c = grid3_index(grid, ci, cj, ck);
                        for (int l = offset[c]; l < offset[c] + counter[c]; l++) ORC_RHO_PAIR(index_list[l])
                    }
        }
#undef ORC_RHO_PAIR
        float pressure = fmaxf(prm->gas_const * (rho - prm->resting_rho), 0.0f); /* :70 */
        float height = orc_tex_bilinear(tex, prm->uv_scale * pi[0], orc_uvz(prm) * pi[2]); /* :72-73 */
        float wave_force = height * rho;                                          /* :75 */
        pressure += wave_force;                                                   /* :76 */
        rho += wave_force / (prm->gas_const * prm->particle_radius);              /* :77 */
        rho_out[i] = fmaxf(prm->resting_rho, rho);                                /* :79 */
        prs_out[i] = pressure;                                                    /* :80 */
    }
    for (int i = 0; i < n; i++) { p[i].extras[0] = rho_out[i]; p[i].extras[1] = prs_out[i]; }
    free(rho_out); free(prs_out);
}

void orc_sph3_neighbour_count(const orc_particle3* p, int n, float h, const orc_grid3* grid,
                              const int* counter, const int* offset, const int* index_list,
                              int* out_count)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; i++) {
        const float* pi = p[i].pos;
        int cnt = 0;
        if (!grid) {
            for (int j = 0; j < n; j++) {
                const float* pj = p[j].pos;
                if (length3(pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]) < h) cnt++;
            }
        } else {
            cell_range3 q = query_range3(grid, pi, h);
            for (int ci = q.i0; ci <= q.i1; ci++)
                for (int cj = q.j0; cj <= q.j1; cj++)
                    for (int ck = q.k0; ck <= q.k1; ck++) {
                        int c = grid3_index(grid, ci, cj, ck);
                        for (int l = offset[c]; l < offset[c] + counter[c]; l++) {
                            const float* pj = p[index_list[l]].pos;
                            if (length3(pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]) < h) cnt++;
                        }
                    }
        }
        out_count[i] = cnt;
    }
}

/* WaveVelocity, force_comp.glsl:117-128 */
static void wave_velocity(const orc_tex* tex, float u, float v, float dt, float out[3])
{
    const float hs = 0.01f;
    float height = orc_tex_bilinear(tex, u, v);
    float heightX = orc_tex_bilinear(tex, u + hs, v);
    float heightY = orc_tex_bilinear(tex, u, v + hs);
    out[0] = (heightX - height) / dt;
    out[1] = (heightY - height) / dt;
    out[2] = (heightX - heightY) / hs;
}

/* WaveNormal, force_comp.glsl:131-139: cross(dy,dx) with dx=(1,0,a), dy=(0,1,b) = (a, b, -1) */
static void wave_normal(const orc_tex* tex, float u, float v, float out[3])
{
    float height = orc_tex_bilinear(tex, u, v);
    float dxz = orc_tex_bilinear(tex, u + 1.0f, v + 0.0f) - height;
    float dyz = orc_tex_bilinear(tex, u + 0.0f, v + 1.0f) - height;
    /* cross(dy, dx) = (dy.y*dx.z - dy.z*dx.y, dy.z*dx.x - dy.x*dx.z, dy.x*dx.y - dy.y*dx.x) */
    out[0] = 1.0f * dxz - dyz * 0.0f;
    out[1] = dyz * 1.0f - 0.0f * dxz;
    out[2] = 0.0f * 0.0f - 1.0f * 1.0f;
}

/* force_comp.glsl:62-115 */
void orc_sph3_force(orc_particle3* p, int n, const orc_params3* prm, const orc_tex* tex,
                    const orc_grid3* grid, const int* counter, const int* offset,
                    const int* index_list)
{
    const float h = prm->smoothing_coeff * prm->particle_radius; /* :53 */
    const float mass = prm->mass;
    const float spiky = -20.0f / (ORC_PI * pow6f(h));            /* :68 */
    const float laplacian = -spiky;                              /* :69 */
    float* fout = (float*)malloc(sizeof(float) * 4 * (size_t)n);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; i++) {
        const float* pi = p[i].pos;
        const float* vi = p[i].vel;
        const float prs_i = p[i].extras[1];
        float pres[3] = {0, 0, 0}, visc[3] = {0, 0, 0};
#define ORC_FORCE_PAIR(J)                                                                          \
        if ((J) != i) {                                                                            \
            const orc_particle3* q = &p[(J)];                                                      \
            float d0 = pi[0] - q->pos[0], d1 = pi[1] - q->pos[1], d2 = pi[2] - q->pos[2];          \
            float r = length3(d0, d1, d2);                                                         \
            if (r < h) {                                                                           \
                float a = mass * (prs_i + q->extras[1]) / (2.0f * q->extras[0]) * spiky *          \
                          pow2f(h - r);                                       /* :85 */            \
                pres[0] -= a * (d0 / r); pres[1] -= a * (d1 / r); pres[2] -= a * (d2 / r);         \
                float hr = h - r;                                             /* :86 */            \
                visc[0] += mass * (q->vel[0] - vi[0]) / q->extras[0] * laplacian * hr;             \
                visc[1] += mass * (q->vel[1] - vi[1]) / q->extras[0] * laplacian * hr;             \
                visc[2] += mass * (q->vel[2] - vi[2]) / q->extras[0] * laplacian * hr;             \
            }                                                                                      \
        }
        if (!grid) {
            for (int j = 0; j < n; j++) ORC_FORCE_PAIR(j)
        } else {
            cell_range3 qr = query_range3(grid, pi, h);
            for (int ci = qr.i0; ci <= qr.i1; ci++)
                for (int cj = qr.j0; cj <= qr.j1; cj++)
                    for (int ck = qr.k0; ck <= qr.k1; ck++) {
                        int c = grid3_index(grid, ci, cj, ck);
                        for (int l = offset[c]; l < offset[c] + counter[c]; l++) {
                            int j = index_list[l];
                            ORC_FORCE_PAIR(j)
                        }
                    }
        }
#undef ORC_FORCE_PAIR
        float fm[4] = {p[i].force[0], p[i].force[1], p[i].force[2], p[i].force[3]};
        if (pi[1] > prm->crest_threshold) {                      /* :91-95 */
            for (int c = 0; c < 4; c++) fm[c] = fm[c] / 0.25f;   /* force /= BREAKING_MASS_FACTOR (vec4, in memory) */
            for (int c = 0; c < 3; c++) visc[c] *= 0.5f;
        }
        for (int c = 0; c < 3; c++) visc[c] *= prm->visc;        /* :97 */
        float cu = pi[0] * prm->uv_scale, cv = pi[2] * orc_uvz(prm); /* :99 */
        float height = orc_tex_bilinear(tex, cu, cv);           /* :100 */
        /* torque = 0.25*cross(pos, force_mem.xyz) :103-104 */
        float tq[3] = { pi[1] * fm[2] - pi[2] * fm[1], pi[2] * fm[0] - pi[0] * fm[2], pi[0] * fm[1] - pi[1] * fm[0] };
        const float tqc = prm->torque_coeff != 0.0f ? prm->torque_coeff : 0.25f;
        for (int c = 0; c < 3; c++) tq[c] *= tqc;
        float wv[3]; wave_velocity(tex, cu, cv, prm->dt, wv);    /* :107 */
        float drag[3];
        for (int c = 0; c < 3; c++) drag[c] = -0.25f * (vi[c] - wv[c]); /* :107-108 */
        float wn[3]; wave_normal(tex, cu, cv, wn);
        float wavef[3];
        for (int c = 0; c < 3; c++) wavef[c] = -height * wn[c] * 0.5f;  /* :110 */
        float grav[3] = {p[i].extras[0] * 0.0f, p[i].extras[0] * prm->gravity_y, p[i].extras[0] * 0.0f}; /* :113 */
        for (int c = 0; c < 3; c++)
            fout[4 * (size_t)i + c] = pres[c] + visc[c] + grav[c] + tq[c] + drag[c] + wavef[c]; /* :114 */
        fout[4 * (size_t)i + 3] = fm[3];
    }
    for (int i = 0; i < n; i++) memcpy(p[i].force, fout + 4 * (size_t)i, sizeof(float) * 4);
    free(fout);
}

/* integrate_comp.glsl:56-92 + CheckBoundary :135-178 */
void orc_sph3_integrate(orc_particle3* p, int n, const orc_params3* prm, const orc_tex* tex)
{
    const float dt = prm->dt, D = prm->damping;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        orc_particle3* q = &p[i];
        float acc[3], nv[3], np[3];
        for (int c = 0; c < 3; c++) acc[c] = q->force[c] / q->extras[0];     /* :62 */
        for (int c = 0; c < 3; c++) nv[c] = q->vel[c] + dt * acc[c];         /* :63 */
        for (int c = 0; c < 3; c++) np[c] = q->pos[c] + dt * nv[c];          /* :64 */
        const float damp = 1.0f - D * dt;
        for (int c = 0; c < 3; c++) nv[c] *= damp;                           /* :66 */
        if (length3(nv[0], nv[1], nv[2]) > prm->foam_speed) {                /* :69-76 */
            for (int c = 0; c < 4; c++) q->force[c] *= 0.5f;
            q->extras[0] *= 0.1f;
            q->extras[1] *= 0.25f;
            for (int c = 0; c < 3; c++) nv[c] *= 0.1f;
        }
        float tex_height = orc_tex_bilinear(tex, np[0] * prm->uv_scale, np[2] * orc_uvz(prm)); /* :79 */
        if (np[1] < tex_height) np[1] = tex_height - prm->particle_radius;   /* :80-83 */
        /* CheckBoundary :137-168 */
        for (int c = 0; c < 3; c++) {
            if (np[c] < prm->lower[c]) { np[c] = prm->lower[c]; nv[c] *= -D; }
            else if (np[c] > prm->upper[c]) { np[c] = prm->upper[c]; nv[c] *= -D; }
        }
        if (prm->attributes[3] > 0.0f && prm->attributes[3] < 1.0f) {        /* :170-177 */
            if (np[1] > prm->upper[1] + 0.1f) { np[1] = prm->upper[1] + 0.1f; nv[1] *= -D; }
        }
        for (int c = 0; c < 3; c++) { q->vel[c] = nv[c]; q->pos[c] = np[c]; } /* :90-91 */
    }
}

/* make_cube + init_particles, CoupledWaterAnimation/Main.cpp:735-776 */
void orc_make_cube(orc_particle3* p, int nx, int ny, int nz, const orc_params3* prm)
{
    const float spacing = prm->smoothing_coeff * 0.85f * prm->particle_radius; /* :739 */
    const float mid = 0.0f;
    size_t n = 0;
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++)
            for (int k = 0; k < nz; k++) {
                orc_particle3* q = &p[n++];
                q->pos[0] = mid + (float)i * spacing; q->pos[1] = (float)j * spacing;
                q->pos[2] = mid + (float)k * spacing; q->pos[3] = 1.0f;
                for (int c = 0; c < 4; c++) { q->vel[c] = 0.0f; q->force[c] = 0.0f; }
                q->extras[0] = prm->resting_rho; q->extras[1] = 0.0f; q->extras[2] = 500.0f; q->extras[3] = 50.0f;
            }
}

/* ------------------------------------------------------------------------------------------ */
/* a-8, a-9: coupled driver with the as-shipped triple-buffer / texture-unit bookkeeping        */
/* ------------------------------------------------------------------------------------------ */
struct orc_coupled {
    int n, w, h, ch, coupling, use_grid;
    orc_params3 prm;
    orc_particle3* particles;
    float* image[3];
    /* StencilImage2DTripleBuffered.h:36-37 */
    int read_index[2], write_index;
    int unit[3];               /* ImageTexture::mUnit */
    int tex_unit0;             /* physical image bound to GL texture unit 0, -1 = unbound */
    orc_grid3 grid;
    int *counter, *offset, *index_list;
};

static int image_with_unit(const orc_coupled* c, int u)
{
    for (int i = 0; i < 3; i++) if (c->unit[i] == u) return i;
    return -1;
}

/* PingPong, StencilImage2DTripleBuffered.cpp:33-40 (SwapUnits = ImageTexture.cpp:100-103) */
static void wave_pingpong(orc_coupled* c)
{
    int t;
    t = c->write_index; c->write_index = c->read_index[0]; c->read_index[0] = t;
    t = c->read_index[0]; c->read_index[0] = c->read_index[1]; c->read_index[1] = t;
    t = c->unit[c->write_index]; c->unit[c->write_index] = c->unit[c->read_index[0]]; c->unit[c->read_index[0]] = t;
    t = c->unit[c->read_index[0]]; c->unit[c->read_index[0]] = c->unit[c->read_index[1]]; c->unit[c->read_index[1]] = t;
}

/* Reinit, StencilImage2DTripleBuffered.cpp:42-59: two MODE_INIT dispatches (shader writes the
 * image on unit 2) each followed by PingPong. */
void orc_coupled_wave_reinit(orc_coupled* c)
{
    for (int i = 0; i < 2; i++) {
        int out = image_with_unit(c, 2);
        orc_wave_init(c->image[out], c->w, c->h, c->ch, ORC_WAVE_COUPLED, c->prm.attributes[3]);
        wave_pingpong(c);
    }
}

orc_coupled* orc_coupled_create(int n, int wave_w, int wave_h, int wave_ch,
                                const orc_params3* prm, int coupling,
                                int use_grid, const float gmin[3], const float gmax[3],
                                const int gn[3])
{
    orc_coupled* c = (orc_coupled*)calloc(1, sizeof(orc_coupled));
    c->n = n; c->w = wave_w; c->h = wave_h; c->ch = wave_ch; c->coupling = coupling; c->prm = *prm;
    c->particles = (orc_particle3*)calloc((size_t)n, sizeof(orc_particle3));
    for (int i = 0; i < 3; i++) {
        c->image[i] = (float*)calloc((size_t)wave_w * wave_h * wave_ch, sizeof(float));
        c->unit[i] = i;                               /* Init: SetUnit(i) :23 */
    }
    c->read_index[0] = 0; c->read_index[1] = 1; c->write_index = 2;
    c->tex_unit0 = -1;
    c->use_grid = use_grid;
    if (use_grid) {
        orc_grid3_init(&c->grid, gmin, gmax, gn);
        int C = grid3_num_cells(&c->grid);
        c->counter = (int*)calloc((size_t)C, sizeof(int));
        c->offset = (int*)calloc((size_t)C, sizeof(int));
        c->index_list = (int*)calloc((size_t)n, sizeof(int));
    }
    orc_coupled_wave_reinit(c);                       /* Init -> Reinit :30 */
    return c;
}

void orc_coupled_destroy(orc_coupled* c)
{
    if (!c) return;
    free(c->particles);
    for (int i = 0; i < 3; i++) free(c->image[i]);
    free(c->counter); free(c->offset); free(c->index_list);
    free(c);
}

orc_particle3* orc_coupled_particles(orc_coupled* c) { return c->particles; }
void orc_coupled_set_params(orc_coupled* c, const orc_params3* prm) { c->prm = *prm; }
int orc_coupled_sampled_image(const orc_coupled* c) { return c->tex_unit0; }

float* orc_coupled_wave(orc_coupled* c, int role)
{
    /* roles by unit: unit 0 = newest (u^{t-1} input of the next step), 1 = previous, 2 = output */
    return c->image[image_with_unit(c, role)];
}

/* one frame = idle() (Main.cpp:540-561) then the display() texture bind (Main.cpp:413) */
void orc_coupled_step(orc_coupled* c, int nframes)
{
    for (int f = 0; f < nframes; f++) {
        orc_tex tex;
        tex.w = c->w; tex.h = c->h; tex.ch = c->ch;
        if (c->coupling == ORC_COUPLING_LATEST) tex.data = c->image[image_with_unit(c, 0)];
        else tex.data = (c->tex_unit0 >= 0) ? c->image[c->tex_unit0] : NULL;

        const orc_grid3* g = NULL;
        if (c->use_grid) {
            orc_grid3_build(&c->grid, c->particles[0].pos, 16, c->n, NULL, c->counter, c->offset, c->index_list);
            g = &c->grid;
        }
        orc_sph3_rho_pres(c->particles, c->n, &c->prm, &tex, g, c->counter, c->offset, c->index_list); /* :549-551 */
        orc_sph3_force(c->particles, c->n, &c->prm, &tex, g, c->counter, c->offset, c->index_list);    /* :552-554 */
        orc_sph3_integrate(c->particles, c->n, &c->prm, &tex);                                        /* :555-557 */

        /* Module::sComputeAll -> StencilImage2DTripleBuffered::Compute :79-95 */
        int in0 = image_with_unit(c, 0), in1 = image_with_unit(c, 1), out = image_with_unit(c, 2);
        orc_wave_evolve(c->image[in0], c->image[in1], c->image[out], c->w, c->h, c->ch, ORC_WAVE_COUPLED,
                        c->prm.attributes[0], c->prm.attributes[1], c->prm.attributes[2], c->prm.attributes[3]);
        wave_pingpong(c);

        /* display(): wave2d.GetReadImage(0).BindTextureUnit() binds at that image's mUnit (F5) */
        int ri0 = c->read_index[0];
        if (c->unit[ri0] == 0) c->tex_unit0 = ri0;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a-4b, a-5b: 2-D Koschier SPH (SphKoschier2D_grid_cs.glsl / SphWaveKoschier2D_grid_cs.glsl)    */
/* ------------------------------------------------------------------------------------------ */
/* constants :116-136 (identical in both files except WIDTH) */
#define K_PARTICLE_RADIUS 0.025f
#define K_PARTICLE_DIAM (2.0f * K_PARTICLE_RADIUS)
#define K_H (4.0f * K_PARTICLE_RADIUS)
#define K_HSQ (K_H * K_H)
#define K_REST_DENS 1000.0f
#define K_VISC 0.05f
#define K_MASS (K_PARTICLE_DIAM * K_PARTICLE_DIAM * K_REST_DENS)
#define K_DT 0.002f
#define K_GAS_CONST 35000.0f
#define K_M_PI 3.14159265f
#define K_VIEW_WIDTH_DEFAULT (2.0f * 4.8f) /* const VIEW_WIDTH; a parameter here so config C2 can widen the tank */
#define K_VIEW_HEIGHT (2.0f * 4.8f)

static inline float k2_const(void) { return 40.0f / (7.0f * K_M_PI * K_HSQ); }

void orc_params2_default(orc_params2* p, int variant)
{
    p->variant = variant; p->time = 0.0f; p->bottom = 0.3f; p->psi = -1.0f;
    p->init_width = (variant == ORC_SPH2_WAVE) ? 128 : 32;
    p->view_width = K_VIEW_WIDTH_DEFAULT;
}

static inline float psi_of(const orc_params2* p)
{
    return (p->psi < 0.0f) ? K_REST_DENS / (1.5f * k2_const()) : p->psi; /* :136 */
}

/* W_cubic :233-246 */
static inline float W_cubic(float r)
{
    const float k2 = k2_const();
    float q = r / K_H;
    if (q >= 1.0f) return 0.0f;
    if (q <= 0.5f) return k2 * (6.0f * q * q * (q - 1.0f) + 1.0f);
    float q1 = 1.0f - q;
    return k2 * 2.0f * q1 * q1 * q1;
}

/* W_cubic_grad :248-262 */
static inline float W_cubic_grad(float r)
{
    const float k2 = k2_const();
    float q = r / K_H;
    if (q >= 1.0f) return 0.0f;
    if (q <= 0.5f) return 6.0f * k2 * q * (3.0f * q - 2.0f) / K_H;
    float q1 = 1.0f - q;
    return -6.0f * k2 * (q1 * q1) / K_H;
}

/* boundary_sdf: SphKoschier2D_grid_cs.glsl:204-225 (variant 0) /
 * SphWaveKoschier2D_grid_cs.glsl:205-230 (variant 1).  res = (nx, ny, sd, id) */
static void boundary_sdf(const orc_params2* prm, float px, float py, float res[4])
{
    float pl[4][3]; int npl;
    if (prm->variant == ORC_SPH2_WAVE) {
        pl[0][0] = 0.0f; pl[0][1] = 1.0f;
        pl[0][2] = -0.5f * K_VIEW_HEIGHT + 15.0f * K_PARTICLE_RADIUS - K_PARTICLE_RADIUS + prm->bottom;
        pl[1][0] = 1.0f; pl[1][1] = 0.0f; pl[1][2] = -K_PARTICLE_RADIUS;
        pl[2][0] = -1.0f; pl[2][1] = 0.0f; pl[2][2] = prm->view_width - K_PARTICLE_RADIUS;
        pl[3][0] = 0.0f; pl[3][1] = -1.0f; pl[3][2] = K_VIEW_HEIGHT - K_PARTICLE_RADIUS;
        npl = 4;
    } else {
        pl[0][0] = 0.0f; pl[0][1] = 1.0f; pl[0][2] = -K_PARTICLE_RADIUS;
        pl[1][0] = 1.0f; pl[1][1] = 0.0f; pl[1][2] = -K_PARTICLE_RADIUS;
        pl[2][0] = -1.0f; pl[2][1] = 0.0f; pl[2][2] = prm->view_width - K_PARTICLE_RADIUS;
        npl = 3;
    }
    for (int k = 0; k < npl; k++) {
        float d = (pl[k][0] * px + pl[k][1] * py) + pl[k][2];   /* sdPlane: dot(plane.xy,p)+plane.z */
        if (k == 0 || !(res[2] < d)) {                            /* opU: (d1.z<d2.z) ? d1 : d2 */
            res[0] = pl[k][0]; res[1] = pl[k][1]; res[2] = d; res[3] = (float)k;
        }
    }
    if (prm->variant == ORC_SPH2_KOSCHIER) {
        /* orbiting unit circle :212-213,218,223 */
        float cx = 0.5f * prm->view_width + 3.0f * cosf(prm->time);
        float cy = 0.5f * K_VIEW_HEIGHT + 3.0f * sinf(prm->time);
        float qx = px - cx, qy = py - cy;
        float len = length2(qx, qy);
        float d3 = len - 1.0f;
        if (!(res[2] < d3)) { res[0] = qx / len; res[1] = qy / len; res[2] = d3; res[3] = 3.0f; }
    }
}

/* GetWaveNormalHeight :348-364.  returns (n.x, n.y, 0, h), xvel = w[1]/w[0] */
static void wave_normal_height(const orc_params2* prm, const orc_tex* t, float x, float out[4], float* xvel)
{
    float coord = x / prm->view_width;
    float size = (t && t->data) ? (float)t->w : 1.0f; /* textureSize of an unbound sampler: 1 (documented choice) */
    float w[4], we[4], ww[4];
    tex1d_linear(t, coord, w);
    tex1d_linear(t, coord - 1.0f / size, we);
    tex1d_linear(t, coord + 1.0f / size, ww);
    float h = w[0], he = we[0], hw = ww[0];
    float nx = he - hw, ny = 1.0f;
    float len = length2(nx, ny);
    out[0] = nx / len; out[1] = ny / len; out[2] = 0.0f; out[3] = h;
    *xvel = w[1] / w[0];
}

/* InitParticle / init_grid: variant 0 :333-356, variant 1 :368-389 */
void orc_sph2_init(orc_particle2* out, int n, const orc_params2* prm)
{
    int cols = prm->init_width;
    int rows = n / cols;
    for (int ix = 0; ix < n; ix++) {
        int i = ix % cols, j = ix / cols;
        float px, py;
        if (prm->variant == ORC_SPH2_WAVE) {
            float xx = (float)i / (float)cols, yy = (float)j / (float)rows;
            px = prm->view_width * xx; py = (18.0f * K_H) * yy;
            px += K_PARTICLE_RADIUS;
            py += 0.5f * K_VIEW_HEIGHT - 15.0f * K_PARTICLE_RADIUS + K_PARTICLE_RADIUS;
        } else {
            px = K_PARTICLE_DIAM * (float)i; py = K_PARTICLE_DIAM * (float)j;
            px += 0.1f / 6.0f * prm->view_width; py += 0.1f / 6.0f * K_VIEW_HEIGHT;
        }
        orc_particle2* q = &out[ix];
        q->pos[0] = px; q->pos[1] = py; q->pos[2] = 0.0f; q->pos[3] = 1.0f;
        for (int c = 0; c < 4; c++) q->vel[c] = 0.0f;
        q->acc[0] = q->acc[1] = q->acc[2] = 0.0f; q->acc[3] = K_REST_DENS;
    }
}

/* ComputeDensityPressure: variant 0 :359-427, variant 1 :393-487 */
void orc_sph2_density(const orc_particle2* in, orc_particle2* out, int n, const orc_params2* prm,
                      const orc_tex* wave1d, const orc_grid2* g, const int* counter,
                      const int* offset, const int* index_list)
{
    const float PSI = psi_of(prm);
#pragma omp parallel for schedule(dynamic, 64)
    for (int ix = 0; ix < n; ix++) {
        orc_particle2 pi = in[ix];
        if (prm->variant == ORC_SPH2_WAVE) {
            float xvel, w[4];
            wave_normal_height(prm, wave1d, pi.pos[0], w, &xvel);         /* :401-402 */
            const float xv_thresh = 0.05f;
            if (pi.pos[3] == 0.0f && fabsf(xvel) > xv_thresh) pi.pos[3] = 1.0f; /* :406-409 */
            if (pi.pos[3] == 1.0f && fabsf(xvel) < xv_thresh) pi.pos[3] = 0.0f; /* :412-415 */
        }
        float db[4]; boundary_sdf(prm, pi.pos[0], pi.pos[1], db);
        if (db[2] < 0.0f) {                                               /* :419-423 */
            pi.pos[0] -= 0.75f * db[2] * db[0]; pi.pos[1] -= 0.75f * db[2] * db[1];
            float vn = fminf(0.0f, pi.vel[0] * db[0] + pi.vel[1] * db[1]);
            pi.vel[0] -= 1.75f * vn * db[0]; pi.vel[1] -= 1.75f * vn * db[1];
        }
        if (pi.pos[1] > K_VIEW_HEIGHT) { out[ix] = pi; continue; }        /* :427-431 */
        float rho = 0.0f;
        int i0, j0, i1, j1;
        grid2_cell(g, pi.pos[0] - K_H, pi.pos[1] - K_H, &i0, &j0);        /* :436-439 */
        grid2_cell(g, pi.pos[0] + K_H, pi.pos[1] + K_H, &i1, &j1);
        for (int i = i0; i <= i1; i++)
            for (int j = j0; j <= j1; j++) {
                int c = i * g->ncells[1] + j;
                int start = offset[c], count = counter[c];
                for (int l = start; l < start + count; l++) {
                    const orc_particle2* pj = &in[index_list[l]];
                    float rx = pi.pos[0] - pj->pos[0], ry = pi.pos[1] - pj->pos[1];
                    float r2 = fmaf(ry, ry, rx * rx);                     /* dot(rij,rij), canonical */
                    if (r2 < K_HSQ) rho += W_cubic(sqrtf(r2));            /* :456-459 */
                }
            }
        rho = K_MASS * rho;                                               /* :471 */
        if (db[2] < K_H) rho += PSI * W_cubic(fmaxf(0.0f, db[2] + 0.0f * K_PARTICLE_RADIUS)); /* :474-477 */
        rho = fmaxf(K_REST_DENS, rho);                                    /* :479 */
        pi.acc[3] = rho;
        float ratio = rho / K_REST_DENS;
        pi.vel[3] = K_GAS_CONST * (pow3f(ratio) - 1.0f);                  /* :303-307 gamma = 3 */
        out[ix] = pi;
    }
}

/* ComputeForces (+ forward Euler): variant 0 :430-506, variant 1 :490-627 */
void orc_sph2_forces(const orc_particle2* in, orc_particle2* out, int n, const orc_params2* prm,
                     const orc_tex* wave1d, const orc_grid2* g, const int* counter,
                     const int* offset, const int* index_list)
{
    const float PSI = psi_of(prm);
    const float c_visc = -K_VISC * 8.0f * K_MASS;
    const float c_press = K_MASS;
#pragma omp parallel for schedule(dynamic, 64)
    for (int ix = 0; ix < n; ix++) {
        orc_particle2 pi = in[ix];
        if (prm->variant == ORC_SPH2_WAVE && pi.pos[3] == 0.0f) { out[ix] = pi; continue; } /* :494-498 */
        float rho_i = pi.acc[3];
        float ap[2] = {0, 0}, av[2] = {0, 0};
        float acc_press_i = pi.vel[3] / (rho_i * rho_i);
        int i0, j0, i1, j1;
        grid2_cell(g, pi.pos[0] - K_H, pi.pos[1] - K_H, &i0, &j0);
        grid2_cell(g, pi.pos[0] + K_H, pi.pos[1] + K_H, &i1, &j1);
        for (int i = i0; i <= i1; i++)
            for (int j = j0; j <= j1; j++) {
                int c = i * g->ncells[1] + j;
                int start = offset[c], count = counter[c];
                for (int l = start; l < start + count; l++) {
                    int jx = index_list[l];
                    if (jx == ix) continue;
                    const orc_particle2* pj = &in[jx];
                    float rx = pi.pos[0] - pj->pos[0], ry = pi.pos[1] - pj->pos[1];
                    float r = length2(rx, ry);
                    if (r < K_H) {
                        float Wgrad = W_cubic_grad(r);
                        float ux = rx / r, uy = ry / r;
                        float rho_j = pj->acc[3];
                        float s = (acc_press_i + pj->vel[3] / (rho_j * rho_j)) * Wgrad;
                        ap[0] -= s * ux; ap[1] -= s * uy;
                        float vx = pi.vel[0] - pj->vel[0], vy = pi.vel[1] - pj->vel[1];
                        float t = 1.0f / rho_j * (vx * rx + vy * ry) / (r * r + 0.01f * K_HSQ) * Wgrad;
                        av[0] -= t * ux; av[1] -= t * uy;
                    }
                }
            }
        av[0] *= c_visc; av[1] *= c_visc;
        ap[0] *= c_press; ap[1] *= c_press;
        float db[4]; boundary_sdf(prm, pi.pos[0], pi.pos[1], db);
        if (db[2] < K_H) {
            float Wgrad = W_cubic_grad(fmaxf(0.0f, db[2] + 0.0f * K_PARTICLE_RADIUS));
            float s = PSI * acc_press_i * Wgrad;
            ap[0] += s * (-db[0]); ap[1] += s * (-db[1]);
        }
        float atro[2] = {0, 0};
        if (prm->variant == ORC_SPH2_WAVE) {
            float vx, wave[4];
            wave_normal_height(prm, wave1d, pi.pos[0], wave, &vx);       /* :565-567 */
            float hh = wave[3];
            float dh = hh - K_VIEW_HEIGHT / 2.0f;                        /* :586 */
            float wave_mask = smoothstepf(0.0f, 0.2f, fabsf(vx));        /* :587 */
            float dy = pi.pos[1] - hh;                                   /* :599 */
            dy -= 0.5f * dh;                                             /* :600 */
            if (ix % 5 < 4) atro[1] -= wave_mask * 500000.0f * smoothstepf(0.0f, 10.0f, dy); /* :602-603 */
        }
        pi.acc[0] = ap[0] + av[0] + 0.0f + atro[0];                      /* G.xy = (0,-9.8) */
        pi.acc[1] = ap[1] + av[1] + (-9.8f) + atro[1];
        pi.vel[0] += K_DT * pi.acc[0]; pi.vel[1] += K_DT * pi.acc[1];
        pi.pos[0] += K_DT * pi.vel[0]; pi.pos[1] += K_DT * pi.vel[1];
        out[ix] = pi;
    }
}

/* SphUgrid::Compute, SphWave2D/StencilBuffer.cpp:150-179 */
int orc_sph2_step(orc_particle2* buf0, orc_particle2* buf1, int read_index, int n, int substeps,
                  const orc_params2* prm, const orc_tex* wave1d, const orc_grid2* g,
                  int* counter, int* offset, int* index_list, int* cell_of)
{
    orc_particle2* buf[2] = {buf0, buf1};
    for (int s = 0; s < substeps; s++) {
        orc_grid2_build(g, buf[read_index][0].pos, 12, n, cell_of, counter, offset, index_list); /* :163-164 */
        orc_sph2_density(buf[read_index], buf[1 - read_index], n, prm, wave1d, g, counter, offset, index_list);
        read_index = 1 - read_index;                                                             /* PingPong */
        orc_sph2_forces(buf[read_index], buf[1 - read_index], n, prm, wave1d, g, counter, offset, index_list);
        read_index = 1 - read_index;
    }
    return read_index;
}
