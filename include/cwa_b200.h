/*
 * cwa_b200.h -- C ABI of the B200-native CoupledWaterAnimation simulation step.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no CUDA/torch types.  Each
 * entry point names the reference interface it replaces (paths relative to the reference repo).
 * The header-only C++ mirrors of the reference classes (Module / ComputeShader / Buffer /
 * ImageTexture / StencilImage2DTripleBuffered / StencilBuffer / SphUgrid / UniformGridSph2D /
 * UgridParticles3D / ParallelScan) in include/cwa/ are written on top of exactly these symbols.
 *
 * Conventions (mirroring the reference's GL conventions):
 *   - every function returns 0 on success, <0 on error (text via cwa_last_error()); none throws;
 *   - one host thread per context; calls enqueue work on the context's CUDA stream and return
 *     without synchronising unless stated ("synchronises");
 *   - handles are small non-negative ints, -1 is the invalid handle (the GLuint(-1) idiom);
 *   - there is NO CPU fallback: without a CUDA device cwa_create() fails.
 */
#ifndef CWA_B200_H
#define CWA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define CWA_API __declspec(dllexport)
#else
#define CWA_API __attribute__((visibility("default")))
#endif

typedef struct cwa_ctx cwa_ctx;
typedef int cwa_buf;    /* Buffer                        (SphWave2D/Buffer.h:5-24)                  */
typedef int cwa_wave;   /* StencilImage2DTripleBuffered  (StencilImage2DTripleBuffered.h:8-41)       */
typedef int cwa_grid;   /* UniformGridSph2D / UgridParticles3D (UniformGridGpu2D.h:82-142, UniformGridParticles3D.h:50-130) */
typedef int cwa_sph;    /* the 3 SPH compute programs + particle SSBO of Main.cpp:540-557           */
typedef int cwa_stencil1d; /* ImageStencil on a 1-D image  (SphWave2D/StencilImage2D.h:10-66)         */
typedef int cwa_sph2;   /* SphUgrid                      (SphWave2D/StencilBuffer.h:64-79)           */
typedef int cwa_shader; /* ComputeShader                 (CoupledWaterAnimation/ComputeShader.h:7-39) */

/* ---- structs with the reference's memory layout -------------------------------------------- */
/* struct Particle, CoupledWaterAnimation/Main.cpp:167-173 == rho_pres_comp.glsl:14-20 (std430) */
typedef struct { float pos[4], vel[4], force[4], extras[4]; } cwa_particle;      /* 64 B */
/* struct Particle, SphWave2D/Main.cpp:35-40 == SphWaveKoschier2D_grid_cs.glsl:39-44 */
typedef struct { float pos[4], vel[4], acc[4]; } cwa_particle2d;                 /* 48 B */
/* UBO blocks, CoupledWaterAnimation/Main.cpp:184-204 (std140) */
typedef struct { float mass, smoothing_coeff, visc, resting_rho; } cwa_constants_uniform;   /* binding 1 */
typedef struct { float upper[4], lower[4]; } cwa_boundary_uniform;                           /* binding 2 */
typedef struct { float attributes[4], mesh_ws_pos[4]; } cwa_wave_uniforms;                   /* binding 3 */
/* Shader `const`s of the reference promoted to a parameter block (north-star: dt, gravity, k ...).
 * Defaults == rho_pres_comp.glsl:5,41,43; force_comp.glsl:7,52-55; integrate_comp.glsl:8,51,69. */
typedef struct {
    float particle_radius, gas_const, dt, gravity_y;
    float damping, crest_threshold, foam_speed, uv_scale;
    float uv_scale_z;        /* texture t = uv_scale_z * pos.z; 0 (default) = uv_scale, the reference's 2.0*pos.xz */
    float torque_coeff;      /* torque = torque_coeff * cross(pos, force_prev), force_comp.glsl:103-104; 0 (default) = the reference's 0.25 */
    float pad1, pad2;
} cwa_sim_constants;                                                                         /* binding 4 (extension) */
/* UniformGridInfo: 2-D SphWave2D/UniformGridGpu2D.h:89-94, 3-D UniformGrid2D/UniformGridParticles3D.h:60-67 */
typedef struct { float min[4], max[4]; int num_cells[4]; float cell_size[4]; } cwa_grid_info;

enum { CWA_TARGET_SSBO = 0, CWA_TARGET_UBO = 1 };
enum { CWA_UBO_SCENE = 0, CWA_UBO_CONSTANTS = 1, CWA_UBO_BOUNDARY = 2, CWA_UBO_WAVE = 3, CWA_UBO_SIM = 4 };
/* StencilImage2DTripleBuffered.h:26-29 */
enum { CWA_MODE_INIT = 0, CWA_MODE_INIT_FROM_TEXTURE = 1, CWA_MODE_EVOLVE = 2, CWA_MODE_TEST = 10 };
/* which wave shader the stencil object runs: wave_comp.glsl or Wave2DSimp/Wave2D_cs.glsl */
enum { CWA_WAVE_COUPLED = 0, CWA_WAVE_SIMP = 1 };
/* SPH<-wave sampling schedule: as shipped (stale texture-unit-0 binding) or newest */
enum { CWA_COUPLING_AS_SHIPPED = 0, CWA_COUPLING_LATEST = 1 };
/* neighbour search of the 3-D passes: all-pairs as shipped (rho_pres_comp.glsl:60) or uniform grid */
enum { CWA_NEIGHBOURS_ALL_PAIRS = 0, CWA_NEIGHBOURS_GRID = 1 };
enum { CWA_SPH2_KOSCHIER = 0, CWA_SPH2_WAVE = 1 };
enum { CWA_GRID_COUNTER = 0, CWA_GRID_OFFSET = 1, CWA_GRID_INDEX_LIST = 2, CWA_GRID_CELL_OF = 3 };
/* OR into `dim` of cwa_grid_create: linear index (i*Ny + j)*Nz + k instead of the reference's (i*Ny + j)*Nx + k
 * (ugrid_particles_cs.glsl:105-108, sic).  Identical when Nx == Nz; used by slab-local grids of the multi-GPU path. */
enum { CWA_GRID_COMPACT_INDEX = 16 };

/* ---- context --------------------------------------------------------------------------------- */
CWA_API const char* cwa_last_error(void);
CWA_API int  cwa_version(void);
CWA_API int  cwa_create(int device, cwa_ctx** out);                 /* replaces glfw/glew context setup, Main.cpp:990-1019 */
CWA_API void cwa_destroy(cwa_ctx* ctx);
CWA_API int  cwa_synchronize(cwa_ctx* ctx);                         /* glFinish */
CWA_API void* cwa_stream(cwa_ctx* ctx);                             /* cudaStream_t the context enqueues on */
CWA_API int  cwa_device_info(cwa_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
/* device-side timing on the context stream (cudaEvent): used by bench.py for per-kernel roofline */
CWA_API int  cwa_timer_begin(cwa_ctx* ctx);
CWA_API int  cwa_timer_end(cwa_ctx* ctx, float* ms);                /* synchronises */
/* number of kernels this library launched (graph replays count their kernel nodes) */
CWA_API unsigned long long cwa_launch_count(cwa_ctx* ctx);
/* per-kernel device time: every launch between begin/end is bracketed by a CUDA-event pair on the
 * context stream (the GpuTimer of SphWave2D/Timer.cpp:30-72, per kernel instead of per frame) */
CWA_API int  cwa_profile_kernel_count(void);
CWA_API const char* cwa_profile_kernel_name(int id);
CWA_API int  cwa_profile_begin(cwa_ctx* ctx);
CWA_API int  cwa_profile_end(cwa_ctx* ctx, float* ms, int* launches, int cap);   /* synchronises */
/* kernel-variant / staging knobs of the neighbour loops (no reference counterpart: the GLSL has one variant).
 * keys: "nb_config" (8: row-mask kernels, the default; 7: neighbour-list kernels; 0..6: shared-memory-staged "lanes" kernels),
 *       "nb_cap_d", "nb_cap_f" (staged slots of the lanes kernels),
 *       "fused_order" (1: canonical ordering fused into the reorder pass),
 *       "fused_integrate" (1: in a full step the force kernels also run the epilogue + integrate; default 0:
 *       measured 6 us faster per C4 frame while few targets are queued, 120 us slower once clumps dominate),
 *       "pipeline" (frames of ONE cwa_coupled_step call; bit 0: the wave stencil of frame f runs on a side stream next to
 *       the grid build of frame f+1, bit 1: the integrate pass of frame f does the cell hash + count of frame f+1;
 *       default 3; results and every readable array are bit-identical to 0),
 *       "scan_config" (tile shape of the look-back scan: 0 = 256 threads x 16 items, 1 = 512 x 32, 2 = 512 x 32 with
 *       warp-striped loads (default), 3 = 1024 x 16 warp-striped),
 *       "nbr_k" (neighbour-list entries per target, multiple of 4 in [8, 256], default 64; set before the SPH object is used) and
 *       "extreme_candidates" (a target with more candidates is finished by a whole warp, default 192),
 *       "wave_transpose" (1, default: the SPH passes sample a transposed copy of the bound wave level -- the grid runs fastest along
 *       z = the texture's t axis -- rebuilt only when that level changed; same texels, bit-identical results),
 *       "inplace_max" (a clump target with at most this many candidates is finished by its own thread in the density pass, default 640),
 *       "heavy_sub_warp" (1, default: queued clump targets of the force pass get eight lanes each, four per warp, when the queue is longer
 *       than the launch), "allpairs_balanced" (0: tiles of 64 targets; 1 / 2: one CTA per SM for the density pass / both passes, default 2),
 *       "allpairs_cull" (1, default: the balanced all-pairs kernels skip candidate tiles whose bounding box is farther than h from the
 *       CTA's targets; bit-identical results), "graph" (1, default: the all-pairs frame and the 2-D frame replay as one CUDA graph),
 *       "slab_ahead" (1, default: count-ahead across the slab exchange) and
 *       "pdl" (bit mask of the launches that carry the programmatic-dependent-launch attribute: 1 scan, 2 insert, 4 reorder, 8 density,
 *       16 density of queued targets, 32 force, 64 force of queued targets, 128 integrate, 256 the 2-D frame; default all but 2 and 4). */
CWA_API int  cwa_set_tuning(cwa_ctx* ctx, const char* key, int value);

/* ---- Buffer: Init / BufferSubData / BindBufferBase / DebugRead*  (SphWave2D/Buffer.cpp:5-83) -- */
CWA_API int cwa_buffer_create(cwa_ctx* ctx, size_t bytes, const void* host_or_null, cwa_buf* out); /* glNamedBufferStorage */
CWA_API int cwa_buffer_wrap(cwa_ctx* ctx, void* device_ptr, size_t bytes, cwa_buf* out);            /* external memory (CUDA-GL interop map, torch) */
CWA_API int cwa_buffer_destroy(cwa_ctx* ctx, cwa_buf b);                                             /* glDeleteBuffers */
CWA_API int cwa_buffer_sub_data(cwa_ctx* ctx, cwa_buf b, size_t off, size_t bytes, const void* host);/* glNamedBufferSubData */
CWA_API int cwa_buffer_read(cwa_ctx* ctx, cwa_buf b, size_t off, size_t bytes, void* host);          /* glGetNamedBufferSubData; synchronises */
CWA_API int cwa_buffer_read_async(cwa_ctx* ctx, cwa_buf b, size_t off, size_t bytes, void* host);    /* same, enqueued only: `host` (pinned) is valid after cwa_synchronize */
CWA_API int cwa_buffer_copy(cwa_ctx* ctx, cwa_buf src, cwa_buf dst, size_t soff, size_t doff, size_t bytes); /* glCopyNamedBufferSubData */
CWA_API int cwa_buffer_bind_base(cwa_ctx* ctx, int target, int binding, cwa_buf b);                  /* glBindBufferBase */
CWA_API int cwa_buffer_device_ptr(cwa_ctx* ctx, cwa_buf b, void** ptr, size_t* bytes);
/* the context's own parameter blocks (created with the reference defaults, bound at UBO 1..4) */
CWA_API int cwa_default_ubo(cwa_ctx* ctx, int binding, cwa_buf* out);

/* ---- ParallelScan::Compute (SphWave2D/ParallelScan.cpp:43-95 + prefix_sum_cs.glsl) ------------ */
/* exclusive prefix sum of n int32 (any n >= 1; the reference asserts n is a power of two) */
CWA_API int cwa_scan_exclusive(cwa_ctx* ctx, cwa_buf in, cwa_buf out, int n);

/* ---- uniform grid: ctor / Init / CollisionQuery|BuildGrid ------------------------------------- */
/* dim = 2: UniformGridSph2D (UniformGridGpu2D.cpp:156-265, uniform_grid_sph_cs.glsl)
 * dim = 3: UgridParticles3D (UniformGridParticles3D.cpp:92-211, ugrid_particles_cs.glsl) */
CWA_API int cwa_grid_create(cwa_ctx* ctx, int dim, const float* mn, const float* mx, const int* num_cells,
                            int max_particles, cwa_grid* out);
CWA_API int cwa_grid_destroy(cwa_ctx* ctx, cwa_grid g);
CWA_API int cwa_grid_get_info(cwa_ctx* ctx, cwa_grid g, cwa_grid_info* out, int* num_cells_total);
/* build over n particles whose pos vec4 sits at byte offset 0 of each `stride_bytes` record */
CWA_API int cwa_grid_build(cwa_ctx* ctx, cwa_grid g, cwa_buf particles, int stride_bytes, int n);
/* which = CWA_GRID_*; counts ints copied to host; synchronises (DebugReadInt) */
CWA_API int cwa_grid_read(cwa_ctx* ctx, cwa_grid g, int which, int* host, int count);
CWA_API int cwa_grid_buffer(cwa_ctx* ctx, cwa_grid g, int which, cwa_buf* out);   /* mGridCounter / mGridOffset / mIndexList as Buffers */

/* ---- StencilImage2DTripleBuffered: Init / Reinit / ReinitFromTexture / Compute / PingPong ------ */
CWA_API int cwa_wave_create(cwa_ctx* ctx, int w, int h, int channels /*1 scalar | 4 RGBA32F*/, int variant, cwa_wave* out); /* Init() :10-31 (runs Reinit) */
CWA_API int cwa_wave_destroy(cwa_ctx* ctx, cwa_wave w);
CWA_API int cwa_wave_reinit(cwa_ctx* ctx, cwa_wave w);                                   /* Reinit() :42-59 */
CWA_API int cwa_wave_reinit_from_texture(cwa_ctx* ctx, cwa_wave w, const float* rgba, int tw, int th); /* ReinitFromTexture :61-77 */
CWA_API int cwa_wave_compute(cwa_ctx* ctx, cwa_wave w, int nsteps);                      /* Compute() :79-95, nsteps times */
CWA_API int cwa_wave_pingpong(cwa_ctx* ctx, cwa_wave w);                                 /* PingPong() :33-40 */
CWA_API int cwa_wave_set_evolve(cwa_ctx* ctx, cwa_wave w, int evolve);                   /* mEvolve checkbox */
CWA_API int cwa_wave_set_params(cwa_ctx* ctx, cwa_wave w, float lambda, float atten, float beta); /* Wave2D_cs.glsl:17-19 uniforms (variant SIMP) */
CWA_API int cwa_wave_resize(cwa_ctx* ctx, cwa_wave w, int nw, int nh);                   /* DrawGui "Apply" :117-123 */
/* state of the rotation, as the reference exposes it: indices (GetReadImage(i)/GetWriteImage) and units */
CWA_API int cwa_wave_state(cwa_ctx* ctx, cwa_wave w, int read_index[2], int* write_index, int unit[3], int* tex_unit0_image);
/* display(): wave2d.GetReadImage(0).BindTextureUnit() -- binds at that image's mUnit (Main.cpp:413, F5) */
CWA_API int cwa_wave_bind_texture_unit(cwa_ctx* ctx, cwa_wave w);
/* image by physical index 0..2 (ImageTexture), or by role: 0 newest, 1 previous, 2 next output */
CWA_API int cwa_wave_read_image(cwa_ctx* ctx, cwa_wave w, int image, float* host);       /* synchronises */
CWA_API int cwa_wave_read_image_async(cwa_ctx* ctx, cwa_wave w, int image, float* host); /* enqueued only: valid after cwa_synchronize */
CWA_API int cwa_wave_mark_written(cwa_ctx* ctx, cwa_wave w, int image);  /* image was written through a raw device pointer (interop map, NCCL receive) */
CWA_API int cwa_wave_write_image(cwa_ctx* ctx, cwa_wave w, int image, const float* host);
CWA_API int cwa_wave_role_image(cwa_ctx* ctx, cwa_wave w, int role, int* image);
CWA_API int cwa_wave_image_buffer(cwa_ctx* ctx, cwa_wave w, int image, cwa_buf* out);    /* GetTexture() for interop */
CWA_API int cwa_wave_size(cwa_ctx* ctx, cwa_wave w, int* width, int* height, int* channels);

/* ---- 3-D SPH passes on the particle SSBO (Main.cpp:540-557) ------------------------------------ */
/* particles: cwa_particle[n] buffer (SSBO binding 0).  grid = -1 -> all-pairs as shipped. */
CWA_API int cwa_sph_create(cwa_ctx* ctx, cwa_buf particles, int n, cwa_grid grid, cwa_sph* out);
CWA_API int cwa_sph_destroy(cwa_ctx* ctx, cwa_sph s);
/* sampler unit 0 for the SPH passes: an image of `w` (physical index) or -1 = unbound (samples 0) */
CWA_API int cwa_sph_bind_wave(cwa_ctx* ctx, cwa_sph s, cwa_wave w, int image);
CWA_API int cwa_sph_rho_pres(cwa_ctx* ctx, cwa_sph s);      /* compute_programs[0]: rho_pres_comp.glsl  */
CWA_API int cwa_sph_force(cwa_ctx* ctx, cwa_sph s);         /* compute_programs[1]: force_comp.glsl     */
CWA_API int cwa_sph_integrate(cwa_ctx* ctx, cwa_sph s);     /* compute_programs[2]: integrate_comp.glsl */
CWA_API int cwa_sph_step(cwa_ctx* ctx, cwa_sph s, int nsteps); /* the three passes, fused schedule */
/* per-particle neighbour counts (r < h, self included) of the current positions; synchronises */
CWA_API int cwa_sph_neighbour_count(cwa_ctx* ctx, cwa_sph s, int* host);
/* make_cube + init_particles (Main.cpp:735-776) generalised to nx*ny*nz; writes the bound SSBO */
CWA_API int cwa_sph_init_cube(cwa_ctx* ctx, cwa_sph s, int nx, int ny, int nz);

/* ---- per-frame entry points named by the north-star ------------------------------------------- */
/* idle() of Main.cpp:530-562 + the display() texture bind: nframes x (rho, force, integrate, wave) */
CWA_API int cwa_coupled_step(cwa_ctx* ctx, cwa_sph s, cwa_wave w, int nframes, int coupling);
/* bind the objects sph_step / wave_step act on (the globals of Main.cpp:74-75) */
CWA_API int cwa_bind_scene(cwa_ctx* ctx, cwa_sph s, cwa_wave w);
CWA_API int sph_step(cwa_ctx* ctx, int nsteps);             /* ComputeShader::Dispatch x3 -> one call */
CWA_API int wave_step(cwa_ctx* ctx, int nsteps);            /* StencilImage2DTripleBuffered::Compute   */

/* ---- multi-GPU decomposition primitives (no reference counterpart: the reference is single-GPU) ---- */
/* number of live particles at the front of the SSBO (owned + ghosts); <= the count given to cwa_sph_create */
CWA_API int cwa_sph_set_count(cwa_ctx* ctx, cwa_sph s, int n);
/* stable copy of the particles of src[0..n) whose pos[axis] satisfies the predicate into dst[dst_offset..]:
 * kind 0: a <= x < b, 1: x < a, 2: x >= a, 3: !(x < a) && !(x >= b) (NaN stays).  *count = copied; synchronises */
CWA_API int cwa_particles_copy_if(cwa_ctx* ctx, cwa_buf src, int n, int axis, int kind, float a, float b,
                                  cwa_buf dst, int dst_offset, int* count);
/* One exchange per frame and neighbour.  Message = 64-byte records: [0] header {int migrants, int ghosts, int overflow},
 * [1, 1+cap_mig) migrants (left the slab [z_lo, z_hi) through this face; marked dead in the SSBO: pos = NaN, pos.w = -1),
 * [1+cap_mig, 1+cap_mig+cap_ghost) ghosts (within `band` of the face).  msg_left / msg_right = -1: no neighbour there. */
CWA_API int cwa_slab_pack(cwa_ctx* ctx, cwa_buf particles, int n_owned, float z_lo, float z_hi, float band,
                          cwa_buf msg_left, cwa_buf msg_right, int cap_mig, int cap_ghost);
/* cwa_sph_step(1) on the bound wave whose integrate pass ALSO packs the messages of the next exchange (owned = original slots
 * [0, n_owned); same selection, layout and dead-slot marking as cwa_slab_pack, applied to the records while they are in registers).
 * The next cwa_slab_pack call with the same arguments is then a no-op. */
CWA_API int cwa_sph_step_slab(cwa_ctx* ctx, cwa_sph s, int n_owned, float z_lo, float z_hi, float band,
                              cwa_buf msg_left, cwa_buf msg_right, int cap_mig, int cap_ghost);
/* append the received migrants behind the owned range, then the ghosts: the received ones plus the migrants of this
 * rank's own outgoing messages (sent_*; still neighbours here this frame).  counts[4] = {owned range, owned + ghosts,
 * flags (1: a sender overflowed, 2: capacity exceeded), migrants adopted}; the one host synchronisation of a frame */
CWA_API int cwa_slab_unpack(cwa_ctx* ctx, cwa_buf particles, int n_owned, cwa_buf rcv_left, cwa_buf rcv_right,
                            cwa_buf sent_left, cwa_buf sent_right, int cap_mig, int cap_ghost, int* counts);
/* squeeze dead slots out of the owned range (rare; stable; synchronises) */
CWA_API int cwa_slab_compact(cwa_ctx* ctx, cwa_buf particles, int n_owned, cwa_buf scratch, int* n_live);
/* wave field stored as a row block: global rows [row0, row0+rows) of a field h_global rows tall (owned rows + halos) */
CWA_API int cwa_wave_create_block(cwa_ctx* ctx, int w, int h_global, int row0, int rows, int channels, int variant, cwa_wave* out);
/* replicated copy of the GLOBAL last row of physical image `image` (WaveNormal's uv+(0,1) tap clamps to it) */
CWA_API int cwa_wave_last_row_buffer(cwa_ctx* ctx, cwa_wave w, int image, cwa_buf* out);

/* ---- the slab-decomposed coupled frame behind the C ABI (csrc/slab.cu; SURVEY 8e / 8b: cwa_create(device, n_devices)) --------
 * N ranks reproduce the single-GPU idle() of Main.cpp:540-561 (+ the display() bind :413).  A rank = one context + one cwa_slab
 * object; ranks are processes (one per GPU: exchange the 64-byte handles of cwa_slab_export and cwa_slab_connect them) or contexts
 * of ONE process on one or several devices (cwa_slab_mailbox + cwa_slab_connect(direct pointer), stepped with cwa_slab_group_step).
 * Per frame there is no host synchronisation, no collective library call and no size negotiation: messages travel by peer
 * stores into the neighbour's mailbox, arrival is a device-side flag, particle counts live on the device. */
typedef int cwa_slab;
typedef struct {
    int   rank, world;
    int   wave_w, wave_h, wave_ch;   /* the GLOBAL height field */
    int   row_lo, row_hi;            /* owned rows */
    int   store_lo, store_hi;        /* stored rows = owned + sampling halos (what cwa_wave_create_block is given) */
    int   left_store_hi;             /* rows [row_lo, left_store_hi) are the left neighbour's upper halo */
    int   right_store_lo;            /* rows [right_store_lo, row_hi) are the right neighbour's lower halo */
    int   halo_rows_max;             /* sizes the wave mailboxes identically on every rank */
    float z_lo, z_hi, band;          /* owned particles: z_lo <= z < z_hi; ghost layer `band` (2h) on either side of a face */
    int   cap_mig, cap_ghost;        /* message capacities in particles (identical on every rank) */
    int   capacity;                  /* particle records the SSBO holds: owned + ghosts + arrivals */
    int   timeout_ms;                /* a wait for a peer that lasts longer raises an error bit instead of hanging (default 10000) */
} cwa_slab_desc;
/* row blocks / z slabs of rank `rank` (row_bounds: world+1 ascending rows 0..wave_h, or NULL = even split); fills everything but
 * cap_mig, cap_ghost and capacity.  Pure host logic (no device needed). */
CWA_API int cwa_slab_plan(int world, int rank, int wave_w, int wave_h, int wave_ch, double uv_scale_z, double h, const int* row_bounds,
                          cwa_slab_desc* out);
/* s: SPH object on a uniform grid whose particle buffer holds `capacity` records; w: cwa_wave_create_block(store_lo, store_hi - store_lo) */
CWA_API int cwa_slab_create(cwa_ctx* ctx, const cwa_slab_desc* desc, cwa_sph s, cwa_wave w, cwa_slab* out);
CWA_API int cwa_slab_destroy(cwa_ctx* ctx, cwa_slab sl);
CWA_API int cwa_slab_export(cwa_ctx* ctx, cwa_slab sl, void* handle64);                 /* cudaIpcMemHandle_t of the mailbox */
CWA_API int cwa_slab_mailbox(cwa_ctx* ctx, cwa_slab sl, void** ptr, size_t* bytes);     /* same-process peers connect with the pointer */
CWA_API int cwa_slab_connect(cwa_ctx* ctx, cwa_slab sl, int peer_rank, const void* handle64_or_null, void* direct_ptr_or_null);
/* the application (re)wrote particles [0, n_owned) of the SSBO: owned range = n_owned, no ghosts, no free slots */
CWA_API int cwa_slab_set_owned(cwa_ctx* ctx, cwa_slab sl, int n_owned);
/* nframes coupled frames of this rank's slab; every rank makes the same calls.  Enqueues only. */
CWA_API int cwa_slab_step(cwa_ctx* ctx, cwa_slab sl, int nframes, int coupling);
/* all ranks of one process, one host thread (ctxs[r] / slabs[r] = rank r) */
CWA_API int cwa_slab_group_step(cwa_ctx* const* ctxs, const cwa_slab* slabs, int n, int nframes, int coupling);
/* counts[8] = {owned range (dead slots included), owned + ghosts, free slots, error bits (1 sender overflow, 2 capacity, 4 / 8 timeout
 * waiting for particles / wave rows), migrants adopted (low, high), particle messages consumed, wave messages consumed}; synchronises */
CWA_API int cwa_slab_counts(cwa_ctx* ctx, cwa_slab sl, int* counts);
CWA_API int cwa_slab_counts_async(cwa_ctx* ctx, cwa_slab sl, int* pinned_counts4);     /* first four of the above, valid after cwa_synchronize */

/* ---- CUDA-GL interop hand-off (csrc/interop.cu).  The renderer keeps its GL objects; none of this is on the timed path.
 * cwa_gl_available() = 1 when the library was built with cuda_gl_interop.h; the calls need a current GL context on the calling thread
 * and fail with the CUDA error text otherwise. */
CWA_API int cwa_gl_available(void);
/* the particle SSBO / vertex buffer of init_particles() (Main.cpp:779-790): register once, map every frame, build the SPH object on
 * the cwa_buf the map returns (same handle every frame), unmap before the draw calls */
CWA_API int cwa_gl_register_buffer(cwa_ctx* ctx, unsigned gl_buffer, int* resource);
CWA_API int cwa_gl_map_buffer(cwa_ctx* ctx, int resource, cwa_buf* out);
CWA_API int cwa_gl_unmap(cwa_ctx* ctx, int resource);
/* a wave level's GL texture (ImageTexture::GetTexture(), RGBA32F as shipped or R32F; gl_target 0 = GL_TEXTURE_2D) and the per-frame
 * copy that replaces display()'s GetReadImage(0).BindTextureUnit() (Main.cpp:413) */
CWA_API int cwa_gl_register_image(cwa_ctx* ctx, unsigned gl_texture, unsigned gl_target, int* resource);
CWA_API int cwa_gl_copy_wave_to_image(cwa_ctx* ctx, int resource, cwa_wave w, int image);
CWA_API int cwa_gl_unregister(cwa_ctx* ctx, int resource);
/* ReinitFromTexture (StencilImage2DTripleBuffered.cpp:61-77) from the decoded bytes of an RGBA8 init texture (a PNG of init-textures/) */
CWA_API int cwa_wave_reinit_from_rgba8(cwa_ctx* ctx, cwa_wave w, const unsigned char* rgba8, int tw, int th);

/* ---- 2-D Koschier SPH on the uniform grid: SphUgrid (SphWave2D/StencilBuffer.cpp:138-179) ------ */
CWA_API int cwa_sph2_create(cwa_ctx* ctx, int n, int variant, cwa_grid grid, cwa_sph2* out); /* Init + Reinit (MODE_INIT) */
CWA_API int cwa_sph2_destroy(cwa_ctx* ctx, cwa_sph2 s);
CWA_API int cwa_sph2_reinit(cwa_ctx* ctx, cwa_sph2 s);                                   /* StencilBuffer::Reinit :38-51 */
CWA_API int cwa_sph2_set_substeps(cwa_ctx* ctx, cwa_sph2 s, int substeps);               /* SetSubsteps */
CWA_API int cwa_sph2_set_uniforms(cwa_ctx* ctx, cwa_sph2 s, float time, float bottom, float psi /*<0: default*/, int init_width);
CWA_API int cwa_sph2_set_view_width(cwa_ctx* ctx, cwa_sph2 s, float view_width);        /* shader const VIEW_WIDTH = 9.6 (extension: wider tanks for C2) */
CWA_API int cwa_sph2_bind_wave1d(cwa_ctx* ctx, cwa_sph2 s, cwa_buf rgba_or_minus1, int width); /* sampler1D wave_tex, unit 0 */
CWA_API int cwa_sph2_compute(cwa_ctx* ctx, cwa_sph2 s, int nframes);                     /* SphUgrid::Compute */
CWA_API int cwa_sph2_read(cwa_ctx* ctx, cwa_sph2 s, cwa_particle2d* host);               /* GetReadBuffer(); synchronises */
CWA_API int cwa_sph2_write(cwa_ctx* ctx, cwa_sph2 s, const cwa_particle2d* host);
CWA_API int cwa_sph2_read_buffer(cwa_ctx* ctx, cwa_sph2 s, cwa_buf* out);

/* ---- ImageStencil + the 1-D wave shaders of the 2-D app ---------------------------------------------------------
 * Shallow1D_cs.glsl (two-phase Lax-Wendroff shallow water, RGBA32F texel = (h, uh, hm, uhm), double buffered, modes 2,3 per frame:
 * SphWave2D/Main.cpp:77-97) and Wave1D_cs.glsl (damped wave equation, texel = (u, v, a, -), triple buffered, 10 substeps:
 * Main.cpp:63-75).  The object reproduces ImageStencil's bookkeeping (mReadIndex / mWriteIndex / per-image unit, PingPong as
 * written); the shaders address the images by unit.  A whole Compute() call is one kernel launch. */
enum { CWA_STENCIL1D_SHALLOW = 0, CWA_STENCIL1D_WAVE = 1 };
enum { CWA_BC_REFLECT = 0, CWA_BC_FREE = 1, CWA_BC_FIXED = 2 };       /* shader const BC = FREE (Shallow1D_cs.glsl:84-87), promoted */
CWA_API int cwa_stencil1d_create(cwa_ctx* ctx, int shader, int width, cwa_stencil1d* out);   /* SetShader + SetNumBuffers + SetGridSize + Init (ends with Reinit) */
CWA_API int cwa_stencil1d_destroy(cwa_ctx* ctx, cwa_stencil1d s);
CWA_API int cwa_stencil1d_reinit(cwa_ctx* ctx, cwa_stencil1d s);                             /* Reinit :85-105 */
CWA_API int cwa_stencil1d_reinit_from_texture(cwa_ctx* ctx, cwa_stencil1d s, const float* rgba, int width); /* ReinitFromTexture :122-140; synchronises */
CWA_API int cwa_stencil1d_compute(cwa_ctx* ctx, cwa_stencil1d s, int nframes);               /* Compute :142-164, nframes times */
CWA_API int cwa_stencil1d_compute_func(cwa_ctx* ctx, cwa_stencil1d s, int mode);             /* ComputeFunc(mode) :107-120, e.g. Splash = mode 1 */
/* uniforms at locations 2..5 (lambda; dx for the shallow-water shader, atten for the wave shader; beta; boundary) + the BC const */
CWA_API int cwa_stencil1d_set_params(cwa_ctx* ctx, cwa_stencil1d s, float lambda, float dx_or_atten, float beta,
                                     float boundary0, float boundary1, int bc);
CWA_API int cwa_stencil1d_pingpong(cwa_ctx* ctx, cwa_stencil1d s);                           /* PingPong :67-83 (after a ComputeShader dispatch) */
CWA_API int cwa_stencil1d_set_substeps(cwa_ctx* ctx, cwa_stencil1d s, int substeps);         /* SetSubsteps */
CWA_API int cwa_stencil1d_set_iterate(cwa_ctx* ctx, cwa_stencil1d s, int iterate);           /* mIterate (GUI checkbox) */
CWA_API int cwa_stencil1d_state(cwa_ctx* ctx, cwa_stencil1d s, int* num_images, int read_index[2], int* write_index, int unit[3]);
CWA_API int cwa_stencil1d_image_buffer(cwa_ctx* ctx, cwa_stencil1d s, int image, cwa_buf* out); /* GetReadImage(i) = image read_index[i]; bind with cwa_sph2_bind_wave1d */
CWA_API int cwa_stencil1d_read_image(cwa_ctx* ctx, cwa_stencil1d s, int image, float* host_rgba);        /* synchronises */
CWA_API int cwa_stencil1d_write_image(cwa_ctx* ctx, cwa_stencil1d s, int image, const float* host_rgba); /* synchronises */

/* ---- parameter reflection (the reference's UniformGui lists a program's uniforms with glGetActiveUniform) --------
 * name -> field of the block currently bound at UBO 1..4: "mass", "smoothing_coeff", "visc", "resting_rho", "upper.x" .. "lower.w",
 * "wave.lambda", "wave.atten", "wave.beta", "wave.type", "mesh_ws_pos.x" .., "particle_radius", "gas_const", "dt", "gravity_y",
 * "damping", "crest_threshold", "foam_speed", "uv_scale", "uv_scale_z", "torque_coeff". */
CWA_API int cwa_param_count(void);
CWA_API int cwa_param_info(int index, const char** name, int* ubo_binding, int* byte_offset, const char** origin);
CWA_API int cwa_param_set(cwa_ctx* ctx, const char* name, float value);
CWA_API int cwa_param_get(cwa_ctx* ctx, const char* name, float* value);                 /* synchronises */
/* ---- checkpoint: one little-endian file {256-byte header: magic "CWACKPT1", frame, sizes, triple-buffer bookkeeping,
 * the four parameter blocks; particle SSBO; the three physical wave images}.  Loading into objects of the same sizes resumes the run
 * bit for bit (layout: csrc/state.cu, reader/writer for tools: coupledwateranimation_b200/checkpoint.py). */
CWA_API int cwa_checkpoint_save(cwa_ctx* ctx, cwa_sph s, cwa_wave w, unsigned long long frame, const char* path);   /* synchronises */
CWA_API int cwa_checkpoint_load(cwa_ctx* ctx, cwa_sph s, cwa_wave w, const char* path, unsigned long long* frame);  /* synchronises */

/* ---- ComputeShader: Init / SetMode / SetGridSize / Dispatch (ComputeShader.cpp:9-56) ----------- */
/* glsl_filename selects the CUDA kernel set that replaces that shader; unknown names fail like
 * InitShader() returning -1.  Dispatch acts on the currently bound buffers/images like
 * glDispatchCompute.  Supported: rho_pres_comp.glsl, force_comp.glsl, integrate_comp.glsl,
 * wave_comp.glsl, Wave2D_cs.glsl, prefix_sum_cs.glsl, Shallow1D_cs.glsl, Wave1D_cs.glsl (the last two dispatch on the
 * cwa_stencil1d object given to cwa_shader_bind_object, with that object's parameters). */
CWA_API int cwa_shader_create(cwa_ctx* ctx, const char* glsl_filename, cwa_shader* out);
CWA_API int cwa_shader_set_mode(cwa_ctx* ctx, cwa_shader s, int mode);
CWA_API int cwa_shader_set_uniform_i(cwa_ctx* ctx, cwa_shader s, int location, int v);
CWA_API int cwa_shader_set_uniform_f(cwa_ctx* ctx, cwa_shader s, int location, float v);
CWA_API int cwa_shader_bind_object(cwa_ctx* ctx, cwa_shader s, int object_handle);       /* the cwa_sph / cwa_wave it drives */
CWA_API int cwa_shader_dispatch(cwa_ctx* ctx, cwa_shader s, int gx, int gy, int gz);

#ifdef __cplusplus
}
#endif
#endif /* CWA_B200_H */
