// internal.cuh -- shared host/device declarations of libcwa_b200 (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <cstdlib>
#include <vector>

#include "../../include/cwa_b200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing: no exceptions cross the C ABI
// ---------------------------------------------------------------------------------------------
void cwa_set_error(const char* fmt, ...);

#define CWA_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            cwa_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            (void)cudaGetLastError();   /* reported: a non-sticky error must not resurface in a later, unrelated call */ \
            return -2;                                                                          \
        }                                                                                       \
    } while (0)

#define CWA_CHECK(cond, ...)                                                                    \
    do {                                                                                        \
        if (!(cond)) { cwa_set_error(__VA_ARGS__); return -1; }                                 \
    } while (0)

#define CWA_TRY(expr)                                                                           \
    do { int _r = (expr); if (_r != 0) return _r; } while (0)

// ---------------------------------------------------------------------------------------------
// device-visible parameter views
// ---------------------------------------------------------------------------------------------
// Pointers to the currently bound UBO-equivalents (std140 blocks of Main.cpp:184-204 + the
// promoted shader constants).  Passed by value to the kernels, read through the constant cache.
struct ParamPtrs {
    const cwa_constants_uniform* constants;   // binding 1
    const cwa_boundary_uniform*  boundary;    // binding 2
    const cwa_wave_uniforms*     wave;        // binding 3
    const cwa_sim_constants*     sim;         // binding 4
};

// sampler2D wave_tex (binding 0): LINEAR / CLAMP_TO_EDGE, .r channel.  data == nullptr: unbound.
// Multi-GPU: `data` may hold only the rows [row0, row0 + h) of a field that is h_global rows tall
// (row-block decomposition); `last_row` then points at a copy of the GLOBAL last row, which
// WaveNormal's uv + (0,1) tap clamps to from everywhere (force_comp.glsl:136).
struct TexView {
    const float* data;
    int w, h, ch;
    int row0, h_global;
    const float* last_row;
    // local single-channel fast path (tex_view_is_local): texel (i, j) = tdata[i * tsi + j * tsj].  Row-major as stored
    // (tsi = 1, tsj = w), or the TRANSPOSED sampling copy (tsi = h, tsj = 1): the uniform grid runs fastest along z, which
    // maps to the texture's t axis, so a warp of cell-ordered particles walks down a texture COLUMN -- 32 rows = 32 cache
    // lines per tap in the row-major image, a handful in the transposed one.
    const float* tdata;
    int tsi, tsj;
};

// UniformGridInfo as the kernels see it.
struct GridView {
    float min[3];
    float cell[3];
    float max[3];
    float inv_cell[3];   // 1/cell, only for the conservative neighbour-range estimate (never for the hash)
    int   n[3];          // Nx, Ny, Nz (Nz = 1 in 2-D)
    int   dim;           // 2 or 3
    int   num_cells;     // allocated/scanned linear cells
    int   kstride;       // 3-D: stride of k in the linear index -- Nx as in the reference (sic), Nz when compact
    // linear index: 3-D (i*Ny + j)*Nx + k (ugrid_particles_cs.glsl:105-108), 2-D i*Ny + j
    // (uniform_grid_sph_cs.glsl:149-152).  A "row" is the run of cells sharing everything but the
    // fastest coordinate (k in 3-D, j in 2-D); rows are contiguous in the sorted particle order.
};

// ---------------------------------------------------------------------------------------------
// host objects behind the handles
// ---------------------------------------------------------------------------------------------
struct BufferObj {
    void*  ptr   = nullptr;
    size_t bytes = 0;
    bool   owned = false;
    bool   live  = false;
};

struct GridObj {
    bool live = false;
    int  dim = 3;
    cwa_grid_info info{};
    GridView view{};
    int  max_particles = 0;
    int  n_built = 0;              // particle count of the last build
    bool cleared_ahead = false;    // counter + scan state were cleared behind the last build's scan: the next build waits for ev_pipe[1] instead of clearing
    // one allocation, cleared by a single memset per build: [counter C][ticket][tile_state]
    int* counter = nullptr;        // int[C]
    int* ticket = nullptr;         // scan tile ticket
    unsigned long long* tile_state = nullptr;
    size_t clear_bytes = 0;
    int* offset = nullptr;         // int[C+1]  (offset[C] = number of inserted particles)
    int* cell_of = nullptr;        // int[max_particles]  (-1: not inserted)
    int* rank = nullptr;           // int[max_particles]  arrival rank inside the cell
    int* arrival = nullptr;        // int[max_particles]  index list in arrival order (input of the canonical ordering)
    int* index_list = nullptr;     // int[max_particles]  canonical: ascending id within a cell
    cwa_buf buf_counter = -1, buf_offset = -1, buf_index = -1, buf_cell_of = -1;
};

struct WaveObj {
    bool live = false;
    int  w = 64, h = 64, ch = 1, variant = CWA_WAVE_COUPLED;
    float* image[3] = {nullptr, nullptr, nullptr};
    cwa_buf image_buf[3] = {-1, -1, -1};
    // StencilImage2DTripleBuffered.h:35-37 and ImageTexture::mUnit
    int  read_index[2] = {0, 1};
    int  write_index = 2;
    int  unit[3] = {0, 1, 2};
    int  tex_unit0 = -1;           // physical image bound to GL texture unit 0 (-1: unbound)
    bool evolve = true;
    float* simp_params = nullptr;  // device float4 (lambda, atten, beta, 0) for variant SIMP
    // row-block decomposition (multi-GPU): the arrays hold global rows [row0, row0 + h) of a field
    // that is h_global rows tall; last_row[i] = copy of the global last row of physical image i
    int  row0 = 0, h_global = 64;
    float* last_row[3] = {nullptr, nullptr, nullptr};
    cwa_buf last_row_buf[3] = {-1, -1, -1};
    // TMA descriptors (scalar fast path): per physical image, halo box and plain box
    CUtensorMap tmap_halo[3];
    CUtensorMap tmap_core[3];
    bool tma_ok = false;
    // transposed sampling copy of ONE image (the one the SPH passes sample), refreshed when that image changed
    unsigned long long version[3] = {1, 1, 1};    // bumped by every library path that writes image i
    float* imageT = nullptr;
    int    imageT_of = -1;
    unsigned long long imageT_version = 0;
    bool   raw_exposed = false;                   // an application holds a raw pointer to an image: contents can change unseen ...
    bool   explicit_touch = false;                // ... unless it reports its writes with cwa_wave_mark_written
};

struct SphObj {
    bool live = false;
    cwa_buf particles = -1;
    int  n = 0;
    const int* n_dev = nullptr;    // slab decomposition: the live count (owned + ghosts) is device-resident; `n` is then only the launch bound
    int  capacity = 0;             // particles the scratch arrays were sized for (cwa_sph_set_count limit)
    cwa_grid grid = -1;            // -1: all-pairs
    cwa_wave wave = -1;            // sampler binding
    int  wave_image = -1;          // physical image, -1 unbound
    // cell-ordered snapshot (grid mode) -- see DESIGN.md "data layout"
    float4 *posS = nullptr, *velS = nullptr, *forceS = nullptr, *miscS = nullptr;
    float  *xyzS = nullptr;                        // x | y | z coordinate streams of the snapshot (each xyz_stride floats): the density pass's candidates
    size_t  xyz_stride = 0;
    float4 *pack = nullptr;                      // per slot two float4: (pos.xyz, p) at 2s and (vel.xyz, rho) at 2s+1 -- one 32-byte sector
    float4 *scratch = nullptr;                   // all-pairs mode: pass results before they are committed to the SSBO
    float4 *boxes = nullptr;                     // all-pairs mode: bounding boxes of every 512 consecutive particles (inside the scratch allocation)
    float4 *pairP = nullptr;                     // neighbour sums of the force pass: (pres.xyz, visc.x)
    float2 *pairV = nullptr;                     //                                   (visc.y, visc.z)
    int   *nbr_list = nullptr, *nbr_count = nullptr;   // neighbour lists of the density pass ([slot][K]) and true counts
    // slab frames: ids of the particles the last exchange brought in, hashed and counted by the unpack kernel (count-ahead across the exchange)
    const int* arrivals = nullptr; const int* arrivals_count = nullptr; int arrivals_max = 0;
    bool   nbr_lists_valid = false;
    bool   nbr_rows_fmt = false;               // nbr_list holds row masks ([tile of 32][9 rows][lane] (first slot, accept mask)) instead of index lists
    int    nbr_k_alloc = 0, nbr_k_used = 0;                 // list capacity the array was sized for / the density pass built the lists with
    int   *heavy_queue = nullptr;                           // targets finished one warp each: [0,cap) density pass, [cap,2cap) force pass
    int   *heavy_cnt = nullptr;                             // the two queue counters (density, force); reset by the reorder kernel of every snapshot
    // count-ahead (frames in the middle of one cwa_coupled_step call): the integrate pass hashes the NEW position of the
    // particle in cell-ordered slot s, counts it into the grid's (pre-cleared) counter and leaves (cell, arrival rank) here,
    // so the next frame's grid build starts at the scan
    int   *cell_next = nullptr, *rank_next = nullptr;
    bool   counts_ahead = false;                            // the grid counter already holds the counts of the current positions
    cudaEvent_t wait_before_sampling = nullptr;             // transient: the first kernel that samples the wave field waits for it
    // slab decomposition: the integrate pass of cwa_sph_step_slab already packed the migrant / ghost messages of the NEXT exchange
    // (same selection as slab_pack_kernel, on the records it just wrote); cwa_slab_pack with the same arguments is then a no-op
    struct SlabPacked {
        bool valid = false;
        cwa_buf particles = -1, msg_left = -1, msg_right = -1;
        int n_owned = 0, cap_mig = 0, cap_ghost = 0;
        float z_lo = 0.f, z_hi = 0.f, band = 0.f;
    } slab_packed;
    // CUDA graphs of one all-pairs coupled frame (cwa_coupled_step, tuning "graph"): the shipped scene is 6 launches of a few
    // microseconds each, i.e. launch latency; one graph per state of the wave rotation / sampler binding (they cycle with period 3)
    struct FrameGraph {
        bool valid = false;
        long long key[16] = {};
        void* exec = nullptr;                               // cudaGraphExec_t
        unsigned nodes = 0;
    } frame_graph[6];
    unsigned long long consts_epoch = 0;                    // params_epoch the prepared constants were derived from
    void*  consts = nullptr;                     // Sph3Const prepared on the device once per dispatch
    bool snapshot_valid = false;
    bool pair_sums_valid = false;
};

struct Sph2Obj {
    bool live = false;
    int  n = 0, variant = CWA_SPH2_WAVE, substeps = 1;
    cwa_grid grid = -1;
    cwa_buf buffer[2] = {-1, -1};
    int  read_index = 0, write_index = 1;
    float time = 0.0f, bottom = 0.3f, psi = -1.0f, view_width = 2.0f * 4.8f;
    int  init_width = 128;
    cwa_buf wave1d = -1;
    int  wave1d_width = 0;
    float4 *posS = nullptr, *velS = nullptr, *accS = nullptr;   // cell-ordered snapshot of the read buffer
    // CUDA graph of one frame (22 launches of a few microseconds on L2-resident data at 65 k particles): captured when the same frame
    // -- same buffers, uniforms, grid -- is asked for the second time in a row, replayed while nothing changes
    long long graph_key[12] = {};
    long long last_key[12] = {};
    void* graph_exec = nullptr;                                 // cudaGraphExec_t
    unsigned graph_nodes = 0;
};

// ImageStencil driving Shallow1D_cs / Wave1D_cs (SphWave2D/StencilImage2D.h:10-66)
struct Stencil1dObj {
    bool live = false;
    int  shader = 0, w = 1, num_images = 2;
    float4* image[3] = {nullptr, nullptr, nullptr};      // RGBA32F texels
    cwa_buf image_buf[3] = {-1, -1, -1};
    int  read_index[2] = {0, 1}, write_index = 1, unit[3] = {0, 1, 2};
    int  substeps = 1, mode_iter_first = 2, mode_iter_last = 2;
    bool iterate = true;
    float lambda = 0.001f, dx_or_atten = 0.1f, beta = 0.001f, boundary[2] = {0.0f, 0.0f};
    int  bc = CWA_BC_FREE;
};

struct ShaderObj {
    bool live = false;
    std::string name;
    int kind = -1;
    int mode = 0;
    int object = -1;
    int   ui[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float uf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// kernels of the hot path, as reported by the per-kernel profile (bench.py roofline section)
enum KernelId {
    KID_CLEAR = 0, KID_HASH_COUNT, KID_SCAN, KID_INSERT, KID_CELL_ORDER, KID_REORDER,
    KID_DENSITY, KID_FORCE, KID_INTEGRATE, KID_WAVE, KID_OTHER, KID_HEAVY, KID_EXCHANGE, KID_COUNT
};

struct ProfRec { int id; cudaEvent_t a, b; };

// Kernel-variant / staging knobs (cwa_set_tuning).  Per CONTEXT: two contexts of one host (two scene streams, two devices) can
// hold different settings.  -1 = not set yet: the default comes from the environment (CWA_NB_CONFIG, ...) on first use.
struct CtxTuning {
    int config = -1, cap_d = -1, cap_f = -1, fused_order = -1, fused_integrate = -1, pipeline = -1, nbr_k = -1, extreme = -1;
    int scan_config = -1, wave_transpose = -1, graph = -1, inplace_max = -1, allpairs_bal = -1, slab_ahead = -1, heavy8 = -1, pdl = -1, ap_cull = -1;
};

struct SlabObj;

// scratch of the multi-GPU primitives in multi.cu (flags, scan output, scan state); owned by the context, freed in cwa_destroy
struct MultiScratch {
    int* flags = nullptr;
    int* pos = nullptr;           // n + 1 entries
    int* ticket = nullptr;
    unsigned long long* state = nullptr;
    size_t cap = 0, tiles = 0;
};

// a GL object registered with cudaGraphicsGLRegister* (interop.cu)
struct GlResource {
    bool live = false;
    bool is_image = false;
    void* res = nullptr;              // cudaGraphicsResource_t
    bool mapped = false;
    cwa_buf mapped_buf = -1;
};

struct cwa_ctx {
    MultiScratch multi;
    std::vector<GlResource> gl_resources;
    CtxTuning tune;
    std::vector<SlabObj*> slabs;
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    int device = 0;
    int sm_count = 148;
    int cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream[2] = {nullptr, nullptr};   // frame pipelining inside cwa_coupled_step: [0] wave stencil, [1] grid clears
    cudaEvent_t  ev_pipe[4] = {nullptr, nullptr, nullptr, nullptr};   // scan done, clear done, integrate done, wave done
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
    unsigned long long launches = 0;
    unsigned long long params_epoch = 1;   // bumped whenever a parameter block may have changed (UBO write / bind)
    std::vector<BufferObj> buffers;
    std::vector<GridObj>   grids;
    std::vector<WaveObj>   waves;
    std::vector<SphObj>    sphs;
    std::vector<Sph2Obj>   sph2s;
    std::vector<Stencil1dObj> stencil1ds;
    std::vector<ShaderObj> shaders;
    cwa_buf ssbo_binding[16];
    cwa_buf ubo_binding[8];
    cwa_buf default_ubo[8];
    cwa_sph  bound_sph = -1;
    cwa_wave bound_wave = -1;
    // scratch for cwa_scan_exclusive
    int* scan_ticket = nullptr;
    unsigned long long* scan_state = nullptr;
    size_t scan_state_tiles = 0;
    void* encode_tiled = nullptr;  // PFN cuTensorMapEncodeTiled
    std::vector<const void*> smem_attr_done;   // kernels whose dynamic shared-memory limit was raised on THIS context's device
};

// handle helpers (defined in api.cu)
BufferObj* get_buffer(cwa_ctx* ctx, cwa_buf b);
GridObj*   get_grid(cwa_ctx* ctx, cwa_grid g);
WaveObj*   get_wave(cwa_ctx* ctx, cwa_wave w);
SphObj*    get_sph(cwa_ctx* ctx, cwa_sph s);
Sph2Obj*   get_sph2(cwa_ctx* ctx, cwa_sph2 s);
int        new_buffer(cwa_ctx* ctx, void* ptr, size_t bytes, bool owned);
ParamPtrs  current_params(cwa_ctx* ctx);

// cross-module internals
int  scan_exclusive_launch(cwa_ctx* ctx, const int* in, int* out, int n, int* ticket,
                           unsigned long long* tile_state);          // grid.cu; out has n+1 entries
size_t scan_num_tiles(int n);
struct GridBuildOpts {
    bool canonical_order = true;       // false: stop after the arrival-order insert (the caller's fused kernel ranks and reorders in one pass)
    // count-ahead: the counter already holds this build's counts; cell id and arrival rank of the particle that sat in
    // cell-ordered slot s of the PREVIOUS build are in ahead_cell[s] / ahead_rank[s] (-1: not inserted), s < n
    const int* ahead_cell = nullptr;
    const int* ahead_rank = nullptr;
    bool clear_after_scan = false;     // clear counter + scan state right after this scan (side stream): the next build is counted ahead ...
    bool clear_is_for_next_build = false;   // ... or (slab frames) the next build starts from the cleared arrays instead of clearing them itself
    const int* n_dev = nullptr;        // device-resident particle count (slab decomposition): the host `n` is then only the launch bound
    // count-ahead across a slab exchange: particles that ARRIVED after the counting integrate pass (migrants, ghosts) were hashed and
    // counted by the unpack kernel (cell_of[id], rank[id]); their ids are arrivals[0 .. *arrivals_count)
    const int* arrivals = nullptr;
    const int* arrivals_count = nullptr;
    int arrivals_max = 0;
};
int  grid_build_internal(cwa_ctx* ctx, GridObj* g, const void* particles, int stride_bytes, int n, const GridBuildOpts& opts = GridBuildOpts());
TexView wave_tex_view(cwa_ctx* ctx, cwa_wave w, int image);           // wave.cu
int  wave_sampling_copy(cwa_ctx* ctx, cwa_wave w, int image, TexView* tex);   // points tex->tdata at the (refreshed) transposed copy when it pays
int  stencil1d_dispatch_mode(cwa_ctx* ctx, int handle, int mode, int shader);   // stencil1d.cu: one dispatch, no PingPong
void wave_touch_buffer(cwa_ctx* ctx, cwa_buf b, bool raw);             // a buffer was written through the Buffer API / its raw pointer handed out
int  wave_step_internal(cwa_ctx* ctx, WaveObj* w);
int  wave_image_with_unit(const WaveObj* w, int unit);                 // physical image bound to image unit 0 (newest) / 1 / 2 (next output)
void wave_pingpong_internal(WaveObj* w);                               // PingPong(): host bookkeeping only
int  wave_dispatch_mode(cwa_ctx* ctx, WaveObj* w, int mode);           // kernel for uMode on units 0/1/2, no rotation                   // one EVOLVE dispatch + PingPong
// slab pack fused into the integrate pass (multi.cu: cwa_sph_step_slab); all-zero = off
struct SlabPackArgs {
    int n_owned = 0;
    const int* n_owned_dev = nullptr;  // device-resident owned range (csrc/slab.cu); overrides n_owned
    int* free_list = nullptr;          // slots freed by migrants (csrc/slab.cu): the next unpack reuses them, so the owned range does not grow
    int* free_count = nullptr;
    float z_lo = 0.f, z_hi = 0.f, band = 0.f;
    float4* msg_l = nullptr;
    float4* msg_r = nullptr;
    int cap_mig = 0, cap_ghost = 0;
};
int  sph_passes_internal(cwa_ctx* ctx, SphObj* s, TexView tex, int which /*bit0 rho, bit1 force, bit2 integrate*/, bool count_ahead = false,
                         const SlabPackArgs* slab = nullptr);
void slab_destroy_all(cwa_ctx* ctx);                                  // slab.cu
void sph_invalidate_for_buffer(cwa_ctx* ctx, cwa_buf particles);       // the particle buffer was written behind the SPH object's back

// Every entry point runs against ITS context's device, whatever device the calling thread had current (a host may hold contexts
// on several devices: the multi-GPU slab group of csrc/slab.cu does); the previous device is restored on return.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const cwa_ctx* ctx)
    {
        if (!ctx) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != ctx->device) { cudaSetDevice(ctx->device); prev = cur; }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: remember it per context, not in a process-wide static
// (a host may hold contexts on several devices)
template <class K>
static inline int ensure_dynamic_smem(cwa_ctx* ctx, K kernel, int bytes)
{
    const void* key = reinterpret_cast<const void*>(kernel);
    for (const void* k : ctx->smem_attr_done) if (k == key) return 0;
    CWA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    ctx->smem_attr_done.push_back(key);
    return 0;
}

// Everything launched inside the scope (kernels, memsets, the KScope events around them) goes to `s` instead of the
// context's main stream; ordering against the main stream is the caller's business (events).
struct StreamScope {
    cwa_ctx* ctx; cudaStream_t saved;
    StreamScope(cwa_ctx* c, cudaStream_t s) : ctx(c), saved(c->stream) { c->stream = s; }
    ~StreamScope() { ctx->stream = saved; }
};

// RAII bracket around one launch (or memset): counts it and, while a profile is being taken,
// records a CUDA-event pair on the launching stream.
struct KScope {
    cwa_ctx* ctx; int slot;
    KScope(cwa_ctx* c, int id) : ctx(c), slot(-1)
    {
        ctx->launches++;
        if (!ctx->profiling) return;
        cudaEvent_t e[2];
        for (int k = 0; k < 2; k++) {
            if (!ctx->ev_pool.empty()) { e[k] = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
            else cudaEventCreate(&e[k]);
        }
        cudaEventRecord(e[0], ctx->stream);
        slot = (int)ctx->prof.size();
        ctx->prof.push_back(ProfRec{id, e[0], e[1]});
    }
    ~KScope() { if (slot >= 0) cudaEventRecord(ctx->prof[slot].b, ctx->stream); }
};

// ---------------------------------------------------------------------------------------------
// device helpers shared by the kernels
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Programmatic dependent launch (sm_90+): the kernels of a frame run back to back on one stream, each a few tens of microseconds, and
// every boundary costs the drain of the last wave plus the ramp-up of the next grid.  A kernel of the chain starts with CWA_PDL_ENTER():
// `launch_dependents` lets the NEXT kernel's CTAs be scheduled as soon as every CTA of this one has started (they fill the SMs the last
// wave leaves idle), `wait` holds them until the PREVIOUS grid has completed and its writes are visible -- so nothing before the wait may
// touch global memory another kernel writes.  Without the launch attribute (cwa_launch with pdl off, or <<<>>>) both are no-ops.
// The compiler may hoist read-only loads (__ldg, const __restrict__: "invariant for the kernel's lifetime") above an inline-asm barrier; a load
// that runs before the wait reads what the previous frame left there (seen: the queue length of sph3_force_heavy_kernel fetched before the force
// pass had filled the queue).  ptxas does the same with ld.global.nc, whatever the PTX order.  So the wait is followed by a branch on a run-time value
// (CWA_PDL_ENTER: `if (%smid == ~0) return;`, never taken) and the whole kernel body is control-dependent on it -- global loads are not speculated above a branch.  tools/check_pdl_sass.py (run by tests/test_pdl_sass.py)
// verifies in the SASS that no kernel touches memory before its ACQBULK.
__device__ __forceinline__ bool cwa_pdl_enter()
{
    unsigned smid;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;\n\tmov.u32 %0, %%smid;" : "=r"(smid) :: "memory");
    return smid == 0xffffffffu;                       // never (%smid < %nsmid), but neither the optimizer nor ptxas can fold it
}
#define CWA_PDL_ENTER() do { if (cwa_pdl_enter()) return; } while (0)

// the kernels of the chain, as bits of the tuning value `pdl` (which launches carry the attribute)
enum { PDL_SCAN = 1, PDL_INSERT = 2, PDL_REORDER = 4, PDL_DENSITY = 8, PDL_DENSITY_HEAVY = 16, PDL_FORCE = 32, PDL_FORCE_HEAVY = 64, PDL_INTEGRATE = 128,
       PDL_GRID2 = 256 /* the 2-D frame (SphUgrid): hash, insert, cell order, reorder, density, forces */ };
#define CWA_PDL_DEFAULT (511 & ~(PDL_INSERT | PDL_REORDER))   // measured (profiles/r2/tuning.md): those two cost more than they gain
static inline bool cwa_pdl_enabled(cwa_ctx* c, int bit)
{
    if (c->tune.pdl < 0) { const char* e = getenv("CWA_PDL"); c->tune.pdl = (e && *e) ? (atoi(e) & 511) : CWA_PDL_DEFAULT; }
    return bit != 0 && (c->tune.pdl & bit) != 0 && !c->profiling;      // (a profile brackets every launch with events: nothing to overlap)
}

// <<<grid, block, smem, ctx->stream>>> with the programmatic-serialization attribute; the kernel must begin with CWA_PDL_ENTER()
template <typename... KArgs, typename... Args>
static inline cudaError_t cwa_launch(cwa_ctx* ctx, int pdl_bit, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = cwa_pdl_enabled(ctx, pdl_bit) ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256).  The 32-byte records of the hot path --
// the (pos, p | vel, rho) pack of a cell-ordered slot, half a particle record, a pair of candidate positions
// -- are exactly one sector, so one request moves what two 128-bit requests moved before.
// `p` must be 32-byte aligned.
struct f4x2 { float4 a, b; };
__device__ __forceinline__ f4x2 cwa_ldg256(const float4* p)
{
    f4x2 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
        : "l"(p));
    return r;
}
__device__ __forceinline__ void cwa_stg256(float4* p, const float4 a, const float4 b)
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// float -> cell coordinate, clamp in the float domain first (NaN -> 0).  Restates
// ivec(floor(q)); clamp(cell, 0, n-1) of uniform_grid_sph_cs.glsl:144-145 with a defined
// result for NaN/huge inputs.  `q` must come from a true IEEE division (__fdiv_rn).
__device__ __forceinline__ int cwa_cell_coord(float q, int n)
{
    float f = floorf(q);
    if (!(f >= 0.0f)) return 0;
    if (f > (float)(n - 1)) return n - 1;
    return (int)f;
}

__device__ __forceinline__ void cwa_cell3(const GridView& g, float x, float y, float z, int& i, int& j, int& k)
{
    i = cwa_cell_coord(__fdiv_rn(__fsub_rn(x, g.min[0]), g.cell[0]), g.n[0]);
    j = cwa_cell_coord(__fdiv_rn(__fsub_rn(y, g.min[1]), g.cell[1]), g.n[1]);
    k = cwa_cell_coord(__fdiv_rn(__fsub_rn(z, g.min[2]), g.cell[2]), g.n[2]);
}

__device__ __forceinline__ void cwa_cell2(const GridView& g, float x, float y, int& i, int& j)
{
    i = cwa_cell_coord(__fdiv_rn(__fsub_rn(x, g.min[0]), g.cell[0]), g.n[0]);
    j = cwa_cell_coord(__fdiv_rn(__fsub_rn(y, g.min[1]), g.cell[1]), g.n[1]);
}

// canonical length (shared with the oracle): explicit fma chain, IEEE sqrt
__device__ __forceinline__ float cwa_len2sq(float x, float y) { return __fmaf_rn(y, y, __fmul_rn(x, x)); }
__device__ __forceinline__ float cwa_len3sq(float x, float y, float z)
{
    return __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
}

__device__ __forceinline__ int cwa_tex_index(float f, int n)
{
    if (!(f >= 0.0f)) return 0;
    if (f > (float)(n - 1)) return n - 1;
    return (int)f;
}

// texel (i, j) of a (possibly row-block) view, j a GLOBAL row: the local block (row-major, or its transposed sampling copy),
// else the replicated global last row, else the nearest local row (only reachable when the caller under-sized the sampling halo)
__device__ __forceinline__ float cwa_texel(const TexView& t, int i, int j)
{
    int l = j - t.row0;
    if ((unsigned)l >= (unsigned)t.h) {
        if (j == t.h_global - 1 && t.last_row != nullptr) return __ldg(t.last_row + (size_t)i * t.ch);
        l = l < 0 ? 0 : t.h - 1;
    }
    return __ldg(t.tdata + (size_t)i * t.tsi + (size_t)l * t.tsj);
}

// texture(wave_tex, (s,t)).r : GL_LINEAR, GL_CLAMP_TO_EDGE, LOD 0, full FP32 weights
__device__ __forceinline__ float cwa_tex_bilinear(const TexView& t, float s, float tt)
{
    if (t.data == nullptr) return 0.0f;
    const int W = t.w, H = t.h_global;
    float u = __fsub_rn(__fmul_rn(s, (float)W), 0.5f);
    float v = __fsub_rn(__fmul_rn(tt, (float)H), 0.5f);
    float fu = floorf(u), fv = floorf(v);
    float a = __fsub_rn(u, fu), b = __fsub_rn(v, fv);
    int i0 = cwa_tex_index(fu, W), i1 = cwa_tex_index(fu + 1.0f, W);
    int j0 = cwa_tex_index(fv, H), j1 = cwa_tex_index(fv + 1.0f, H);
    float t00 = cwa_texel(t, i0, j0);
    float t10 = cwa_texel(t, i1, j0);
    float t01 = cwa_texel(t, i0, j1);
    float t11 = cwa_texel(t, i1, j1);
    float r0 = __fadd_rn(t00, __fmul_rn(a, __fsub_rn(t10, t00)));
    float r1 = __fadd_rn(t01, __fmul_rn(a, __fsub_rn(t11, t01)));
    return __fadd_rn(r0, __fmul_rn(b, __fsub_rn(r1, r0)));
}

// Same sampler for a field that is entirely local, single-channel and 32-bit indexable (row0 == 0,
// h == h_global, ch == 1, data != nullptr): identical arithmetic and clamping (NaN -> texel 0), without the
// row-block lookups, channel stride and 64-bit index arithmetic of the general view.
__device__ __forceinline__ float cwa_tex_bilinear_local(const TexView& t, float s, float tt)
{
    const float fW = (float)t.w, fH = (float)t.h;
    const float u = __fsub_rn(__fmul_rn(s, fW), 0.5f);
    const float v = __fsub_rn(__fmul_rn(tt, fH), 0.5f);
    const float fu = floorf(u), fv = floorf(v);
    const float a = __fsub_rn(u, fu), b = __fsub_rn(v, fv);
    const float wm = fW - 1.0f, hm = fH - 1.0f;
    const int i0 = (int)fminf(fmaxf(fu, 0.0f), wm), i1 = (int)fminf(fmaxf(fu + 1.0f, 0.0f), wm);
    const int j0 = (int)fminf(fmaxf(fv, 0.0f), hm), j1 = (int)fminf(fmaxf(fv + 1.0f, 0.0f), hm);
    const float* p0 = t.tdata + j0 * t.tsj;
    const float* p1 = t.tdata + j1 * t.tsj;
    const int o0 = i0 * t.tsi, o1 = i1 * t.tsi;
    const float t00 = __ldg(p0 + o0), t10 = __ldg(p0 + o1), t01 = __ldg(p1 + o0), t11 = __ldg(p1 + o1);
    const float r0 = __fadd_rn(t00, __fmul_rn(a, __fsub_rn(t10, t00)));
    const float r1 = __fadd_rn(t01, __fmul_rn(a, __fsub_rn(t11, t01)));
    return __fadd_rn(r0, __fmul_rn(b, __fsub_rn(r1, r0)));
}

template <bool LOCAL>
__device__ __forceinline__ float cwa_tex_sample(const TexView& t, float s, float tt)
{
    if (LOCAL) return cwa_tex_bilinear_local(t, s, tt);
    return cwa_tex_bilinear(t, s, tt);
}

static inline bool tex_view_is_local(const TexView& t)
{
    return t.data != nullptr && t.ch == 1 && t.row0 == 0 && t.h == t.h_global && (long long)t.w * t.h < (1ll << 30);
}

__device__ __forceinline__ float cwa_smoothstep(float e0, float e1, float x)
{
    float t = __fdiv_rn(__fsub_rn(x, e0), __fsub_rn(e1, e0));
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    return __fmul_rn(__fmul_rn(t, t), __fsub_rn(3.0f, __fmul_rn(2.0f, t)));
}

#endif // __CUDACC__
