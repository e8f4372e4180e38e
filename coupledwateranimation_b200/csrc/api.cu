// api.cu -- context, Buffer objects, bindings, timers and the ComputeShader-style dispatch layer
// of libcwa_b200.  Host logic only; kernels live in wave.cu / grid.cu / sph3.cu / sph2.cu.
#include "internal.cuh"

#include <cudaTypedefs.h>

#include <mutex>

// ---------------------------------------------------------------------------------------------
// error text
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void cwa_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* cwa_last_error(void) { return g_err; }
extern "C" int cwa_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------
// handle tables
// ---------------------------------------------------------------------------------------------
BufferObj* get_buffer(cwa_ctx* ctx, cwa_buf b)
{
    if (!ctx || b < 0 || b >= (int)ctx->buffers.size() || !ctx->buffers[b].live) return nullptr;
    return &ctx->buffers[b];
}
GridObj* get_grid(cwa_ctx* ctx, cwa_grid g)
{
    if (!ctx || g < 0 || g >= (int)ctx->grids.size() || !ctx->grids[g].live) return nullptr;
    return &ctx->grids[g];
}
WaveObj* get_wave(cwa_ctx* ctx, cwa_wave w)
{
    if (!ctx || w < 0 || w >= (int)ctx->waves.size() || !ctx->waves[w].live) return nullptr;
    return &ctx->waves[w];
}
SphObj* get_sph(cwa_ctx* ctx, cwa_sph s)
{
    if (!ctx || s < 0 || s >= (int)ctx->sphs.size() || !ctx->sphs[s].live) return nullptr;
    return &ctx->sphs[s];
}
Sph2Obj* get_sph2(cwa_ctx* ctx, cwa_sph2 s)
{
    if (!ctx || s < 0 || s >= (int)ctx->sph2s.size() || !ctx->sph2s[s].live) return nullptr;
    return &ctx->sph2s[s];
}

int new_buffer(cwa_ctx* ctx, void* ptr, size_t bytes, bool owned)
{
    BufferObj b;
    b.ptr = ptr; b.bytes = bytes; b.owned = owned; b.live = true;
    for (size_t i = 0; i < ctx->buffers.size(); i++)
        if (!ctx->buffers[i].live) { ctx->buffers[i] = b; return (int)i; }
    ctx->buffers.push_back(b);
    return (int)ctx->buffers.size() - 1;
}

ParamPtrs current_params(cwa_ctx* ctx)
{
    auto ptr_of = [&](int binding) -> const void* {
        BufferObj* b = get_buffer(ctx, ctx->ubo_binding[binding]);
        if (!b) b = get_buffer(ctx, ctx->default_ubo[binding]);
        return b ? b->ptr : nullptr;
    };
    ParamPtrs p;
    p.constants = (const cwa_constants_uniform*)ptr_of(CWA_UBO_CONSTANTS);
    p.boundary  = (const cwa_boundary_uniform*)ptr_of(CWA_UBO_BOUNDARY);
    p.wave      = (const cwa_wave_uniforms*)ptr_of(CWA_UBO_WAVE);
    p.sim       = (const cwa_sim_constants*)ptr_of(CWA_UBO_SIM);
    return p;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" int cwa_create(int device, cwa_ctx** out)
{
    CWA_CHECK(out != nullptr, "cwa_create: out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cwa_set_error("cwa_create: no CUDA device (%s); libcwa_b200 has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return -3;
    }
    CWA_CHECK(device >= 0 && device < count, "cwa_create: device %d out of range (%d devices)", device, count);
    CWA_CUDA(cudaSetDevice(device));
    cwa_ctx* ctx = new cwa_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    CWA_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major; ctx->cc_minor = prop.minor;
    ctx->total_mem = prop.totalGlobalMem;
    if (prop.major < 10) {
        cwa_set_error("cwa_create: device %d is sm_%d%d; libcwa_b200 is built for sm_100a only", device, prop.major, prop.minor);
        delete ctx;
        return -3;
    }
    CWA_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CWA_CUDA(cudaEventCreate(&ctx->ev0));
    CWA_CUDA(cudaEventCreate(&ctx->ev1));
    for (int i = 0; i < 2; i++) CWA_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream[i], cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) CWA_CUDA(cudaEventCreateWithFlags(&ctx->ev_pipe[i], cudaEventDisableTiming));
    for (int i = 0; i < 16; i++) ctx->ssbo_binding[i] = -1;
    for (int i = 0; i < 8; i++) { ctx->ubo_binding[i] = -1; ctx->default_ubo[i] = -1; }

    // cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (ge == cudaSuccess && qres == cudaDriverEntryPointSuccess) ctx->encode_tiled = fn;
        else { (void)cudaGetLastError(); ctx->encode_tiled = nullptr; }
    }

    // default parameter blocks == the reference's defaults (Main.cpp:184-204 + shader consts)
    cwa_constants_uniform cu = {0.02f, 2.0f, 3000.0f, 1000.0f};
    cwa_boundary_uniform bu = {{0.48f, 1.0f, 0.48f, 500.0f}, {0.0f, -0.02f, 0.0f, 50.0f}};
    cwa_wave_uniforms wu = {{0.01f, 0.985f, 0.001f, 1.0f}, {2.0f, 0.35f, -1.0f, 0.0f}};
    cwa_sim_constants sc = {0.005f, 4000.0f, 0.00005f, -9806.65f, 0.3f, 0.01f, 25.0f, 2.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    int rc = 0;
    rc |= cwa_buffer_create(ctx, sizeof(cu), &cu, &ctx->default_ubo[CWA_UBO_CONSTANTS]);
    rc |= cwa_buffer_create(ctx, sizeof(bu), &bu, &ctx->default_ubo[CWA_UBO_BOUNDARY]);
    rc |= cwa_buffer_create(ctx, sizeof(wu), &wu, &ctx->default_ubo[CWA_UBO_WAVE]);
    rc |= cwa_buffer_create(ctx, sizeof(sc), &sc, &ctx->default_ubo[CWA_UBO_SIM]);
    if (rc != 0) { delete ctx; return -2; }
    for (int i = 1; i <= 4; i++) ctx->ubo_binding[i] = ctx->default_ubo[i];

    // scratch for the stand-alone scan entry point
    ctx->scan_state_tiles = 1 << 16;
    CWA_CUDA(cudaMalloc(&ctx->scan_ticket, sizeof(int) * 2 + sizeof(unsigned long long) * ctx->scan_state_tiles));
    ctx->scan_state = (unsigned long long*)(ctx->scan_ticket + 2);
    *out = ctx;
    return 0;
}

extern "C" void cwa_destroy(cwa_ctx* ctx)
{
    DeviceGuard _dg(ctx);
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2; i++) cudaStreamSynchronize(ctx->side_stream[i]);
    slab_destroy_all(ctx);
    for (size_t i = 0; i < ctx->sphs.size(); i++) if (ctx->sphs[i].live) cwa_sph_destroy(ctx, (int)i);
    for (size_t i = 0; i < ctx->sph2s.size(); i++) if (ctx->sph2s[i].live) cwa_sph2_destroy(ctx, (int)i);
    for (size_t i = 0; i < ctx->waves.size(); i++) if (ctx->waves[i].live) cwa_wave_destroy(ctx, (int)i);
    for (size_t i = 0; i < ctx->stencil1ds.size(); i++) if (ctx->stencil1ds[i].live) cwa_stencil1d_destroy(ctx, (int)i);
    for (size_t i = 0; i < ctx->grids.size(); i++) if (ctx->grids[i].live) cwa_grid_destroy(ctx, (int)i);
    for (auto& b : ctx->buffers) if (b.live && b.owned && b.ptr) cudaFree(b.ptr);
    if (ctx->scan_ticket) cudaFree(ctx->scan_ticket);
    cudaFree(ctx->multi.flags); cudaFree(ctx->multi.pos); cudaFree(ctx->multi.ticket);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->ev_pipe[i]);
    for (int i = 0; i < 2; i++) { cudaStreamSynchronize(ctx->side_stream[i]); cudaStreamDestroy(ctx->side_stream[i]); }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int cwa_synchronize(cwa_ctx* ctx)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx, "null context");
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" void* cwa_stream(cwa_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int cwa_device_info(cwa_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx, "null context");
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (total_mem) *total_mem = ctx->total_mem;
    return 0;
}

extern "C" int cwa_timer_begin(cwa_ctx* ctx)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx, "null context");
    CWA_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return 0;
}

extern "C" int cwa_timer_end(cwa_ctx* ctx, float* ms)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && ms, "null argument");
    CWA_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    CWA_CUDA(cudaEventSynchronize(ctx->ev1));
    CWA_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return 0;
}

extern "C" unsigned long long cwa_launch_count(cwa_ctx* ctx) { return ctx ? ctx->launches : 0ull; }

// per-kernel device timing (CUDA-event pairs around every launch on the context stream)
static const char* const g_kernel_names[KID_COUNT] = {
    "clear(memset)", "grid_hash_count", "scan_lookback", "grid_insert", "grid_cell_order", "reorder",
    "density", "force", "integrate", "wave_evolve", "other", "heavy_targets", "exchange"};

extern "C" int cwa_profile_kernel_count(void) { return KID_COUNT; }
extern "C" const char* cwa_profile_kernel_name(int id) { return (id >= 0 && id < KID_COUNT) ? g_kernel_names[id] : ""; }

extern "C" int cwa_profile_begin(cwa_ctx* ctx)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx, "null context");
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto& r : ctx->prof) { ctx->ev_pool.push_back(r.a); ctx->ev_pool.push_back(r.b); }
    ctx->prof.clear();
    ctx->profiling = true;
    return 0;
}

extern "C" int cwa_profile_end(cwa_ctx* ctx, float* ms, int* launches, int cap)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && ms && launches && cap >= KID_COUNT, "cwa_profile_end: need arrays of %d entries", KID_COUNT);
    ctx->profiling = false;
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < cap; i++) { ms[i] = 0.0f; launches[i] = 0; }
    for (auto& r : ctx->prof) {
        float t = 0.0f;
        CWA_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.id] += t; launches[r.id]++;
        ctx->ev_pool.push_back(r.a); ctx->ev_pool.push_back(r.b);
    }
    ctx->prof.clear();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Buffer  (SphWave2D/Buffer.cpp:5-83)
// ---------------------------------------------------------------------------------------------
extern "C" int cwa_buffer_create(cwa_ctx* ctx, size_t bytes, const void* host, cwa_buf* out)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && out, "null argument");
    CWA_CHECK(bytes > 0, "cwa_buffer_create: zero size");
    void* p = nullptr;
    CWA_CUDA(cudaMalloc(&p, bytes));
    if (host) CWA_CUDA(cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    else CWA_CUDA(cudaMemsetAsync(p, 0, bytes, ctx->stream));
    *out = new_buffer(ctx, p, bytes, true);
    return 0;
}

extern "C" int cwa_buffer_wrap(cwa_ctx* ctx, void* device_ptr, size_t bytes, cwa_buf* out)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && out && device_ptr, "null argument");
    cudaPointerAttributes attr;
    CWA_CUDA(cudaPointerGetAttributes(&attr, device_ptr));
    CWA_CHECK(attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged,
              "cwa_buffer_wrap: pointer is not device memory");
    *out = new_buffer(ctx, device_ptr, bytes, false);
    return 0;
}

extern "C" int cwa_buffer_destroy(cwa_ctx* ctx, cwa_buf b)
{
    DeviceGuard _dg(ctx);
    BufferObj* o = get_buffer(ctx, b);
    CWA_CHECK(o, "invalid buffer handle %d", b);
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    if (o->owned && o->ptr) CWA_CUDA(cudaFree(o->ptr));
    o->live = false; o->ptr = nullptr;
    for (int i = 0; i < 16; i++) if (ctx->ssbo_binding[i] == b) ctx->ssbo_binding[i] = -1;
    for (int i = 0; i < 8; i++) if (ctx->ubo_binding[i] == b) ctx->ubo_binding[i] = ctx->default_ubo[i];
    return 0;
}

extern "C" int cwa_buffer_sub_data(cwa_ctx* ctx, cwa_buf b, size_t off, size_t bytes, const void* host)
{
    DeviceGuard _dg(ctx);
    BufferObj* o = get_buffer(ctx, b);
    CWA_CHECK(o, "invalid buffer handle %d", b);
    CWA_CHECK(host && off + bytes <= o->bytes, "cwa_buffer_sub_data: range [%zu,%zu) outside buffer of %zu bytes", off, off + bytes, o->bytes);
    CWA_CUDA(cudaMemcpyAsync((char*)o->ptr + off, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    // a particle upload invalidates any cell-ordered snapshot built from this buffer
    sph_invalidate_for_buffer(ctx, b);
    for (int i = 0; i < 8; i++) if (ctx->ubo_binding[i] == b) ctx->params_epoch++;      // a parameter block changed
    wave_touch_buffer(ctx, b, false);                                                    // a wave image written through its buffer handle
    return 0;
}

extern "C" int cwa_buffer_read(cwa_ctx* ctx, cwa_buf b, size_t off, size_t bytes, void* host)
{
    DeviceGuard _dg(ctx);
    BufferObj* o = get_buffer(ctx, b);
    CWA_CHECK(o, "invalid buffer handle %d", b);
    CWA_CHECK(host && off + bytes <= o->bytes, "cwa_buffer_read: range outside buffer");
    CWA_CUDA(cudaMemcpyAsync(host, (const char*)o->ptr + off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Read-back that does not block: enqueued behind the work already on the context stream; the host memory (pinned, or the copy
// degrades to a synchronous one) holds the data after the next cwa_synchronize().  The GL analogue is glGetNamedBufferSubData
// into a persistently mapped / pixel-pack buffer followed by a fence.
extern "C" int cwa_buffer_read_async(cwa_ctx* ctx, cwa_buf b, size_t off, size_t bytes, void* host)
{
    DeviceGuard _dg(ctx);
    BufferObj* o = get_buffer(ctx, b);
    CWA_CHECK(o, "invalid buffer handle %d", b);
    CWA_CHECK(host && off + bytes <= o->bytes, "cwa_buffer_read_async: range outside buffer");
    CWA_CUDA(cudaMemcpyAsync(host, (const char*)o->ptr + off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

extern "C" int cwa_buffer_copy(cwa_ctx* ctx, cwa_buf src, cwa_buf dst, size_t soff, size_t doff, size_t bytes)
{
    DeviceGuard _dg(ctx);
    BufferObj* s = get_buffer(ctx, src);
    BufferObj* d = get_buffer(ctx, dst);
    CWA_CHECK(s && d, "invalid buffer handle");
    CWA_CHECK(soff + bytes <= s->bytes && doff + bytes <= d->bytes, "cwa_buffer_copy: range outside buffer");
    CWA_CUDA(cudaMemcpyAsync((char*)d->ptr + doff, (const char*)s->ptr + soff, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    for (int i = 0; i < 8; i++) if (ctx->ubo_binding[i] == dst) ctx->params_epoch++;
    sph_invalidate_for_buffer(ctx, dst);
    wave_touch_buffer(ctx, dst, false);
    return 0;
}

extern "C" int cwa_buffer_bind_base(cwa_ctx* ctx, int target, int binding, cwa_buf b)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx, "null context");
    CWA_CHECK(b == -1 || get_buffer(ctx, b), "invalid buffer handle %d", b);
    if (target == CWA_TARGET_SSBO) {
        CWA_CHECK(binding >= 0 && binding < 16, "SSBO binding %d out of range", binding);
        ctx->ssbo_binding[binding] = b;
    } else if (target == CWA_TARGET_UBO) {
        CWA_CHECK(binding >= 0 && binding < 8, "UBO binding %d out of range", binding);
        ctx->ubo_binding[binding] = (b == -1) ? ctx->default_ubo[binding] : b;
        ctx->params_epoch++;
    } else {
        CWA_CHECK(false, "unknown buffer target %d", target);
    }
    return 0;
}

extern "C" int cwa_buffer_device_ptr(cwa_ctx* ctx, cwa_buf b, void** ptr, size_t* bytes)
{
    DeviceGuard _dg(ctx);
    BufferObj* o = get_buffer(ctx, b);
    CWA_CHECK(o, "invalid buffer handle %d", b);
    if (ptr) *ptr = o->ptr;
    if (bytes) *bytes = o->bytes;
    wave_touch_buffer(ctx, b, true);              // a raw pointer to a wave image leaves the library: its contents can change unseen
    sph_invalidate_for_buffer(ctx, b);
    return 0;
}

extern "C" int cwa_default_ubo(cwa_ctx* ctx, int binding, cwa_buf* out)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && out, "null argument");
    CWA_CHECK(binding >= 1 && binding <= 4, "no default block for UBO binding %d", binding);
    *out = ctx->default_ubo[binding];
    return 0;
}

// ---------------------------------------------------------------------------------------------
// ParallelScan::Compute
// ---------------------------------------------------------------------------------------------
extern "C" int cwa_scan_exclusive(cwa_ctx* ctx, cwa_buf in, cwa_buf out, int n)
{
    DeviceGuard _dg(ctx);
    BufferObj* i = get_buffer(ctx, in);
    BufferObj* o = get_buffer(ctx, out);
    CWA_CHECK(i && o, "invalid buffer handle");
    CWA_CHECK(n >= 1, "cwa_scan_exclusive: n must be >= 1");
    CWA_CHECK((size_t)n * 4 <= i->bytes && (size_t)n * 4 <= o->bytes, "cwa_scan_exclusive: buffers smaller than n ints");
    CWA_CHECK(scan_num_tiles(n) <= ctx->scan_state_tiles, "cwa_scan_exclusive: n too large");
    CWA_CUDA(cudaMemsetAsync(ctx->scan_ticket, 0, sizeof(int) * 2 + sizeof(unsigned long long) * scan_num_tiles(n), ctx->stream));
    // out may hold only n entries: the total (entry n) is dropped when it does not fit
    bool has_total = (size_t)(n + 1) * 4 <= o->bytes && o->ptr != i->ptr;
    (void)has_total;
    return scan_exclusive_launch(ctx, (const int*)i->ptr, (int*)o->ptr, has_total ? n : -n, ctx->scan_ticket, ctx->scan_state);
}

// ---------------------------------------------------------------------------------------------
// scene binding and the north-star entry points
// ---------------------------------------------------------------------------------------------
extern "C" int cwa_bind_scene(cwa_ctx* ctx, cwa_sph s, cwa_wave w)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx, "null context");
    CWA_CHECK(s == -1 || get_sph(ctx, s), "invalid sph handle %d", s);
    CWA_CHECK(w == -1 || get_wave(ctx, w), "invalid wave handle %d", w);
    ctx->bound_sph = s;
    ctx->bound_wave = w;
    return 0;
}

extern "C" int sph_step(cwa_ctx* ctx, int nsteps)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx, "null context");
    CWA_CHECK(get_sph(ctx, ctx->bound_sph), "sph_step: no SPH object bound (cwa_bind_scene)");
    return cwa_sph_step(ctx, ctx->bound_sph, nsteps);
}

extern "C" int wave_step(cwa_ctx* ctx, int nsteps)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx, "null context");
    CWA_CHECK(get_wave(ctx, ctx->bound_wave), "wave_step: no wave object bound (cwa_bind_scene)");
    return cwa_wave_compute(ctx, ctx->bound_wave, nsteps);
}

// ---------------------------------------------------------------------------------------------
// ComputeShader  (CoupledWaterAnimation/ComputeShader.cpp:9-56)
// ---------------------------------------------------------------------------------------------
enum ShaderKind { SK_RHO = 0, SK_FORCE, SK_INTEGRATE, SK_WAVE, SK_WAVE_SIMP, SK_PREFIX, SK_SHALLOW1D, SK_WAVE1D };

extern "C" int cwa_shader_create(cwa_ctx* ctx, const char* glsl_filename, cwa_shader* out)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && out && glsl_filename, "null argument");
    *out = -1;
    std::string n(glsl_filename);
    size_t slash = n.find_last_of("/\\");
    if (slash != std::string::npos) n = n.substr(slash + 1);
    int kind = -1;
    if (n == "rho_pres_comp.glsl") kind = SK_RHO;
    else if (n == "force_comp.glsl") kind = SK_FORCE;
    else if (n == "integrate_comp.glsl") kind = SK_INTEGRATE;
    else if (n == "wave_comp.glsl") kind = SK_WAVE;
    else if (n == "Wave2D_cs.glsl") kind = SK_WAVE_SIMP;
    else if (n == "prefix_sum_cs.glsl") kind = SK_PREFIX;
    else if (n == "Shallow1D_cs.glsl") kind = SK_SHALLOW1D;
    else if (n == "Wave1D_cs.glsl") kind = SK_WAVE1D;
    // InitShader() returns -1 for a shader it cannot build (InitShader.cpp:105)
    CWA_CHECK(kind >= 0, "cwa_shader_create: no CUDA kernel set replaces shader '%s'", glsl_filename);
    ShaderObj s;
    s.live = true; s.name = n; s.kind = kind;
    if (kind == SK_PREFIX) { s.ui[0] = 0; s.ui[1] = 2; s.ui[2] = 2; }   // phase, stride, n defaults (prefix_sum_cs.glsl:9-11)
    ctx->shaders.push_back(s);
    *out = (int)ctx->shaders.size() - 1;
    return 0;
}

static ShaderObj* get_shader(cwa_ctx* ctx, cwa_shader s)
{
    if (!ctx || s < 0 || s >= (int)ctx->shaders.size() || !ctx->shaders[s].live) return nullptr;
    return &ctx->shaders[s];
}

extern "C" int cwa_shader_set_mode(cwa_ctx* ctx, cwa_shader s, int mode)
{
    DeviceGuard _dg(ctx);
    ShaderObj* o = get_shader(ctx, s);
    CWA_CHECK(o, "invalid shader handle %d", s);
    o->mode = mode;
    return 0;
}

extern "C" int cwa_shader_set_uniform_i(cwa_ctx* ctx, cwa_shader s, int location, int v)
{
    DeviceGuard _dg(ctx);
    ShaderObj* o = get_shader(ctx, s);
    CWA_CHECK(o, "invalid shader handle %d", s);
    CWA_CHECK(location >= 0 && location < 8, "uniform location %d out of range", location);
    o->ui[location] = v;
    return 0;
}

extern "C" int cwa_shader_set_uniform_f(cwa_ctx* ctx, cwa_shader s, int location, float v)
{
    DeviceGuard _dg(ctx);
    ShaderObj* o = get_shader(ctx, s);
    CWA_CHECK(o, "invalid shader handle %d", s);
    CWA_CHECK(location >= 0 && location < 8, "uniform location %d out of range", location);
    o->uf[location] = v;
    return 0;
}

extern "C" int cwa_shader_bind_object(cwa_ctx* ctx, cwa_shader s, int object_handle)
{
    DeviceGuard _dg(ctx);
    ShaderObj* o = get_shader(ctx, s);
    CWA_CHECK(o, "invalid shader handle %d", s);
    o->object = object_handle;
    return 0;
}

int prefix_sum_level_launch(cwa_ctx* ctx, int* x, int n, int phase, int stride, int nthreads);   // grid.cu

extern "C" int cwa_shader_dispatch(cwa_ctx* ctx, cwa_shader s, int gx, int gy, int gz)
{
    DeviceGuard _dg(ctx);
    ShaderObj* o = get_shader(ctx, s);
    CWA_CHECK(o, "invalid shader handle %d", s);
    (void)gy; (void)gz;
    switch (o->kind) {
    case SK_RHO: case SK_FORCE: case SK_INTEGRATE: {
        int h = o->object >= 0 ? o->object : ctx->bound_sph;
        SphObj* sp = get_sph(ctx, h);
        CWA_CHECK(sp, "dispatch of %s: no SPH object bound to the shader", o->name.c_str());
        if (o->kind == SK_RHO) return cwa_sph_rho_pres(ctx, h);
        if (o->kind == SK_FORCE) return cwa_sph_force(ctx, h);
        return cwa_sph_integrate(ctx, h);
    }
    case SK_WAVE: case SK_WAVE_SIMP: {
        int h = o->object >= 0 ? o->object : ctx->bound_wave;
        WaveObj* w = get_wave(ctx, h);
        CWA_CHECK(w, "dispatch of %s: no wave object bound to the shader", o->name.c_str());
        // Dispatch alone runs the kernel for the current uMode on the images at units 0/1/2 and
        // does NOT rotate (StencilImage2DTripleBuffered::Compute rotates after Dispatch).
        return wave_dispatch_mode(ctx, w, o->mode);
    }
    case SK_SHALLOW1D: case SK_WAVE1D: {
        // one dispatch in the current uMode on the ImageStencil object bound to the shader (cwa_shader_bind_object); like the wave
        // shaders it does not rotate -- ImageStencil::Compute / ComputeFunc call PingPong after Dispatch (cwa_stencil1d_pingpong)
        return stencil1d_dispatch_mode(ctx, o->object, o->mode, o->kind == SK_SHALLOW1D ? CWA_STENCIL1D_SHALLOW : CWA_STENCIL1D_WAVE);
    }
    case SK_PREFIX: {
        // one Blelloch level of prefix_sum_cs.glsl:18-43 on the buffer at SSBO binding 1
        BufferObj* b = get_buffer(ctx, ctx->ssbo_binding[1]);
        CWA_CHECK(b, "prefix_sum_cs dispatch: nothing bound at SSBO binding 1");
        CWA_CHECK((size_t)o->ui[2] * 4 <= b->bytes, "prefix_sum_cs dispatch: n exceeds the bound buffer");
        return prefix_sum_level_launch(ctx, (int*)b->ptr, o->ui[2], o->ui[0], o->ui[1], gx * 1024);
    }
    }
    CWA_CHECK(false, "dispatch: unknown shader kind");
    return -1;
}
