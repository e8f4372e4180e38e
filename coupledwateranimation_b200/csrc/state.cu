// state.cu -- two host-side services around the simulation state:
//   * a name -> field table over the parameter blocks (what SphWave2D/UniformGui.cpp does with glGetActiveUniform: a GUI / CLI can
//     list and edit h, mass, rho0, k, mu, dt, gravity, the box ... at run time without knowing the block layouts);
//   * a checkpoint: one little-endian file {header, the four parameter blocks, particle SSBO, the three wave levels, the
//     triple-buffer bookkeeping, frame counter} that restores a run bit for bit (the reference has no save / resume at all).
#include "internal.cuh"

#include <cstdio>

// ---------------------------------------------------------------------------------------------
// parameter reflection
// ---------------------------------------------------------------------------------------------
struct ParamField { const char* name; int ubo; int offset_floats; const char* origin; };
static const ParamField g_params[] = {
    // ConstantsUniform (binding 1), Main.cpp:184-190
    {"mass", CWA_UBO_CONSTANTS, 0, "ConstantsUniform.mass"},
    {"smoothing_coeff", CWA_UBO_CONSTANTS, 1, "ConstantsUniform.smoothing_coeff (h = smoothing_coeff * particle_radius)"},
    {"visc", CWA_UBO_CONSTANTS, 2, "ConstantsUniform.visc"},
    {"resting_rho", CWA_UBO_CONSTANTS, 3, "ConstantsUniform.resting_rho"},
    // BoundaryUniform (binding 2), Main.cpp:192-197
    {"upper.x", CWA_UBO_BOUNDARY, 0, "BoundaryUniform.upper"}, {"upper.y", CWA_UBO_BOUNDARY, 1, "BoundaryUniform.upper"},
    {"upper.z", CWA_UBO_BOUNDARY, 2, "BoundaryUniform.upper"}, {"upper.w", CWA_UBO_BOUNDARY, 3, "BoundaryUniform.upper"},
    {"lower.x", CWA_UBO_BOUNDARY, 4, "BoundaryUniform.lower"}, {"lower.y", CWA_UBO_BOUNDARY, 5, "BoundaryUniform.lower"},
    {"lower.z", CWA_UBO_BOUNDARY, 6, "BoundaryUniform.lower"}, {"lower.w", CWA_UBO_BOUNDARY, 7, "BoundaryUniform.lower"},
    // WaveUniforms (binding 3), Main.cpp:199-204
    {"wave.lambda", CWA_UBO_WAVE, 0, "WaveUniforms.attributes.x"}, {"wave.atten", CWA_UBO_WAVE, 1, "WaveUniforms.attributes.y"},
    {"wave.beta", CWA_UBO_WAVE, 2, "WaveUniforms.attributes.z"}, {"wave.type", CWA_UBO_WAVE, 3, "WaveUniforms.attributes.w"},
    {"mesh_ws_pos.x", CWA_UBO_WAVE, 4, "WaveUniforms.mesh_ws_pos"}, {"mesh_ws_pos.y", CWA_UBO_WAVE, 5, "WaveUniforms.mesh_ws_pos"},
    {"mesh_ws_pos.z", CWA_UBO_WAVE, 6, "WaveUniforms.mesh_ws_pos"}, {"mesh_ws_pos.w", CWA_UBO_WAVE, 7, "WaveUniforms.mesh_ws_pos"},
    // shader constants promoted to parameters (binding 4)
    {"particle_radius", CWA_UBO_SIM, 0, "const PARTICLE_RADIUS, rho_pres_comp.glsl:5"},
    {"gas_const", CWA_UBO_SIM, 1, "const GAS_CONST (k), rho_pres_comp.glsl:41"},
    {"dt", CWA_UBO_SIM, 2, "const TIME_STEP, integrate_comp.glsl:8"},
    {"gravity_y", CWA_UBO_SIM, 3, "const G.y, force_comp.glsl:52"},
    {"damping", CWA_UBO_SIM, 4, "const DAMPING, integrate_comp.glsl:51"},
    {"crest_threshold", CWA_UBO_SIM, 5, "literal 0.01, force_comp.glsl:91"},
    {"foam_speed", CWA_UBO_SIM, 6, "literal 25.0, integrate_comp.glsl:69"},
    {"uv_scale", CWA_UBO_SIM, 7, "literal 2.0 in 2.0*pos.xz, rho_pres_comp.glsl:72"},
    {"uv_scale_z", CWA_UBO_SIM, 8, "extension: t = uv_scale_z * pos.z (0 = uv_scale)"},
    {"torque_coeff", CWA_UBO_SIM, 9, "literal 0.25, force_comp.glsl:103 (0 = 0.25)"},
};
static const int g_num_params = (int)(sizeof(g_params) / sizeof(g_params[0]));

static const ParamField* find_param(const char* name)
{
    if (!name) return nullptr;
    for (int i = 0; i < g_num_params; i++) if (std::strcmp(g_params[i].name, name) == 0) return &g_params[i];
    return nullptr;
}

extern "C" int cwa_param_count(void) { return g_num_params; }

extern "C" int cwa_param_info(int index, const char** name, int* ubo_binding, int* byte_offset, const char** origin)
{
    CWA_CHECK(index >= 0 && index < g_num_params, "cwa_param_info: index %d out of range", index);
    if (name) *name = g_params[index].name;
    if (ubo_binding) *ubo_binding = g_params[index].ubo;
    if (byte_offset) *byte_offset = g_params[index].offset_floats * 4;
    if (origin) *origin = g_params[index].origin;
    return 0;
}

// writes the field of the block CURRENTLY BOUND at that UBO binding (like a GUI slider writing through glProgramUniform / the UBO)
extern "C" int cwa_param_set(cwa_ctx* ctx, const char* name, float value)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx, "null context");
    const ParamField* f = find_param(name);
    CWA_CHECK(f, "cwa_param_set: unknown parameter '%s'", name ? name : "(null)");
    return cwa_buffer_sub_data(ctx, ctx->ubo_binding[f->ubo], (size_t)f->offset_floats * 4, 4, &value);
}

extern "C" int cwa_param_get(cwa_ctx* ctx, const char* name, float* value)          // synchronises
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && value, "null argument");
    const ParamField* f = find_param(name);
    CWA_CHECK(f, "cwa_param_get: unknown parameter '%s'", name ? name : "(null)");
    return cwa_buffer_read(ctx, ctx->ubo_binding[f->ubo], (size_t)f->offset_floats * 4, 4, value);
}

// ---------------------------------------------------------------------------------------------
// checkpoint
// ---------------------------------------------------------------------------------------------
// File layout (little-endian, every section starts on a 16-byte boundary):
//   [0]    char magic[8] = "CWACKPT1"; uint32 version = 1; uint32 header_bytes = 256
//   [16]   uint64 frame; uint32 n_particles; uint32 particle_bytes = 64; uint32 wave_w, wave_h, wave_ch, wave_variant
//   [48]   int32 read_index[2], write_index, unit[3], tex_unit0, evolve          (StencilImage2DTripleBuffered + ImageTexture::mUnit)
//   [80]   cwa_constants_uniform (16 B); [96] cwa_boundary_uniform (32 B); [128] cwa_wave_uniforms (32 B); [160] cwa_sim_constants (48 B)
//   [208]  zero padding to 256
//   [256]  particle SSBO: n_particles x {pos, vel, force, extras}
//   then   wave image 0, 1, 2: wave_h x wave_w x wave_ch floats each (physical images, not roles)
struct CkptHeader {
    char     magic[8];
    uint32_t version, header_bytes;
    uint64_t frame;
    uint32_t n_particles, particle_bytes, wave_w, wave_h, wave_ch, wave_variant;
    int32_t  read_index[2], write_index, unit[3], tex_unit0, evolve;
    cwa_constants_uniform constants;
    cwa_boundary_uniform boundary;
    cwa_wave_uniforms wave;
    cwa_sim_constants sim;
    char     pad[256 - 208];
};
static_assert(sizeof(CkptHeader) == 256, "checkpoint header layout");

extern "C" int cwa_checkpoint_save(cwa_ctx* ctx, cwa_sph hs, cwa_wave hw, unsigned long long frame, const char* path)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, hs);
    WaveObj* w = get_wave(ctx, hw);
    CWA_CHECK(s && w && path, "cwa_checkpoint_save: invalid sph (%d) / wave (%d) handle or path", hs, hw);
    CWA_CHECK(w->row0 == 0 && w->h == w->h_global, "cwa_checkpoint_save: row-block wave objects are saved per rank through their own files only when whole");
    BufferObj* pb = get_buffer(ctx, s->particles);
    CWA_CHECK(pb, "cwa_checkpoint_save: particle buffer vanished");
    CkptHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, "CWACKPT1", 8);
    h.version = 1; h.header_bytes = 256; h.frame = frame;
    h.n_particles = (uint32_t)s->n; h.particle_bytes = 64;
    h.wave_w = (uint32_t)w->w; h.wave_h = (uint32_t)w->h; h.wave_ch = (uint32_t)w->ch; h.wave_variant = (uint32_t)w->variant;
    h.read_index[0] = w->read_index[0]; h.read_index[1] = w->read_index[1]; h.write_index = w->write_index;
    for (int i = 0; i < 3; i++) h.unit[i] = w->unit[i];
    h.tex_unit0 = w->tex_unit0; h.evolve = w->evolve ? 1 : 0;
    CWA_TRY(cwa_buffer_read(ctx, ctx->ubo_binding[CWA_UBO_CONSTANTS], 0, sizeof(h.constants), &h.constants));
    CWA_TRY(cwa_buffer_read(ctx, ctx->ubo_binding[CWA_UBO_BOUNDARY], 0, sizeof(h.boundary), &h.boundary));
    CWA_TRY(cwa_buffer_read(ctx, ctx->ubo_binding[CWA_UBO_WAVE], 0, sizeof(h.wave), &h.wave));
    CWA_TRY(cwa_buffer_read(ctx, ctx->ubo_binding[CWA_UBO_SIM], 0, sizeof(h.sim), &h.sim));
    const size_t pbytes = (size_t)s->n * 64, ibytes = (size_t)w->w * w->h * w->ch * 4;
    std::vector<char> host(pbytes > ibytes ? pbytes : ibytes);
    FILE* f = std::fopen(path, "wb");
    CWA_CHECK(f, "cwa_checkpoint_save: cannot open '%s' for writing", path);
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
    if (ok && pbytes) {
        if (cwa_buffer_read(ctx, s->particles, 0, pbytes, host.data()) != 0) { std::fclose(f); return -2; }
        ok = std::fwrite(host.data(), 1, pbytes, f) == pbytes;
    }
    for (int i = 0; ok && i < 3; i++) {
        if (cwa_wave_read_image(ctx, hw, i, (float*)host.data()) != 0) { std::fclose(f); return -2; }
        ok = std::fwrite(host.data(), 1, ibytes, f) == ibytes;
    }
    ok = (std::fclose(f) == 0) && ok;
    CWA_CHECK(ok, "cwa_checkpoint_save: short write to '%s'", path);
    return 0;
}

// Restores particles, wave levels, bookkeeping and the four parameter blocks into EXISTING objects of matching sizes.
extern "C" int cwa_checkpoint_load(cwa_ctx* ctx, cwa_sph hs, cwa_wave hw, const char* path, unsigned long long* frame)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, hs);
    WaveObj* w = get_wave(ctx, hw);
    CWA_CHECK(s && w && path, "cwa_checkpoint_load: invalid sph (%d) / wave (%d) handle or path", hs, hw);
    FILE* f = std::fopen(path, "rb");
    CWA_CHECK(f, "cwa_checkpoint_load: cannot open '%s'", path);
    CkptHeader h;
    if (std::fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "CWACKPT1", 8) != 0 || h.version != 1 || h.header_bytes != 256) {
        std::fclose(f);
        CWA_CHECK(false, "cwa_checkpoint_load: '%s' is not a version-1 checkpoint", path);
    }
    if ((int)h.n_particles != s->n || h.particle_bytes != 64 || (int)h.wave_w != w->w || (int)h.wave_h != w->h || (int)h.wave_ch != w->ch ||
        (int)h.wave_variant != w->variant) {
        std::fclose(f);
        CWA_CHECK(false, "cwa_checkpoint_load: checkpoint holds %u particles and a %ux%ux%u wave (variant %u); the objects hold %d and %dx%dx%d (variant %d)",
                  h.n_particles, h.wave_w, h.wave_h, h.wave_ch, h.wave_variant, s->n, w->w, w->h, w->ch, w->variant);
    }
    for (int i = 0; i < 3; i++) {
        const bool okidx = h.unit[i] >= 0 && h.unit[i] < 3;
        if (!okidx || h.write_index < 0 || h.write_index > 2 || h.read_index[0] < 0 || h.read_index[0] > 2 || h.read_index[1] < 0 || h.read_index[1] > 2 ||
            h.tex_unit0 < -1 || h.tex_unit0 > 2) {
            std::fclose(f);
            CWA_CHECK(false, "cwa_checkpoint_load: corrupt triple-buffer bookkeeping in '%s'", path);
        }
    }
    if (w->row0 != 0 || w->h != w->h_global) {                      // same rule as cwa_checkpoint_save: a row block is not a whole field
        std::fclose(f);
        CWA_CHECK(false, "cwa_checkpoint_load: the wave object is a row block of a larger field");
    }
    const size_t pbytes = (size_t)s->n * 64, ibytes = (size_t)w->w * w->h * w->ch * 4;
    // the whole payload must be there BEFORE any live object is touched: a truncated file leaves the state as it was
    {
        const long here = std::ftell(f);
        bool size_ok = here >= 0 && std::fseek(f, 0, SEEK_END) == 0;
        const long end = size_ok ? std::ftell(f) : -1;
        size_ok = size_ok && end >= 0 && (unsigned long long)end >= (unsigned long long)h.header_bytes + pbytes + 3ull * ibytes &&
                  std::fseek(f, here, SEEK_SET) == 0;
        if (!size_ok) {
            std::fclose(f);
            CWA_CHECK(false, "cwa_checkpoint_load: '%s' is truncated (needs %llu bytes); nothing was restored", path,
                      (unsigned long long)h.header_bytes + pbytes + 3ull * ibytes);
        }
    }
    std::vector<char> host(pbytes > ibytes ? pbytes : ibytes);
    bool ok = true;
    if (pbytes) {
        ok = std::fread(host.data(), 1, pbytes, f) == pbytes;
        if (ok && cwa_buffer_sub_data(ctx, s->particles, 0, pbytes, host.data()) != 0) ok = false;
        if (ok) cudaStreamSynchronize(ctx->stream);                 // `host` is reused below
    }
    for (int i = 0; ok && i < 3; i++) {
        ok = std::fread(host.data(), 1, ibytes, f) == ibytes;
        if (ok && cwa_wave_write_image(ctx, hw, i, (const float*)host.data()) != 0) ok = false;
        if (ok) cudaStreamSynchronize(ctx->stream);
    }
    std::fclose(f);
    CWA_CHECK(ok, "cwa_checkpoint_load: '%s' is truncated", path);
    w->read_index[0] = h.read_index[0]; w->read_index[1] = h.read_index[1]; w->write_index = h.write_index;
    for (int i = 0; i < 3; i++) w->unit[i] = h.unit[i];
    w->tex_unit0 = h.tex_unit0; w->evolve = h.evolve != 0;
    CWA_TRY(cwa_buffer_sub_data(ctx, ctx->ubo_binding[CWA_UBO_CONSTANTS], 0, sizeof(h.constants), &h.constants));
    CWA_TRY(cwa_buffer_sub_data(ctx, ctx->ubo_binding[CWA_UBO_BOUNDARY], 0, sizeof(h.boundary), &h.boundary));
    CWA_TRY(cwa_buffer_sub_data(ctx, ctx->ubo_binding[CWA_UBO_WAVE], 0, sizeof(h.wave), &h.wave));
    CWA_TRY(cwa_buffer_sub_data(ctx, ctx->ubo_binding[CWA_UBO_SIM], 0, sizeof(h.sim), &h.sim));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));                   // the header is on this function's stack
    if (frame) *frame = h.frame;
    return 0;
}
