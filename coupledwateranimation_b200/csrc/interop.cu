// interop.cu -- CUDA-GL interop hand-off: the renderer of the reference stays untouched and keeps its GL objects;
// the simulation step reads / writes them through cudaGraphics* mappings, outside the timed path.
//   particle SSBO as a vertex buffer        CoupledWaterAnimation/Main.cpp:779-790 (glGenBuffers / glBufferData / attrib 0 = pos, stride 64)
//   newest wave level as a texture          Main.cpp:413 (wave2d.GetReadImage(0).BindTextureUnit()), wave_vs.glsl:14,34
//   ReinitFromTexture(GLuint)               StencilImage2DTripleBuffered.cpp:61-77 + wave_comp.glsl:76-80 (init-textures/*.png, RGBA8)
// Compiled when cuda_gl_interop.h is usable (always, with the GL type stand-in of csrc/gl_compat when the machine has no GL headers);
// without a current GL context the register calls fail with the CUDA error text, they never crash.
#include "internal.cuh"

#if defined(__has_include)
#if __has_include(<cuda_gl_interop.h>) && __has_include(<GL/gl.h>)
#define CWA_HAVE_GL_INTEROP 1
#endif
#endif

#ifdef CWA_HAVE_GL_INTEROP
#include <cuda_gl_interop.h>
#endif

static std::vector<GlResource>& gl_table(cwa_ctx* ctx) { return ctx->gl_resources; }
static GlResource* gl_get(cwa_ctx* ctx, int h)
{
    auto& t = gl_table(ctx);
    if (h < 0 || h >= (int)t.size() || !t[h].live) return nullptr;
    return &t[h];
}

extern "C" int cwa_gl_available(void)
{
#ifdef CWA_HAVE_GL_INTEROP
    return 1;
#else
    return 0;
#endif
}

#ifndef CWA_HAVE_GL_INTEROP
#define CWA_GL_UNAVAILABLE() do { cwa_set_error("libcwa_b200 was built without cuda_gl_interop.h"); return -4; } while (0)
#endif

// glGenBuffers name of the particle SSBO / vertex buffer -> a handle that can be mapped as a cwa_buf
extern "C" int cwa_gl_register_buffer(cwa_ctx* ctx, unsigned gl_buffer, int* resource)
{
    CWA_CHECK(ctx && resource, "null argument");
    *resource = -1;
#ifdef CWA_HAVE_GL_INTEROP
    DeviceGuard dg(ctx);
    cudaGraphicsResource_t r = nullptr;
    CWA_CUDA(cudaGraphicsGLRegisterBuffer(&r, (GLuint)gl_buffer, cudaGraphicsRegisterFlagsNone));
    GlResource g; g.live = true; g.is_image = false; g.res = r;
    gl_table(ctx).push_back(g);
    *resource = (int)gl_table(ctx).size() - 1;
    return 0;
#else
    (void)gl_buffer; CWA_GL_UNAVAILABLE();
#endif
}

// glCreateTextures name of a wave level (GL_TEXTURE_2D, RGBA32F as shipped or R32F) -> a handle cwa_gl_copy_wave_to_image writes to
extern "C" int cwa_gl_register_image(cwa_ctx* ctx, unsigned gl_texture, unsigned gl_target, int* resource)
{
    CWA_CHECK(ctx && resource, "null argument");
    *resource = -1;
#ifdef CWA_HAVE_GL_INTEROP
    DeviceGuard dg(ctx);
    cudaGraphicsResource_t r = nullptr;
    CWA_CUDA(cudaGraphicsGLRegisterImage(&r, (GLuint)gl_texture, (GLenum)(gl_target ? gl_target : GL_TEXTURE_2D), cudaGraphicsRegisterFlagsWriteDiscard));
    GlResource g; g.live = true; g.is_image = true; g.res = r;
    gl_table(ctx).push_back(g);
    *resource = (int)gl_table(ctx).size() - 1;
    return 0;
#else
    (void)gl_texture; (void)gl_target; CWA_GL_UNAVAILABLE();
#endif
}

// Map a registered GL buffer and expose it as a cwa_buf (no copy): create the SPH object on it and the vertex buffer the renderer
// draws IS the particle SSBO, as in the reference.  Unmap before the GL draw calls of the frame.
extern "C" int cwa_gl_map_buffer(cwa_ctx* ctx, int resource, cwa_buf* out)
{
    CWA_CHECK(ctx && out, "null argument");
    *out = -1;
#ifdef CWA_HAVE_GL_INTEROP
    DeviceGuard dg(ctx);
    GlResource* g = gl_get(ctx, resource);
    CWA_CHECK(g && !g->is_image, "cwa_gl_map_buffer: handle %d is not a registered GL buffer", resource);
    CWA_CHECK(!g->mapped, "cwa_gl_map_buffer: already mapped");
    cudaGraphicsResource_t r = (cudaGraphicsResource_t)g->res;
    CWA_CUDA(cudaGraphicsMapResources(1, &r, ctx->stream));
    void* p = nullptr; size_t bytes = 0;
    CWA_CUDA(cudaGraphicsResourceGetMappedPointer(&p, &bytes, r));
    if (g->mapped_buf >= 0 && get_buffer(ctx, g->mapped_buf)) {       // the same handle as last frame: objects built on it stay valid
        BufferObj* b = get_buffer(ctx, g->mapped_buf);
        b->ptr = p; b->bytes = bytes;
    } else {
        g->mapped_buf = new_buffer(ctx, p, bytes, false);
    }
    g->mapped = true;
    sph_invalidate_for_buffer(ctx, g->mapped_buf);                    // GL may have written it while it was unmapped
    *out = g->mapped_buf;
    return 0;
#else
    (void)resource; CWA_GL_UNAVAILABLE();
#endif
}

extern "C" int cwa_gl_unmap(cwa_ctx* ctx, int resource)
{
    CWA_CHECK(ctx, "null context");
#ifdef CWA_HAVE_GL_INTEROP
    DeviceGuard dg(ctx);
    GlResource* g = gl_get(ctx, resource);
    CWA_CHECK(g && g->mapped, "cwa_gl_unmap: handle %d is not mapped", resource);
    cudaGraphicsResource_t r = (cudaGraphicsResource_t)g->res;
    CWA_CUDA(cudaGraphicsUnmapResources(1, &r, ctx->stream));
    g->mapped = false;
    return 0;
#else
    (void)resource; CWA_GL_UNAVAILABLE();
#endif
}

// display(): the level GetReadImage(0) holds goes to the GL texture the wave mesh samples (Main.cpp:413): map, copy the device
// image into the texture's array (R32F for scalar fields, RGBA32F for 4-channel ones), unmap.  image = physical index 0..2.
extern "C" int cwa_gl_copy_wave_to_image(cwa_ctx* ctx, int resource, cwa_wave hw, int image)
{
    CWA_CHECK(ctx, "null context");
#ifdef CWA_HAVE_GL_INTEROP
    DeviceGuard dg(ctx);
    GlResource* g = gl_get(ctx, resource);
    WaveObj* w = get_wave(ctx, hw);
    CWA_CHECK(g && g->is_image, "cwa_gl_copy_wave_to_image: handle %d is not a registered GL image", resource);
    CWA_CHECK(w && image >= 0 && image < 3, "cwa_gl_copy_wave_to_image: invalid wave handle %d / image %d", hw, image);
    cudaGraphicsResource_t r = (cudaGraphicsResource_t)g->res;
    CWA_CUDA(cudaGraphicsMapResources(1, &r, ctx->stream));
    cudaArray_t arr = nullptr;
    cudaError_t e = cudaGraphicsSubResourceGetMappedArray(&arr, r, 0, 0);
    if (e == cudaSuccess) {
        const size_t row = (size_t)w->w * w->ch * 4;
        e = cudaMemcpy2DToArrayAsync(arr, 0, 0, w->image[image], row, row, (size_t)w->h, cudaMemcpyDeviceToDevice, ctx->stream);
    }
    cudaGraphicsUnmapResources(1, &r, ctx->stream);
    CWA_CUDA(e);
    return 0;
#else
    (void)resource; (void)hw; (void)image; CWA_GL_UNAVAILABLE();
#endif
}

extern "C" int cwa_gl_unregister(cwa_ctx* ctx, int resource)
{
    CWA_CHECK(ctx, "null context");
#ifdef CWA_HAVE_GL_INTEROP
    DeviceGuard dg(ctx);
    GlResource* g = gl_get(ctx, resource);
    CWA_CHECK(g, "cwa_gl_unregister: invalid handle %d", resource);
    cudaGraphicsResource_t r = (cudaGraphicsResource_t)g->res;
    if (g->mapped) cudaGraphicsUnmapResources(1, &r, ctx->stream);
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    if (g->mapped_buf >= 0) if (BufferObj* b = get_buffer(ctx, g->mapped_buf)) { b->live = false; b->ptr = nullptr; }
    CWA_CUDA(cudaGraphicsUnregisterResource(r));
    g->live = false;
    return 0;
#else
    (void)resource; CWA_GL_UNAVAILABLE();
#endif
}

// ReinitFromTexture with the decoded bytes of an init texture (init-textures/*.png loaded as GL_RGBA8, LoadTexture.cpp): texelFetch
// on a normalised 8-bit texture returns c / 255 per channel (OpenGL 4.5 spec 2.3.5.1), then wave_comp.glsl:76-80 as for float data.
extern "C" int cwa_wave_reinit_from_rgba8(cwa_ctx* ctx, cwa_wave hw, const unsigned char* rgba8, int tw, int th)
{
    CWA_CHECK(ctx && rgba8 && tw >= 1 && th >= 1, "cwa_wave_reinit_from_rgba8: null texture or bad size");
    std::vector<float> f((size_t)tw * th * 4);
    for (size_t i = 0; i < f.size(); i++) f[i] = (float)rgba8[i] / 255.0f;
    return cwa_wave_reinit_from_texture(ctx, hw, f.data(), tw, th);
}
