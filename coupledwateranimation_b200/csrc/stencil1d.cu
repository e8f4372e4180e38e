// stencil1d.cu -- the 1-D wave substrates of the 2-D app for sm_100a:
//   SphWave2D/Shallow1D_cs.glsl   two-phase Lax-Wendroff shallow water, texel (h, uh, hm, uhm), double buffered
//   SphWave2D/Wave1D_cs.glsl      damped 1-D wave equation, texel (u, v, a, -), triple buffered, 10 substeps
//   SphWave2D/StencilImage2D.cpp  ImageStencil: N-buffer ping-pong (mReadIndex / mWriteIndex / per-image unit), Reinit, Compute,
//                                 ComputeFunc, ReinitFromTexture
// The reference issues one dispatch + barrier + PingPong per mode and substep for a 128- or 1024-texel image: 2 (shallow) or 10
// (wave) launches per frame that each do a few hundred nanoseconds of work.  Here a whole Compute() call -- every substep and
// mode of every frame asked for -- is ONE launch: a single CTA keeps all N images in shared memory, runs the dispatches back to
// back with a __syncthreads() where the reference has a glMemoryBarrier, rotates the image roles in registers (PingPong is a
// cyclic shift of the units) and writes the images back once.  Images too wide for shared memory fall back to one launch per
// dispatch.  Arithmetic is spelled with __f*_rn in the GLSL's association order (no FMA contraction), so the iteration is
// bit-identical to the CPU restatement.
#include "internal.cuh"

struct Stencil1dParams {
    float lambda, p3 /* dx (shallow) | atten (wave) */, beta, b0, b1;
    int   bc;
};

enum { S1D_SHALLOW = CWA_STENCIL1D_SHALLOW, S1D_WAVE = CWA_STENCIL1D_WAVE };

__device__ __forceinline__ float4 s1d_load(const float4* img, int w, int x)
{
    return ((unsigned)x < (unsigned)w) ? img[x] : make_float4(0.f, 0.f, 0.f, 0.f);       // imageLoad outside the image: zeros
}
__device__ __forceinline__ float s1d_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
// exp() evaluated in double and rounded once (GLSL leaves its precision implementation-defined; shared with the oracle)
__device__ __forceinline__ float s1d_exp(float x) { return (float)exp((double)x); }

#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))
#define DIV(a, b) __fdiv_rn((a), (b))

// EnforceBC (Shallow1D_cs.glsl:169-235, Wave1D_cs.glsl:132-196), stores in program order
__device__ __forceinline__ void s1d_enforce_bc(float4* out, int w, int coord, float4 c, const Stencil1dParams& p, bool negate_all)
{
    const float boundary_scale = 0.1f;
    if (p.bc == CWA_BC_FIXED) {
        if (coord == 0) { out[0] = make_float4(MUL(p.b0, boundary_scale), 0.f, 0.f, 0.f); return; }
        if (coord == w - 1) { out[coord] = make_float4(MUL(p.b1, boundary_scale), 0.f, 0.f, 0.f); return; }
        out[coord] = c;
        return;
    }
    if (coord == 0 || coord == w - 1) return;          // neighbours write the boundary values
    out[coord] = c;
    if (p.bc == CWA_BC_REFLECT) {
        if (negate_all) c = make_float4(-c.x, -c.y, -c.z, -c.w); else c.y = -c.y;
    }
    if (coord == 1) out[0] = c;
    if (coord == w - 2) out[w - 1] = c;
}

// One invocation of Shallow1D_cs.glsl main() for `coord`.  Stores of several invocations to the same texel are resolved the way a
// lock-step GPU resolves them (later in program order wins): in ITERATE0 the unconditional final store (:160) overwrites every
// EnforceBC copy, so only it is issued.
__device__ __forceinline__ void shallow1d_point(const float4* in, float4* out, int w, int mode, int coord, const Stencil1dParams& p)
{
    const float VIEW_HEIGHT = 2.0f * 4.8f;
    const float lambda = p.lambda, dx = p.p3;
    if (mode == 0 || mode == 1) {
        const float x = DIV((float)coord, (float)(w - 1));
        const float xc = SUB(x, 0.5f);
        const float e = s1d_exp(DIV(MUL(-xc, xc), 0.005f));
        if (mode == 0) {                                                     // InitWave :120-134
            const float h = MUL(MUL(VIEW_HEIGHT, 0.1f), e);
            out[coord] = make_float4(ADD(MUL(VIEW_HEIGHT, 0.5f), h), MUL(MUL(0.5f, fabsf(h)), s1d_sign(xc)), 0.f, 0.f);
        } else {                                                             // Splash :103-118
            float4 v = s1d_load(in, w, coord);
            const float h = MUL(MUL(-VIEW_HEIGHT, 0.1f), e);
            v.x = ADD(v.x, h);
            v.y = ADD(v.y, MUL(MUL(0.2f, fabsf(h)), s1d_sign(xc)));
            out[coord] = v;
        }
        return;
    }
    const float4 c = s1d_load(in, w, coord);
    if (mode == 2) {                                                         // ITERATE0 :147-160
        const float4 e = s1d_load(in, w, coord + 1);
        float4 r = c;
        r.z = SUB(DIV(ADD(c.x, e.x), 2.0f), DIV(MUL(DIV(lambda, 2.0f), SUB(e.y, c.y)), dx));
        // (eUH^2/eH + 0.5 G eH^2 - cUH^2/cH - 0.5 G cH^2), left to right
        const float G = 9.8f;
        float s = ADD(DIV(MUL(e.y, e.y), e.x), MUL(MUL(MUL(0.5f, G), e.x), e.x));
        s = SUB(s, DIV(MUL(c.y, c.y), c.x));
        s = SUB(s, MUL(MUL(MUL(0.5f, G), c.x), c.x));
        r.w = SUB(DIV(ADD(c.y, e.y), 2.0f), DIV(MUL(DIV(lambda, 2.0f), s), dx));
        out[coord] = r;
    } else if (mode == 3) {                                                  // ITERATE1 :161-167
        const float4 ww = s1d_load(in, w, coord - 1);
        float4 r = c;
        r.x = SUB(c.x, DIV(MUL(lambda, SUB(c.w, ww.w)), dx));
        const float G = 9.8f;
        float s = ADD(DIV(MUL(c.w, c.w), c.z), MUL(MUL(MUL(0.5f, G), c.z), c.z));
        s = SUB(s, DIV(MUL(ww.w, ww.w), ww.z));
        s = SUB(s, MUL(MUL(MUL(0.5f, G), ww.z), ww.z));
        r.y = SUB(c.y, DIV(MUL(lambda, s), dx));
        s1d_enforce_bc(out, w, coord, r, p, false);
    }
}

// One invocation of Wave1D_cs.glsl main().  MODE_INIT_1 stores the initial profile and then the result of one step taken from the
// INPUT images; texels 0 and w-1 end up with the neighbours' boundary copies (later in program order), so with the FREE / REFLECT
// conditions the first store of those two invocations is not issued (w >= 3).
__device__ __forceinline__ void wave1d_point(const float4* in0, const float4* in1, float4* out, int w, int mode, int coord,
                                             const Stencil1dParams& p)
{
    const float lambda = p.lambda, atten = p.p3, beta = p.beta;
    if (mode == 0 || mode == 1) {                                            // InitWave :98-121
        const int cen = (int)MUL(0.25f, (float)w) + mode;
        const int x = cen - coord;
        float d0 = 0.0f;
        if (p.bc == CWA_BC_FIXED) {
            const float a = MUL(p.b0, 0.1f), b = MUL(p.b1, 0.1f), t = DIV((float)coord, (float)(w - 1));
            d0 = ADD(MUL(a, SUB(1.0f, t)), MUL(b, t));                       // mix(a, b, t)
        }
        const float4 v = make_float4(ADD(d0, MUL(0.1f, s1d_exp(DIV((float)(-x * x), 5000.0f)))), 0.f, 0.f, 0.f);
        const bool overwritten = (mode == 1) && (p.bc != CWA_BC_FIXED) && (w >= 3) && (coord == 0 || coord == w - 1);
        if (!overwritten) out[coord] = v;
        if (mode == 0) return;
        const float4 c1 = s1d_load(in1, w, coord), c0 = s1d_load(in0, w, coord);
        const float4 e0 = s1d_load(in0, w, coord + 1), w0 = s1d_load(in0, w, coord - 1);
        const float hl = MUL(0.5f, lambda);
        float4 r;
        r.x = SUB(c0.x, MUL(hl, ADD(SUB(e0.x, MUL(2.0f, c0.x)), w0.x)));
        r.w = SUB(c0.w, MUL(hl, ADD(SUB(e0.w, MUL(2.0f, c0.w)), w0.w)));
        r.y = DIV(SUB(r.x, c1.x), 2.0f);                                     // ComputeVelocityAcceleration :73-79
        r.z = ADD(SUB(r.x, MUL(2.0f, c0.x)), c1.x);
        s1d_enforce_bc(out, w, coord, r, p, true);
        return;
    }
    if (mode != 2) return;
    const float4 c1 = s1d_load(in1, w, coord), c0 = s1d_load(in0, w, coord);
    const float4 e0 = s1d_load(in0, w, coord + 1), w0 = s1d_load(in0, w, coord - 1);
    const float kc = SUB(SUB(2.0f, MUL(2.0f, lambda)), beta), k1 = SUB(1.0f, beta);
    float4 r;                                                                // IterateWave :123-130
    r.x = MUL(atten, SUB(ADD(MUL(kc, c0.x), MUL(lambda, ADD(e0.x, w0.x))), MUL(k1, c1.x)));
    r.w = MUL(atten, SUB(ADD(MUL(kc, c0.w), MUL(lambda, ADD(e0.w, w0.w))), MUL(k1, c1.w)));
    r.y = DIV(SUB(r.x, c1.x), 2.0f);
    r.z = ADD(SUB(r.x, MUL(2.0f, c0.x)), c1.x);
    s1d_enforce_bc(out, w, coord, r, p, true);
}

#undef MUL
#undef ADD
#undef SUB
#undef DIV

template <int SHADER>
__device__ __forceinline__ void s1d_point(const float4* in0, const float4* in1, float4* out, int w, int mode, int coord,
                                          const Stencil1dParams& p)
{
    if (SHADER == S1D_SHALLOW) shallow1d_point(in0, out, w, mode, coord, p);
    else wave1d_point(in0, in1, out, w, mode, coord, p);
}

// One launch per dispatch (wide images): grid-stride over the texels, images in global memory.
template <int SHADER>
__global__ void __launch_bounds__(256)
stencil1d_dispatch_kernel(const float4* __restrict__ in0, const float4* __restrict__ in1, float4* __restrict__ out, int w, int mode,
                          Stencil1dParams p)
{
    const int coord = blockIdx.x * blockDim.x + threadIdx.x;
    if (coord < w) s1d_point<SHADER>(in0, in1, out, w, mode, coord, p);
}

// The whole Compute() call in one launch: one CTA, the N images live in shared memory, dispatch d runs mode
// mode0 + d % nmodes on the images that currently hold units 0 / 1 / N-1; PingPong = cyclic shift (out -> unit 0 -> unit 1 -> out).
constexpr int S1D_THREADS = 1024;
template <int SHADER>
__global__ void __launch_bounds__(S1D_THREADS)
stencil1d_fused_kernel(float4* __restrict__ g0, float4* __restrict__ g1, float4* __restrict__ g2, int w, int mode0, int nmodes,
                       int ndispatch, Stencil1dParams p)
{
    extern __shared__ float4 s1d_smem[];
    constexpr int N = (SHADER == S1D_SHALLOW) ? 2 : 3;
    float4* gimg[3] = {g0, g1, g2};              // by ROLE at entry: [0] unit 0, [1] unit 1 (wave) / output (shallow), [2] output (wave)
    float4* simg[3] = {s1d_smem, s1d_smem + w, s1d_smem + 2 * (size_t)w};
    for (int k = 0; k < N; k++)
        for (int x = threadIdx.x; x < w; x += S1D_THREADS) simg[k][x] = gimg[k][x];
    __syncthreads();
    int r0 = 0, r1 = 1, ro = N - 1;              // shared-memory image holding unit 0 / unit 1 / the output unit
    for (int d = 0; d < ndispatch; d++) {
        const int mode = mode0 + d % nmodes;
        for (int x = threadIdx.x; x < w; x += S1D_THREADS) s1d_point<SHADER>(simg[r0], simg[r1], simg[ro], w, mode, x, p);
        __syncthreads();                         // glMemoryBarrier
        if (N == 2) { const int t = r0; r0 = ro; ro = t; r1 = ro; }
        else { const int t = ro; ro = r1; r1 = r0; r0 = t; }
    }
    for (int k = 0; k < N; k++)                  // storage image k never moves; only the roles did
        for (int x = threadIdx.x; x < w; x += S1D_THREADS) gimg[k][x] = simg[k][x];
}

// ---------------------------------------------------------------------------------------------
// host object: ImageStencil (SphWave2D/StencilImage2D.h:10-66, .cpp:4-164)
// ---------------------------------------------------------------------------------------------
static Stencil1dObj* get_s1d(cwa_ctx* ctx, cwa_stencil1d h)
{
    if (!ctx || h < 0 || h >= (int)ctx->stencil1ds.size() || !ctx->stencil1ds[h].live) return nullptr;
    return &ctx->stencil1ds[h];
}

static int s1d_image_with_unit(const Stencil1dObj* s, int u)
{
    for (int i = 0; i < s->num_images; i++) if (s->unit[i] == u) return i;
    return -1;
}

// PingPong, StencilImage2D.cpp:67-83 (as written: the index arrays and the units follow different permutations for N = 3)
static void s1d_pingpong(Stencil1dObj* s)
{
    if (s->num_images == 1) return;
    const int nread = s->num_images - 1;
    std::swap(s->write_index, s->read_index[0]);
    for (int i = 0; i < nread - 1; i++) std::swap(s->read_index[i], s->read_index[i + 1]);
    std::swap(s->unit[s->write_index], s->unit[s->read_index[0]]);
    for (int i = 0; i < nread - 1; i++) std::swap(s->unit[s->read_index[i]], s->unit[s->read_index[i + 1]]);
}

static Stencil1dParams s1d_params(const Stencil1dObj* s)
{
    Stencil1dParams p;
    p.lambda = s->lambda; p.p3 = s->dx_or_atten; p.beta = s->beta; p.b0 = s->boundary[0]; p.b1 = s->boundary[1]; p.bc = s->bc;
    return p;
}

// `ndispatch` dispatches, dispatch d in mode mode0 + d % nmodes, each followed by PingPong
static int s1d_run(cwa_ctx* ctx, Stencil1dObj* s, int mode0, int nmodes, int ndispatch)
{
    if (ndispatch <= 0) return 0;
    const int N = s->num_images, w = s->w;
    const Stencil1dParams p = s1d_params(s);
    const size_t smem = (size_t)N * w * sizeof(float4);
    if (smem <= 200 * 1024) {
        // the kernel wants the images by role; hand it the storage images permuted so that slot k holds unit k's image
        float4* by_unit[3] = {nullptr, nullptr, nullptr};
        for (int u = 0; u < N; u++) by_unit[u] = s->image[s1d_image_with_unit(s, u)];
        KScope k(ctx, KID_WAVE);
        if (s->shader == S1D_SHALLOW) {
            CWA_TRY(ensure_dynamic_smem(ctx, stencil1d_fused_kernel<S1D_SHALLOW>, 200 * 1024));
            stencil1d_fused_kernel<S1D_SHALLOW><<<1, S1D_THREADS, smem, ctx->stream>>>(by_unit[0], by_unit[1], nullptr, w, mode0, nmodes, ndispatch, p);
        } else {
            CWA_TRY(ensure_dynamic_smem(ctx, stencil1d_fused_kernel<S1D_WAVE>, 200 * 1024));
            stencil1d_fused_kernel<S1D_WAVE><<<1, S1D_THREADS, smem, ctx->stream>>>(by_unit[0], by_unit[1], by_unit[2], w, mode0, nmodes, ndispatch, p);
        }
        CWA_CUDA(cudaGetLastError());
        for (int d = 0; d < ndispatch; d++) s1d_pingpong(s);
        return 0;
    }
    for (int d = 0; d < ndispatch; d++) {
        const int mode = mode0 + d % nmodes;
        const float4* in0 = s->image[s1d_image_with_unit(s, 0)];
        const float4* in1 = (N == 3) ? s->image[s1d_image_with_unit(s, 1)] : in0;
        float4* out = s->image[s1d_image_with_unit(s, N - 1)];
        { KScope k(ctx, KID_WAVE);
          if (s->shader == S1D_SHALLOW) stencil1d_dispatch_kernel<S1D_SHALLOW><<<ceil_div(w, 256), 256, 0, ctx->stream>>>(in0, in1, out, w, mode, p);
          else stencil1d_dispatch_kernel<S1D_WAVE><<<ceil_div(w, 256), 256, 0, ctx->stream>>>(in0, in1, out, w, mode, p); }
        CWA_CUDA(cudaGetLastError());
        s1d_pingpong(s);
    }
    return 0;
}

// ComputeShader::Dispatch on a bound ImageStencil object: one dispatch, roles untouched
int stencil1d_dispatch_mode(cwa_ctx* ctx, int handle, int mode, int shader)
{
    Stencil1dObj* s = get_s1d(ctx, handle);
    CWA_CHECK(s, "dispatch of a 1-D wave shader: no ImageStencil object bound to the shader (cwa_shader_bind_object)");
    CWA_CHECK(s->shader == shader, "dispatch: the bound ImageStencil object was created for the other 1-D shader");
    CWA_CHECK(mode >= 0 && mode <= 3, "dispatch: unsupported uMode %d", mode);
    const int N = s->num_images, w = s->w;
    const Stencil1dParams p = s1d_params(s);
    const float4* in0 = s->image[s1d_image_with_unit(s, 0)];
    const float4* in1 = (N == 3) ? s->image[s1d_image_with_unit(s, 1)] : in0;
    float4* out = s->image[s1d_image_with_unit(s, N - 1)];
    KScope k(ctx, KID_WAVE);
    if (s->shader == S1D_SHALLOW) stencil1d_dispatch_kernel<S1D_SHALLOW><<<ceil_div(w, 256), 256, 0, ctx->stream>>>(in0, in1, out, w, mode, p);
    else stencil1d_dispatch_kernel<S1D_WAVE><<<ceil_div(w, 256), 256, 0, ctx->stream>>>(in0, in1, out, w, mode, p);
    CWA_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int cwa_stencil1d_pingpong(cwa_ctx* ctx, cwa_stencil1d h)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s, "invalid stencil1d handle %d", h);
    s1d_pingpong(s);
    return 0;
}

extern "C" int cwa_stencil1d_create(cwa_ctx* ctx, int shader, int width, cwa_stencil1d* out)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && out, "null argument");
    *out = -1;
    CWA_CHECK(shader == S1D_SHALLOW || shader == S1D_WAVE, "cwa_stencil1d_create: unknown shader %d", shader);
    CWA_CHECK(width >= 1 && width <= (1 << 24), "cwa_stencil1d_create: bad width %d", width);
    Stencil1dObj s;
    s.live = true; s.shader = shader; s.w = width;
    if (shader == S1D_SHALLOW) {                 // InitShallowWaterEquation, SphWave2D/Main.cpp:77-97 + Shallow1D_cs.glsl:23-26
        s.num_images = 2; s.mode_iter_first = 2; s.mode_iter_last = 3; s.substeps = 1;
        s.lambda = 0.001f; s.dx_or_atten = 0.1f; s.beta = 0.001f;
    } else {                                     // InitWaveEquation, SphWave2D/Main.cpp:63-75 + Wave1D_cs.glsl:19-22
        s.num_images = 3; s.mode_iter_first = 2; s.mode_iter_last = 2; s.substeps = 10;
        s.lambda = 0.01f; s.dx_or_atten = 0.9995f; s.beta = 0.001f;
    }
    // SetNumBuffers :38-65, Init :12-36 (SetUnit(i); fresh texture storage is treated as zeros)
    for (int i = 0; i < s.num_images - 1; i++) s.read_index[i] = i;
    s.write_index = s.num_images - 1;
    for (int i = 0; i < s.num_images; i++) {
        s.unit[i] = i;
        CWA_CUDA(cudaMalloc(&s.image[i], (size_t)width * sizeof(float4)));
        CWA_CUDA(cudaMemsetAsync(s.image[i], 0, (size_t)width * sizeof(float4), ctx->stream));
        s.image_buf[i] = new_buffer(ctx, s.image[i], (size_t)width * sizeof(float4), false);
    }
    ctx->stencil1ds.push_back(s);
    *out = (int)ctx->stencil1ds.size() - 1;
    return cwa_stencil1d_reinit(ctx, *out);      // Init() ends with Reinit() :35
}

extern "C" int cwa_stencil1d_destroy(cwa_ctx* ctx, cwa_stencil1d h)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s, "invalid stencil1d handle %d", h);
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < s->num_images; i++) {
        cudaFree(s->image[i]);
        if (BufferObj* b = get_buffer(ctx, s->image_buf[i])) b->live = false;
    }
    s->live = false;
    return 0;
}

// Reinit :85-105: one init dispatch per read image (modes mMODE_INIT_FIRST + i), PingPong after each
extern "C" int cwa_stencil1d_reinit(cwa_ctx* ctx, cwa_stencil1d h)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s, "invalid stencil1d handle %d", h);
    const int nread = s->num_images - 1;
    return s1d_run(ctx, s, 0, nread, nread);
}

// ReinitFromTexture :122-140 (mode -1: texelFetch(uInitImage, coord) -> output image), then PingPong
extern "C" int cwa_stencil1d_reinit_from_texture(cwa_ctx* ctx, cwa_stencil1d h, const float* rgba, int width)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s && rgba && width >= 1, "cwa_stencil1d_reinit_from_texture: invalid handle %d or texture", h);
    float4* out = s->image[s1d_image_with_unit(s, s->num_images - 1)];
    const int n = width < s->w ? width : s->w;
    CWA_CUDA(cudaMemsetAsync(out, 0, (size_t)s->w * sizeof(float4), ctx->stream));            // texelFetch outside the texture: zeros
    CWA_CUDA(cudaMemcpyAsync(out, rgba, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));                                             // `rgba` may be pageable
    s1d_pingpong(s);
    return 0;
}

// Compute :142-164: nframes x substeps x (modes MODE_ITERATE_FIRST..LAST), one launch
extern "C" int cwa_stencil1d_compute(cwa_ctx* ctx, cwa_stencil1d h, int nframes)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s && nframes >= 0, "invalid stencil1d handle %d", h);
    if (!s->iterate) return 0;                   // mIterate == false :144
    const int nmodes = s->mode_iter_last - s->mode_iter_first + 1;
    long long nd = (long long)nframes * s->substeps * nmodes;
    while (nd > 0) {                             // bounded launches (a watchdog-friendly 64 K dispatches each)
        const int chunk = nd > 65536 ? 65536 - 65536 % nmodes : (int)nd;
        CWA_TRY(s1d_run(ctx, s, s->mode_iter_first, nmodes, chunk));
        nd -= chunk;
    }
    return 0;
}

// ComputeFunc(mode) :107-120: one dispatch in `mode` (e.g. Splash = MODE_INIT_1 of Shallow1D) + PingPong
extern "C" int cwa_stencil1d_compute_func(cwa_ctx* ctx, cwa_stencil1d h, int mode)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s, "invalid stencil1d handle %d", h);
    CWA_CHECK(mode >= 0 && mode <= 3, "cwa_stencil1d_compute_func: unsupported uMode %d", mode);
    return s1d_run(ctx, s, mode, 1, 1);
}

extern "C" int cwa_stencil1d_set_params(cwa_ctx* ctx, cwa_stencil1d h, float lambda, float dx_or_atten, float beta,
                                        float boundary0, float boundary1, int bc)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s, "invalid stencil1d handle %d", h);
    CWA_CHECK(bc == CWA_BC_REFLECT || bc == CWA_BC_FREE || bc == CWA_BC_FIXED, "cwa_stencil1d_set_params: unknown boundary condition %d", bc);
    s->lambda = lambda; s->dx_or_atten = dx_or_atten; s->beta = beta; s->boundary[0] = boundary0; s->boundary[1] = boundary1; s->bc = bc;
    return 0;
}

extern "C" int cwa_stencil1d_set_substeps(cwa_ctx* ctx, cwa_stencil1d h, int substeps)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s && substeps >= 0, "invalid stencil1d handle %d or substep count", h);
    s->substeps = substeps;
    return 0;
}

extern "C" int cwa_stencil1d_set_iterate(cwa_ctx* ctx, cwa_stencil1d h, int iterate)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s, "invalid stencil1d handle %d", h);
    s->iterate = iterate != 0;
    return 0;
}

extern "C" int cwa_stencil1d_state(cwa_ctx* ctx, cwa_stencil1d h, int* num_images, int read_index[2], int* write_index, int unit[3])
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s, "invalid stencil1d handle %d", h);
    if (num_images) *num_images = s->num_images;
    if (read_index) for (int i = 0; i < s->num_images - 1; i++) read_index[i] = s->read_index[i];
    if (write_index) *write_index = s->write_index;
    if (unit) for (int i = 0; i < s->num_images; i++) unit[i] = s->unit[i];
    return 0;
}

// storage image `image` (0..N-1) as a Buffer: GetReadImage(i) is image read_index[i]; bind it with cwa_sph2_bind_wave1d
extern "C" int cwa_stencil1d_image_buffer(cwa_ctx* ctx, cwa_stencil1d h, int image, cwa_buf* out)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s && out, "invalid stencil1d handle %d", h);
    CWA_CHECK(image >= 0 && image < s->num_images, "image index %d out of range", image);
    *out = s->image_buf[image];
    return 0;
}

extern "C" int cwa_stencil1d_read_image(cwa_ctx* ctx, cwa_stencil1d h, int image, float* host)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s && host, "invalid stencil1d handle %d", h);
    CWA_CHECK(image >= 0 && image < s->num_images, "image index %d out of range", image);
    CWA_CUDA(cudaMemcpyAsync(host, s->image[image], (size_t)s->w * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cwa_stencil1d_write_image(cwa_ctx* ctx, cwa_stencil1d h, int image, const float* host)
{
    DeviceGuard _dg(ctx);
    Stencil1dObj* s = get_s1d(ctx, h);
    CWA_CHECK(s && host, "invalid stencil1d handle %d", h);
    CWA_CHECK(image >= 0 && image < s->num_images, "image index %d out of range", image);
    CWA_CUDA(cudaMemcpyAsync(s->image[image], host, (size_t)s->w * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
