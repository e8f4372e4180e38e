// wave.cu -- triple-buffered 2-D wave-equation stencil for sm_100a.
//   replaces  wave_comp.glsl:51-205 / Wave2DSimp/Wave2D_cs.glsl:37-103 (kernels) and
//             StencilImage2DTripleBuffered.cpp:4-95 (+ ImageTexture SwapUnits) (host object).
//
// Fast path (scalar field, W % 4 == 0): persistent CTAs walk 128x16 tiles; each tile's u^{t-1}
// halo box (136x18) and u^{t-2} box (128x16) are fetched by TMA (cp.async.bulk.tensor.2d) into a
// 3-stage shared-memory ring guarded by mbarriers; every thread produces 4 consecutive cells and
// stores them with one 128-bit store.  Clamp-to-edge is applied when the halo is READ (TMA
// zero-fills out-of-bounds elements, which is not edge replication).
// The arithmetic is written with explicit round-to-nearest intrinsics in the GLSL's association
// order, so a step is bit-identical to the CPU oracle.
#include "internal.cuh"

#include <cudaTypedefs.h>

constexpr int WT_W = 128;               // tile width  (cells)
#ifndef CWA_WT_H
#define CWA_WT_H 16
#endif
#ifndef CWA_WT_STAGES
#define CWA_WT_STAGES 3
#endif
#ifndef CWA_WT_THREADS
#define CWA_WT_THREADS 128
#endif
#ifndef CWA_WT_CTAS_PER_SM
#define CWA_WT_CTAS_PER_SM 4
#endif
constexpr int WT_H = CWA_WT_H;          // tile height (cells)
constexpr int WT_HALO_W = WT_W + 8;     // halo box: 4 cells of left pad keep float4 alignment
constexpr int WT_HALO_H = WT_H + 2;
constexpr int WT_STAGES = CWA_WT_STAGES;
constexpr int WT_THREADS = CWA_WT_THREADS;
constexpr int WT_WARPS = WT_THREADS / 32;
static_assert(WT_H % WT_WARPS == 0, "tile rows must divide evenly among the warps");
constexpr int WT_HALO_BYTES = WT_HALO_W * WT_HALO_H * 4;                 // 9792
constexpr int WT_HALO_BYTES_PAD = ((WT_HALO_BYTES + 127) / 128) * 128;   // 9856
constexpr int WT_CORE_BYTES = WT_W * WT_H * 4;                           // 8192
constexpr int WT_STAGE_BYTES = WT_HALO_BYTES_PAD + WT_CORE_BYTES;        // 18048
constexpr int WT_SMEM_BYTES = WT_STAGES * WT_STAGE_BYTES + 128;          // + alignment slack

// ---------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier + TMA)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// one cell of EvolveWave (wave_comp.glsl:171-187 / Wave2D_cs.glsl:78-84), oracle-identical order
// ---------------------------------------------------------------------------------------------
struct WaveCoef { float kc, lambda, k1, atten; int wake; int mid_x; };

__device__ __forceinline__ WaveCoef wave_coef(const float4 a, int variant, int W)
{
    WaveCoef c;
    c.lambda = a.x; c.atten = a.y;
    c.kc = __fsub_rn(__fsub_rn(2.0f, __fmul_rn(4.0f, a.x)), a.z);   // (2 - 4*lambda - beta)
    c.k1 = __fsub_rn(1.0f, a.z);                                    // (1 - beta)
    c.wake = (variant == CWA_WAVE_COUPLED) && (a.w > 0.0f) && (a.w < 1.0f);
    c.mid_x = W / 2;                                                // CoordOnLine :159-169
    return c;
}

__device__ __forceinline__ float wave_cell(const WaveCoef& k, float c0, float n0, float s0, float e0, float w0,
                                           float c1, bool red_channel, int gx)
{
    float sum = __fadd_rn(__fadd_rn(__fadd_rn(n0, s0), e0), w0);
    float v = __fsub_rn(__fadd_rn(__fmul_rn(k.kc, c0), __fmul_rn(k.lambda, sum)), __fmul_rn(k.k1, c1));
    v = __fmul_rn(v, k.atten);
    if (red_channel && k.wake && v > 0.0001f && gx == k.mid_x) v = __fadd_rn(v, 0.001f);   // :177-184
    return v;
}

// ---------------------------------------------------------------------------------------------
// TMA kernel (scalar field)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WT_THREADS)
wave_evolve_tma_kernel(const __grid_constant__ CUtensorMap tm_halo,   // u^{t-1}, box 136x18
                       const __grid_constant__ CUtensorMap tm_core,   // u^{t-2}, box 128x16
                       float* __restrict__ out, int W, int H, const float4* __restrict__ attr,
                       int variant, int tiles_x, int num_tiles)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_bar[WT_STAGES];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WT_STAGES; s++) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const WaveCoef k = wave_coef(__ldg(attr), variant, W);
    const int first = blockIdx.x;
    const int my_count = (first < num_tiles) ? (num_tiles - first + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    auto issue = [&](int it) {
        const int t = first + it * (int)gridDim.x;
        const int x0 = (t % tiles_x) * WT_W, y0 = (t / tiles_x) * WT_H;
        const int s = it % WT_STAGES;
        unsigned char* st = smem + (size_t)s * WT_STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], (uint32_t)(WT_HALO_BYTES + WT_CORE_BYTES));
        tma_load_2d(st, &tm_halo, x0 - 4, y0 - 1, &full_bar[s]);
        tma_load_2d(st + WT_HALO_BYTES_PAD, &tm_core, x0, y0, &full_bar[s]);
    };

    if (tid == 0) {
        const int pre = my_count < WT_STAGES ? my_count : WT_STAGES;
        for (int it = 0; it < pre; it++) issue(it);
    }

    for (int it = 0; it < my_count; it++) {
        const int t = first + it * (int)gridDim.x;
        const int x0 = (t % tiles_x) * WT_W, y0 = (t / tiles_x) * WT_H;
        const int s = it % WT_STAGES;
        const float* halo = reinterpret_cast<const float*>(smem + (size_t)s * WT_STAGE_BYTES);
        const float* core = reinterpret_cast<const float*>(smem + (size_t)s * WT_STAGE_BYTES + WT_HALO_BYTES_PAD);
        mbar_wait(&full_bar[s], (uint32_t)((it / WT_STAGES) & 1));

        const int gx = x0 + 4 * lane;
#pragma unroll
        for (int rr = 0; rr < WT_H / WT_WARPS; rr++) {
            const int r = warp + WT_WARPS * rr;
            const int gy = y0 + r;
            if (gy < H && gx < W) {
                const int hr = r + 1;
                const int hn = (gy == H - 1) ? hr : hr + 1;      // clamp_coord(+1) :189-193
                const int hs = (gy == 0) ? hr : hr - 1;          // clamp_coord(-1)
                const float4 c0 = *reinterpret_cast<const float4*>(halo + hr * WT_HALO_W + 4 + 4 * lane);
                const float4 n0 = *reinterpret_cast<const float4*>(halo + hn * WT_HALO_W + 4 + 4 * lane);
                const float4 s0 = *reinterpret_cast<const float4*>(halo + hs * WT_HALO_W + 4 + 4 * lane);
                const float4 c1 = *reinterpret_cast<const float4*>(core + r * WT_W + 4 * lane);
                const float wl = (gx == 0) ? c0.x : halo[hr * WT_HALO_W + 3 + 4 * lane];
                const float er = (gx + 3 == W - 1) ? c0.w : halo[hr * WT_HALO_W + 8 + 4 * lane];
                float4 o;
                o.x = wave_cell(k, c0.x, n0.x, s0.x, c0.y, wl, c1.x, true, gx + 0);
                o.y = wave_cell(k, c0.y, n0.y, s0.y, c0.z, c0.x, c1.y, true, gx + 1);
                o.z = wave_cell(k, c0.z, n0.z, s0.z, c0.w, c0.y, c1.z, true, gx + 2);
                o.w = wave_cell(k, c0.w, n0.w, s0.w, er, c0.z, c1.w, true, gx + 3);
                *reinterpret_cast<float4*>(out + (size_t)gy * W + gx) = o;
            }
        }
        __syncthreads();                                         // every thread is done with stage s
        if (tid == 0 && it + WT_STAGES < my_count) issue(it + WT_STAGES);
    }
}

// ---------------------------------------------------------------------------------------------
// generic kernel: RGBA32F-compatible images (ch = 4) and widths that are not a multiple of 4
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
wave_evolve_generic_kernel(const float* __restrict__ u0, const float* __restrict__ u1, float* __restrict__ out,
                           int W, int H, int ch, const float4* __restrict__ attr, int variant)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const WaveCoef k = wave_coef(__ldg(attr), variant, W);
    const int xe = min(x + 1, W - 1), xw = max(x - 1, 0), yn = min(y + 1, H - 1), ys = max(y - 1, 0);
    for (int c = 0; c < ch; c++) {
        const float c1 = __ldg(u1 + ((size_t)y * W + x) * ch + c);
        const float c0 = __ldg(u0 + ((size_t)y * W + x) * ch + c);
        const float n0 = __ldg(u0 + ((size_t)yn * W + x) * ch + c);
        const float s0 = __ldg(u0 + ((size_t)ys * W + x) * ch + c);
        const float e0 = __ldg(u0 + ((size_t)y * W + xe) * ch + c);
        const float w0 = __ldg(u0 + ((size_t)y * W + xw) * ch + c);
        out[((size_t)y * W + x) * ch + c] = wave_cell(k, c0, n0, s0, e0, w0, c1, c == 0, x);
    }
}

// InitWave: wave_comp.glsl:82-140 / Wave2D_cs.glsl:66-76
__global__ void __launch_bounds__(256)
wave_init_kernel(float* __restrict__ out, int W, int H_local, int ch, const float4* __restrict__ attr, int variant,
                 int row0, int H)
{
    // (x, y) are GLOBAL image coordinates; a row block stores row y at local row y - row0
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int yl = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || yl >= H_local) return;
    const int y = yl + row0;
    const float type = __ldg(attr).w;
    int c0x, c0y, c1x, c1y, c2x = 0, c2y = 0;
    bool third = false;
    float peak, e0;
    const float fw = (float)W, fh = (float)H;
    if (variant == CWA_WAVE_SIMP) {
        c0x = (int)__fmul_rn(0.25f, fw); c0y = (int)__fmul_rn(0.25f, fh);
        c1x = (int)__fmul_rn(0.75f, fw); c1y = (int)__fmul_rn(0.75f, fh);
        peak = 0.5f; e0 = 3.0f;
    } else {
        e0 = 5.0f;
        if (type == 1.0f) {
            c0x = (int)__fmul_rn(0.25f, fw); c0y = (int)__fmul_rn(0.25f, fh);
            c1x = (int)__fmul_rn(0.75f, fw); c1y = (int)__fmul_rn(0.75f, fh);
            peak = 0.5f;
        } else if (type == 0.0f) {
            c0x = (int)__fmul_rn(0.25f, fw); c0y = H;
            c2x = (int)__fmul_rn(0.5f, fw);  c2y = H;
            c1x = (int)__fmul_rn(0.75f, fw); c1y = H;
            peak = 1.0f; third = true;
        } else {
            c0x = (int)__fmul_rn(0.5f, fw); c0y = (int)__fmul_rn(0.1f, fh);
            c1x = c0x; c1y = c0y; c2x = c0x; c2y = c0y;
            peak = 0.1f; third = true;
        }
    }
    auto dist = [&](int cx, int cy) {
        const float dx = __fsub_rn((float)x, (float)cx), dy = __fsub_rn((float)y, (float)cy);
        return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    };
    float d = fminf(dist(c0x, c0y), dist(c1x, c1y));
    if (third) d = fminf(d, dist(c2x, c2y));
    const float v = __fmul_rn(peak, cwa_smoothstep(e0, 0.0f, d));
    float* o = out + ((size_t)yl * W + x) * ch;
    o[0] = v;
    for (int c = 1; c < ch; c++) o[c] = 0.0f;
}

// InitFromImage: texelFetch(uInitImage, coord*ivec2(2,1)) (wave_comp.glsl:76-80) or coord (Wave2D_cs.glsl:60-64)
__global__ void __launch_bounds__(256)
wave_init_from_texture_kernel(float* __restrict__ out, int W, int H, int ch, const float* __restrict__ tex,
                              int tw, int th, int xmul)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const int sx = x * xmul, sy = y;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);          // out-of-range texelFetch: robust-access zero
    if (sx < tw && sy < th) v = *reinterpret_cast<const float4*>(tex + ((size_t)sy * tw + sx) * 4);
    float* o = out + ((size_t)y * W + x) * ch;
    o[0] = v.x;
    if (ch == 4) { o[1] = v.y; o[2] = v.z; o[3] = v.w; }
}

// Transposed sampling copy: out[x * H + y] = in[y * W + x] (32 x 32 tiles through shared memory, both sides coalesced)
__global__ void __launch_bounds__(256)
wave_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int W, int H)
{
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int x = x0 + tx, y = y0 + r;
        if (x < W && y < H) tile[r][tx] = __ldg(in + (size_t)y * W + x);
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int x = x0 + r, y = y0 + tx;
        if (x < W && y < H) out[(size_t)x * H + y] = tile[tx][r];
    }
}

// ---------------------------------------------------------------------------------------------
// host object
// ---------------------------------------------------------------------------------------------
static bool wave_transpose_on(cwa_ctx* ctx)
{
    if (ctx->tune.wave_transpose < 0) { const char* e = getenv("CWA_WAVE_TRANSPOSE"); ctx->tune.wave_transpose = e ? (atoi(e) != 0) : 1; }
    return ctx->tune.wave_transpose != 0;
}

void wave_touch_buffer(cwa_ctx* ctx, cwa_buf b, bool raw)
{
    for (auto& w : ctx->waves) {
        if (!w.live) continue;
        for (int i = 0; i < 3; i++)
            if (w.image_buf[i] == b) { w.version[i]++; if (raw) w.raw_exposed = true; }
    }
}

// The SPH passes sample ONE image per frame.  When the field is local, scalar and large enough for the access pattern to
// matter, they read a transposed copy (TexView::tdata); it is rebuilt only when the sampled image changed -- once every three
// frames under the as-shipped binding schedule.  Same texels, same arithmetic: results are bit-identical.
int wave_sampling_copy(cwa_ctx* ctx, cwa_wave h, int image, TexView* tex)
{
    WaveObj* w = get_wave(ctx, h);
    if (!w || image < 0 || image >= 3 || !wave_transpose_on(ctx) || (w->raw_exposed && !w->explicit_touch)) return 0;
    // any scalar view (whole field or row block) that is 32-bit indexable and large enough for the access pattern to matter
    if (tex->ch != 1 || tex->data != w->image[image] || (long long)w->w * w->h < 256 * 256 || (long long)w->w * w->h >= (1ll << 30)) return 0;
    if (w->imageT == nullptr) CWA_CUDA(cudaMalloc(&w->imageT, (size_t)w->w * w->h * 4));
    if (w->imageT_of != image || w->imageT_version != w->version[image]) {
        KScope k(ctx, KID_OTHER);
        wave_transpose_kernel<<<dim3(ceil_div(w->w, 32), ceil_div(w->h, 32)), 256, 0, ctx->stream>>>(w->image[image], w->imageT, w->w, w->h);
        CWA_CUDA(cudaGetLastError());
        w->imageT_of = image; w->imageT_version = w->version[image];
    }
    tex->tdata = w->imageT; tex->tsi = w->h; tex->tsj = 1;
    return 0;
}

int wave_image_with_unit(const WaveObj* w, int u) { for (int i = 0; i < 3; i++) if (w->unit[i] == u) return i; return -1; }
static int image_with_unit(const WaveObj* w, int u)
{
    for (int i = 0; i < 3; i++) if (w->unit[i] == u) return i;
    return -1;
}

static const float4* wave_attr_ptr(cwa_ctx* ctx, WaveObj* w)
{
    if (w->variant == CWA_WAVE_SIMP) return reinterpret_cast<const float4*>(w->simp_params);
    return reinterpret_cast<const float4*>(current_params(ctx).wave);      // WaveUniforms.attributes
}

static int wave_make_tmaps(cwa_ctx* ctx, WaveObj* w)
{
    w->tma_ok = false;
    if (w->ch != 1 || (w->w % 4) != 0 || ctx->encode_tiled == nullptr) return 0;
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ctx->encode_tiled);
    for (int i = 0; i < 3; i++) {
        cuuint64_t gdim[2] = {(cuuint64_t)w->w, (cuuint64_t)w->h};
        cuuint64_t gstride[1] = {(cuuint64_t)w->w * 4};
        cuuint32_t estr[2] = {1, 1};
        cuuint32_t box_h[2] = {WT_HALO_W, WT_HALO_H};
        cuuint32_t box_c[2] = {WT_W, WT_H};
        CUresult r1 = encode(&w->tmap_halo[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w->image[i], gdim, gstride, box_h, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = encode(&w->tmap_core[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w->image[i], gdim, gstride, box_c, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
            cwa_set_error("cuTensorMapEncodeTiled failed (%d, %d) for a %dx%d field", (int)r1, (int)r2, w->w, w->h);
            return -2;
        }
    }
    w->tma_ok = true;
    return 0;
}

static int wave_alloc_images(cwa_ctx* ctx, WaveObj* w)
{
    const size_t bytes = (size_t)w->w * w->h * w->ch * 4;
    for (int i = 0; i < 3; i++) {
        CWA_CUDA(cudaMalloc(&w->image[i], bytes));
        CWA_CUDA(cudaMemsetAsync(w->image[i], 0, bytes, ctx->stream));   // glTextureStorage2D: treat as zero
        w->version[i]++;
        if (w->image_buf[i] >= 0 && get_buffer(ctx, w->image_buf[i])) {
            BufferObj* b = get_buffer(ctx, w->image_buf[i]);
            b->ptr = w->image[i]; b->bytes = bytes;
        } else {
            w->image_buf[i] = new_buffer(ctx, w->image[i], bytes, false);
        }
    }
    return wave_make_tmaps(ctx, w);
}

// PingPong, StencilImage2DTripleBuffered.cpp:33-40
static void wave_pingpong(WaveObj* w)
{
    std::swap(w->write_index, w->read_index[0]);
    std::swap(w->read_index[0], w->read_index[1]);
    std::swap(w->unit[w->write_index], w->unit[w->read_index[0]]);      // SwapUnits, ImageTexture.cpp:100-103
    std::swap(w->unit[w->read_index[0]], w->unit[w->read_index[1]]);
}

void wave_pingpong_internal(WaveObj* w) { wave_pingpong(w); }

int wave_dispatch_mode(cwa_ctx* ctx, WaveObj* w, int mode)
{
    const int in0 = image_with_unit(w, 0), in1 = image_with_unit(w, 1), outi = image_with_unit(w, 2);
    const dim3 gblock(256), ggrid(ceil_div(w->w, 32), ceil_div(w->h, 8));
    const float4* attr = wave_attr_ptr(ctx, w);
    if (mode == CWA_MODE_TEST) return 0;                           // MODE_TEST does nothing (wave_comp.glsl:71-72)
    CWA_CHECK(mode == CWA_MODE_INIT || mode == CWA_MODE_EVOLVE, "wave dispatch: unsupported uMode %d", mode);
    w->version[outi]++;
    KScope kscope(ctx, mode == CWA_MODE_EVOLVE ? KID_WAVE : KID_OTHER);
    if (mode == CWA_MODE_INIT) {
        wave_init_kernel<<<ggrid, gblock, 0, ctx->stream>>>(w->image[outi], w->w, w->h, w->ch, attr, w->variant, w->row0, w->h_global);
    } else if (mode == CWA_MODE_EVOLVE) {
        if (w->tma_ok) {
            CWA_TRY(ensure_dynamic_smem(ctx, wave_evolve_tma_kernel, WT_SMEM_BYTES));
            const int tiles_x = ceil_div(w->w, WT_W), tiles_y = ceil_div(w->h, WT_H);
            const int num_tiles = tiles_x * tiles_y;
            int grid = ctx->sm_count * CWA_WT_CTAS_PER_SM;
            if (grid > num_tiles) grid = num_tiles;
            wave_evolve_tma_kernel<<<grid, WT_THREADS, WT_SMEM_BYTES, ctx->stream>>>(
                w->tmap_halo[in0], w->tmap_core[in1], w->image[outi], w->w, w->h, attr, w->variant, tiles_x, num_tiles);
        } else {
            wave_evolve_generic_kernel<<<ggrid, gblock, 0, ctx->stream>>>(w->image[in0], w->image[in1], w->image[outi],
                                                                         w->w, w->h, w->ch, attr, w->variant);
        }
    }
    CWA_CUDA(cudaGetLastError());
    return 0;
}

int wave_step_internal(cwa_ctx* ctx, WaveObj* w)
{
    CWA_TRY(wave_dispatch_mode(ctx, w, CWA_MODE_EVOLVE));
    wave_pingpong(w);
    return 0;
}

TexView wave_tex_view(cwa_ctx* ctx, cwa_wave h, int image)
{
    TexView t{nullptr, 1, 1, 1, 0, 1, nullptr, nullptr, 1, 1};
    WaveObj* w = get_wave(ctx, h);
    if (w && image >= 0 && image < 3) {
        t.data = w->image[image]; t.w = w->w; t.h = w->h; t.ch = w->ch;
        t.row0 = w->row0; t.h_global = w->h_global; t.last_row = w->last_row[image];
        t.tdata = t.data; t.tsi = w->ch; t.tsj = w->w * w->ch;      // row-major as stored (.r channel of texel (i, j))
    }
    return t;
}

extern "C" int cwa_wave_create(cwa_ctx* ctx, int width, int height, int channels, int variant, cwa_wave* out)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && out, "null argument");
    *out = -1;
    CWA_CHECK(width >= 1 && height >= 1, "cwa_wave_create: bad size %dx%d", width, height);
    CWA_CHECK(channels == 1 || channels == 4, "cwa_wave_create: channels must be 1 (scalar) or 4 (RGBA32F)");
    CWA_CHECK(variant == CWA_WAVE_COUPLED || variant == CWA_WAVE_SIMP, "cwa_wave_create: unknown shader variant %d", variant);
    WaveObj w;
    w.live = true; w.w = width; w.h = height; w.ch = channels; w.variant = variant;
    w.row0 = 0; w.h_global = height;
    const float simp[4] = {0.01f, 0.9995f, 0.001f, 1.0f};            // Wave2D_cs.glsl:17-19
    CWA_CUDA(cudaMalloc(&w.simp_params, 16));
    CWA_CUDA(cudaMemcpyAsync(w.simp_params, simp, 16, cudaMemcpyHostToDevice, ctx->stream));
    CWA_TRY(wave_alloc_images(ctx, &w));
    ctx->waves.push_back(w);
    *out = (int)ctx->waves.size() - 1;
    return cwa_wave_reinit(ctx, *out);                               // Init() ends with Reinit() :30
}

extern "C" int cwa_wave_destroy(cwa_ctx* ctx, cwa_wave h)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 2; i++) CWA_CUDA(cudaStreamSynchronize(ctx->side_stream[i]));
    for (int i = 0; i < 3; i++) {
        cudaFree(w->image[i]);
        if (BufferObj* b = get_buffer(ctx, w->image_buf[i])) b->live = false;
    }
    cudaFree(w->simp_params);
    cudaFree(w->imageT); w->imageT = nullptr;
    for (int i = 0; i < 3; i++) {
        if (w->last_row[i]) cudaFree(w->last_row[i]);
        if (BufferObj* b = get_buffer(ctx, w->last_row_buf[i])) b->live = false;
    }
    w->live = false;
    return 0;
}

extern "C" int cwa_wave_reinit(cwa_ctx* ctx, cwa_wave h)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    for (int i = 0; i < 2; i++) {                                    // Reinit :49-58: two INIT passes
        CWA_TRY(wave_dispatch_mode(ctx, w, CWA_MODE_INIT));
        wave_pingpong(w);
    }
    return 0;
}

extern "C" int cwa_wave_reinit_from_texture(cwa_ctx* ctx, cwa_wave h, const float* rgba, int tw, int th)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w && rgba, "invalid wave handle %d or null texture", h);
    CWA_CHECK(tw >= 1 && th >= 1, "bad texture size");
    float* dtex = nullptr;
    const size_t bytes = (size_t)tw * th * 16;
    CWA_CUDA(cudaMalloc(&dtex, bytes));
    CWA_CUDA(cudaMemcpyAsync(dtex, rgba, bytes, cudaMemcpyHostToDevice, ctx->stream));
    const int outi = image_with_unit(w, 2);
    w->version[outi]++;
    const dim3 gblock(256), ggrid(ceil_div(w->w, 32), ceil_div(w->h, 8));
    { KScope k(ctx, KID_OTHER);
      wave_init_from_texture_kernel<<<ggrid, gblock, 0, ctx->stream>>>(w->image[outi], w->w, w->h, w->ch, dtex, tw, th,
                                                                      w->variant == CWA_WAVE_COUPLED ? 2 : 1); }
    CWA_CUDA(cudaGetLastError());
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    CWA_CUDA(cudaFree(dtex));
    wave_pingpong(w);                                                // ReinitFromTexture :76
    return 0;
}

extern "C" int cwa_wave_compute(cwa_ctx* ctx, cwa_wave h, int nsteps)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    if (!w->evolve) return 0;                                        // Compute() :81
    for (int s = 0; s < nsteps; s++) CWA_TRY(wave_step_internal(ctx, w));
    return 0;
}

extern "C" int cwa_wave_pingpong(cwa_ctx* ctx, cwa_wave h)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    wave_pingpong(w);
    return 0;
}

extern "C" int cwa_wave_set_evolve(cwa_ctx* ctx, cwa_wave h, int evolve)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    w->evolve = evolve != 0;
    return 0;
}

extern "C" int cwa_wave_set_params(cwa_ctx* ctx, cwa_wave h, float lambda, float atten, float beta)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    const float p[4] = {lambda, atten, beta, 1.0f};
    CWA_CUDA(cudaMemcpyAsync(w->simp_params, p, 16, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

extern "C" int cwa_wave_resize(cwa_ctx* ctx, cwa_wave h, int nw, int nh)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    CWA_CHECK(nw >= 1 && nh >= 1, "cwa_wave_resize: bad size");
    CWA_CHECK(w->row0 == 0 && w->h == w->h_global, "cwa_wave_resize: not supported on a row-block wave object");
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 2; i++) CWA_CUDA(cudaStreamSynchronize(ctx->side_stream[i]));   // a pipelined stencil step / grid clear may still run there
    w->tma_ok = false;
    for (int i = 0; i < 3; i++) {
        cudaFree(w->image[i]); w->image[i] = nullptr;
        if (BufferObj* b = get_buffer(ctx, w->image_buf[i])) { b->ptr = nullptr; b->bytes = 0; }    // until wave_alloc_images re-points them
    }
    cudaFree(w->imageT); w->imageT = nullptr; w->imageT_of = -1;
    w->w = nw; w->h = nh; w->h_global = nh;
    return wave_alloc_images(ctx, w);                                // ImageTexture::Resize: new storage, contents cleared
}

extern "C" int cwa_wave_state(cwa_ctx* ctx, cwa_wave h, int read_index[2], int* write_index, int unit[3], int* tex_unit0_image)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    if (read_index) { read_index[0] = w->read_index[0]; read_index[1] = w->read_index[1]; }
    if (write_index) *write_index = w->write_index;
    if (unit) for (int i = 0; i < 3; i++) unit[i] = w->unit[i];
    if (tex_unit0_image) *tex_unit0_image = w->tex_unit0;
    return 0;
}

extern "C" int cwa_wave_bind_texture_unit(cwa_ctx* ctx, cwa_wave h)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    const int ri0 = w->read_index[0];
    if (w->unit[ri0] == 0) w->tex_unit0 = ri0;       // BindTextureUnit binds at the image's own mUnit (ImageTexture.cpp:42-45)
    return 0;
}

static int resolve_image(WaveObj* w, int image)
{
    return (image >= 0 && image < 3) ? image : -1;
}

extern "C" int cwa_wave_read_image(cwa_ctx* ctx, cwa_wave h, int image, float* host)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w && host, "invalid wave handle %d", h);
    const int i = resolve_image(w, image);
    CWA_CHECK(i >= 0, "image index %d out of range", image);
    CWA_CUDA(cudaMemcpyAsync(host, w->image[i], (size_t)w->w * w->h * w->ch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cwa_wave_read_image_async(cwa_ctx* ctx, cwa_wave h, int image, float* host)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w && host, "invalid wave handle %d", h);
    const int i = resolve_image(w, image);
    CWA_CHECK(i >= 0, "image index %d out of range", image);
    CWA_CUDA(cudaMemcpyAsync(host, w->image[i], (size_t)w->w * w->h * w->ch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

extern "C" int cwa_wave_write_image(cwa_ctx* ctx, cwa_wave h, int image, const float* host)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w && host, "invalid wave handle %d", h);
    const int i = resolve_image(w, image);
    CWA_CHECK(i >= 0, "image index %d out of range", image);
    CWA_CUDA(cudaMemcpyAsync(w->image[i], host, (size_t)w->w * w->h * w->ch * 4, cudaMemcpyHostToDevice, ctx->stream));
    w->version[i]++;
    return 0;
}

// An application that writes an image through a raw device pointer (cwa_buffer_device_ptr: CUDA-GL interop maps, NCCL receives
// into halo rows) reports it here; without such reports the library stops keeping derived copies of that object's images.
extern "C" int cwa_wave_mark_written(cwa_ctx* ctx, cwa_wave h, int image)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    const int i = resolve_image(w, image);
    CWA_CHECK(i >= 0, "image index %d out of range", image);
    w->explicit_touch = true;
    w->version[i]++;
    return 0;
}

extern "C" int cwa_wave_role_image(cwa_ctx* ctx, cwa_wave h, int role, int* image)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w && image, "invalid wave handle %d", h);
    CWA_CHECK(role >= 0 && role < 3, "role %d out of range", role);
    *image = image_with_unit(w, role);
    return 0;
}

extern "C" int cwa_wave_image_buffer(cwa_ctx* ctx, cwa_wave h, int image, cwa_buf* out)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w && out, "invalid wave handle %d", h);
    CWA_CHECK(image >= 0 && image < 3, "image index %d out of range", image);
    *out = w->image_buf[image];
    return 0;
}

extern "C" int cwa_wave_size(cwa_ctx* ctx, cwa_wave h, int* width, int* height, int* channels)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w, "invalid wave handle %d", h);
    if (width) *width = w->w;
    if (height) *height = w->h;
    if (channels) *channels = w->ch;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// row-block decomposition (multi-GPU): this rank stores global rows [row0, row0+rows)
// of a field that is h_global rows tall (its owned block plus the sampling halos).  EVOLVE clamps
// at the local array border; the caller overwrites the halo rows with the neighbours' owned rows
// after every step, so only genuinely global borders keep the clamp.
// ---------------------------------------------------------------------------------------------
extern "C" int cwa_wave_create_block(cwa_ctx* ctx, int width, int h_global, int row0, int rows, int channels, int variant, cwa_wave* out)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && out, "null argument");
    *out = -1;
    CWA_CHECK(width >= 1 && h_global >= 1 && rows >= 1 && row0 >= 0 && row0 + rows <= h_global,
              "cwa_wave_create_block: rows [%d,%d) outside a field of %d rows", row0, row0 + rows, h_global);
    CWA_CHECK(channels == 1 || channels == 4, "cwa_wave_create_block: channels must be 1 or 4");
    CWA_CHECK(variant == CWA_WAVE_COUPLED || variant == CWA_WAVE_SIMP, "cwa_wave_create_block: unknown shader variant %d", variant);
    WaveObj w;
    w.live = true; w.w = width; w.h = rows; w.ch = channels; w.variant = variant;
    w.row0 = row0; w.h_global = h_global;
    const float simp[4] = {0.01f, 0.9995f, 0.001f, 1.0f};
    CWA_CUDA(cudaMalloc(&w.simp_params, 16));
    CWA_CUDA(cudaMemcpyAsync(w.simp_params, simp, 16, cudaMemcpyHostToDevice, ctx->stream));
    CWA_TRY(wave_alloc_images(ctx, &w));
    for (int i = 0; i < 3; i++) {
        const size_t bytes = (size_t)width * channels * 4;
        CWA_CUDA(cudaMalloc(&w.last_row[i], bytes));
        CWA_CUDA(cudaMemsetAsync(w.last_row[i], 0, bytes, ctx->stream));
        w.last_row_buf[i] = new_buffer(ctx, w.last_row[i], bytes, false);
    }
    ctx->waves.push_back(w);
    *out = (int)ctx->waves.size() - 1;
    return cwa_wave_reinit(ctx, *out);
}

extern "C" int cwa_wave_last_row_buffer(cwa_ctx* ctx, cwa_wave h, int image, cwa_buf* out)
{
    DeviceGuard _dg(ctx);
    WaveObj* w = get_wave(ctx, h);
    CWA_CHECK(w && out, "invalid wave handle %d", h);
    CWA_CHECK(image >= 0 && image < 3, "image index %d out of range", image);
    CWA_CHECK(w->last_row_buf[image] >= 0, "cwa_wave_last_row_buffer: not a row-block wave object");
    *out = w->last_row_buf[image];
    return 0;
}
