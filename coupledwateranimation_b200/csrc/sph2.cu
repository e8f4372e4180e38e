// sph2.cu -- 2-D Koschier-style WCSPH on the uniform grid (SphWave2D) for sm_100a.
//   SphKoschier2D_grid_cs.glsl / SphWaveKoschier2D_grid_cs.glsl:
//     mode 0 InitParticle            -> sph2_init_kernel
//     mode 1 ComputeDensityPressure  -> sph2_density_kernel
//     mode 2 ComputeForces+integrate -> sph2_forces_kernel
//   SphUgrid::Compute (StencilBuffer.cpp:150-179) -> cwa_sph2_compute
//
// The reference walks per-cell index lists and gathers 48-B structs through them for every
// neighbour.  Here the read buffer is gathered ONCE per pass into cell order (float4 reorder), the
// targets are processed in cell order and every neighbour row (cells that differ only in j) is a
// contiguous run of the cell-ordered arrays.  Particles outside the grid extents are never
// inserted (strict point_in_aabb) but are still updated as targets, exactly like the shader.
#include "internal.cuh"

namespace k2d {
// constants: SphWaveKoschier2D_grid_cs.glsl:116-136 (same values in SphKoschier2D_grid_cs.glsl)
__device__ constexpr float PARTICLE_RADIUS = 0.025f;
__device__ constexpr float PARTICLE_DIAM = 2.0f * PARTICLE_RADIUS;
__device__ constexpr float H = 4.0f * PARTICLE_RADIUS;
__device__ constexpr float HSQ = H * H;
__device__ constexpr float REST_DENS = 1000.0f;
__device__ constexpr float VISC = 0.05f;
__device__ constexpr float MASS = PARTICLE_DIAM * PARTICLE_DIAM * REST_DENS;
__device__ constexpr float DT = 0.002f;
__device__ constexpr float GAS_CONST = 35000.0f;
__device__ constexpr float M_PI_F = 3.14159265f;
__device__ constexpr float VIEW_HEIGHT = 2.0f * 4.8f;   // VIEW_WIDTH (also 9.6) is Sph2Params::view_width
__device__ constexpr float K2 = 40.0f / (7.0f * M_PI_F * HSQ);
}  // namespace k2d

struct Sph2Params {
    int variant;
    float time, bottom, psi;
    int init_width;
    float view_width;          // const VIEW_WIDTH = 2*4.8 in the shaders; a parameter so C2 can widen the tank
    const float4* wave1d;      // sampler1D wave_tex (RGBA32F), nullptr = unbound
    int wave1d_width;
};

struct P2 { float4 pos, vel, acc; };

__device__ __forceinline__ float k2_W_cubic(float r)                  // :233-246
{
    const float q = r / k2d::H;
    if (q >= 1.0f) return 0.0f;
    if (q <= 0.5f) return k2d::K2 * (6.0f * q * q * (q - 1.0f) + 1.0f);
    const float q1 = 1.0f - q;
    return k2d::K2 * 2.0f * q1 * q1 * q1;
}

__device__ __forceinline__ float k2_W_cubic_grad(float r)             // :248-262
{
    const float q = r / k2d::H;
    if (q >= 1.0f) return 0.0f;
    if (q <= 0.5f) return 6.0f * k2d::K2 * q * (3.0f * q - 2.0f) / k2d::H;
    const float q1 = 1.0f - q;
    return -6.0f * k2d::K2 * (q1 * q1) / k2d::H;
}

// boundary_sdf -> (nx, ny, sd, id)
__device__ __forceinline__ float4 k2_boundary_sdf(const Sph2Params& prm, float px, float py)
{
    using namespace k2d;
    float4 res = make_float4(0.f, 0.f, 0.f, 0.f);                     // (set by both branches below before the first opU)
    auto opU = [&](float nx, float ny, float d, float id) {
        if (!(res.z < d)) res = make_float4(nx, ny, d, id);           // (d1.z<d2.z) ? d1 : d2
    };
    if (prm.variant == CWA_SPH2_WAVE) {
        const float c0 = -0.5f * VIEW_HEIGHT + 15.0f * PARTICLE_RADIUS - PARTICLE_RADIUS + prm.bottom;
        res = make_float4(0.0f, 1.0f, (0.0f * px + 1.0f * py) + c0, 0.0f);
        opU(1.0f, 0.0f, (1.0f * px + 0.0f * py) + (-PARTICLE_RADIUS), 1.0f);
        opU(-1.0f, 0.0f, (-1.0f * px + 0.0f * py) + (prm.view_width - PARTICLE_RADIUS), 2.0f);
        opU(0.0f, -1.0f, (0.0f * px + -1.0f * py) + (VIEW_HEIGHT - PARTICLE_RADIUS), 3.0f);
    } else {
        res = make_float4(0.0f, 1.0f, (0.0f * px + 1.0f * py) + (-PARTICLE_RADIUS), 0.0f);
        opU(1.0f, 0.0f, (1.0f * px + 0.0f * py) + (-PARTICLE_RADIUS), 1.0f);
        opU(-1.0f, 0.0f, (-1.0f * px + 0.0f * py) + (prm.view_width - PARTICLE_RADIUS), 2.0f);
        const float cx = 0.5f * prm.view_width + 3.0f * cosf(prm.time), cy = 0.5f * VIEW_HEIGHT + 3.0f * sinf(prm.time);
        const float qx = px - cx, qy = py - cy;
        const float len = sqrtf(cwa_len2sq(qx, qy));
        opU(qx / len, qy / len, len - 1.0f, 3.0f);
    }
    return res;
}

// texture(wave_tex, coord) on a 1-D RGBA32F LINEAR/CLAMP_TO_EDGE texture; unbound -> (0,0,0,1)
__device__ __forceinline__ float4 k2_tex1d(const Sph2Params& prm, float s)
{
    if (prm.wave1d == nullptr) return make_float4(0.f, 0.f, 0.f, 1.f);
    const int W = prm.wave1d_width;
    const float u = s * (float)W - 0.5f;
    const float fu = floorf(u);
    const float a = u - fu;
    const int i0 = cwa_tex_index(fu, W), i1 = cwa_tex_index(fu + 1.0f, W);
    const float4 x0 = __ldg(prm.wave1d + i0), x1 = __ldg(prm.wave1d + i1);
    return make_float4(x0.x + a * (x1.x - x0.x), x0.y + a * (x1.y - x0.y), x0.z + a * (x1.z - x0.z), x0.w + a * (x1.w - x0.w));
}

// GetWaveNormalHeight :348-364 -> (n.x, n.y, 0, h), xvel
__device__ __forceinline__ float4 k2_wave_normal_height(const Sph2Params& prm, float x, float& xvel)
{
    const float coord = x / prm.view_width;
    const float size = (prm.wave1d != nullptr) ? (float)prm.wave1d_width : 1.0f;
    const float4 w = k2_tex1d(prm, coord);
    const float4 we = k2_tex1d(prm, coord - 1.0f / size);
    const float4 ww = k2_tex1d(prm, coord + 1.0f / size);
    const float nx = we.x - ww.x, ny = 1.0f;
    const float len = sqrtf(cwa_len2sq(nx, ny));
    xvel = w.y / w.x;
    return make_float4(nx / len, ny / len, 0.0f, w.x);
}

// InitParticle / init_grid
__global__ void __launch_bounds__(256) sph2_init_kernel(float4* __restrict__ out, int n, Sph2Params prm)
{
    using namespace k2d;
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    if (ix >= n) return;
    const int cols = prm.init_width, rows = n / cols;
    const int i = ix % cols, j = ix / cols;
    float px, py;
    if (prm.variant == CWA_SPH2_WAVE) {
        px = __fmul_rn(prm.view_width, __fdiv_rn((float)i, (float)cols));
        py = __fmul_rn(18.0f * H, __fdiv_rn((float)j, (float)rows));
        px = __fadd_rn(px, PARTICLE_RADIUS);
        py = __fadd_rn(py, 0.5f * VIEW_HEIGHT - 15.0f * PARTICLE_RADIUS + PARTICLE_RADIUS);
    } else {
        px = __fadd_rn(__fmul_rn(PARTICLE_DIAM, (float)i), __fmul_rn(0.1f / 6.0f, prm.view_width));
        py = __fadd_rn(__fmul_rn(PARTICLE_DIAM, (float)j), 0.1f / 6.0f * VIEW_HEIGHT);
    }
    out[(size_t)ix * 3 + 0] = make_float4(px, py, 0.0f, 1.0f);
    out[(size_t)ix * 3 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    out[(size_t)ix * 3 + 2] = make_float4(0.f, 0.f, 0.f, REST_DENS);
}

// gather the read buffer into cell order: slot s <- particle index_list[s]   (3 lanes / particle)
__global__ void __launch_bounds__(256)
sph2_reorder_kernel(const float4* __restrict__ aos, const int* __restrict__ index_list, const int* __restrict__ offset,
                    int num_cells, float4* __restrict__ posS, float4* __restrict__ velS, float4* __restrict__ accS)
{
    CWA_PDL_ENTER();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = __ldg(offset + num_cells);              // number of inserted particles
    const int s = t / 3, q = t - 3 * s;
    if (s >= m) return;
    const float4 v = __ldg(aos + (size_t)__ldg(index_list + s) * 3 + q);
    if (q == 0) posS[s] = v; else if (q == 1) velS[s] = v; else accS[s] = v;
}

// Canonical order + gather in one pass, as sph3_order_reorder_kernel does for the 3-D frame (thread per arrival slot): the particle that
// arrived at slot s is ranked among its cell peers by counting the smaller ids (ascending id inside a cell == the CPU twin), its id goes
// to index_list[rank slot] and its 48-byte record into the cell-ordered arrays; the record loads are issued before the rank chain.
__global__ void __launch_bounds__(256)
sph2_order_reorder_kernel(const float4* __restrict__ aos, const int* __restrict__ arrival, const int* __restrict__ cell_of,
                          const int* __restrict__ offset, int num_cells, int* __restrict__ index_list,
                          float4* __restrict__ posS, float4* __restrict__ velS, float4* __restrict__ accS)
{
    CWA_PDL_ENTER();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= __ldg(offset + num_cells)) return;           // inserted particles
    const int id = __ldg(arrival + s);
    const float4 r0 = __ldg(aos + (size_t)id * 3), r1 = __ldg(aos + (size_t)id * 3 + 1), r2 = __ldg(aos + (size_t)id * 3 + 2);
    const int c = __ldg(cell_of + id);
    const int b = __ldg(offset + c), e = __ldg(offset + c + 1);
    int smaller = 0;
#pragma unroll 4
    for (int q = b; q < e; q++) smaller += (__ldg(arrival + q) < id) ? 1 : 0;
    const int t = b + smaller;
    index_list[t] = id;
    posS[t] = r0; velS[t] = r1; accS[t] = r2;
}

// target enumeration shared by both passes: first the inserted particles in cell order, then the
// particles the grid rejected (cell_of == -1) in a second launch over the original order.
__device__ __forceinline__ bool sph2_target(int t, int n, int m, bool tail, const int* __restrict__ index_list,
                                            const int* __restrict__ cell_of, int& ix)
{
    if (!tail) { if (t >= m) return false; ix = __ldg(index_list + t); return true; }
    if (t >= n) return false;
    if (__ldg(cell_of + t) >= 0) return false;
    ix = t;
    return true;
}

// L lanes per target (a power of two <= 32): C2 has 65 536 targets with ~144 candidates each -- one thread per target leaves a B200 with
// 14 warps per SM, each on a serial chain of 144 dependent-latency steps.  The lanes of a target split every row of candidates and meet
// in a shuffle reduction over their own group (every exit before it is taken by all lanes of the group: they hold the same target).
template <int S2_LANES>
__device__ __forceinline__ unsigned s2_group_mask() { return (S2_LANES == 32 ? 0xffffffffu : ((1u << (S2_LANES & 31)) - 1u)) << ((threadIdx.x & 31u) & ~(unsigned)(S2_LANES - 1)); }

template <int S2_LANES>
__global__ void __launch_bounds__(128)
sph2_density_kernel(const float4* __restrict__ in, float4* __restrict__ out, int n, GridView g,
                    const int* __restrict__ offset, const int* __restrict__ index_list, const int* __restrict__ cell_of,
                    const float4* __restrict__ posS, Sph2Params prm, int main_blocks)
{
    CWA_PDL_ENTER();
    using namespace k2d;
    // one launch for both target lists: blocks [0, main_blocks) take the inserted particles in cell order, the rest the rejected ones
    const int tail = (int)blockIdx.x >= main_blocks;
    const int gt = (blockIdx.x - (tail ? main_blocks : 0)) * blockDim.x + threadIdx.x;
    const int t = gt / S2_LANES, sub = gt % S2_LANES;
    const int m = __ldg(offset + g.num_cells);
    int ix;
    if (!sph2_target(t, n, m, tail != 0, index_list, cell_of, ix)) return;
    P2 pi;
    pi.pos = __ldg(in + (size_t)ix * 3); pi.vel = __ldg(in + (size_t)ix * 3 + 1); pi.acc = __ldg(in + (size_t)ix * 3 + 2);
    const float PSI = (prm.psi < 0.0f) ? REST_DENS / (1.5f * K2) : prm.psi;

    if (prm.variant == CWA_SPH2_WAVE) {                                   // :400-415
        float xvel;
        (void)k2_wave_normal_height(prm, pi.pos.x, xvel);
        const float xv_thresh = 0.05f;
        if (pi.pos.w == 0.0f && fabsf(xvel) > xv_thresh) pi.pos.w = 1.0f;
        if (pi.pos.w == 1.0f && fabsf(xvel) < xv_thresh) pi.pos.w = 0.0f;
    }
    const float4 db = k2_boundary_sdf(prm, pi.pos.x, pi.pos.y);
    if (db.z < 0.0f) {                                                    // :419-423
        pi.pos.x -= 0.75f * db.z * db.x; pi.pos.y -= 0.75f * db.z * db.y;
        const float vn = fminf(0.0f, pi.vel.x * db.x + pi.vel.y * db.y);
        pi.vel.x -= 1.75f * vn * db.x; pi.vel.y -= 1.75f * vn * db.y;
    }
    float4* o = out + (size_t)ix * 3;
    if (pi.pos.y > VIEW_HEIGHT) { if (sub == 0) { o[0] = pi.pos; o[1] = pi.vel; o[2] = pi.acc; } return; }   // :427-431

    float rho = 0.0f;
    int i0, j0, i1, j1;
    cwa_cell2(g, pi.pos.x - H, pi.pos.y - H, i0, j0);                     // :436-439
    cwa_cell2(g, pi.pos.x + H, pi.pos.y + H, i1, j1);
    for (int i = i0; i <= i1; i++) {
        const int base = i * g.n[1];
        const int g0 = __ldg(offset + base + j0), g1 = __ldg(offset + base + j1 + 1);
        for (int q = g0 + sub; q < g1; q += S2_LANES) {
            const float4 pj = __ldg(posS + q);
            const float r2 = cwa_len2sq(pi.pos.x - pj.x, pi.pos.y - pj.y);
            if (r2 < HSQ) rho += k2_W_cubic(sqrtf(r2));                   // :456-459
        }
    }
    {
        const unsigned gm = s2_group_mask<S2_LANES>();
#pragma unroll
        for (int d = 1; d < S2_LANES; d <<= 1) rho += __shfl_xor_sync(gm, rho, d);
    }
    if (sub != 0) return;
    rho = MASS * rho;                                                     // :471
    if (db.z < H) rho += PSI * k2_W_cubic(fmaxf(0.0f, db.z + 0.0f * PARTICLE_RADIUS));
    rho = fmaxf(REST_DENS, rho);
    pi.acc.w = rho;
    const float ratio = rho / REST_DENS;
    pi.vel.w = GAS_CONST * (ratio * ratio * ratio - 1.0f);                // Tait, gamma = 3 :303-307
    o[0] = pi.pos; o[1] = pi.vel; o[2] = pi.acc;
}

template <int S2_LANES>
__global__ void __launch_bounds__(128)
sph2_forces_kernel(const float4* __restrict__ in, float4* __restrict__ out, int n, GridView g,
                   const int* __restrict__ offset, const int* __restrict__ index_list, const int* __restrict__ cell_of,
                   const float4* __restrict__ posS, const float4* __restrict__ velS, const float4* __restrict__ accS,
                   Sph2Params prm, int main_blocks)
{
    CWA_PDL_ENTER();
    using namespace k2d;
    const int tail = (int)blockIdx.x >= main_blocks;
    const int gt = (blockIdx.x - (tail ? main_blocks : 0)) * blockDim.x + threadIdx.x;
    const int t = gt / S2_LANES, sub = gt % S2_LANES;
    const int m = __ldg(offset + g.num_cells);
    int ix;
    if (!sph2_target(t, n, m, tail != 0, index_list, cell_of, ix)) return;
    const int self_slot = tail ? -1 : t;
    P2 pi;
    pi.pos = __ldg(in + (size_t)ix * 3); pi.vel = __ldg(in + (size_t)ix * 3 + 1); pi.acc = __ldg(in + (size_t)ix * 3 + 2);
    float4* o = out + (size_t)ix * 3;
    if (prm.variant == CWA_SPH2_WAVE && pi.pos.w == 0.0f) { if (sub == 0) { o[0] = pi.pos; o[1] = pi.vel; o[2] = pi.acc; } return; }   // :494-498
    const float PSI = (prm.psi < 0.0f) ? REST_DENS / (1.5f * K2) : prm.psi;
    const float rho_i = pi.acc.w;
    const float c_visc = -VISC * 8.0f * MASS, c_press = MASS;
    const float acc_press_i = pi.vel.w / (rho_i * rho_i);
    float apx = 0.f, apy = 0.f, avx = 0.f, avy = 0.f;
    int i0, j0, i1, j1;
    cwa_cell2(g, pi.pos.x - H, pi.pos.y - H, i0, j0);
    cwa_cell2(g, pi.pos.x + H, pi.pos.y + H, i1, j1);
    for (int i = i0; i <= i1; i++) {
        const int base = i * g.n[1];
        const int g0 = __ldg(offset + base + j0), g1 = __ldg(offset + base + j1 + 1);
        for (int q = g0 + sub; q < g1; q += S2_LANES) {
            if (q == self_slot) continue;                                  // jx != ix
            const float4 pj = __ldg(posS + q);
            const float rx = pi.pos.x - pj.x, ry = pi.pos.y - pj.y;
            const float r = sqrtf(cwa_len2sq(rx, ry));
            if (r < H) {
                const float4 vj = __ldg(velS + q);
                const float rho_j = __ldg(accS + q).w;
                const float Wgrad = k2_W_cubic_grad(r);
                const float ux = rx / r, uy = ry / r;
                const float s = (acc_press_i + vj.w / (rho_j * rho_j)) * Wgrad;
                apx -= s * ux; apy -= s * uy;
                const float vx = pi.vel.x - vj.x, vy = pi.vel.y - vj.y;
                const float tt = 1.0f / rho_j * (vx * rx + vy * ry) / (r * r + 0.01f * HSQ) * Wgrad;
                avx -= tt * ux; avy -= tt * uy;
            }
        }
    }
    {
        const unsigned gm = s2_group_mask<S2_LANES>();
#pragma unroll
        for (int d = 1; d < S2_LANES; d <<= 1) {
            apx += __shfl_xor_sync(gm, apx, d); apy += __shfl_xor_sync(gm, apy, d);
            avx += __shfl_xor_sync(gm, avx, d); avy += __shfl_xor_sync(gm, avy, d);
        }
    }
    if (sub != 0) return;
    avx *= c_visc; avy *= c_visc;
    apx *= c_press; apy *= c_press;
    const float4 db = k2_boundary_sdf(prm, pi.pos.x, pi.pos.y);
    if (db.z < H) {
        const float Wgrad = k2_W_cubic_grad(fmaxf(0.0f, db.z + 0.0f * PARTICLE_RADIUS));
        const float s = PSI * acc_press_i * Wgrad;
        apx += s * (-db.x); apy += s * (-db.y);
    }
    float atx = 0.f, aty = 0.f;
    if (prm.variant == CWA_SPH2_WAVE) {                                    // :565-603
        float vx;
        const float4 wave = k2_wave_normal_height(prm, pi.pos.x, vx);
        const float hh = wave.w;
        const float dh = hh - VIEW_HEIGHT / 2.0f;
        const float wave_mask = cwa_smoothstep(0.0f, 0.2f, fabsf(vx));
        float dy = pi.pos.y - hh;
        dy -= 0.5f * dh;
        if (ix % 5 < 4) aty -= wave_mask * 500000.0f * cwa_smoothstep(0.0f, 10.0f, dy);
    }
    pi.acc.x = apx + avx + 0.0f + atx;
    pi.acc.y = apy + avy + (-9.8f) + aty;
    pi.vel.x += DT * pi.acc.x; pi.vel.y += DT * pi.acc.y;
    pi.pos.x += DT * pi.vel.x; pi.pos.y += DT * pi.vel.y;
    o[0] = pi.pos; o[1] = pi.vel; o[2] = pi.acc;
}

// ---------------------------------------------------------------------------------------------
// host object: SphUgrid
// ---------------------------------------------------------------------------------------------
static Sph2Params sph2_params(cwa_ctx* ctx, Sph2Obj* s)
{
    Sph2Params p;
    p.variant = s->variant; p.time = s->time; p.bottom = s->bottom; p.psi = s->psi; p.init_width = s->init_width;
    p.view_width = s->view_width;
    BufferObj* wb = get_buffer(ctx, s->wave1d);
    p.wave1d = wb ? (const float4*)wb->ptr : nullptr;
    p.wave1d_width = wb ? s->wave1d_width : 0;
    return p;
}

static void sph2_pingpong(Sph2Obj* s) { std::swap(s->read_index, s->write_index); }
static void sph2_drop_graph(Sph2Obj* s);   // StencilBuffer::PingPong :32-36

extern "C" int cwa_sph2_reinit(cwa_ctx* ctx, cwa_sph2 h)
{
    DeviceGuard _dg(ctx);
    Sph2Obj* s = get_sph2(ctx, h);
    CWA_CHECK(s, "invalid sph2 handle %d", h);
    BufferObj* wb = get_buffer(ctx, s->buffer[s->write_index]);
    CWA_CHECK(wb, "sph2: buffer vanished");
    CWA_CHECK(s->init_width > 0 && s->n / s->init_width > 0, "sph2: init lattice width %d does not fit %d particles", s->init_width, s->n);
    { KScope k(ctx, KID_OTHER);
      sph2_init_kernel<<<ceil_div(s->n, 256), 256, 0, ctx->stream>>>((float4*)wb->ptr, s->n, sph2_params(ctx, s)); }   // MODE_INIT :43-48
    CWA_CUDA(cudaGetLastError());
    sph2_pingpong(s);                                                        // :50
    return 0;
}

extern "C" int cwa_sph2_create(cwa_ctx* ctx, int n, int variant, cwa_grid grid, cwa_sph2* out)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && out, "null argument");
    *out = -1;
    CWA_CHECK(n > 0, "cwa_sph2_create: n must be positive");
    CWA_CHECK(variant == CWA_SPH2_KOSCHIER || variant == CWA_SPH2_WAVE, "cwa_sph2_create: unknown variant %d", variant);
    GridObj* g = get_grid(ctx, grid);
    CWA_CHECK(g && g->dim == 2, "cwa_sph2_create: grid handle %d is not a 2-D grid", grid);
    CWA_CHECK(g->max_particles >= n, "cwa_sph2_create: grid capacity %d < %d particles", g->max_particles, n);
    Sph2Obj s;
    s.live = true; s.n = n; s.variant = variant; s.grid = grid;
    s.init_width = (variant == CWA_SPH2_WAVE) ? 128 : 32;
    for (int i = 0; i < 2; i++) CWA_TRY(cwa_buffer_create(ctx, (size_t)n * sizeof(cwa_particle2d), nullptr, &s.buffer[i]));
    CWA_CUDA(cudaMalloc(&s.posS, (size_t)n * 16));
    CWA_CUDA(cudaMalloc(&s.velS, (size_t)n * 16));
    CWA_CUDA(cudaMalloc(&s.accS, (size_t)n * 16));
    ctx->sph2s.push_back(s);
    *out = (int)ctx->sph2s.size() - 1;
    return cwa_sph2_reinit(ctx, *out);                                       // StencilBuffer::Init ends with Reinit :29
}

extern "C" int cwa_sph2_destroy(cwa_ctx* ctx, cwa_sph2 h)
{
    DeviceGuard _dg(ctx);
    Sph2Obj* s = get_sph2(ctx, h);
    CWA_CHECK(s, "invalid sph2 handle %d", h);
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    sph2_drop_graph(s);
    cudaFree(s->posS); cudaFree(s->velS); cudaFree(s->accS);
    for (int i = 0; i < 2; i++) cwa_buffer_destroy(ctx, s->buffer[i]);
    s->live = false;
    return 0;
}

extern "C" int cwa_sph2_set_substeps(cwa_ctx* ctx, cwa_sph2 h, int substeps)
{
    DeviceGuard _dg(ctx);
    Sph2Obj* s = get_sph2(ctx, h);
    CWA_CHECK(s && substeps >= 0, "invalid sph2 handle %d or substeps", h);
    s->substeps = substeps;
    return 0;
}

extern "C" int cwa_sph2_set_uniforms(cwa_ctx* ctx, cwa_sph2 h, float time, float bottom, float psi, int init_width)
{
    DeviceGuard _dg(ctx);
    Sph2Obj* s = get_sph2(ctx, h);
    CWA_CHECK(s, "invalid sph2 handle %d", h);
    s->time = time; s->bottom = bottom; s->psi = psi;
    if (init_width > 0) s->init_width = init_width;
    return 0;
}

extern "C" int cwa_sph2_set_view_width(cwa_ctx* ctx, cwa_sph2 h, float view_width)
{
    DeviceGuard _dg(ctx);
    Sph2Obj* s = get_sph2(ctx, h);
    CWA_CHECK(s && view_width > 0.0f, "invalid sph2 handle %d or view width", h);
    s->view_width = view_width;
    return 0;
}

extern "C" int cwa_sph2_bind_wave1d(cwa_ctx* ctx, cwa_sph2 h, cwa_buf rgba, int width)
{
    DeviceGuard _dg(ctx);
    Sph2Obj* s = get_sph2(ctx, h);
    CWA_CHECK(s, "invalid sph2 handle %d", h);
    if (rgba == -1) { s->wave1d = -1; s->wave1d_width = 0; return 0; }
    BufferObj* b = get_buffer(ctx, rgba);
    CWA_CHECK(b && width >= 1 && (size_t)width * 16 <= b->bytes, "cwa_sph2_bind_wave1d: buffer smaller than %d RGBA32F texels", width);
    s->wave1d = rgba; s->wave1d_width = width;
    return 0;
}

// one frame of SphUgrid::Compute (StencilBuffer.cpp:150-179): substeps x (grid build, density pass, forces pass)
static int sph2_frame(cwa_ctx* ctx, Sph2Obj* s, GridObj* g, const Sph2Params& prm, int lanes)
{
    const int n = s->n;
    const int blocks = ceil_div((long long)n * lanes, 128);
    for (int sub = 0; sub < s->substeps; sub++) {
        const float4* rd = (const float4*)get_buffer(ctx, s->buffer[s->read_index])->ptr;
        float4* wr = (float4*)get_buffer(ctx, s->buffer[s->write_index])->ptr;
        GridBuildOpts opts;
        opts.canonical_order = false;                                    // ranked and gathered in one pass below
        CWA_TRY(grid_build_internal(ctx, g, rd, 48, n, opts));           // mGrid.CollisionQuery() :163-164
        if (n > 0) { KScope k(ctx, KID_REORDER);
          cwa_launch(ctx, PDL_GRID2, sph2_order_reorder_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, rd, g->arrival, g->cell_of, g->offset, g->view.num_cells, g->index_list, s->posS, s->velS, s->accS); }
        {                                                                // mode 1 :169-176 (inserted particles + the rejected ones: one launch)
            KScope k(ctx, KID_DENSITY);
#define CWA_S2_DENS(L) cwa_launch(ctx, PDL_GRID2, sph2_density_kernel<L>, dim3(2 * blocks), dim3(128), 0, rd, wr, n, g->view, g->offset, g->index_list, g->cell_of, s->posS, prm, blocks)
            if (lanes == 4) CWA_S2_DENS(4); else if (lanes == 16) CWA_S2_DENS(16); else if (lanes == 32) CWA_S2_DENS(32); else CWA_S2_DENS(8);
#undef CWA_S2_DENS
        }
        sph2_pingpong(s);
        rd = (const float4*)get_buffer(ctx, s->buffer[s->read_index])->ptr;
        wr = (float4*)get_buffer(ctx, s->buffer[s->write_index])->ptr;
        // mode 2 reads the density output through the SAME (now stale) grid lists
        { KScope k(ctx, KID_REORDER);
          cwa_launch(ctx, PDL_GRID2, sph2_reorder_kernel, dim3(ceil_div((long long)n * 3, 256)), dim3(256), 0, rd, g->index_list, g->offset, g->view.num_cells, s->posS, s->velS, s->accS); }
        {
            KScope k(ctx, KID_FORCE);
#define CWA_S2_FORCE(L) cwa_launch(ctx, PDL_GRID2, sph2_forces_kernel<L>, dim3(2 * blocks), dim3(128), 0, rd, wr, n, g->view, g->offset, g->index_list, g->cell_of, s->posS, s->velS, s->accS, prm, blocks)
            if (lanes == 4) CWA_S2_FORCE(4); else if (lanes == 16) CWA_S2_FORCE(16); else if (lanes == 32) CWA_S2_FORCE(32); else CWA_S2_FORCE(8);
#undef CWA_S2_FORCE
        }
        sph2_pingpong(s);
    }
    CWA_CUDA(cudaGetLastError());
    return 0;
}

static void sph2_drop_graph(Sph2Obj* s)
{
    if (s->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)s->graph_exec);
    s->graph_exec = nullptr; s->graph_nodes = 0;
    memset(s->graph_key, 0, sizeof(s->graph_key));
}

extern "C" int cwa_sph2_compute(cwa_ctx* ctx, cwa_sph2 h, int nframes)
{
    DeviceGuard _dg(ctx);
    Sph2Obj* s = get_sph2(ctx, h);
    CWA_CHECK(s, "invalid sph2 handle %d", h);
    GridObj* g = get_grid(ctx, s->grid);
    CWA_CHECK(g, "sph2: grid vanished");
    const Sph2Params prm = sph2_params(ctx, s);
    // lanes per target of the neighbour kernels (tuning CWA_S2_LANES: 4, 8, 16 or 32)
    static const int lanes_env = [] { const char* e = getenv("CWA_S2_LANES"); const int v = e ? atoi(e) : 8; return (v == 4 || v == 8 || v == 16 || v == 32) ? v : 8; }();
    const int lanes = lanes_env;
    if (ctx->tune.graph < 0) { const char* e = getenv("CWA_GRAPH"); ctx->tune.graph = (e && atoi(e) == 0) ? 0 : 1; }
    for (int f = 0; f < nframes; f++) {
        // what a frame depends on besides the contents of the buffers: 2 x substeps ping-pongs leave the roles where they were, so while
        // nothing below changes every frame is the same launch sequence
        long long key[12] = {};
        key[0] = (long long)(intptr_t)get_buffer(ctx, s->buffer[s->read_index])->ptr; key[1] = (long long)(intptr_t)get_buffer(ctx, s->buffer[s->write_index])->ptr;
        key[2] = s->n; key[3] = s->substeps; key[4] = (long long)(intptr_t)g->counter; key[5] = lanes; key[6] = prm.variant;
        memcpy(&key[7], &prm.time, 4); memcpy((char*)&key[7] + 4, &prm.bottom, 4); memcpy(&key[8], &prm.psi, 4); memcpy((char*)&key[8] + 4, &prm.view_width, 4);
        key[9] = prm.init_width; key[10] = (long long)(intptr_t)prm.wave1d; key[11] = prm.wave1d_width;
        const bool use_graph = ctx->tune.graph != 0 && !ctx->profiling && s->n > 0;
        if (use_graph && s->graph_exec != nullptr && memcmp(key, s->graph_key, sizeof(key)) == 0) {
            CWA_CUDA(cudaGraphLaunch((cudaGraphExec_t)s->graph_exec, ctx->stream));
            ctx->launches += s->graph_nodes;
            continue;
        }
        if (use_graph && memcmp(key, s->last_key, sizeof(key)) == 0) {   // the second identical frame in a row: worth a capture
            sph2_drop_graph(s);
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const unsigned long long l0 = ctx->launches;
                const int rc = sph2_frame(ctx, s, g, prm, lanes);
                cudaGraph_t graph = nullptr;
                const cudaError_t ee = cudaStreamEndCapture(ctx->stream, &graph);
                cudaGraphExec_t exec = nullptr;
                if (rc == 0 && ee == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                    s->graph_exec = exec; s->graph_nodes = (unsigned)(ctx->launches - l0);
                    memcpy(s->graph_key, key, sizeof(key));
                }
                if (graph) cudaGraphDestroy(graph);
                (void)cudaGetLastError();
                ctx->launches = l0;
                if (rc != 0) return rc;
                if (s->graph_exec != nullptr) {                          // the captured frame has not run yet
                    CWA_CUDA(cudaGraphLaunch((cudaGraphExec_t)s->graph_exec, ctx->stream));
                    ctx->launches += s->graph_nodes;
                    continue;
                }
            } else {
                (void)cudaGetLastError();
            }
        }
        memcpy(s->last_key, key, sizeof(key));
        CWA_TRY(sph2_frame(ctx, s, g, prm, lanes));
    }
    return 0;
}

extern "C" int cwa_sph2_read(cwa_ctx* ctx, cwa_sph2 h, cwa_particle2d* host)
{
    DeviceGuard _dg(ctx);
    Sph2Obj* s = get_sph2(ctx, h);
    CWA_CHECK(s && host, "invalid sph2 handle %d", h);
    return cwa_buffer_read(ctx, s->buffer[s->read_index], 0, (size_t)s->n * sizeof(cwa_particle2d), host);
}

extern "C" int cwa_sph2_write(cwa_ctx* ctx, cwa_sph2 h, const cwa_particle2d* host)
{
    DeviceGuard _dg(ctx);
    Sph2Obj* s = get_sph2(ctx, h);
    CWA_CHECK(s && host, "invalid sph2 handle %d", h);
    return cwa_buffer_sub_data(ctx, s->buffer[s->read_index], 0, (size_t)s->n * sizeof(cwa_particle2d), host);
}

extern "C" int cwa_sph2_read_buffer(cwa_ctx* ctx, cwa_sph2 h, cwa_buf* out)
{
    DeviceGuard _dg(ctx);
    Sph2Obj* s = get_sph2(ctx, h);
    CWA_CHECK(s && out, "invalid sph2 handle %d", h);
    *out = s->buffer[s->read_index];
    return 0;
}
