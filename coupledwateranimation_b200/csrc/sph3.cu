// sph3.cu -- the three 3-D SPH passes of CoupledWaterAnimation for sm_100a.
//   rho_pres_comp.glsl:49-81   -> sph3_density_*   (poly6 density, EOS pressure, wave coupling)
//   force_comp.glsl:62-139     -> sph3_force_* + force_epilogue (spiky pressure, viscosity, crest
//                                 rule, torque, wave drag / normal force, gravity)
//   integrate_comp.glsl:56-178 -> sph3_integrate_* (symplectic Euler, foam rule, surface clamp, box)
//
// Grid mode (the north-star path): particles are gathered into cell order (coalesced float4 reorder + x | y | z coordinate streams);
// the neighbour rows of a target -- same (i, j), consecutive k -- are contiguous runs of the cell-ordered arrays.  Three families of
// neighbour kernels share that snapshot (cwa_set_tuning "nb_config", all parity-tested):
//   8 (default)  row-mask kernels: one thread per target walks its narrowed 3 x 3 rows in ONE loop, two candidates per trip on packed
//                FP32 (FFMA2), accepted candidates recorded as one 32-bit mask per row; the force pass walks the set bits.
//   7            round 1's index-list kernels (row loop nest, one store per accepted pair).
//   0..6         "lanes" kernels: a CTA owns P targets x L lanes, neighbour-row windows staged in shared memory with 1-D bulk async copies
//                (cp.async.bulk, TMA engine, mbarrier completion), lanes interleave over consecutive candidates, warp-shuffle reduction.
// Clump targets (hundreds of candidates) are finished in place by the density pass and one warp each by the force pass.
// All-pairs mode reproduces the shipped O(N^2) loops with shared-memory tiles, one CTA per SM, several targets per thread.
#include "internal.cuh"
#include <math_constants.h>
#include <climits>
#include <cstdlib>

#define CWA_PI 3.141592741f   // rho_pres_comp.glsl:8

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s3_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s3_mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s3_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void s3_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s3_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s3_mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s3_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void s3_mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "S3_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra S3_WAIT_DONE;\n"
        "bra S3_WAIT_LOOP;\n"
        "S3_WAIT_DONE:\n"
        "}\n" ::"r"(s3_smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (bytes % 16 == 0, both addresses 16-B aligned)
__device__ __forceinline__ void s3_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s3_smem_u32(dst)), "l"(src), "r"(bytes), "r"(s3_smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// constants derived from the parameter blocks (prepared on the device once per dispatch)
// ---------------------------------------------------------------------------------------------
struct Sph3Const {
    float h, h2, accept_r2;        // accept pair  <=>  r2 <= accept_r2  <=>  sqrt_rn(r2) < h
    float mass, rho0, visc_coeff;
    float poly6;                   // mass*315 / (64*PI*h^9)
    float spiky, laplacian;        // -20/(PI*h^6), +20/(PI*h^6)
    float gas, radius, dt, gravity_y, damping, crest, foam_speed, uv_scale, uv_scale_z, torque;
    float wave_type;
    float upper[3], lower[3];
};

// largest float t with sqrt_rn(t) < h: the pair test "length(delta) < h" of the shaders becomes an
// exact compare on the squared distance (correctly rounded sqrt is monotonic).
__device__ __forceinline__ float accept_threshold(float h)
{
    if (!(h > 0.0f)) return -1.0f;
    float t = __fmul_rn(h, h);
    for (int it = 0; it < 4 && __fsqrt_rn(t) >= h; it++) t = __uint_as_float(__float_as_uint(t) - 1u);
    for (int it = 0; it < 4; it++) {
        float u = __uint_as_float(__float_as_uint(t) + 1u);
        if (__fsqrt_rn(u) < h) t = u; else break;
    }
    return t;
}

__device__ __forceinline__ Sph3Const load_consts(const ParamPtrs& prm)
{
    Sph3Const c;
    const cwa_constants_uniform cu = *prm.constants;
    const cwa_sim_constants sc = *prm.sim;
    c.mass = cu.mass; c.rho0 = cu.resting_rho; c.visc_coeff = cu.visc;
    c.h = __fmul_rn(cu.smoothing_coeff, sc.particle_radius);        // rho_pres_comp.glsl:54
    c.h2 = __fmul_rn(c.h, c.h);
    c.accept_r2 = accept_threshold(c.h);
    const float h2 = c.h2, h4 = __fmul_rn(h2, h2), h8 = __fmul_rn(h4, h4);
    const float h9 = __fmul_rn(h8, c.h), h6 = __fmul_rn(h4, h2);
    c.poly6 = __fdiv_rn(__fmul_rn(cu.mass, 315.0f), __fmul_rn(__fmul_rn(64.0f, CWA_PI), h9));
    c.spiky = __fdiv_rn(-20.0f, __fmul_rn(CWA_PI, h6));            // force_comp.glsl:68
    c.laplacian = -c.spiky;                                        // :69
    c.gas = sc.gas_const; c.radius = sc.particle_radius; c.dt = sc.dt; c.gravity_y = sc.gravity_y;
    c.damping = sc.damping; c.crest = sc.crest_threshold; c.foam_speed = sc.foam_speed; c.uv_scale = sc.uv_scale;
    c.uv_scale_z = (sc.uv_scale_z != 0.0f) ? sc.uv_scale_z : sc.uv_scale;
    c.torque = (sc.torque_coeff != 0.0f) ? sc.torque_coeff : 0.25f;
    c.wave_type = prm.wave->attributes[3];
#pragma unroll
    for (int a = 0; a < 3; a++) { c.upper[a] = prm.boundary->upper[a]; c.lower[a] = prm.boundary->lower[a]; }
    return c;
}

__global__ void sph3_prepare_kernel(ParamPtrs prm, Sph3Const* __restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = load_consts(prm);
}

// ---------------------------------------------------------------------------------------------
// pair terms
// ---------------------------------------------------------------------------------------------
// rho_pres_comp.glsl:62-67
__device__ __forceinline__ void pair_density(float accept_r2, float h2, float poly6, float px, float py, float pz,
                                             const float4 q, float& rho)
{
    const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
    const float r2 = cwa_len3sq(dx, dy, dz);
    if (r2 <= accept_r2) {
        const float d = h2 - r2;
        rho = fmaf(poly6, d * d * d, rho);
    }
}

// force_comp.glsl:79-87 (caller excludes j == i and rejects r >= h)
__device__ __forceinline__ void pair_force(const Sph3Const& c, float px, float py, float pz, float prs_i,
                                           float vx, float vy, float vz, const float4 qa, const float4 qb,
                                           float& fpx, float& fpy, float& fpz, float& fvx, float& fvy, float& fvz)
{
    const float dx = px - qa.x, dy = py - qa.y, dz = pz - qa.z;
    const float r2 = cwa_len3sq(dx, dy, dz);
    // MUFU.RSQ / MUFU.RCP without the denormal fix-up sequences of rsqrtf() / __frcp_rn() (14 of the 50 instructions of a pair):
    // r2 of an accepted pair lies in (0, h^2] ~ 1e-4 and rho >= rho0, nowhere near the denormal range, and 2^-22 relative error is
    // two orders of magnitude inside the 1e-4 tolerance.  r2 == 0 (coincident particles) still gives inf * 0 = NaN.
    float inv_r, inv_rho;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_r) : "f"(r2));
    const float r = r2 * inv_r;                                  // r = 0 -> NaN like normalize(0)
    const float hr = c.h - r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv_rho) : "f"(qb.w));
    // pres_force -= mass*(p_i+p_j)/(2 rho_j) * spiky * (h-r)^2 * normalize(delta)
    const float a = c.mass * (prs_i + qa.w) * (0.5f * inv_rho) * c.spiky * (hr * hr) * inv_r;
    fpx = fmaf(-a, dx, fpx); fpy = fmaf(-a, dy, fpy); fpz = fmaf(-a, dz, fpz);
    // visc_force += mass*(v_j - v_i)/rho_j * laplacian * (h-r)
    const float b = c.mass * inv_rho * c.laplacian * hr;
    fvx = fmaf(b, qb.x - vx, fvx); fvy = fmaf(b, qb.y - vy, fvy); fvz = fmaf(b, qb.z - vz, fvz);
}

// ---------------------------------------------------------------------------------------------
// per-particle epilogues (shared by grid and all-pairs kernels)
// ---------------------------------------------------------------------------------------------
// rho_pres_comp.glsl:70-80
template <bool LOCAL = false>
__device__ __forceinline__ void density_epilogue(const Sph3Const& c, const TexView& tex, float px, float pz, float rho,
                                                 float& rho_out, float& prs_out)
{
    float pressure = fmaxf(c.gas * (rho - c.rho0), 0.0f);
    const float height = cwa_tex_sample<LOCAL>(tex, c.uv_scale * px, c.uv_scale_z * pz);
    const float wave_force = height * rho;
    pressure += wave_force;
    rho += __fdiv_rn(wave_force, c.gas * c.radius);
    rho_out = fmaxf(c.rho0, rho);
    prs_out = pressure;
}

// force_comp.glsl:90-114.  fprev = particles[i].force as stored by the previous frame;
// (fp*, fv*) = the neighbour sums of :74-88.
template <bool LOCAL = false>
__device__ __forceinline__ float4 force_epilogue(const Sph3Const& c, const TexView& tex, float px, float py, float pz,
                                                 float vx, float vy, float vz, float rho_i, float4 fprev,
                                                 float fpx, float fpy, float fpz, float fvx, float fvy, float fvz)
{
    if (py > c.crest) {                                            // :91-95
        // x / 0.25 == x * 4 bit for bit (power-of-two scaling is exact, overflow and subnormals included)
        fprev.x = __fmul_rn(fprev.x, 4.0f); fprev.y = __fmul_rn(fprev.y, 4.0f);
        fprev.z = __fmul_rn(fprev.z, 4.0f); fprev.w = __fmul_rn(fprev.w, 4.0f);
        fvx *= 0.5f; fvy *= 0.5f; fvz *= 0.5f;
    }
    fvx *= c.visc_coeff; fvy *= c.visc_coeff; fvz *= c.visc_coeff; // :97
    const float cu = px * c.uv_scale, cv = pz * c.uv_scale_z;        // :99
    const float height = cwa_tex_sample<LOCAL>(tex, cu, cv);            // :100
    // torque = 0.25 * cross(pos, force_prev.xyz)  :103-104
    const float tx = c.torque * (py * fprev.z - pz * fprev.y);
    const float ty = c.torque * (pz * fprev.x - px * fprev.z);
    const float tz = c.torque * (px * fprev.y - py * fprev.x);
    // WaveVelocity :117-128
    const float hX = cwa_tex_sample<LOCAL>(tex, cu + 0.01f, cv);
    const float hY = cwa_tex_sample<LOCAL>(tex, cu, cv + 0.01f);
    const float wvx = __fdiv_rn(hX - height, c.dt), wvy = __fdiv_rn(hY - height, c.dt), wvz = __fdiv_rn(hX - hY, 0.01f);
    const float dgx = -0.25f * (vx - wvx), dgy = -0.25f * (vy - wvy), dgz = -0.25f * (vz - wvz);   // :107-108
    // WaveNormal :131-139: cross(dy,dx) = (dx.z, dy.z, -1)
    const float nx = cwa_tex_sample<LOCAL>(tex, cu + 1.0f, cv) - height;
    const float ny = cwa_tex_sample<LOCAL>(tex, cu, cv + 1.0f) - height;
    const float wfx = -height * nx * 0.5f, wfy = -height * ny * 0.5f, wfz = -height * -1.0f * 0.5f;  // :110
    const float gy = rho_i * c.gravity_y;                                                         // :113
    float4 f;
    f.x = fpx + fvx + 0.0f + tx + dgx + wfx;                       // :114
    f.y = fpy + fvy + gy + ty + dgy + wfy;
    f.z = fpz + fvz + 0.0f + tz + dgz + wfz;
    f.w = fprev.w;
    return f;
}

// integrate_comp.glsl:56-92 + CheckBoundary :135-178
template <bool LOCAL = false>
__device__ __forceinline__ void integrate_particle(const Sph3Const& c, const TexView& tex, float4& pos, float4& vel,
                                                   float4& force, float& rho, float& prs)
{
    const float ax = __fdiv_rn(force.x, rho), ay = __fdiv_rn(force.y, rho), az = __fdiv_rn(force.z, rho);
    float nvx = vel.x + c.dt * ax, nvy = vel.y + c.dt * ay, nvz = vel.z + c.dt * az;
    float npx = pos.x + c.dt * nvx, npy = pos.y + c.dt * nvy, npz = pos.z + c.dt * nvz;
    const float damp = 1.0f - c.damping * c.dt;
    nvx *= damp; nvy *= damp; nvz *= damp;
    if (__fsqrt_rn(cwa_len3sq(nvx, nvy, nvz)) > c.foam_speed) {    // :69-76
        force.x *= 0.5f; force.y *= 0.5f; force.z *= 0.5f; force.w *= 0.5f;
        rho *= 0.1f; prs *= 0.25f;
        nvx *= 0.1f; nvy *= 0.1f; nvz *= 0.1f;
    }
    const float th = cwa_tex_sample<LOCAL>(tex, npx * c.uv_scale, npz * c.uv_scale_z);   // :79
    if (npy < th) npy = th - c.radius;                                            // :80-83
    const float D = c.damping;
    if (npx < c.lower[0]) { npx = c.lower[0]; nvx *= -D; } else if (npx > c.upper[0]) { npx = c.upper[0]; nvx *= -D; }
    if (npy < c.lower[1]) { npy = c.lower[1]; nvy *= -D; } else if (npy > c.upper[1]) { npy = c.upper[1]; nvy *= -D; }
    if (npz < c.lower[2]) { npz = c.lower[2]; nvz *= -D; } else if (npz > c.upper[2]) { npz = c.upper[2]; nvz *= -D; }
    if (c.wave_type > 0.0f && c.wave_type < 1.0f) {                               // :170-177
        if (npy > c.upper[1] + 0.1f) { npy = c.upper[1] + 0.1f; nvy *= -D; }
    }
    pos.x = npx; pos.y = npy; pos.z = npz;
    vel.x = nvx; vel.y = nvy; vel.z = nvz;
}

// ---------------------------------------------------------------------------------------------
// (3) coalesced float4 reorder into cell order
// ---------------------------------------------------------------------------------------------
// 4 lanes per particle: lane q moves vec4 q of the 64-B struct, so the gather reads whole 64-B
// records and the four cell-ordered arrays are written with 128-B contiguous runs per 8 slots.
__global__ void __launch_bounds__(256)
sph3_reorder_kernel(const float4* __restrict__ aos, const int* __restrict__ index_list, const int* __restrict__ count,
                    float4* __restrict__ posS, float4* __restrict__ velS, float4* __restrict__ forceS,
                    float4* __restrict__ miscS, int* __restrict__ heavy_cnt, float* __restrict__ xyzS, size_t xyz_stride)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) { heavy_cnt[0] = 0; heavy_cnt[1] = 0; }   // clump queues of the density / force passes that follow this snapshot
    const int s = t >> 2, q = t & 3;
    const int n = __ldg(count);                  // inserted particles (NaN positions are left out)
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < n) {
        const int i = __ldg(index_list + s);
        v = __ldg(aos + (size_t)i * 4 + q);
    }
    // lane 3 (extras) needs pos.w and vel.w from lanes 0 and 1 of its quad
    const unsigned lane = threadIdx.x & 31u, quad = lane & ~3u;
    const float pw = __shfl_sync(0xffffffffu, v.w, quad + 0);
    const float vw = __shfl_sync(0xffffffffu, v.w, quad + 1);
    if (s >= n) return;
    if (q == 0) { posS[s] = v; xyzS[s] = v.x; xyzS[xyz_stride + s] = v.y; xyzS[2 * xyz_stride + s] = v.z; }
    else if (q == 1) velS[s] = v;
    else if (q == 2) forceS[s] = v;
    else miscS[s] = make_float4(pw, vw, v.z, v.w);      // (pos.w, vel.w, extras.z, extras.w)
}

// Canonical ordering + reorder in one pass (thread per arrival slot): the particle that arrived at slot s
// is ranked among its cell peers by counting the smaller ids (ascending id inside a cell == the CPU twin,
// SURVEY F7), its id goes to index_list[rank slot] and its 64-B record into the cell-ordered arrays.  The
// record loads are issued first, so they overlap the rank chain (cell id -> offsets -> arrival list).
__global__ void __launch_bounds__(256)
sph3_order_reorder_kernel(const float4* __restrict__ aos, const int* __restrict__ arrival, const int* __restrict__ cell_of,
                          const int* __restrict__ offset, const int* __restrict__ count, int* __restrict__ index_list,
                          float4* __restrict__ posS, float4* __restrict__ velS, float4* __restrict__ forceS,
                          float4* __restrict__ miscS, int* __restrict__ heavy_cnt, float* __restrict__ xyzS, size_t xyz_stride)
{
    CWA_PDL_ENTER();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) { heavy_cnt[0] = 0; heavy_cnt[1] = 0; }   // clump queues of the density / force passes that follow this snapshot
    if (s >= __ldg(count)) return;                 // inserted particles (NaN positions are left out)
    const int id = __ldg(arrival + s);
    const float4* rec = aos + (size_t)id * 4;
    const f4x2 h0 = cwa_ldg256(rec), h1 = cwa_ldg256(rec + 2);   // the 64-byte record: two sectors, two requests
    const float4 r0 = h0.a, r1 = h0.b, r2 = h1.a, r3 = h1.b;
    const int c = __ldg(cell_of + id);
    const int b = __ldg(offset + c), e = __ldg(offset + c + 1);
    int smaller = 0;
#pragma unroll 4
    for (int q = b; q < e; q++) smaller += (__ldg(arrival + q) < id) ? 1 : 0;
    const int t = b + smaller;
    index_list[t] = id;
    posS[t] = r0; velS[t] = r1; forceS[t] = r2;
    miscS[t] = make_float4(r0.w, r1.w, r3.z, r3.w);          // (pos.w, vel.w, extras.z, extras.w)
    xyzS[t] = r0.x; xyzS[xyz_stride + t] = r0.y; xyzS[2 * xyz_stride + t] = r0.z;   // coordinate streams of the density pass
}

// ---------------------------------------------------------------------------------------------
// grid-mode neighbour kernels
// ---------------------------------------------------------------------------------------------
constexpr int NB_WINDOWS = 9;

struct BlockWindows {
    int lo[NB_WINDOWS];      // first cell-ordered slot covered by the window
    int hi[NB_WINDOWS];      // one past the last slot
    int soff[NB_WINDOWS];    // start inside the staging buffer (slots); -1 = not staged
};

// Windows of the CTA: for row offset (di,dj) the neighbour cells of every target lie between
// lin(first target) + off - 1 and lin(last target) + off + 1 in linear cell order.  Thread w < 9
// fetches the bounds of window w; thread 0 then packs the windows into the staging buffer.
__device__ __forceinline__ void window_bounds(const GridView& g, const int* __restrict__ offset, const float4 pfirst,
                                              const float4 plast, int w, BlockWindows* bw)
{
    int i, j, k;
    cwa_cell3(g, pfirst.x, pfirst.y, pfirst.z, i, j, k);
    const int lin_f = (i * g.n[1] + j) * g.kstride + k;
    cwa_cell3(g, plast.x, plast.y, plast.z, i, j, k);
    const int lin_l = (i * g.n[1] + j) * g.kstride + k;
    const int si = g.n[1] * g.kstride, sj = g.kstride;
    const int off = (w / 3 - 1) * si + (w % 3 - 1) * sj;
    int lo_lin = lin_f + off - 1, hi_lin = lin_l + off + 1;
    lo_lin = max(lo_lin, 0); hi_lin = min(hi_lin, g.num_cells - 1);
    int lo = 0, hi = 0;
    if (hi_lin >= lo_lin) { lo = __ldg(offset + lo_lin); hi = __ldg(offset + hi_lin + 1); }
    bw->lo[w] = lo; bw->hi[w] = hi;
}

__device__ __forceinline__ uint32_t window_pack(int cap_slots, bool reach_ok, BlockWindows* bw)
{
    int acc = 0;
    for (int w = 0; w < NB_WINDOWS; w++) {
        const int len = bw->hi[w] - bw->lo[w];
        if (reach_ok && len > 0 && acc + len <= cap_slots) { bw->soff[w] = acc; acc += len; }
        else bw->soff[w] = -1;
    }
    return (uint32_t)acc;
}

// Conservative cell coordinate for the neighbour query: floor((x - min) * inv_cell + bias) clamped.
// The hash itself uses the exact IEEE division (cwa_cell3); here a slightly LARGER cell range than
// ComputeCellIndex(pos -+ h) is harmless: cells outside the exact range only hold particles farther
// than h on that axis, which the distance test rejects, so the accepted set is unchanged.
__device__ __forceinline__ int approx_cell(float x, float mn, float inv, float bias, int n)
{
    const float f = floorf(fmaf(x - mn, inv, bias));
    return (int)fminf(fmaxf(f, 0.0f), (float)(n - 1));           // NaN -> 0
}
#define CWA_RANGE_EPS 1e-3f

struct Query3 { int i0, i1, j0, j1, k0, k1, ci, cj; };

__device__ __forceinline__ Query3 make_query(const GridView& g, float x, float y, float z, float h)
{
    Query3 q;
    q.i0 = approx_cell(x - h, g.min[0], g.inv_cell[0], -CWA_RANGE_EPS, g.n[0]);
    q.i1 = approx_cell(x + h, g.min[0], g.inv_cell[0], +CWA_RANGE_EPS, g.n[0]);
    q.j0 = approx_cell(y - h, g.min[1], g.inv_cell[1], -CWA_RANGE_EPS, g.n[1]);
    q.j1 = approx_cell(y + h, g.min[1], g.inv_cell[1], +CWA_RANGE_EPS, g.n[1]);
    q.k0 = approx_cell(z - h, g.min[2], g.inv_cell[2], -CWA_RANGE_EPS, g.n[2]);
    q.k1 = approx_cell(z + h, g.min[2], g.inv_cell[2], +CWA_RANGE_EPS, g.n[2]);
    q.ci = approx_cell(x, g.min[0], g.inv_cell[0], 0.0f, g.n[0]);
    q.cj = approx_cell(y, g.min[1], g.inv_cell[1], 0.0f, g.n[1]);
    return q;
}

// staged-window lookup for the row (i,j) of a query: returns the index shift into the staging
// buffer, or INT_MIN when the row range must be read from global memory
__device__ __forceinline__ int row_shift(const BlockWindows& bw, const Query3& q, int i, int j, int g0, int g1)
{
    const int di = i - q.ci + 1, dj = j - q.cj + 1;
    if ((unsigned)di < 3u && (unsigned)dj < 3u) {
        const int w = di * 3 + dj;
        if (bw.soff[w] >= 0 && g0 >= bw.lo[w] && g1 <= bw.hi[w]) return bw.soff[w] - bw.lo[w];
    }
    return INT_MIN;
}

template <int P, int L>
__global__ void __launch_bounds__(P * L)
sph3_density_grid_kernel(const float4* __restrict__ posS, const float4* __restrict__ velS,
                         float4* __restrict__ pack, int n_max, GridView g,
                         const int* __restrict__ offset, const Sph3Const* __restrict__ cc, TexView tex, int cap_slots)
{
    extern __shared__ float4 stage[];
    __shared__ BlockWindows bw;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * P;
    const int n = min(n_max, __ldg(offset + g.num_cells));   // inserted particles
    if (t0 >= n) return;
    const int nt = min(P, n - t0);
    const float h = cc->h, accept_r2 = cc->accept_r2, h2 = cc->h2, poly6 = cc->poly6;
    const bool reach_ok = (h <= g.cell[0]) && (h <= g.cell[1]) && (h <= g.cell[2]);

    // own target first: its load overlaps the window setup below
    const int slot = t0 + tid / L, sub = tid % L;
    const bool active = (tid / L) < nt;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    Query3 q{0, -1, 0, -1, 0, 0, 0, 0};
    if (active) {
        p = __ldg(posS + slot);
        q = make_query(g, p.x, p.y, p.z, h);
    }
    if (tid < NB_WINDOWS) window_bounds(g, offset, __ldg(posS + t0), __ldg(posS + t0 + nt - 1), tid, &bw);
    if (tid == 0) s3_mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t slots = window_pack(cap_slots, reach_ok, &bw);
        if (slots > 0) {
            s3_mbar_expect_tx(&bar, slots * 16u);
            for (int w = 0; w < NB_WINDOWS; w++)
                if (bw.soff[w] >= 0) s3_bulk_g2s(stage + bw.soff[w], posS + bw.lo[w], (uint32_t)(bw.hi[w] - bw.lo[w]) * 16u, &bar);
        } else {
            s3_mbar_arrive(&bar);
        }
    }
    __syncthreads();                                   // windows packed, barrier armed
    s3_mbar_wait(&bar, 0);

    float rho = 0.0f;
    for (int i = q.i0; i <= q.i1; i++) {
        for (int j = q.j0; j <= q.j1; j++) {
            const int base = (i * g.n[1] + j) * g.kstride;
            const int g0 = __ldg(offset + base + q.k0), g1 = __ldg(offset + base + q.k1 + 1);
            const int shift = row_shift(bw, q, i, j, g0, g1);
            if (shift != INT_MIN) {
                const float4* src = stage + shift;
                for (int c = g0 + sub; c < g1; c += L) pair_density(accept_r2, h2, poly6, p.x, p.y, p.z, src[c], rho);
            } else {
                for (int c = g0 + sub; c < g1; c += L) pair_density(accept_r2, h2, poly6, p.x, p.y, p.z, __ldg(posS + c), rho);
            }
        }
    }
#pragma unroll
    for (int d = 1; d < L; d <<= 1) rho += __shfl_xor_sync(0xffffffffu, rho, d);
    if (active && sub == 0) {
        float rho_out, prs_out;
        density_epilogue(*cc, tex, p.x, p.z, rho, rho_out, prs_out);
        const float4 v = __ldg(velS + slot);
        cwa_stg256(pack + 2 * (size_t)slot, make_float4(p.x, p.y, p.z, prs_out), make_float4(v.x, v.y, v.z, rho_out));
    }
}

// neighbour sums of force_comp.glsl:74-88 only; the per-particle terms (:90-114) are evaluated by
// the element-wise kernels below, where every lane has work.
template <int P, int L>
__global__ void __launch_bounds__(P * L)
sph3_force_grid_kernel(const float4* __restrict__ pack,
                       float4* __restrict__ pairP, float2* __restrict__ pairV, int n_max, GridView g,
                       const int* __restrict__ offset, const Sph3Const* __restrict__ cc, int cap_slots)
{
    extern __shared__ float4 stage[];                 // [cap_slots] records of two float4: (pos, p), (vel, rho)
    __shared__ BlockWindows bw;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * P;
    const int n = min(n_max, __ldg(offset + g.num_cells));   // inserted particles
    if (t0 >= n) return;
    const int nt = min(P, n - t0);
    const Sph3Const c = *cc;
    const bool reach_ok = (c.h <= g.cell[0]) && (c.h <= g.cell[1]) && (c.h <= g.cell[2]);

    const int slot = t0 + tid / L, sub = tid % L;
    const bool active = (tid / L) < nt;
    float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa;
    Query3 q{0, -1, 0, -1, 0, 0, 0, 0};
    if (active) {
        pa = __ldg(pack + 2 * (size_t)slot);
        pb = __ldg(pack + 2 * (size_t)slot + 1);
        q = make_query(g, pa.x, pa.y, pa.z, c.h);
    }
    if (tid < NB_WINDOWS) window_bounds(g, offset, __ldg(pack + 2 * (size_t)t0), __ldg(pack + 2 * (size_t)(t0 + nt - 1)), tid, &bw);
    if (tid == 0) s3_mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t slots = window_pack(cap_slots, reach_ok, &bw);
        if (slots > 0) {
            s3_mbar_expect_tx(&bar, slots * 32u);
            for (int w = 0; w < NB_WINDOWS; w++)
                if (bw.soff[w] >= 0)
                    s3_bulk_g2s(stage + 2 * bw.soff[w], pack + 2 * (size_t)bw.lo[w], (uint32_t)(bw.hi[w] - bw.lo[w]) * 32u, &bar);
        } else {
            s3_mbar_arrive(&bar);
        }
    }
    __syncthreads();
    s3_mbar_wait(&bar, 0);

    float fpx = 0.f, fpy = 0.f, fpz = 0.f, fvx = 0.f, fvy = 0.f, fvz = 0.f;
    for (int i = q.i0; i <= q.i1; i++) {
        for (int j = q.j0; j <= q.j1; j++) {
            const int base = (i * g.n[1] + j) * g.kstride;
            const int g0 = __ldg(offset + base + q.k0), g1 = __ldg(offset + base + q.k1 + 1);
            const int shift = row_shift(bw, q, i, j, g0, g1);
            // records of the row: staged (shared memory) or straight from global memory, two float4 per slot
            const float4* sa = (shift != INT_MIN) ? (const float4*)(stage + 2 * shift) : pack;
            for (int cb = g0 + sub; cb < g1; cb += 32 * L) {
                // phase 1: mark accepted candidates (up to 32 per lane) -- cheap, runs on every candidate
                unsigned mask = 0u;
                int cnd = cb;
#pragma unroll 4
                for (int t = 0; t < 32 && cnd < g1; t++, cnd += L) {
                    const float4 qa = sa[2 * cnd];
                    const float r2 = cwa_len3sq(pa.x - qa.x, pa.y - qa.y, pa.z - qa.z);
                    if (r2 <= c.accept_r2 && cnd != slot) mask |= (1u << t);
                }
                // phase 2: evaluate only the accepted ones
                while (mask) {
                    const int t = __ffs(mask) - 1;
                    mask &= mask - 1u;
                    const int cj = cb + t * L;
                    pair_force(c, pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, sa[2 * cj], sa[2 * cj + 1], fpx, fpy, fpz, fvx, fvy, fvz);
                }
            }
        }
    }
#pragma unroll
    for (int d = 1; d < L; d <<= 1) {
        fpx += __shfl_xor_sync(0xffffffffu, fpx, d); fpy += __shfl_xor_sync(0xffffffffu, fpy, d);
        fpz += __shfl_xor_sync(0xffffffffu, fpz, d); fvx += __shfl_xor_sync(0xffffffffu, fvx, d);
        fvy += __shfl_xor_sync(0xffffffffu, fvy, d); fvz += __shfl_xor_sync(0xffffffffu, fvz, d);
    }
    if (active && sub == 0) {
        pairP[slot] = make_float4(fpx, fpy, fpz, fvx);
        pairV[slot] = make_float2(fvy, fvz);
    }
}

// ---------------------------------------------------------------------------------------------
// neighbour-list kernels (nb_config 7, round 1's default): ONE thread per target
// ---------------------------------------------------------------------------------------------
// With h <= cell the query pos -+ h touches at most 3 x 3 rows (i,j) of cells, each a contiguous run
// [offset[k0], offset[k1+1]) of the cell-ordered arrays.  The density pass loads the (up to) 18 row bounds
// of its target up front (independent loads), parks them in a small shared-memory table and walks the nine
// rows; every accepted candidate is also appended to the target's neighbour list, which the force pass
// walks instead of testing the candidates a second time (positions do not change between the two passes).
//
// Work per target varies by orders of magnitude once particles clump (cells with hundreds of particles): a thread
// left alone with such a target would outlast the rest of the kernel.  Targets with more than EXTREME_CANDIDATES
// candidates (density pass) or more than K neighbours (force pass) are therefore pushed to a device-side queue and
// finished by the "heavy" kernels, one WARP per target, spread over the whole GPU.
constexpr int RT_ROWS = 9;
constexpr int TILE_P = 128;               // targets per tile (= CTA size of the list kernels)
constexpr int EXTREME_CANDIDATES = 192;   // a target with more candidates than this (or more 31-slot row segments than table entries) is a clump target; tuning "extreme_candidates", up to 9 x 31 = 279
constexpr int EXTREME_MARK = 1 << 30;     // neighbour count written for such a target: routes it to the heavy force kernel

struct RowBounds { int b[RT_ROWS], e[RT_ROWS]; int total; bool wide; };

// Conservative query of a target.  When every cell is between (1 + 2.5e-3) h and 1.25 h wide, pos -+ h always
// spans the 3 x 3 x 3 block around the target's cell (a superset of ComputeCellIndex(pos -+ h): a neighbour is
// closer than h <= cell / (1 + 2.5e-3) on every axis, the margin covers the rounding of this estimate against
// the hash), so one cell coordinate per axis is enough; otherwise the per-axis ranges of make_query.
__device__ __forceinline__ Query3 list_query(const GridView& g, float x, float y, float z, float h)
{
    const float hm = h * (1.0f + 2.5e-3f), hx = h * 1.25f;
    if (hm <= g.cell[0] && hm <= g.cell[1] && hm <= g.cell[2] && (g.cell[0] <= hx || g.n[0] == 1) && (g.cell[1] <= hx || g.n[1] == 1) && (g.cell[2] <= hx || g.n[2] == 1)) {
        Query3 q;
        const int ci = approx_cell(x, g.min[0], g.inv_cell[0], 0.0f, g.n[0]);
        const int cj = approx_cell(y, g.min[1], g.inv_cell[1], 0.0f, g.n[1]);
        const int ck = approx_cell(z, g.min[2], g.inv_cell[2], 0.0f, g.n[2]);
        q.i0 = max(ci - 1, 0); q.i1 = min(ci + 1, g.n[0] - 1);
        q.j0 = max(cj - 1, 0); q.j1 = min(cj + 1, g.n[1] - 1);
        q.k0 = max(ck - 1, 0); q.k1 = min(ck + 1, g.n[2] - 1);
        q.ci = ci; q.cj = cj;
        return q;
    }
    return make_query(g, x, y, z, h);
}

__device__ __forceinline__ RowBounds rows_load(const GridView& g, const int* __restrict__ offset, const Query3& q)
{
    RowBounds rb;
    const int ni = q.i1 - q.i0 + 1, nj = q.j1 - q.j0 + 1;
    rb.wide = (ni > 3) || (nj > 3);                    // h > cell: the generic loops handle the target
#pragma unroll
    for (int r = 0; r < RT_ROWS; r++) {
        const int a = r / 3, b = r % 3;
        rb.b[r] = 0; rb.e[r] = 0;
        if (!rb.wide && a < ni && b < nj) {
            const int base = ((q.i0 + a) * g.n[1] + (q.j0 + b)) * g.kstride;
            rb.b[r] = __ldg(offset + base + q.k0);
            rb.e[r] = __ldg(offset + base + q.k1 + 1);
        }
    }
    rb.total = 0;
#pragma unroll
    for (int r = 0; r < RT_ROWS; r++) rb.total += rb.e[r] - rb.b[r];
    return rb;
}

// Density pass + neighbour lists.  Candidates are read through L1 with 256-bit loads (a pair of cell-ordered
// positions is one sector); shared memory only holds the row table, which leaves most of the SM's 228 KB to
// the L1 cache.  An accepted candidate (self included) is appended to the target's own row of K entries in
// global memory ([slot][K], K = 64 by default: two private 128-byte lines that stay in L2 until the force pass reads
// them) with one predicated store; count[slot] keeps counting past K, and the force pass re-scans the grid for such a target.
template <bool LOCAL>
__global__ void __launch_bounds__(TILE_P)
sph3_density_list_kernel(const float4* __restrict__ posS, const float4* __restrict__ velS, float4* __restrict__ pack,
                         int* __restrict__ nbr_list, int* __restrict__ nbr_count,
                         int* __restrict__ heavy_queue, int* __restrict__ heavy_count,
                         int n_max, GridView g, const int* __restrict__ offset, const Sph3Const* __restrict__ cc, TexView tex,
                         const int K, const int extreme_candidates)
{
    __shared__ int2 tab[RT_ROWS * TILE_P];             // (first slot, length) of the 3 x 3 rows of every target
    const int tid = threadIdx.x;
    const int slot = blockIdx.x * TILE_P + tid;
    const int n = min(n_max, __ldg(offset + g.num_cells));   // inserted particles
    if (slot >= n) return;
    const float h = cc->h, accept_r2 = cc->accept_r2, h2 = cc->h2, poly6 = cc->poly6;
    const float4 p = __ldg(posS + slot);
    const Query3 q = list_query(g, p.x, p.y, p.z, h);
    const RowBounds rb = rows_load(g, offset, q);
    if (rb.total > extreme_candidates) {               // a big clump: finished by sph3_density_heavy_kernel, one warp per target
        const int qi = atomicAdd(heavy_count, 1);
        if (qi < n_max) heavy_queue[qi] = slot;
        return;
    }
#pragma unroll
    for (int r = 0; r < RT_ROWS; r++) tab[r * TILE_P + tid] = make_int2(rb.b[r], rb.e[r] - rb.b[r]);

    float rho = 0.0f;
    int cnt = 0;
    int* gl = nbr_list + (size_t)slot * K;             // the target's own row of K entries
    asm volatile("" : "+l"(gl));                       // keep the row pointer in a register pair: entry address = one IMAD.WIDE
    const int last = K - 1;
    // The append is ONE predicated store, no branch and no capacity test: entry min(cnt, K-1) is written, so a target with more
    // than K neighbours keeps overwriting its last entry -- its count ends above K, which tells the force pass to ignore the list.
    // The predicate is formed inside the asm from the same compare the accumulation uses (r2 <= accept_r2 [and valid]).
    auto step = [&](const float4 qp, int j, bool valid) {
        float r2 = cwa_len3sq(p.x - qp.x, p.y - qp.y, p.z - qp.z);
        r2 = valid ? r2 : CUDART_INF_F;                // a slot outside the row can never be accepted: one compare serves both
        const bool ok = r2 <= accept_r2;
        const float d = ok ? (h2 - r2) : 0.0f;
        rho = fmaf(poly6, d * d * d, rho);             // weight 0 when rejected: no branch around the arithmetic
        int* const w = gl + min(cnt, last);
        asm volatile("{\n\t.reg .pred q;\n\tsetp.le.f32 q, %2, %3;\n\t@q st.global.b32 [%0], %1;\n\t}"
                     :: "l"(w), "r"(j), "f"(r2), "f"(accept_r2) : "memory");
        cnt += ok ? 1 : 0;
    };
    // Four candidates per trip.  `j` is even: the two 256-bit loads fetch the pairs (j, j+1) and (j+2, j+3),
    // both issued before the first use.  Candidates outside [first, first + len) -- the slot before an odd row
    // start, the slots after the row end (the cell-ordered array is padded by 4) -- are masked out: candidate u of the
    // trip is valid iff lo <= u < hi, two trip-level bounds compared against constants.
    auto quad = [&](int j, int first, int len, bool masked) {
        const f4x2 lo2 = cwa_ldg256(posS + j), hi2 = cwa_ldg256(posS + j + 2);
        const float4 qp[4] = {lo2.a, lo2.b, hi2.a, hi2.b};
        const int lo = first - j, hi = first + len - j;
#pragma unroll
        for (int u = 0; u < 4; u++) step(qp[u], j + u, !masked || (lo <= u && u < hi));
    };
#pragma unroll 1
    for (int r = 0; r < RT_ROWS; r++) {
        const int2 e = tab[r * TILE_P + tid];
        if (e.y <= 0) continue;
        const int end = e.x + e.y;
        int j = e.x & ~1;
        quad(j, e.x, e.y, true);                       // first quad: the row may start at an odd slot
        for (j += 4; j + 4 <= end; j += 4) quad(j, e.x, e.y, false);
        if (j < end) quad(j, e.x, e.y, true);          // last quad: reads past the end of the row
    }
    if (rb.wide) {                                     // h > cell: generic query, one candidate at a time
        for (int i = q.i0; i <= q.i1; i++)
            for (int j = q.j0; j <= q.j1; j++) {
                const int base = (i * g.n[1] + j) * g.kstride;
                const int g0 = __ldg(offset + base + q.k0), g1 = __ldg(offset + base + q.k1 + 1);
                for (int c = g0; c < g1; c++) step(__ldg(posS + c), c, true);
            }
    }
    float rho_out, prs_out;
    density_epilogue<LOCAL>(*cc, tex, p.x, p.z, rho, rho_out, prs_out);
    const float4 v = __ldg(velS + slot);
    cwa_stg256(pack + 2 * (size_t)slot, make_float4(p.x, p.y, p.z, prs_out), make_float4(v.x, v.y, v.z, rho_out));
    nbr_count[slot] = cnt;
}

// ---------------------------------------------------------------------------------------------
// Row-mask neighbour kernels (nb_config 8, the default).  Same thread-per-target walk as the list kernels above, rebuilt around
// what ncu showed about sph3_density_list_kernel on C4 (profiles/r2/density_r1_vs_r2.md): the L1 data pipe, not the issue slots,
// was the limit (l1tex wavefronts 77 % of peak) and 45 % of its wavefronts were the neighbour-list appends -- one 4-byte store per
// accepted pair, every lane into its own 256-byte list row, i.e. one wavefront and one 32-byte sector per ENTRY.
//  * Neighbours are recorded as one 32-bit ACCEPT MASK per row of the query (bit b = slot first + b of the row), kept in a register
//    while the row is walked: accepting a candidate is one predicated OR, no address arithmetic, no store.  A target leaves at most
//    nine (first slot, mask) pairs, written once, coalesced ([tile of 32 targets][row][lane]): 72 bytes per target instead of
//    ~54 scattered ones, and the force pass reads them back with nine coalesced loads instead of 32-line index gathers.
//  * ONE loop over all rows of the target instead of a loop nest.  In a nest every lane waits at the end of each row for the longest row of the
//    warp (cost = sum over rows of the warp-wide maximum); in the flat loop a lane moves on to its next row on its own (cost = the
//    warp-wide maximum of the TOTAL, which varies far less than the individual rows).  The row advance is branch-free (in almost
//    every trip SOME lane of the warp advances: a branch would run both sides every time with a handful of lanes active).
//  * two candidates per trip (one 256-bit load = one sector), the next pair loaded before the current one is used; rows are six candidates
//    long on average, so four-wide trips spent a third of their slots outside the row.
//  * the query is narrowed per row with the target's position inside its cell: a neighbour row (i', j') whose footprint is farther than h
//    in (x, y) is dropped, and its k range keeps cell ck -+ 1 only when that cell's nearest point is closer than h.  Cells of ~h: 20.6 of
//    the 27 cells survive on average.  The test is conservative (margin 2.5e-3 h against the rounding of the hash), and a dropped cell only
//    holds particles farther than h, so the accepted set -- and with it every result -- is unchanged.
// Targets the scheme does not cover -- more than EXTREME_CANDIDATES candidates, a row longer than 31 slots, cells smaller than h
// (more than 3 x 3 rows) -- go to the heavy kernels (one warp per target, generic loops).
// ---------------------------------------------------------------------------------------------
constexpr int INPLACE_MARK = EXTREME_MARK + 1; // neighbour count of a clump target the density pass finished in place, without masks
constexpr int INPLACE_MAX = 640;          // candidates up to which the density pass finishes a clump target in place (tuning "inplace_max"; beyond: one warp per target)
// packed FP32 pairs (sm_100: FADD2 / FMUL2 / FFMA2; each half is rounded like the scalar instruction)
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// 1 << n with PTX semantics (a shift by 32 or more gives 0): the rows of a clump target finished without masks are longer than a mask
__device__ __forceinline__ unsigned shl_clamp(unsigned v, int n) { unsigned r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }
__device__ __forceinline__ unsigned long long f2_ldg(const float* p) { unsigned long long r; asm("ld.global.nc.b64 %0, [%1];" : "=l"(r) : "l"(p)); return r; }

constexpr int ROW_MASK_BITS = 31;         // longest row a mask records: with an odd first slot the pair loop shifts by (length + 1) - 1 at most
struct FlatRows { int b[RT_ROWS], e[RT_ROWS]; int total; bool wide, longrow; int cell_lin; };   // cell_lin: the target's cell (fast path), -1 = generic query

__device__ __forceinline__ FlatRows flat_rows_load(const GridView& g, const int* __restrict__ offset, float x, float y, float z, float h)
{
    FlatRows fr;
    const float hm = h * (1.0f + 2.5e-3f), hx = h * 1.25f;
    // (an axis with ONE layer of cells may be as thick as it likes: its single cell is the whole range; e.g. columns over a shallow sheet)
    const bool fast = hm <= g.cell[0] && hm <= g.cell[1] && hm <= g.cell[2] && (g.cell[0] <= hx || g.n[0] == 1) && (g.cell[1] <= hx || g.n[1] == 1) && (g.cell[2] <= hx || g.n[2] == 1);
    fr.total = 0;
    fr.cell_lin = -1;
#pragma unroll
    for (int r = 0; r < RT_ROWS; r++) { fr.b[r] = 0; fr.e[r] = 0; }
    if (!fast) {                                           // generic query: rows_load semantics (more than 3 x 3 rows -> heavy kernels)
        const Query3 q = make_query(g, x, y, z, h);
        const RowBounds rb = rows_load(g, offset, q);
#pragma unroll
        for (int r = 0; r < RT_ROWS; r++) { fr.b[r] = rb.b[r]; fr.e[r] = rb.e[r]; }
        fr.total = rb.total; fr.wide = rb.wide;
    } else {
        fr.wide = false;
        const int ci = approx_cell(x, g.min[0], g.inv_cell[0], 0.0f, g.n[0]);
        const int cj = approx_cell(y, g.min[1], g.inv_cell[1], 0.0f, g.n[1]);
        const int ck = approx_cell(z, g.min[2], g.inv_cell[2], 0.0f, g.n[2]);
        const int kl = max(ck - 1, 0), kh = min(ck + 1, g.n[2] - 1);
        fr.cell_lin = (ci * g.n[1] + cj) * g.kstride + ck;
        // squared distance (minus the margin) from the target to the lower / upper neighbour slab of its cell, per axis
        const float m = 2.5e-3f * h, h2 = h * h;
        float dl[3], dh[3];
        const float pos[3] = {x, y, z};
        const int cc[3] = {ci, cj, ck};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float lo = fmaf((float)cc[a], g.cell[a], g.min[a]);
            const float l = fmaxf(pos[a] - lo - m, 0.0f), u = fmaxf(lo + g.cell[a] - pos[a] - m, 0.0f);
            dl[a] = l * l; dh[a] = u * u;
        }
#pragma unroll
        for (int r = 0; r < RT_ROWS; r++) {
            const int di = r / 3 - 1, dj = r % 3 - 1;
            const int i = ci + di, j = cj + dj;
            const float s = (di < 0 ? dl[0] : (di > 0 ? dh[0] : 0.0f)) + (dj < 0 ? dl[1] : (dj > 0 ? dh[1] : 0.0f));
            if (i >= 0 && i < g.n[0] && j >= 0 && j < g.n[1] && s < h2) {
                const int k0 = (s + dl[2] < h2) ? kl : ck;
                const int k1 = (s + dh[2] < h2) ? kh : ck;
                const int base = (i * g.n[1] + j) * g.kstride;
                fr.b[r] = __ldg(offset + base + k0);
                fr.e[r] = __ldg(offset + base + k1 + 1);
            }
        }
#pragma unroll
        for (int r = 0; r < RT_ROWS; r++) fr.total += fr.e[r] - fr.b[r];
    }
    fr.longrow = false;
#pragma unroll
    for (int r = 0; r < RT_ROWS; r++) fr.longrow = fr.longrow || (fr.e[r] - fr.b[r] > ROW_MASK_BITS);
    return fr;
}

// (first slot, accept mask) pairs of a target: [tile of 32 targets][row][lane]
__device__ __forceinline__ int2* rows_of(int2* nbr_rows, int slot) { return nbr_rows + ((size_t)(slot >> 5) * RT_ROWS) * 32 + (slot & 31); }

template <bool LOCAL>
__global__ void __launch_bounds__(TILE_P)
sph3_density_flat_kernel(const float4* __restrict__ posS, const float* __restrict__ xyzS, const size_t xyz_stride,
                         const float4* __restrict__ velS, float4* __restrict__ pack,
                         int2* __restrict__ nbr_rows, int* __restrict__ nbr_count,
                         int* __restrict__ heavy_queue, int* __restrict__ heavy_count,
                         int n_max, GridView g, const int* __restrict__ offset, const Sph3Const* __restrict__ cc, TexView tex,
                         const int extreme_candidates, const int inplace_max)
{
    CWA_PDL_ENTER();
    __shared__ int2 tab[(RT_ROWS + 1) * TILE_P];       // non-empty rows of every target, compacted: (first slot, end slot), later (first slot, mask)
    const int tid = threadIdx.x;
    const int slot = blockIdx.x * TILE_P + tid;
    const int n = min(n_max, __ldg(offset + g.num_cells));   // inserted particles
    if (slot >= n) return;
    const float h = cc->h, accept_r2 = cc->accept_r2, h2 = cc->h2, poly6 = cc->poly6;
    const float4 p = __ldg(posS + slot);
    const FlatRows fr = flat_rows_load(g, offset, p.x, p.y, p.z, h);
    // Three kinds of targets.  maskable: the usual case, neighbours recorded as row masks.  in place: a clump target (more candidates than
    // `extreme_candidates`, or a row too long for a mask) walks its rows like everybody else, only without masks: the density loop has no
    // divergent part, clump targets come in runs of consecutive slots (lanes read the same candidates), and the per-target set-up of the
    // one-warp-per-target kernel -- half its instructions for targets of a few hundred candidates -- is not paid.  The FORCE pass of such a
    // target goes to sph3_force_heavy_kernel (measured, profiles/r2/tuning.md: one thread alone with the divergent pair term of a few
    // hundred candidates is a 50-150 us tail, in its own warp or from a queue).  heavy: a generic wide query or more than `inplace_max`
    // candidates -- sph3_density_heavy_kernel, one warp per target.
    // (a row longer than a mask is cut into segments of ROW_MASK_BITS slots -- table entries of their own -- while the table has room)
    int segs = 0;
#pragma unroll
    for (int r = 0; r < RT_ROWS; r++) segs += (fr.e[r] - fr.b[r] + ROW_MASK_BITS - 1) / ROW_MASK_BITS;
    const bool maskable = !fr.wide && segs <= RT_ROWS && fr.total <= extreme_candidates;
    if (!maskable && (fr.wide || fr.cell_lin < 0 || fr.total > inplace_max)) {
        const int qi = atomicAdd(heavy_count, 1);
        if (qi < n_max) heavy_queue[qi] = slot;
        return;
    }
    int nr = 0;
    if (maskable && fr.longrow) {
#pragma unroll
        for (int r = 0; r < RT_ROWS; r++)
            for (int b = fr.b[r]; b < fr.e[r]; b += ROW_MASK_BITS) { tab[nr * TILE_P + tid] = make_int2(b, min(b + ROW_MASK_BITS, fr.e[r])); nr++; }
    } else {
#pragma unroll
        for (int r = 0; r < RT_ROWS; r++)
            if (fr.e[r] > fr.b[r]) { tab[nr * TILE_P + tid] = make_int2(fr.b[r], fr.e[r]); nr++; }
    }

    // The candidates come as three coordinate streams (x | y | z, written by the reorder pass next to posS): a pair of consecutive slots
    // is one aligned 64-bit word per coordinate -- 24 bytes per pair through the L1 data pipe (the limit of this kernel, profiles/r2)
    // instead of the 32 of two float4 -- and lands in a register pair, so the whole distance evaluation runs on Blackwell's packed
    // FP32 instructions (sub / mul / fma .f32x2 = FADD2 / FMUL2 / FFMA2: both candidates of the pair per issue slot, each lane rounded
    // exactly like the scalar instruction, so the accept test is bit-identical to cwa_len3sq of the oracle).
    unsigned long long rho2 = 0ull;                    // partial sums of the two slots of a pair
    unsigned mask = 0u;
    const unsigned long long px2 = f2_pack(p.x, p.x), py2 = f2_pack(p.y, p.y), pz2 = f2_pack(p.z, p.z);
    const unsigned long long h22 = f2_pack(h2, h2), poly2 = f2_pack(poly6, poly6);
    const float* const xs = xyzS; const float* const ys = xyzS + xyz_stride; const float* const zs = xyzS + 2 * xyz_stride;
    // one pair (slots j, j + 1) against the target; slot j counts iff j >= first (head), slot j + 1 iff j + 1 < end (tail)
    auto step2 = [&](unsigned long long X, unsigned long long Y, unsigned long long Z, int j, int first, int end, unsigned bit1) {
        const unsigned long long dx = f2_sub(px2, X), dy = f2_sub(py2, Y), dz = f2_sub(pz2, Z);
        const unsigned long long r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
        const unsigned long long hd = f2_sub(h22, r2);
        float r2a, r2b, hda, hdb, da, db;
        f2_unpack(r2, r2a, r2b); f2_unpack(hd, hda, hdb);
        asm("{\n\t.reg .pred v, q;\n\tsetp.ge.s32 v, %4, %5;\n\tsetp.le.and.f32 q, %2, %3, v;\n\t@q or.b32 %0, %0, %6;\n\t"
            "selp.f32 %1, %7, 0f00000000, q;\n\t}"
            : "+r"(mask), "=f"(da) : "f"(r2a), "f"(accept_r2), "r"(j), "r"(first), "r"(bit1 >> 1), "f"(hda));
        asm("{\n\t.reg .pred v, q;\n\tsetp.lt.s32 v, %4, %5;\n\tsetp.le.and.f32 q, %2, %3, v;\n\t@q or.b32 %0, %0, %6;\n\t"
            "selp.f32 %1, %7, 0f00000000, q;\n\t}"
            : "+r"(mask), "=f"(db) : "f"(r2b), "f"(accept_r2), "r"(j + 1), "r"(end), "r"(bit1), "f"(hdb));
        const unsigned long long d = f2_pack(da, db);
        rho2 = f2_fma(poly2, f2_mul(f2_mul(d, d), d), rho2);
    };
    if (nr > 0) {
        int2* tp = tab + tid;                          // the lane's current row in the table; rows are TILE_P entries apart
        const int2* const tp_last = tab + (nr - 1) * TILE_P + tid;
        int2 row = *tp;
        int j = row.x & ~1;
        unsigned long long AX = f2_ldg(xs + j), AY = f2_ldg(ys + j), AZ = f2_ldg(zs + j), BX = AX, BY = AY, BZ = AZ;
        // One stage: find where the NEXT pair comes from (the same row, or the first pair of the lane's next row), issue its loads into
        // NXT, then evaluate the pair held in CUR; a finished row leaves its mask in the table.  Two stages per trip ping-pong between
        // A and B, so no pair is ever copied.  The table has one spare row: the entry behind the last row is read, never used.
#define CWA_FLAT_STAGE(CX, CY, CZ, NX, NY, NZ)                                                     \
        {                                                                                          \
            const bool adv = j + 2 >= row.y;                                                       \
            const bool more = !adv || tp != tp_last;                                               \
            int2* const tp_cur = tp;                                                               \
            tp += adv ? TILE_P : 0;                                                                \
            const int2 cand = *tp;                                                                 \
            const int2 rown = adv ? cand : row;                                                    \
            const int jn = adv ? (cand.x & ~1) : j + 2;                                            \
            if (more) { NX = f2_ldg(xs + jn); NY = f2_ldg(ys + jn); NZ = f2_ldg(zs + jn); }        \
            step2(CX, CY, CZ, j, row.x, row.y, shl_clamp(1u, j + 1 - row.x));                      \
            if (adv) { tp_cur->y = (int)mask; mask = 0u; }                                         \
            if (!more) break;                                                                      \
            j = jn; row = rown;                                                                    \
        }
        while (true) {
            CWA_FLAT_STAGE(AX, AY, AZ, BX, BY, BZ)
            CWA_FLAT_STAGE(BX, BY, BZ, AX, AY, AZ)
        }
#undef CWA_FLAT_STAGE
    }
    float rho_lo, rho_hi;
    f2_unpack(rho2, rho_lo, rho_hi);
    const float rho = rho_lo + rho_hi;
    float rho_out, prs_out;
    density_epilogue<LOCAL>(*cc, tex, p.x, p.z, rho, rho_out, prs_out);
    const float4 v = __ldg(velS + slot);
    cwa_stg256(pack + 2 * (size_t)slot, make_float4(p.x, p.y, p.z, prs_out), make_float4(v.x, v.y, v.z, rho_out));
    if (!maskable) { nbr_count[slot] = INPLACE_MARK; return; }
    // the rows that accepted somebody, compacted
    int2* out = rows_of(nbr_rows, slot);
    int oc = 0;
    for (int r = 0; r < nr; r++) {
        const int2 e = tab[r * TILE_P + tid];
        if (e.y != 0) { out[oc * 32] = e; oc++; }
    }
    nbr_count[slot] = oc;
}

// Candidate enumeration of the heavy kernels (one warp per target).  The rows of a query are short where particles pile up on a wall
// (two or three cells along k, each crowded), so walking them one after the other leaves a single load in flight per lane and a partly
// filled warp at every row end.  Instead lanes 0..nrows-1 fetch the row bounds, a warp scan turns the row lengths into a flat index
// space [0, total) (prefix + first slot of each row parked in 2 x 33 ints of shared memory per warp), and lane l takes the flat
// candidates l, l + 32, ... four at a time: four independent loads in flight, every lane busy until the last trip.
struct HeavyRows { int total; };
constexpr int HEAVY_WARPS = 4;            // warps per CTA of the heavy kernels (128 threads)

__device__ __forceinline__ int heavy_rows_setup(const GridView& g, const int* __restrict__ offset, const Query3& q, int nj, int nrows, int lane,
                                                int* sP, int* sG)
{
    int g0 = 0, len = 0;
    if (lane < nrows) {
        const int base = ((q.i0 + lane / nj) * g.n[1] + (q.j0 + lane % nj)) * g.kstride;
        g0 = __ldg(offset + base + q.k0);
        len = __ldg(offset + base + q.k1 + 1) - g0;
    }
    int incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncwarp();                          // the previous target's reads of sP / sG are done
    sP[lane] = incl - len; sG[lane] = g0;
    if (lane == 31) sP[32] = incl;
    __syncwarp();
    return __shfl_sync(0xffffffffu, incl, 31);
}

// flat candidate f -> cell-ordered slot; `r` is the lane's current row and only moves forward (f grows)
__device__ __forceinline__ int heavy_candidate(const int* sP, const int* sG, int f, int& r)
{
    while (f >= sP[r + 1]) r++;
    return sG[r] + (f - sP[r]);
}

// Extreme targets of the density pass (a clump: hundreds to thousands of candidates; left to one thread each they
// would outlast the rest of the kernel): one WARP per queued target, spread over the whole GPU.  Lanes 0..8 fetch the
// bounds of the nine rows in parallel; then lane l takes every 32nd candidate of each row (coalesced 16-byte loads),
// and the partial sums are combined with warp shuffles (fixed order: run-to-run deterministic); lane 0 runs the
// per-particle epilogue.  No neighbour list is written: the count is EXTREME_MARK, which routes the target to
// sph3_force_heavy_kernel.
template <bool LOCAL>
__global__ void __launch_bounds__(128)
sph3_density_heavy_kernel(const float4* __restrict__ posS, const float4* __restrict__ velS, float4* __restrict__ pack,
                          int* __restrict__ nbr_count, const int* __restrict__ heavy_queue, const int* __restrict__ heavy_count, int cap,
                          GridView g, const int* __restrict__ offset, const Sph3Const* __restrict__ cc, TexView tex)
{
    CWA_PDL_ENTER();
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = min(__ldg(heavy_count), cap);
    const float h = cc->h, accept_r2 = cc->accept_r2, h2 = cc->h2, poly6 = cc->poly6;
    __shared__ int s_rows[HEAVY_WARPS][2][33];
    int* const sP = s_rows[(threadIdx.x >> 5) % HEAVY_WARPS][0];
    int* const sG = s_rows[(threadIdx.x >> 5) % HEAVY_WARPS][1];
    for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total; e += nwarps) {
        const int slot = __ldg(heavy_queue + e);
        const float4 p = __ldg(posS + slot);
        const Query3 q = list_query(g, p.x, p.y, p.z, h);
        const int nj = q.j1 - q.j0 + 1, nrows = (q.i1 - q.i0 + 1) * nj;
        float part = 0.0f;
        if (nrows <= 32) {
            const int ncand = heavy_rows_setup(g, offset, q, nj, nrows, lane, sP, sG);
            int r = 0;
            for (int f0 = lane; f0 < ncand; f0 += 128) {
                int c[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int f = f0 + 32 * u;
                    ok[u] = f < ncand;
                    c[u] = ok[u] ? heavy_candidate(sP, sG, f, r) : slot;
                }
                float4 qp[4];
#pragma unroll
                for (int u = 0; u < 4; u++) qp[u] = __ldg(posS + c[u]);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float r2 = cwa_len3sq(p.x - qp[u].x, p.y - qp[u].y, p.z - qp[u].z);
                    const float d = (ok[u] && r2 <= accept_r2) ? (h2 - r2) : 0.0f;
                    part = fmaf(poly6, d * d * d, part);
                }
            }
        } else
        for (int r0 = 0; r0 < nrows; r0 += 32) {           // generic wide query (h > cell): 32 rows per round, row after row
            int g0 = 0, g1 = 0;
            if (r0 + lane < nrows) {
                const int r = r0 + lane;
                const int base = ((q.i0 + r / nj) * g.n[1] + (q.j0 + r % nj)) * g.kstride;
                g0 = __ldg(offset + base + q.k0); g1 = __ldg(offset + base + q.k1 + 1);
            }
            const int rows_here = min(32, nrows - r0);
            for (int r = 0; r < rows_here; r++) {
                const int b = __shfl_sync(0xffffffffu, g0, r), en = __shfl_sync(0xffffffffu, g1, r);
#pragma unroll 4
                for (int c = b + lane; c < en; c += 32) {
                    const float4 qp = __ldg(posS + c);
                    const float r2 = cwa_len3sq(p.x - qp.x, p.y - qp.y, p.z - qp.z);
                    const float d = (r2 <= accept_r2) ? (h2 - r2) : 0.0f;
                    part = fmaf(poly6, d * d * d, part);
                }
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        if (lane == 0) {
            float rho_out, prs_out;
            density_epilogue<LOCAL>(*cc, tex, p.x, p.z, part, rho_out, prs_out);
            const float4 v = __ldg(velS + slot);
            cwa_stg256(pack + 2 * (size_t)slot, make_float4(p.x, p.y, p.z, prs_out), make_float4(v.x, v.y, v.z, rho_out));
            nbr_count[slot] = EXTREME_MARK;
        }
    }
}

// Arguments of the fused tail of a full step (force epilogue + integrate + write-back of the 64-byte record): when
// sph_step runs all three passes the force kernels finish the particle themselves, so the neighbour sums never
// travel through memory and the separate integrate kernel (and its re-read of the pack) disappears.
struct FinishArgs {
    const float4* forceS;        // previous frame's force (torque term, crest rule)
    const float4* miscS;         // (pos.w, vel.w, extras.z, extras.w): carried through unchanged
    const int*    index_list;    // cell-ordered slot -> particle id
    float4*       aos;           // the particle SSBO
    TexView       tex;
};

template <bool LOCAL>
__device__ __forceinline__ void finish_particle(const Sph3Const& c, const FinishArgs& fa, int slot, const float4 pa, const float4 pb,
                                                float fpx, float fpy, float fpz, float fvx, float fvy, float fvz)
{
    const float4 m = __ldg(fa.miscS + slot);
    float4 f = force_epilogue<LOCAL>(c, fa.tex, pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, pb.w, __ldg(fa.forceS + slot), fpx, fpy, fpz, fvx, fvy, fvz);
    float4 pos = make_float4(pa.x, pa.y, pa.z, m.x), vel = make_float4(pb.x, pb.y, pb.z, m.y);
    float rho = pb.w, prs = pa.w;
    integrate_particle<LOCAL>(c, fa.tex, pos, vel, f, rho, prs);
    float4* o = fa.aos + (size_t)__ldg(fa.index_list + slot) * 4;
    cwa_stg256(o, pos, vel);
    cwa_stg256(o + 2, f, make_float4(rho, prs, m.z, m.w));
}

// Force pass over the neighbour lists of the density pass: neighbour sums of force_comp.glsl:74-88.  One thread
// per target, no shared memory; a neighbour's (pos, p | vel, rho) record is one sector fetched with one 256-bit
// load, and consecutive cell-ordered targets share most of their neighbours, so the gathers hit L1/L2.  A target
// whose list overflowed (more than K neighbours) or that the density pass marked extreme is queued for the heavy kernel.
#ifndef CWA_FORCE_MINB
#define CWA_FORCE_MINB 6                  // resident CTAs per SM the force list kernel is compiled for (76 registers; 8 -> 64: measured, tuning.md)
#endif
template <bool FUSED, bool LOCAL>
__global__ void __launch_bounds__(TILE_P, CWA_FORCE_MINB)
sph3_force_list_kernel(const float4* __restrict__ pack, const int* __restrict__ nbr_list, const int* __restrict__ nbr_count,
                       int* __restrict__ heavy_queue, int* __restrict__ heavy_count,
                       float4* __restrict__ pairP, float2* __restrict__ pairV, int n_max,
                       GridView g, const int* __restrict__ offset, const Sph3Const* __restrict__ cc, FinishArgs fa, const int K)
{
    const int slot = blockIdx.x * TILE_P + threadIdx.x;
    const int n = min(n_max, __ldg(offset + g.num_cells));   // inserted particles
    if (slot >= n) return;
    const int cnt = __ldg(nbr_count + slot);
    if (cnt > K) {                                     // list overflow / extreme target: sph3_force_heavy_kernel
        const int qi = atomicAdd(heavy_count, 1);
        if (qi < n_max) heavy_queue[qi] = slot;     // (a pass dispatched twice without a grid build in between re-queues: same results)
        return;
    }
    const Sph3Const c = *cc;
    const f4x2 own = cwa_ldg256(pack + 2 * (size_t)slot);
    const float4 pa = own.a, pb = own.b;
    float fpx = 0.f, fpy = 0.f, fpz = 0.f, fvx = 0.f, fvy = 0.f, fvz = 0.f;
    const int4* lp = reinterpret_cast<const int4*>(nbr_list + (size_t)slot * K);
    int4 jn = (cnt > 0) ? __ldg(lp) : make_int4(slot, slot, slot, slot);
    for (int e = 0; e < cnt; e += 4) {
        const int4 j4 = jn;
        if (e + 4 < cnt) jn = __ldg(lp + (e >> 2) + 1);                           // next quad of indices: overlaps the gathers below
        int j[4] = {j4.x, j4.y, j4.z, j4.w};
        float4 qa[4], qb[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (e + u >= cnt) j[u] = slot;                // tail of the last quad: skipped like self
#pragma unroll
        for (int u = 0; u < 4; u++) { const f4x2 rec = cwa_ldg256(pack + 2 * (size_t)j[u]); qa[u] = rec.a; qb[u] = rec.b; }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (j[u] != slot) pair_force(c, pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, qa[u], qb[u], fpx, fpy, fpz, fvx, fvy, fvz);
    }
    if (FUSED) {
        finish_particle<LOCAL>(c, fa, slot, pa, pb, fpx, fpy, fpz, fvx, fvy, fvz);
    } else {
        pairP[slot] = make_float4(fpx, fpy, fpz, fvx);
        pairV[slot] = make_float2(fvy, fvz);
    }
}

// Force pass over the row masks of sph3_density_flat_kernel: neighbour sums of force_comp.glsl:74-88, one thread per target.
// The (first slot, mask) pairs are copied into shared memory with coalesced loads; the lane then walks the set bits of its
// masks in ONE loop (branch-free move to the next row), gathering each neighbour's (pos, p | vel, rho) record -- one sector --
// one neighbour ahead of the pair evaluation.
template <bool FUSED, bool LOCAL>
__global__ void __launch_bounds__(TILE_P)
sph3_force_rows_kernel(const float4* __restrict__ pack, const int2* __restrict__ nbr_rows, const int* __restrict__ nbr_count,
                       int* __restrict__ heavy_queue, int* __restrict__ heavy_count,
                       float4* __restrict__ pairP, float2* __restrict__ pairV, int n_max,
                       GridView g, const int* __restrict__ offset, const Sph3Const* __restrict__ cc, FinishArgs fa)
{
    CWA_PDL_ENTER();
    __shared__ int2 tab[(RT_ROWS + 1) * TILE_P];
    const int tid = threadIdx.x;
    const int slot = blockIdx.x * TILE_P + tid;
    const int n = min(n_max, __ldg(offset + g.num_cells));   // inserted particles
    if (slot >= n) return;
    // A target has a dozen neighbours: the loop below is short, and the chain of dependent loads in front of it (row count -> rows ->
    // first neighbour record) was 40 % of the kernel's stall samples (ncu source view).  So everything the thread will need is requested
    // at once: all nine row slots (entries behind the count are stale and never used), the count and the target's own record.
    const int2* rows = rows_of(const_cast<int2*>(nbr_rows), slot);
    int2 rw[RT_ROWS];
#pragma unroll
    for (int r = 0; r < RT_ROWS; r++) rw[r] = __ldg(rows + r * 32);
    const f4x2 own = cwa_ldg256(pack + 2 * (size_t)slot);
    const int nr = __ldg(nbr_count + slot);
    if (nr > RT_ROWS) {                                // clump target (INPLACE_MARK or EXTREME_MARK): sph3_force_heavy_kernel
        const int qi = atomicAdd(heavy_count, 1);
        if (qi < n_max) heavy_queue[qi] = slot;        // (a pass dispatched twice without a grid build in between re-queues: same results)
        return;
    }
#pragma unroll
    for (int r = 0; r < RT_ROWS; r++)
        if (r < nr) tab[r * TILE_P + tid] = rw[r];
    const Sph3Const c = *cc;
    const float4 pa = own.a, pb = own.b;
    float fpx = 0.f, fpy = 0.f, fpz = 0.f, fvx = 0.f, fvy = 0.f, fvz = 0.f;
    {
        // The lane walks the set bits of its masks in ONE loop; moving to the next row is branch-free (pointer + 0 or one table row, entry
        // read unconditionally and selected).  Records are gathered TWO neighbours ahead of the pair evaluation (three buffers in
        // rotation): the gathers are L2 / DRAM latency and one thread has little else to overlap them with.
        int trow = -1;                                 // current table row of the lane
        const int last = nr - 1;
        int first = 0;
        unsigned mask = 0u;
        auto next = [&](int& j) -> bool {
            const bool adv = mask == 0u;
            const bool none = adv && trow == last;
            const bool go = adv && !none;
            trow += go ? 1 : 0;
            const int2 cand = tab[max(trow, 0) * TILE_P + tid];
            if (go) { first = cand.x; mask = (unsigned)cand.y; }
            j = first + __ffs((int)mask) - 1;
            mask &= mask - 1u;
            return !none;
        };
        int j0, j1, j2;
        bool v0 = next(j0), v1 = next(j1), v2;
        f4x2 X0 = own, X1 = own, X2 = own;
        if (v0) X0 = cwa_ldg256(pack + 2 * (size_t)j0);
        if (v1) X1 = cwa_ldg256(pack + 2 * (size_t)j1);
#define CWA_ROWS_STAGE(JC, VC, XC, JN, VN, XN)                                                     \
        {                                                                                          \
            if (!VC) break;                                                                        \
            VN = next(JN);                                                                         \
            if (VN) XN = cwa_ldg256(pack + 2 * (size_t)JN);                                        \
            if (JC != slot) pair_force(c, pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, XC.a, XC.b, fpx, fpy, fpz, fvx, fvy, fvz); \
        }
        while (true) {
            CWA_ROWS_STAGE(j0, v0, X0, j2, v2, X2)
            CWA_ROWS_STAGE(j1, v1, X1, j0, v0, X0)
            CWA_ROWS_STAGE(j2, v2, X2, j1, v1, X1)
        }
#undef CWA_ROWS_STAGE
    }
    if (FUSED) {
        finish_particle<LOCAL>(c, fa, slot, pa, pb, fpx, fpy, fpz, fvx, fvy, fvz);
    } else {
        pairP[slot] = make_float4(fpx, fpy, fpz, fvx);
        pairV[slot] = make_float2(fvy, fvz);
    }
}

// Queued targets of the force pass (list overflow, clumps): one WARP per target re-scans the grid; lanes 0..8 fetch the row bounds in
// parallel, lane l takes every 32nd candidate of a row, the six sums are combined with warp shuffles.
template <bool FUSED, bool LOCAL>
__global__ void __launch_bounds__(128)
sph3_force_heavy_kernel(const float4* __restrict__ pack, const int* __restrict__ heavy_queue, const int* __restrict__ heavy_count, int cap,
                        float4* __restrict__ pairP, float2* __restrict__ pairV, GridView g,
                        const int* __restrict__ offset, const Sph3Const* __restrict__ cc, FinishArgs fa, const int sub_warp_heavy, const int sub_warp_max)
{
    CWA_PDL_ENTER();
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = min(__ldg(heavy_count), cap);
    const Sph3Const c = *cc;
    // fused tail: lane k keeps the result of the warp's k-th target, and the (up to 32) kept targets are finished by
    // all lanes at once -- the per-particle epilogue is as long as a short neighbour scan, one lane at a time would
    // leave the warp 31/32 idle when many targets are queued
    int kept = 0, k_slot = 0;
    float4 k_pa = make_float4(0.f, 0.f, 0.f, 0.f), k_pb = k_pa;
    float k0 = 0.f, k1 = 0.f, k2 = 0.f, k3 = 0.f, k4 = 0.f, k5 = 0.f;
    __shared__ int s_rows[HEAVY_WARPS][2][33];
    int* const sP = s_rows[(threadIdx.x >> 5) % HEAVY_WARPS][0];
    int* const sG = s_rows[(threadIdx.x >> 5) % HEAVY_WARPS][1];
    // one target, the whole warp: lane l takes every 32nd candidate
    auto whole_warp = [&](const int slot) {
        const f4x2 own = cwa_ldg256(pack + 2 * (size_t)slot);
        const float4 pa = own.a, pb = own.b;
        const Query3 q = list_query(g, pa.x, pa.y, pa.z, c.h);
        const int nj = q.j1 - q.j0 + 1, nrows = (q.i1 - q.i0 + 1) * nj;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, s5 = 0.f;
        if (nrows <= 32) {
            const int ncand = heavy_rows_setup(g, offset, q, nj, nrows, lane, sP, sG);
            int r = 0;
            for (int f0 = lane; f0 < ncand; f0 += 128) {
                int cand[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int f = f0 + 32 * u;
                    cand[u] = (f < ncand) ? heavy_candidate(sP, sG, f, r) : slot;        // past the end: the target itself, skipped below
                }
                f4x2 rec[4];
#pragma unroll
                for (int u = 0; u < 4; u++) rec[u] = cwa_ldg256(pack + 2 * (size_t)cand[u]);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float r2 = cwa_len3sq(pa.x - rec[u].a.x, pa.y - rec[u].a.y, pa.z - rec[u].a.z);
                    if (r2 <= c.accept_r2 && cand[u] != slot)
                        pair_force(c, pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, rec[u].a, rec[u].b, s0, s1, s2, s3, s4, s5);
                }
            }
        } else
        for (int r0 = 0; r0 < nrows; r0 += 32) {
            int g0 = 0, g1 = 0;
            if (r0 + lane < nrows) {
                const int r = r0 + lane;
                const int base = ((q.i0 + r / nj) * g.n[1] + (q.j0 + r % nj)) * g.kstride;
                g0 = __ldg(offset + base + q.k0); g1 = __ldg(offset + base + q.k1 + 1);
            }
            const int rows_here = min(32, nrows - r0);
            for (int r = 0; r < rows_here; r++) {
                const int b = __shfl_sync(0xffffffffu, g0, r), en = __shfl_sync(0xffffffffu, g1, r);
#pragma unroll 2
                for (int k = b + lane; k < en; k += 32) {
                    const f4x2 rec = cwa_ldg256(pack + 2 * (size_t)k);
                    const float r2 = cwa_len3sq(pa.x - rec.a.x, pa.y - rec.a.y, pa.z - rec.a.z);
                    if (r2 <= c.accept_r2 && k != slot)
                        pair_force(c, pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, rec.a, rec.b, s0, s1, s2, s3, s4, s5);
                }
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, d); s1 += __shfl_xor_sync(0xffffffffu, s1, d);
            s2 += __shfl_xor_sync(0xffffffffu, s2, d); s3 += __shfl_xor_sync(0xffffffffu, s3, d);
            s4 += __shfl_xor_sync(0xffffffffu, s4, d); s5 += __shfl_xor_sync(0xffffffffu, s5, d);
        }
        if (!FUSED) {
            if (lane == 0) {
                pairP[slot] = make_float4(s0, s1, s2, s3);
                pairV[slot] = make_float2(s4, s5);
            }
        } else {
            if (lane == kept) { k_slot = slot; k_pa = pa; k_pb = pb; k0 = s0; k1 = s1; k2 = s2; k3 = s3; k4 = s4; k5 = s5; }
            if (++kept == 32) {
                finish_particle<LOCAL>(c, fa, k_slot, k_pa, k_pb, k0, k1, k2, k3, k4, k5);
                kept = 0;
            }
        }
    };
    // Cells of ~h (every query is the 3 x 3 x 3 block: nine rows) and the separate integrate kernel: FOUR targets per warp, eight lanes each.
    // A clump target has a few hundred candidates -- three trips of a whole warp -- so with one target per warp half the instructions were
    // per-target overhead (set-up 17 %, candidate mapping 20 %, shuffle reduction 9 %: ncu at frame 3000, profiles/r2); four targets share
    // those instructions.  The divergent pair term costs the same (a trip still tests 128 candidates).  Targets of more than `sub_warp_max` (= `inplace_max`: the ones the density pass queued as well)
    // candidates (the few cells of hundreds of particles early in a run) keep the whole warp: eight lanes on thousands of candidates are a tail.
    {
        const float hm = c.h * (1.0f + 2.5e-3f), hx = c.h * 1.25f;
        const bool fast = hm <= g.cell[0] && hm <= g.cell[1] && hm <= g.cell[2] && (g.cell[0] <= hx || g.n[0] == 1) && (g.cell[1] <= hx || g.n[1] == 1) &&
                          (g.cell[2] <= hx || g.n[2] == 1);
        // (fewer targets than warps: a warp each finishes sooner; sub_warp_heavy == 2, the tests: whatever the queue length)
        if (!FUSED && fast && (sub_warp_heavy == 2 || (sub_warp_heavy == 1 && total > nwarps))) {
            constexpr int G = 8;
            const int gl = lane & (G - 1), grp = lane >> 3;
            const unsigned gmask = 0xffu << (grp * G);
            int* const gP = sP + grp * 8;                              // 4 groups x (<= 8 prefix entries) share the warp's 33 + 33 ints:
            int* const gG = sG + grp * 8;                              // entry r < 8 of the group in gP / gG, row 8 and the total in registers
            // (the warp's four targets are CONSECUTIVE queue entries -- neighbours in cell order, their candidates are the same cache lines;
            //  measured with the four a stride apart: frame 3040 of C4 536 us instead of 489)
            for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4; base < total; base += nwarps * 4) {
                const int e = base + grp;
                const bool valid = e < total;
                const int slot = __ldg(heavy_queue + (valid ? e : base));
                const f4x2 own = cwa_ldg256(pack + 2 * (size_t)slot);
                const float4 pa = own.a, pb = own.b;
                const int ci = approx_cell(pa.x, g.min[0], g.inv_cell[0], 0.0f, g.n[0]);
                const int cj = approx_cell(pa.y, g.min[1], g.inv_cell[1], 0.0f, g.n[1]);
                const int ck = approx_cell(pa.z, g.min[2], g.inv_cell[2], 0.0f, g.n[2]);
                const int kl = max(ck - 1, 0), kh = min(ck + 1, g.n[2] - 1);
                // rows 0..7: one per lane; row 8 (i + 1, j + 1): every lane loads it (same address: one transaction)
                auto row_bounds = [&](int r, int& b, int& len) {
                    const int i = ci + r / 3 - 1, j = cj + r % 3 - 1;
                    b = 0; len = 0;
                    if (i >= 0 && i < g.n[0] && j >= 0 && j < g.n[1]) {
                        const int rowbase = (i * g.n[1] + j) * g.kstride;
                        b = __ldg(offset + rowbase + kl);
                        len = __ldg(offset + rowbase + kh + 1) - b;
                    }
                };
                int b_own, len_own, b8, len8;
                row_bounds(gl, b_own, len_own);
                row_bounds(8, b8, len8);
                int incl = len_own;
#pragma unroll
                for (int d = 1; d < G; d <<= 1) {
                    const int t = __shfl_up_sync(gmask, incl, d, G);
                    if (gl >= d) incl += t;
                }
                const int total8 = __shfl_sync(gmask, incl, G - 1, G);          // candidates of rows 0..7
                const int ncand_all = total8 + len8;
                const bool big = valid && ncand_all > sub_warp_max;             // a handful of these per frame: the whole warp, below
                const int ncand = big ? 0 : ncand_all;
                __syncwarp(gmask);                                              // the group's reads of the previous target's table are done
                gP[gl] = incl - len_own; gG[gl] = b_own;
                __syncwarp(gmask);
                auto candidate = [&](int f, int& r) {                           // flat index -> cell-ordered slot; r only moves forward
                    if (f >= total8) return b8 + (f - total8);
                    while (r < G - 1 && f >= gP[r + 1]) r++;
                    return gG[r] + (f - gP[r]);
                };
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, s5 = 0.f;
                int r = 0;
                for (int fb = gl; fb < ncand; fb += 4 * G) {
                    int cand[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int f = fb + G * u;
                        cand[u] = (f < ncand) ? candidate(f, r) : slot;              // past the end: the target itself, skipped below
                    }
                    f4x2 rec[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) rec[u] = cwa_ldg256(pack + 2 * (size_t)cand[u]);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const float r2 = cwa_len3sq(pa.x - rec[u].a.x, pa.y - rec[u].a.y, pa.z - rec[u].a.z);
                        if (r2 <= c.accept_r2 && cand[u] != slot)
                            pair_force(c, pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, rec[u].a, rec[u].b, s0, s1, s2, s3, s4, s5);
                    }
                }
#pragma unroll
                for (int d = G / 2; d >= 1; d >>= 1) {
                    s0 += __shfl_xor_sync(gmask, s0, d, G); s1 += __shfl_xor_sync(gmask, s1, d, G);
                    s2 += __shfl_xor_sync(gmask, s2, d, G); s3 += __shfl_xor_sync(gmask, s3, d, G);
                    s4 += __shfl_xor_sync(gmask, s4, d, G); s5 += __shfl_xor_sync(gmask, s5, d, G);
                }
                if (gl == 0 && valid && !big) {
                    pairP[slot] = make_float4(s0, s1, s2, s3);
                    pairV[slot] = make_float2(s4, s5);
                }
                unsigned bigm = __ballot_sync(0xffffffffu, big && gl == 0);
                for (; bigm; bigm &= bigm - 1) {                                // (warp-uniform; whole_warp syncs the warp around its table)
                    const int slot_b = __shfl_sync(0xffffffffu, slot, __ffs(bigm) - 1);
                    __syncwarp();
                    whole_warp(slot_b);
                    __syncwarp();
                }
            }
            return;
        }
    }
    for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total; e += nwarps) whole_warp(__ldg(heavy_queue + e));
    if (FUSED && lane < kept) finish_particle<LOCAL>(c, fa, k_slot, k_pa, k_pb, k0, k1, k2, k3, k4, k5);
}

// force epilogue + integrate on the cell-ordered snapshot; one thread per particle writes the full
// 64-B record back to the particle SSBO in its original order.
// AHEAD (frames in the middle of one cwa_coupled_step call): the thread also does the NEXT frame's cell hash + count on
// the position it just wrote -- the same cwa_cell3 on the same stored floats as grid_hash_count_kernel, warp-aggregated
// (threads are in cell order, so most lanes of a warp share a few cells) -- and leaves cell id and arrival rank indexed
// by its slot: the next grid build starts at the scan and never re-reads the particle records for the hash.
// MODE 2 / 3 (slab decomposition, cwa_sph_step_slab; 3 = with count-ahead): an OWNED particle (original slot < n_owned) that ends the frame within `band` of a
// slab face, or beyond it, is also copied into the message for that neighbour -- the selection, message layout and dead-slot marking
// of slab_pack_kernel (multi.cu), applied to the record in registers instead of a second pass over the SSBO.
#define CWA_DEAD_W_SPH (-1.0f)
template <bool LOCAL, int MODE>
__global__ void __launch_bounds__(128, 10)          // latency-bound gathers and scatters: favour occupancy over registers
sph3_finalize_integrate_sorted_kernel(const float4* __restrict__ pack,
                                      const float4* __restrict__ forceS, const float4* __restrict__ miscS,
                                      const float4* __restrict__ pairP, const float2* __restrict__ pairV,
                                      const int* __restrict__ index_list, const int* __restrict__ count, float4* __restrict__ aos,
                                      const Sph3Const* __restrict__ cc, TexView tex,
                                      GridView g, int n, int* __restrict__ counter, int* __restrict__ cell_next, int* __restrict__ rank_next,
                                      SlabPackArgs sp)
{
    CWA_PDL_ENTER();
    constexpr bool AHEAD = (MODE == 1 || MODE == 3);  // 3: slab pack AND count-ahead (the particles that stay owned; arrivals are counted by the unpack kernel)
    constexpr bool SLAB = (MODE == 2 || MODE == 3);
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < __ldg(count);
    if (!AHEAD && !live) return;
    int cell = -2;                                  // -2: slot beyond the inserted particles, -1: position became NaN
    if (live) {
        const Sph3Const c = *cc;
        const f4x2 ab = cwa_ldg256(pack + 2 * (size_t)s);
        const float4 a = ab.a, b = ab.b, m = __ldg(miscS + s), pp = __ldg(pairP + s);
        const float2 pv = __ldg(pairV + s);
        float4 f = force_epilogue<LOCAL>(c, tex, a.x, a.y, a.z, b.x, b.y, b.z, b.w, __ldg(forceS + s), pp.x, pp.y, pp.z, pp.w, pv.x, pv.y);
        float4 pos = make_float4(a.x, a.y, a.z, m.x), vel = make_float4(b.x, b.y, b.z, m.y);
        float rho = b.w, prs = a.w;
        integrate_particle<LOCAL>(c, tex, pos, vel, f, rho, prs);
        const int id = __ldg(index_list + s);
        float4* o = aos + (size_t)id * 4;                          // 64-byte record = two 256-bit stores (two full sectors)
        const float4 ex = make_float4(rho, prs, m.z, m.w);
        float4 pos_out = pos;
        bool stays = true;                                      // still in this rank's owned range after the frame (always, outside slab frames)
        if (SLAB) stays = id < (sp.n_owned_dev != nullptr ? __ldg(sp.n_owned_dev) : sp.n_owned);   // ghosts are dropped
        if (SLAB && stays) {
            const float z = pos.z;                                 // NaN z: every test below is false -> stays
            float4* msg = nullptr;
            bool migrate = false;
            if (sp.msg_l != nullptr && z < sp.z_lo + sp.band) { msg = sp.msg_l; migrate = z < sp.z_lo; }
            else if (sp.msg_r != nullptr && z >= sp.z_hi - sp.band) { msg = sp.msg_r; migrate = z >= sp.z_hi; }
            if (msg != nullptr) {
                int* hdr = reinterpret_cast<int*>(msg);
                const int mslot = atomicAdd(hdr + (migrate ? 0 : 1), 1);
                const int cap = migrate ? sp.cap_mig : sp.cap_ghost;
                if (mslot >= cap) {
                    atomicExch(hdr + 2, 1);                        // overflow: reported to the host, particle stays put
                } else {
                    float4* d = msg + 4 * (size_t)(1 + (migrate ? 0 : sp.cap_mig) + mslot);
                    cwa_stg256(d, pos, vel);
                    cwa_stg256(d + 2, f, ex);
                    if (migrate) {
                        stays = false;
                        pos_out = make_float4(__int_as_float(0x7fffffff), __int_as_float(0x7fffffff), __int_as_float(0x7fffffff), CWA_DEAD_W_SPH);
                        if (sp.free_list != nullptr) sp.free_list[atomicAdd(sp.free_count, 1)] = id;   // the slot is reused by the next unpack
                    }
                }
            }
        }
        cwa_stg256(o, pos_out, vel);
        cwa_stg256(o + 2, f, ex);
        if (AHEAD) {
            cell = -1;
            if (stays && pos.x == pos.x && pos.y == pos.y && pos.z == pos.z) { // as grid_hash_count_kernel<3>
                int ci, cj, ck;
                cwa_cell3(g, pos.x, pos.y, pos.z, ci, cj, ck);
                cell = (ci * g.n[1] + cj) * g.kstride + ck;
            }
        }
    }
    if (AHEAD) {
        // Warp aggregation by RUNS: the threads are in cell order and a particle moves a fraction of a cell per frame, so lanes
        // that end up in the same cell are almost always neighbours.  One shuffle + one ballot find the runs of equal cells, the
        // first lane of a run adds the run length to the cell's counter and the others take base + position in the run (a cell
        // that shows up in two separate runs simply gets two atomics).  Cheaper than __match_any_sync, which loops over the
        // distinct values of the warp.
        const unsigned lane = threadIdx.x & 31u;
        const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0u || cell != prev);
        const int start = 31 - __clz(heads & (0xffffffffu >> (31u - lane)));           // first lane of my run
        const unsigned after = heads & ~(0xffffffffu >> (31u - lane));                  // run heads behind me
        const int end = after ? (__ffs(after) - 1) : 32;                                // one past the last lane of my run
        int base = 0;
        if (cell >= 0 && (int)lane == start) base = atomicAdd(&counter[cell], end - start);
        base = __shfl_sync(0xffffffffu, base, start);
        if (s < n) {
            cell_next[s] = cell;
            rank_next[s] = base + ((int)lane - start);
        }
    }
}

// individually dispatched passes: commit one pass result to the SSBO
__global__ void __launch_bounds__(256)
sph3_scatter_rho_pres_kernel(const float4* __restrict__ pack,
                             const int* __restrict__ index_list, const int* __restrict__ count, float4* __restrict__ aos)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= __ldg(count)) return;
    const int i = __ldg(index_list + s);
    float2* e = reinterpret_cast<float2*>(aos + (size_t)i * 4 + 3);
    *e = make_float2(__ldg(pack + 2 * (size_t)s + 1).w, __ldg(pack + 2 * (size_t)s).w);      // extras[0] = rho, extras[1] = pressure
}

__global__ void __launch_bounds__(256)
sph3_finalize_force_sorted_kernel(const float4* __restrict__ pack,
                                  const float4* __restrict__ forceS, const float4* __restrict__ pairP,
                                  const float2* __restrict__ pairV, const int* __restrict__ index_list, const int* __restrict__ count,
                                  float4* __restrict__ aos, const Sph3Const* __restrict__ cc, TexView tex)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= __ldg(count)) return;
    const Sph3Const c = *cc;
    const f4x2 ab = cwa_ldg256(pack + 2 * (size_t)s);
    const float4 a = ab.a, b = ab.b, pp = __ldg(pairP + s);
    const float2 pv = __ldg(pairV + s);
    const float4 f = force_epilogue(c, tex, a.x, a.y, a.z, b.x, b.y, b.z, b.w, __ldg(forceS + s), pp.x, pp.y, pp.z, pp.w, pv.x, pv.y);
    aos[(size_t)__ldg(index_list + s) * 4 + 2] = f;
}

// integrate directly on the SSBO (original order); used when the pass is dispatched on its own
__global__ void __launch_bounds__(256)
sph3_integrate_aos_kernel(float4* __restrict__ aos, int n, const Sph3Const* __restrict__ cc, TexView tex)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Sph3Const c = *cc;
    float4* o = aos + (size_t)i * 4;
    float4 pos = o[0], vel = o[1], f = o[2];
    const float4 e = o[3];
    float rho = e.x, prs = e.y;
    integrate_particle(c, tex, pos, vel, f, rho, prs);
    o[0] = pos; o[1] = vel; o[2] = f; o[3] = make_float4(rho, prs, e.z, e.w);
}

// ---------------------------------------------------------------------------------------------
// all-pairs kernels (the shipped O(N^2) loops): TT targets x 4 candidate quarters per CTA
// ---------------------------------------------------------------------------------------------
constexpr int AP_TT = 64;        // targets per CTA
constexpr int AP_L = 4;          // lanes per target (each scans one quarter of every tile)
constexpr int AP_TILE = 256;     // candidates per shared-memory tile

__global__ void __launch_bounds__(AP_TT * AP_L)
sph3_density_allpairs_kernel(const float4* __restrict__ aos, int n, const Sph3Const* __restrict__ cc, TexView tex,
                             float2* __restrict__ out_rp)
{
    __shared__ float4 tile[AP_TILE];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * AP_TT + tid / AP_L, sub = tid % AP_L;
    const float accept_r2 = cc->accept_r2, h2 = cc->h2, poly6 = cc->poly6;
    const bool active = i < n;
    const float4 p = active ? aos[(size_t)i * 4] : make_float4(0.f, 0.f, 0.f, 0.f);
    float rho = 0.0f;
    for (int j0 = 0; j0 < n; j0 += AP_TILE) {
        const int j = j0 + tid;
        tile[tid] = (j < n) ? aos[(size_t)j * 4] : make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);
        __syncthreads();
#pragma unroll 8
        for (int q = sub; q < AP_TILE; q += AP_L) pair_density(accept_r2, h2, poly6, p.x, p.y, p.z, tile[q], rho);
        __syncthreads();
    }
#pragma unroll
    for (int d = 1; d < AP_L; d <<= 1) rho += __shfl_xor_sync(0xffffffffu, rho, d);
    if (active && sub == 0) {
        float rho_out, prs_out;
        density_epilogue(*cc, tex, p.x, p.z, rho, rho_out, prs_out);
        out_rp[i] = make_float2(rho_out, prs_out);     // committed to the SSBO after the pass
    }
}

__global__ void __launch_bounds__(256)
sph3_commit_rho_pres_kernel(const float2* __restrict__ rp, int n, float4* __restrict__ aos)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    *reinterpret_cast<float2*>(aos + (size_t)i * 4 + 3) = rp[i];
}

__global__ void __launch_bounds__(AP_TT * AP_L)
sph3_force_allpairs_kernel(const float4* __restrict__ aos, int n, const Sph3Const* __restrict__ cc, TexView tex,
                           float4* __restrict__ out_force)
{
    __shared__ float4 tileA[AP_TILE];
    __shared__ float4 tileB[AP_TILE];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * AP_TT + tid / AP_L, sub = tid % AP_L;
    const Sph3Const c = *cc;
    const bool active = i < n;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), v = p, e = p;
    if (active) { p = aos[(size_t)i * 4]; v = aos[(size_t)i * 4 + 1]; e = aos[(size_t)i * 4 + 3]; }
    float fpx = 0.f, fpy = 0.f, fpz = 0.f, fvx = 0.f, fvy = 0.f, fvz = 0.f;
    for (int j0 = 0; j0 < n; j0 += AP_TILE) {
        const int j = j0 + tid;
        if (j < n) {
            const float4 qp = aos[(size_t)j * 4], qv = aos[(size_t)j * 4 + 1], qe = aos[(size_t)j * 4 + 3];
            tileA[tid] = make_float4(qp.x, qp.y, qp.z, qe.y);
            tileB[tid] = make_float4(qv.x, qv.y, qv.z, qe.x);
        } else {
            tileA[tid] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);
            tileB[tid] = make_float4(0.f, 0.f, 0.f, 1.f);
        }
        __syncthreads();
        // same two-phase scheme as the grid kernel: 64 candidates per lane and tile
        unsigned long long mask = 0ull;
#pragma unroll 8
        for (int t = 0; t < AP_TILE / AP_L; t++) {
            const float4 qa = tileA[sub + t * AP_L];
            const float r2 = cwa_len3sq(p.x - qa.x, p.y - qa.y, p.z - qa.z);
            if (r2 <= c.accept_r2 && (j0 + sub + t * AP_L) != i) mask |= (1ull << t);
        }
        while (mask) {
            const int t = __ffsll((long long)mask) - 1;
            mask &= mask - 1ull;
            const int q = sub + t * AP_L;
            pair_force(c, p.x, p.y, p.z, e.y, v.x, v.y, v.z, tileA[q], tileB[q], fpx, fpy, fpz, fvx, fvy, fvz);
        }
        __syncthreads();
    }
#pragma unroll
    for (int d = 1; d < AP_L; d <<= 1) {
        fpx += __shfl_xor_sync(0xffffffffu, fpx, d); fpy += __shfl_xor_sync(0xffffffffu, fpy, d);
        fpz += __shfl_xor_sync(0xffffffffu, fpz, d); fvx += __shfl_xor_sync(0xffffffffu, fvx, d);
        fvy += __shfl_xor_sync(0xffffffffu, fvy, d); fvz += __shfl_xor_sync(0xffffffffu, fvz, d);
    }
    if (active && sub == 0) {
        const float4 fprev = aos[(size_t)i * 4 + 2];
        out_force[i] = force_epilogue(c, tex, p.x, p.y, p.z, v.x, v.y, v.z, e.x, fprev, fpx, fpy, fpz, fvx, fvy, fvz);
    }
}

// SM-balanced, register-tiled all-pairs kernels (used when the targets of one SM fit one CTA: the shipped scene has 138 per SM).
// The kernels above cut the targets into tiles of 64: 20 480 targets = 320 CTAs on 148 SMs = 2.16 per SM, i.e. the SMs that get three
// CTAs set the time and the rest idle a third of it; and every pair test pays its own shared-memory load and loop step (16 instructions
// per test: the kernels are issue-bound).  Here ONE CTA per SM owns ceil(n / SMs) consecutive targets; a thread tests every candidate
// it reads against T targets held in registers, and L lanes share a group of T targets (L need not be a power of two: the partial sums
// meet in shared memory, fixed order).  The density accumulation is branch-free (weight 0 when rejected).
constexpr int APB_THREADS = 1024;
#ifndef CWA_APB_TILE_D
#define CWA_APB_TILE_D 1024
#endif
#ifndef CWA_APB_TILE
#define CWA_APB_TILE 512
#endif
#ifndef CWA_APB_BOX
#define CWA_APB_BOX 512
#endif
constexpr int APB_TILE_D = CWA_APB_TILE_D;      // candidates per tile, density pass (a barrier pair per tile: few, long tiles)
constexpr int APB_TILE = CWA_APB_TILE;         // candidates per tile, force pass: at most 64 per lane (its accept mask) for L >= 8
constexpr int APB_TD = 4;             // targets per thread, density pass
constexpr int APB_TF = 2;             // targets per thread, force pass

// Tile boxes.  Both passes cut the candidates into tiles of consecutive particles, and consecutive particles of the shipped scene are
// neighbours in space (a lattice in id order that deforms slowly), as are a CTA's ~138 consecutive targets: most (CTA, tile) combinations cannot
// hold a pair within h.  A small kernel computes the bounding box of every APB_BOX particles (NaN coordinates left out: they are never accepted);
// a CTA skips -- without loading it -- every tile whose box is farther than h from the box of its own targets, and a thread skips the tiles
// farther than h from each of its T targets.  Exact: a skipped tile holds no accepted pair (margin 1 % on h^2, a thousand times the rounding
// of the distance), so sums and summation order are those of the unculled kernels; in the worst case (fully mixed particles) nothing is
// skipped and the cost is the box tests.
constexpr int APB_BOX = CWA_APB_BOX;          // particles per box; both tile sizes are multiples
__global__ void __launch_bounds__(128)
sph3_allpairs_boxes_kernel(const float4* __restrict__ aos, int n, float4* __restrict__ boxes)
{
    __shared__ float red[6][4];
    const int tid = threadIdx.x;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int q = tid; q < APB_BOX; q += 128) {
        const int j = blockIdx.x * APB_BOX + q;
        if (j < n) {
            const float4 p = aos[(size_t)j * 4];
            lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);      // fminf / fmaxf drop a NaN operand
            hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], d));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], d));
        }
    if ((tid & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { red[a][tid >> 5] = lo[a]; red[3 + a][tid >> 5] = hi[a]; }
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++)
            for (int w = 1; w < 4; w++) { lo[a] = fminf(lo[a], red[a][w]); hi[a] = fmaxf(hi[a], red[3 + a][w]); }
        boxes[2 * blockIdx.x] = make_float4(lo[0], lo[1], lo[2], 0.f);
        boxes[2 * blockIdx.x + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
}

struct Box3 { float lo[3], hi[3]; };
// box of the particles [j0, j0 + count) from the APB_BOX-granular boxes (count a multiple of APB_BOX)
__device__ __forceinline__ Box3 apb_tile_box(const float4* __restrict__ boxes, int j0, int count, int n)
{
    Box3 b;
#pragma unroll
    for (int a = 0; a < 3; a++) { b.lo[a] = CUDART_INF_F; b.hi[a] = -CUDART_INF_F; }
    for (int t = j0 / APB_BOX; t < (j0 + count) / APB_BOX && t * APB_BOX < n; t++) {
        const float4 l = __ldg(boxes + 2 * t), h = __ldg(boxes + 2 * t + 1);
        b.lo[0] = fminf(b.lo[0], l.x); b.lo[1] = fminf(b.lo[1], l.y); b.lo[2] = fminf(b.lo[2], l.z);
        b.hi[0] = fmaxf(b.hi[0], h.x); b.hi[1] = fmaxf(b.hi[1], h.y); b.hi[2] = fmaxf(b.hi[2], h.z);
    }
    return b;
}
// squared distance between two boxes / a point and a box; anything NaN compares false below: not skipped
__device__ __forceinline__ float apb_box_box_d2(const Box3& a, const Box3& b)
{
    float d2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) { const float d = fmaxf(fmaxf(a.lo[k] - b.hi[k], b.lo[k] - a.hi[k]), 0.0f); d2 = fmaf(d, d, d2); }
    return d2;
}
__device__ __forceinline__ float apb_point_box_d2(float x, float y, float z, const Box3& b)
{
    const float p[3] = {x, y, z};
    float d2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) { const float d = fmaxf(fmaxf(b.lo[k] - p[k], p[k] - b.hi[k]), 0.0f); d2 = fmaf(d, d, d2); }
    return d2;
}
// box of the CTA's own targets [t_first, t_end): every thread gets the same six numbers (one block reduction at kernel start)
__device__ __forceinline__ Box3 apb_cta_box(const float4* __restrict__ aos, int t_first, int t_end, float (*red)[32])
{
    const int tid = threadIdx.x;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int t = t_first + tid; t < t_end; t += blockDim.x) {
        const float4 p = aos[(size_t)t * 4];
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], d));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], d));
        }
    if ((tid & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { red[a][tid >> 5] = lo[a]; red[3 + a][tid >> 5] = hi[a]; }
    }
    __syncthreads();
    Box3 b;
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float l = CUDART_INF_F, h = -CUDART_INF_F;
        for (int w = 0; w < nw; w++) { l = fminf(l, red[a][w]); h = fmaxf(h, red[3 + a][w]); }
        b.lo[a] = l; b.hi[a] = h;
    }
    __syncthreads();
    return b;
}

template <int T>
__global__ void __launch_bounds__(APB_THREADS)
sph3_density_allpairs_bal_kernel(const float4* __restrict__ aos, int n, int tpc, int L, const Sph3Const* __restrict__ cc, TexView tex,
                                 float2* __restrict__ out_rp, const float4* __restrict__ boxes)
{
    __shared__ float4 tile[APB_TILE_D];
    __shared__ float part[T][APB_THREADS];
    __shared__ float red[6][32];
    const int tid = threadIdx.x;
    const int grp = tid / L, sub = tid - grp * L;
    const int t0 = blockIdx.x * tpc + grp * T;                    // first target of the thread's group
    const int t_end = min(n, (blockIdx.x + 1) * tpc);             // one past the CTA's last target
    const bool active = grp * T < tpc && t0 < t_end;
    const float accept_r2 = cc->accept_r2, h2 = cc->h2, poly6 = cc->poly6;
    float px[T], py[T], pz[T], rho[T];
#pragma unroll
    for (int k = 0; k < T; k++) {
        const bool on = active && t0 + k < t_end;
        const float4 p = on ? aos[(size_t)(t0 + k) * 4] : make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);   // parked: every test fails
        px[k] = p.x; py[k] = p.y; pz[k] = p.z; rho[k] = 0.0f;
    }
    const float cull_r2 = h2 * 1.01f;
    const bool cull = boxes != nullptr;
    Box3 mine;
    if (cull) mine = apb_cta_box(aos, blockIdx.x * tpc, t_end, red);
    for (int j0 = 0; j0 < n; j0 += APB_TILE_D) {
        bool near = active;
        if (cull) {
            const Box3 tb = apb_tile_box(boxes, j0, APB_TILE_D, n);
            if (apb_box_box_d2(mine, tb) > cull_r2) continue;                 // uniform over the CTA: no load, no barrier
            bool any = false;
#pragma unroll
            for (int k = 0; k < T; k++) any = any || !(apb_point_box_d2(px[k], py[k], pz[k], tb) > cull_r2);
            near = active && any;
        }
        if (tid < APB_TILE_D) {
            const int j = j0 + tid;
            tile[tid] = (j < n) ? aos[(size_t)j * 4] : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, 0.f);
        }
        __syncthreads();
        if (near) {
#pragma unroll 2
            for (int q = sub; q < APB_TILE_D; q += L) {
                const float4 c = tile[q];
#pragma unroll
                for (int k = 0; k < T; k++) {
                    const float r2 = cwa_len3sq(px[k] - c.x, py[k] - c.y, pz[k] - c.z);
                    const float d = (r2 <= accept_r2) ? (h2 - r2) : 0.0f;
                    rho[k] = fmaf(poly6, d * d * d, rho[k]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < T; k++) part[k][tid] = rho[k];
    __syncthreads();
    if (active && sub < T && t0 + sub < t_end) {                  // lane k of the group finishes target k
        float sum = 0.0f;
        for (int l = 0; l < L; l++) sum += part[sub][tid - sub + l];
        const float4 p = aos[(size_t)(t0 + sub) * 4];
        float rho_out, prs_out;
        density_epilogue(*cc, tex, p.x, p.z, sum, rho_out, prs_out);
        out_rp[t0 + sub] = make_float2(rho_out, prs_out);         // committed to the SSBO after the pass
    }
}

template <int T>
__global__ void __launch_bounds__(APB_THREADS)
sph3_force_allpairs_bal_kernel(const float4* __restrict__ aos, int n, int tpc, int L, const Sph3Const* __restrict__ cc, TexView tex,
                               float4* __restrict__ out_force, const float4* __restrict__ boxes)
{
    __shared__ float4 tileA[APB_TILE];
    __shared__ float4 tileB[APB_TILE];
    __shared__ float part[6][APB_THREADS];             // reused for each of the thread's T targets; first the CTA's box reduction
    const int tid = threadIdx.x;
    const int grp = tid / L, sub = tid - grp * L;
    const int t0 = blockIdx.x * tpc + grp * T;
    const int t_end = min(n, (blockIdx.x + 1) * tpc);
    const bool active = grp * T < tpc && t0 < t_end;
    const Sph3Const c = *cc;
    float4 p[T], v[T];
    float prs[T], acc[T][6];
#pragma unroll
    for (int k = 0; k < T; k++) {
        const bool on = active && t0 + k < t_end;
        p[k] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f); v[k] = make_float4(0.f, 0.f, 0.f, 0.f); prs[k] = 0.f;
        if (on) { p[k] = aos[(size_t)(t0 + k) * 4]; v[k] = aos[(size_t)(t0 + k) * 4 + 1]; prs[k] = aos[(size_t)(t0 + k) * 4 + 3].y; }
#pragma unroll
        for (int a = 0; a < 6; a++) acc[k][a] = 0.f;
    }
    const int per_lane = (APB_TILE + L - 1) / L;       // <= 64: the launcher keeps L >= 8
    const float cull_r2 = c.h2 * 1.01f;
    const bool cull = boxes != nullptr;
    Box3 mine;
    if (cull) mine = apb_cta_box(aos, blockIdx.x * tpc, t_end, reinterpret_cast<float (*)[32]>(&part[0][0]));
    for (int j0 = 0; j0 < n; j0 += APB_TILE) {
        bool near = active;
        if (cull) {
            const Box3 tb = apb_tile_box(boxes, j0, APB_TILE, n);
            if (apb_box_box_d2(mine, tb) > cull_r2) continue;                 // uniform over the CTA: no load, no barrier
            bool any = false;
#pragma unroll
            for (int k = 0; k < T; k++) any = any || !(apb_point_box_d2(p[k].x, p[k].y, p[k].z, tb) > cull_r2);
            near = active && any;
        }
        if (tid < APB_TILE) {
            const int j = j0 + tid;
            if (j < n) {
                const float4 qp = aos[(size_t)j * 4], qv = aos[(size_t)j * 4 + 1], qe = aos[(size_t)j * 4 + 3];
                tileA[tid] = make_float4(qp.x, qp.y, qp.z, qe.y);
                tileB[tid] = make_float4(qv.x, qv.y, qv.z, qe.x);
            } else {
                tileA[tid] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, 0.f);
                tileB[tid] = make_float4(0.f, 0.f, 0.f, 1.f);
            }
        }
        __syncthreads();
        if (near) {
            // two phases: mark the accepted candidates of the lane for each of its targets (cheap, every candidate, one shared-memory read
            // for all T of them), then evaluate only those
            unsigned long long mask[T];
#pragma unroll
            for (int k = 0; k < T; k++) mask[k] = 0ull;
#pragma unroll 2
            for (int t = 0; t < per_lane; t++) {
                const int q = sub + t * L;
                if (q < APB_TILE) {
                    const float4 qa = tileA[q];
#pragma unroll
                    for (int k = 0; k < T; k++) {
                        const float r2 = cwa_len3sq(p[k].x - qa.x, p[k].y - qa.y, p[k].z - qa.z);
                        if (r2 <= c.accept_r2 && (j0 + q) != t0 + k) mask[k] |= (1ull << t);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < T; k++) {
                unsigned long long m = mask[k];
                while (m) {
                    const int t = __ffsll((long long)m) - 1;
                    m &= m - 1ull;
                    const int q = sub + t * L;
                    pair_force(c, p[k].x, p[k].y, p[k].z, prs[k], v[k].x, v[k].y, v[k].z, tileA[q], tileB[q],
                               acc[k][0], acc[k][1], acc[k][2], acc[k][3], acc[k][4], acc[k][5]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < T; k++) {                                 // lane k of the group finishes target k
#pragma unroll
        for (int a = 0; a < 6; a++) part[a][tid] = acc[k][a];
        __syncthreads();
        if (active && sub == k && t0 + k < t_end) {
            const int i = t0 + k;
            float sm[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int l = 0; l < L; l++)
#pragma unroll
                for (int a = 0; a < 6; a++) sm[a] += part[a][tid - sub + l];
            const float4 pi = aos[(size_t)i * 4], vi = aos[(size_t)i * 4 + 1], ei = aos[(size_t)i * 4 + 3], fprev = aos[(size_t)i * 4 + 2];
            out_force[i] = force_epilogue(c, tex, pi.x, pi.y, pi.z, vi.x, vi.y, vi.z, ei.x, fprev, sm[0], sm[1], sm[2], sm[3], sm[4], sm[5]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
sph3_commit_force_kernel(const float4* __restrict__ f, int n, float4* __restrict__ aos)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    aos[(size_t)i * 4 + 2] = f[i];
}

// neighbour counts for the "neighbour set identical" assertions
__global__ void __launch_bounds__(256)
sph3_neighbour_count_grid_kernel(const float4* __restrict__ posS, const int* __restrict__ index_list, const int* __restrict__ count,
                                 GridView g, const int* __restrict__ offset, const Sph3Const* __restrict__ cc,
                                 int* __restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= __ldg(count)) return;
    const float h = cc->h, accept_r2 = cc->accept_r2;
    const float4 p = __ldg(posS + s);
    // exact reference range (ComputeCellIndex of pos -+ h), deliberately NOT the conservative one
    int i0, j0, k0, i1, j1, k1;
    cwa_cell3(g, p.x - h, p.y - h, p.z - h, i0, j0, k0);
    cwa_cell3(g, p.x + h, p.y + h, p.z + h, i1, j1, k1);
    int cnt = 0;
    for (int i = i0; i <= i1; i++)
        for (int j = j0; j <= j1; j++) {
            const int base = (i * g.n[1] + j) * g.kstride;
            const int g0 = __ldg(offset + base + k0), g1 = __ldg(offset + base + k1 + 1);
            for (int q = g0; q < g1; q++) {
                const float4 o = __ldg(posS + q);
                if (cwa_len3sq(p.x - o.x, p.y - o.y, p.z - o.z) <= accept_r2) cnt++;
            }
        }
    out[__ldg(index_list + s)] = cnt;
}

__global__ void __launch_bounds__(256)
sph3_neighbour_count_allpairs_kernel(const float4* __restrict__ aos, int n, const Sph3Const* __restrict__ cc,
                                     int* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float accept_r2 = cc->accept_r2;
    const float4 p = aos[(size_t)i * 4];
    int cnt = 0;
    for (int j = 0; j < n; j++) {
        const float4 o = __ldg(aos + (size_t)j * 4);
        if (cwa_len3sq(p.x - o.x, p.y - o.y, p.z - o.z) <= accept_r2) cnt++;
    }
    out[i] = cnt;
}

// make_cube + init_particles, CoupledWaterAnimation/Main.cpp:735-776
__global__ void __launch_bounds__(256)
sph3_init_cube_kernel(float4* __restrict__ aos, int nx, int ny, int nz, ParamPtrs prm)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = nx * ny * nz;
    if (t >= n) return;
    const float spacing = __fmul_rn(__fmul_rn(prm.constants->smoothing_coeff, 0.85f), prm.sim->particle_radius);   // :739
    const int k = t % nz, j = (t / nz) % ny, i = t / (nz * ny);       // i outer, j, k inner :743-753
    aos[(size_t)t * 4 + 0] = make_float4(__fadd_rn(0.0f, __fmul_rn((float)i, spacing)), __fmul_rn((float)j, spacing),
                                         __fadd_rn(0.0f, __fmul_rn((float)k, spacing)), 1.0f);
    aos[(size_t)t * 4 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    aos[(size_t)t * 4 + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
    aos[(size_t)t * 4 + 3] = make_float4(prm.constants->resting_rho, 0.0f, 500.0f, 50.0f);   // :774
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
constexpr int DENS_CAP_MAX = 3072;        // staged slots (16 B each): at most 48 KB
constexpr int FORCE_CAP_MAX = 2048;       // staged slots (32 B each): at most 64 KB
constexpr int DENS_CAP_DEFAULT = 2048;    // 32 KB/CTA: measured best occupancy/staging trade-off of the lanes kernels on C4 (profiles/r1/tuning.md)
constexpr int FORCE_CAP_DEFAULT = 1536;   // 48 KB/CTA

static int env_int(const char* name, int dflt, int lo, int hi)
{
    const char* e = getenv(name);
    if (!e) return dflt;
    int v = atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}

// Tuning knobs of the neighbour kernels.  Defaults come from the environment (CWA_NB_CONFIG, CWA_NB_CAP_D / _F,
// CWA_FUSED_ORDER) the first time they are needed; cwa_set_tuning() overrides them at run time (used by the
// parity tests to cover every kernel variant in one process).
//   nb_config: 8 (default) = row-mask kernels, one target per thread (sph3_density_flat_kernel + sph3_force_rows_kernel);
//              7 = neighbour-list kernels with the round-1 density pass (row loop nest, four candidates per trip);
//              0..6 = "lanes" kernels (targets per CTA x lanes per target: 0: 128x4, 1: 128x2, 2: 64x4, 3: 256x1,
//              4: 128x1, 5: 64x2, 6: 256x2): neighbour rows staged in shared memory by TMA bulk copies, lanes of a
//              target combined with warp shuffles, the force pass scans the candidates again
//   cap_d / cap_f: staging budgets of the lanes kernels in slots (0 disables staging)
constexpr int NB_CONFIG_DEFAULT = 8;
constexpr int NBR_K_DEFAULT = 64;         // neighbour-list entries per target (self included); longer lists fall back to a grid scan
constexpr int NBR_K_MAX = 256;
static int nb_config(cwa_ctx* c) { if (c->tune.config < 0) c->tune.config = env_int("CWA_NB_CONFIG", NB_CONFIG_DEFAULT, 0, 8); return c->tune.config; }
static int dens_cap(cwa_ctx* c) { if (c->tune.cap_d < 0) c->tune.cap_d = env_int("CWA_NB_CAP_D", DENS_CAP_DEFAULT, 0, DENS_CAP_MAX); return c->tune.cap_d; }
static int force_cap(cwa_ctx* c) { if (c->tune.cap_f < 0) c->tune.cap_f = env_int("CWA_NB_CAP_F", FORCE_CAP_DEFAULT, 0, FORCE_CAP_MAX); return c->tune.cap_f; }
static bool fused_integrate(cwa_ctx* c) { if (c->tune.fused_integrate < 0) c->tune.fused_integrate = env_int("CWA_FUSED_INTEGRATE", 0, 0, 1); return c->tune.fused_integrate != 0; }
// pipeline (cwa_coupled_step with several frames per call): bit 0 = the wave stencil of frame f runs on a side stream next to the
// grid build of frame f+1; bit 1 = count-ahead (integrate of frame f does the cell hash + count of frame f+1)
static int pipeline_mode(cwa_ctx* c) { if (c->tune.pipeline < 0) c->tune.pipeline = env_int("CWA_PIPELINE", 3, 0, 3); return c->tune.pipeline; }
// nbr_k: list capacity per target (multiple of 4, <= 256); extreme: candidate count above which the density pass hands a target to a whole warp
static int nbr_k(cwa_ctx* c) { if (c->tune.nbr_k < 0) c->tune.nbr_k = env_int("CWA_NBR_K", NBR_K_DEFAULT, 8, NBR_K_MAX) & ~3; return c->tune.nbr_k; }
static int extreme_candidates(cwa_ctx* c) { if (c->tune.extreme < 0) c->tune.extreme = env_int("CWA_EXTREME", EXTREME_CANDIDATES, 16, 1 << 20); return c->tune.extreme; }
static int inplace_max(cwa_ctx* c) { if (c->tune.inplace_max < 0) c->tune.inplace_max = env_int("CWA_INPLACE_MAX", INPLACE_MAX, 0, 1 << 20); return c->tune.inplace_max; }
static int allpairs_balanced(cwa_ctx* c) { if (c->tune.allpairs_bal < 0) c->tune.allpairs_bal = env_int("CWA_ALLPAIRS_BALANCED", 2, 0, 2); return c->tune.allpairs_bal; }
static bool allpairs_cull(cwa_ctx* c) { if (c->tune.ap_cull < 0) c->tune.ap_cull = env_int("CWA_ALLPAIRS_CULL", 1, 0, 1); return c->tune.ap_cull != 0; }
static int heavy_sub_warp(cwa_ctx* c) { if (c->tune.heavy8 < 0) c->tune.heavy8 = env_int("CWA_HEAVY8", 1, 0, 2); return c->tune.heavy8; }
static bool fused_order(cwa_ctx* c) { if (c->tune.fused_order < 0) c->tune.fused_order = env_int("CWA_FUSED_ORDER", 1, 0, 1); return c->tune.fused_order != 0; }

extern "C" int cwa_set_tuning(cwa_ctx* ctx, const char* key, int value)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && key, "null argument");
    const std::string k(key);
    if (k == "nb_config") { CWA_CHECK(value >= 0 && value <= 8, "nb_config %d out of range", value); ctx->tune.config = value; }
    else if (k == "nb_cap_d") { CWA_CHECK(value >= 0 && value <= DENS_CAP_MAX, "nb_cap_d %d out of range", value); ctx->tune.cap_d = value; }
    else if (k == "nb_cap_f") { CWA_CHECK(value >= 0 && value <= FORCE_CAP_MAX, "nb_cap_f %d out of range", value); ctx->tune.cap_f = value; }
    else if (k == "fused_order") { ctx->tune.fused_order = value ? 1 : 0; }
    else if (k == "fused_integrate") { ctx->tune.fused_integrate = value ? 1 : 0; }
    else if (k == "nbr_k") { CWA_CHECK(value >= 8 && value <= NBR_K_MAX && value % 4 == 0, "nbr_k %d: multiple of 4 in [8, %d]", value, NBR_K_MAX); ctx->tune.nbr_k = value; }
    else if (k == "heavy_sub_warp") { ctx->tune.heavy8 = value < 0 ? 0 : (value > 2 ? 2 : value); }
    else if (k == "allpairs_cull") { ctx->tune.ap_cull = value ? 1 : 0; }
    else if (k == "pdl") { ctx->tune.pdl = value & 511; }
    else if (k == "slab_ahead") { ctx->tune.slab_ahead = value ? 1 : 0; }
    else if (k == "allpairs_balanced") { CWA_CHECK(value >= 0 && value <= 2, "allpairs_balanced %d: 0 off, 1 density pass, 2 both passes", value); ctx->tune.allpairs_bal = value; }
    else if (k == "inplace_max") { CWA_CHECK(value >= 0, "inplace_max %d negative", value); ctx->tune.inplace_max = value; }
    else if (k == "extreme_candidates") { CWA_CHECK(value >= 16, "extreme_candidates %d too small", value); ctx->tune.extreme = value; }
    else if (k == "wave_transpose") { ctx->tune.wave_transpose = value ? 1 : 0; }
    else if (k == "scan_config") { CWA_CHECK(value >= 0 && value <= 3, "scan_config %d out of range", value); ctx->tune.scan_config = value; }
    else if (k == "graph") { ctx->tune.graph = value ? 1 : 0; }
    else if (k == "pipeline") { CWA_CHECK(value >= 0 && value <= 3, "pipeline %d out of range", value); ctx->tune.pipeline = value; }
    else CWA_CHECK(false, "cwa_set_tuning: unknown key '%s'", key);
    return 0;
}

template <int P, int L>
static int launch_density(cwa_ctx* ctx, SphObj* s, GridObj* g, TexView tex)
{
    CWA_TRY(ensure_dynamic_smem(ctx, sph3_density_grid_kernel<P, L>, DENS_CAP_MAX * 16));
    KScope k(ctx, KID_DENSITY);
    sph3_density_grid_kernel<P, L><<<ceil_div(s->n, P), P * L, dens_cap(ctx) * 16, ctx->stream>>>(
        s->posS, s->velS, s->pack, s->n, g->view, g->offset, (const Sph3Const*)s->consts, tex, dens_cap(ctx));
    return 0;
}

template <int P, int L>
static int launch_force(cwa_ctx* ctx, SphObj* s, GridObj* g)
{
    CWA_TRY(ensure_dynamic_smem(ctx, sph3_force_grid_kernel<P, L>, FORCE_CAP_MAX * 32));
    KScope k(ctx, KID_FORCE);
    sph3_force_grid_kernel<P, L><<<ceil_div(s->n, P), P * L, force_cap(ctx) * 32, ctx->stream>>>(
        s->pack, s->pairP, s->pairV, s->n, g->view, g->offset, (const Sph3Const*)s->consts, force_cap(ctx));
    return 0;
}

// rows variant: one thread per target (CWA_NB_CONFIG 7: P = 128, 8: P = 64, 9: P = 256)

// heavy kernels: a fixed grid of warps walks the device-side queue (no host round trip)
static int heavy_grid(cwa_ctx* ctx) { return ctx->sm_count * 16; }     // x 4 warps: every resident warp slot of the GPU

// bytes of the neighbour structure of `n` targets: index lists [slot][K] (nb_config 7) or row masks [tile of 32][9 rows][lane] (nb_config 8)
static size_t nbr_bytes(int n, int K)
{
    const size_t per = (size_t)K * 4 > (size_t)RT_ROWS * 8 ? (size_t)K * 4 : (size_t)RT_ROWS * 8;
    return (((size_t)(n > 0 ? n : 1) + 31) / 32 * 32) * per;
}

static int launch_density_list(cwa_ctx* ctx, SphObj* s, GridObj* g, TexView tex, int variant)   // 0: index lists, 1: row masks
{
    const Sph3Const* cc = (const Sph3Const*)s->consts;
    const int ntiles = ceil_div(s->n, TILE_P);
    const bool local = tex_view_is_local(tex);
    int* const hc = s->heavy_cnt;                        // queue counter, zeroed by the reorder pass of this snapshot
    const int K = nbr_k(ctx);
    if (s->nbr_k_alloc < K) {                            // the list capacity was raised after the object was created
        CWA_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(s->nbr_list);
        CWA_CUDA(cudaMalloc(&s->nbr_list, nbr_bytes(s->capacity, K)));
        s->nbr_k_alloc = K;
    }
    s->nbr_k_used = K;
    const bool flat = variant != 0;
    s->nbr_rows_fmt = flat;
    if (flat) {
        KScope k(ctx, KID_DENSITY);
        int2* rows = reinterpret_cast<int2*>(s->nbr_list);
        if (local)
            cwa_launch(ctx, PDL_DENSITY, sph3_density_flat_kernel<true>, dim3(ntiles), dim3(TILE_P), 0,
                s->posS, s->xyzS, s->xyz_stride, s->velS, s->pack, rows, s->nbr_count, s->heavy_queue, hc, s->n, g->view, g->offset, cc, tex, extreme_candidates(ctx), inplace_max(ctx));
        else
            cwa_launch(ctx, PDL_DENSITY, sph3_density_flat_kernel<false>, dim3(ntiles), dim3(TILE_P), 0,
                s->posS, s->xyzS, s->xyz_stride, s->velS, s->pack, rows, s->nbr_count, s->heavy_queue, hc, s->n, g->view, g->offset, cc, tex, extreme_candidates(ctx), inplace_max(ctx));
    } else {
      KScope k(ctx, KID_DENSITY);
      if (local)
          sph3_density_list_kernel<true><<<ntiles, TILE_P, 0, ctx->stream>>>(
              s->posS, s->velS, s->pack, s->nbr_list, s->nbr_count, s->heavy_queue, hc, s->n, g->view, g->offset, cc, tex, K, extreme_candidates(ctx));
      else
          sph3_density_list_kernel<false><<<ntiles, TILE_P, 0, ctx->stream>>>(
              s->posS, s->velS, s->pack, s->nbr_list, s->nbr_count, s->heavy_queue, hc, s->n, g->view, g->offset, cc, tex, K, extreme_candidates(ctx)); }
    { KScope k(ctx, KID_HEAVY);
      if (local)
          cwa_launch(ctx, PDL_DENSITY_HEAVY, sph3_density_heavy_kernel<true>, dim3(heavy_grid(ctx)), dim3(128), 0,
              s->posS, s->velS, s->pack, s->nbr_count, s->heavy_queue, hc, s->n, g->view, g->offset, cc, tex);
      else
          cwa_launch(ctx, PDL_DENSITY_HEAVY, sph3_density_heavy_kernel<false>, dim3(heavy_grid(ctx)), dim3(128), 0,
              s->posS, s->velS, s->pack, s->nbr_count, s->heavy_queue, hc, s->n, g->view, g->offset, cc, tex); }
    s->nbr_lists_valid = true;
    return 0;
}

// fused = true: full step, the force kernels also run the force epilogue + integrate and write the SSBO records
static int launch_force_list(cwa_ctx* ctx, SphObj* s, GridObj* g, bool fused, float4* aos, TexView tex)
{
    int* fq = s->heavy_queue + s->capacity;              // second half: the force pass's queue
    const Sph3Const* cc = (const Sph3Const*)s->consts;
    const FinishArgs fa{s->forceS, s->miscS, g->index_list, aos, tex};
    const int blocks = ceil_div(s->n, TILE_P);
    const bool local = tex_view_is_local(tex);
#define CWA_FORCE_LIST(F, L) sph3_force_list_kernel<F, L><<<blocks, TILE_P, 0, ctx->stream>>>( \
        s->pack, s->nbr_list, s->nbr_count, fq, s->heavy_cnt + 1, s->pairP, s->pairV, s->n, g->view, g->offset, cc, fa, s->nbr_k_used)
#define CWA_FORCE_HEAVY(F, L) cwa_launch(ctx, PDL_FORCE_HEAVY, sph3_force_heavy_kernel<F, L>, dim3(heavy_grid(ctx)), dim3(128), 0, \
        s->pack, fq, s->heavy_cnt + 1, s->n, s->pairP, s->pairV, g->view, g->offset, cc, fa, heavy_sub_warp(ctx), inplace_max(ctx))
#define CWA_FORCE_ROWS(F, L) cwa_launch(ctx, PDL_FORCE, sph3_force_rows_kernel<F, L>, dim3(blocks), dim3(TILE_P), 0, \
        s->pack, reinterpret_cast<const int2*>(s->nbr_list), s->nbr_count, fq, s->heavy_cnt + 1, s->pairP, s->pairV, s->n, g->view, g->offset, cc, fa)
    if (s->nbr_rows_fmt) {
      KScope k(ctx, KID_FORCE);
      if (!fused) CWA_FORCE_ROWS(false, false); else if (local) CWA_FORCE_ROWS(true, true); else CWA_FORCE_ROWS(true, false);
    } else {
      KScope k(ctx, KID_FORCE);
      if (!fused) CWA_FORCE_LIST(false, false); else if (local) CWA_FORCE_LIST(true, true); else CWA_FORCE_LIST(true, false); }
    { KScope k(ctx, KID_HEAVY);
      if (!fused) CWA_FORCE_HEAVY(false, false); else if (local) CWA_FORCE_HEAVY(true, true); else CWA_FORCE_HEAVY(true, false); }
#undef CWA_FORCE_LIST
#undef CWA_FORCE_ROWS
#undef CWA_FORCE_HEAVY
    return 0;
}

static int sph_snapshot(cwa_ctx* ctx, SphObj* s, bool count_next_ahead = false, bool clear_for_next_build = false)
{
    GridObj* g = get_grid(ctx, s->grid);
    BufferObj* pb = get_buffer(ctx, s->particles);
    CWA_CHECK(g && pb, "sph: grid or particle buffer vanished");
    GridBuildOpts opts;
    opts.canonical_order = !fused_order(ctx);
    if (s->counts_ahead) {
        opts.ahead_cell = s->cell_next; opts.ahead_rank = s->rank_next;
        opts.arrivals = s->arrivals; opts.arrivals_count = s->arrivals_count; opts.arrivals_max = s->arrivals_max;
    }
    opts.clear_after_scan = count_next_ahead || clear_for_next_build;
    opts.clear_is_for_next_build = clear_for_next_build && !count_next_ahead;
    opts.n_dev = s->n_dev;
    s->counts_ahead = false;
    CWA_TRY(grid_build_internal(ctx, g, pb->ptr, 64, s->n, opts));
    if (fused_order(ctx)) {
        KScope k(ctx, KID_REORDER);
        cwa_launch(ctx, PDL_REORDER, sph3_order_reorder_kernel, dim3(ceil_div(s->n > 0 ? s->n : 1, 256)), dim3(256), 0,
            (const float4*)pb->ptr, g->arrival, g->cell_of, g->offset, g->offset + g->view.num_cells, g->index_list,
            s->posS, s->velS, s->forceS, s->miscS, s->heavy_cnt, s->xyzS, s->xyz_stride);
    } else {
        KScope k(ctx, KID_REORDER);
        sph3_reorder_kernel<<<ceil_div((long long)(s->n > 0 ? s->n : 1) * 4, 256), 256, 0, ctx->stream>>>(
            (const float4*)pb->ptr, g->index_list, g->offset + g->view.num_cells, s->posS, s->velS, s->forceS, s->miscS, s->heavy_cnt, s->xyzS, s->xyz_stride);
    }
    CWA_CUDA(cudaGetLastError());
    s->snapshot_valid = true;
    s->pair_sums_valid = false;
    s->nbr_lists_valid = false;
    return 0;
}

// Derived constants are recomputed only when a parameter block may have changed: every library call that writes or
// rebinds a UBO bumps ctx->params_epoch.  A block that lives in wrapped (application-owned) memory can change behind
// the library's back, so it is re-read every time.
static int sph_prepare(cwa_ctx* ctx, SphObj* s)
{
    bool external = false;
    for (int i = 1; i <= 4; i++) {
        BufferObj* o = get_buffer(ctx, ctx->ubo_binding[i]);
        if (o && !o->owned) external = true;
    }
    if (!external && s->consts_epoch == ctx->params_epoch) return 0;
    KScope k(ctx, KID_OTHER);
    sph3_prepare_kernel<<<1, 32, 0, ctx->stream>>>(current_params(ctx), (Sph3Const*)s->consts);
    s->consts_epoch = ctx->params_epoch;
    return 0;
}

static float2* sph_scratch_rp(SphObj* s) { return reinterpret_cast<float2*>(s->scratch); }
static float4* sph_scratch_force(SphObj* s) { return s->scratch; }

// which: bit0 rho_pres, bit1 force, bit2 integrate.  In grid mode a full step (7) keeps every
// intermediate in the cell-ordered snapshot and writes the SSBO once, at the end.
// count_ahead (set by cwa_coupled_step for frames that are followed by another frame of the same call): the integrate pass also
// hashes + counts the new positions for the next frame's grid build.
void sph_invalidate_for_buffer(cwa_ctx* ctx, cwa_buf particles)
{
    for (auto& s : ctx->sphs)
        if (s.live && s.particles == particles) { s.snapshot_valid = false; s.slab_packed.valid = false; s.counts_ahead = false; }
}

int sph_passes_internal(cwa_ctx* ctx, SphObj* s, TexView tex, int which, bool count_ahead, const SlabPackArgs* slab)
{
    s->slab_packed.valid = false;                                  // whatever was packed described the previous positions
    BufferObj* pb = get_buffer(ctx, s->particles);
    CWA_CHECK(pb, "sph: particle buffer vanished");
    float4* aos = (float4*)pb->ptr;
    const int n = s->n;
    if (n == 0) return 0;
    CWA_TRY(sph_prepare(ctx, s));
    const Sph3Const* cc = (const Sph3Const*)s->consts;

    if (s->grid < 0) {                                             // ---- all-pairs, as shipped
        const int blocks = ceil_div(n, AP_TT);
        // one CTA per SM with an equal share of the targets when that share fits a CTA with at least 4 lanes per target
        const int tpc = ceil_div(n, ctx->sm_count);
        auto lanes_for = [&](int T) { const int groups = ceil_div(tpc > 0 ? tpc : 1, T); const int l = APB_THREADS / groups; return l < 64 ? l : 64; };
        const int lanes_d = lanes_for(APB_TD), lanes_f = lanes_for(APB_TF);
        const bool balanced = tpc > 0 && lanes_d >= 4 && lanes_f >= 8 && allpairs_balanced(ctx) != 0;
        const int bal_blocks = balanced ? ceil_div(n, tpc) : 0;
        const bool cull = balanced && allpairs_cull(ctx);             // tile boxes (exact culling of far tiles)
        if (which & 1) {
            if (cull) { KScope k(ctx, KID_OTHER); sph3_allpairs_boxes_kernel<<<ceil_div(n, APB_BOX), 128, 0, ctx->stream>>>(aos, n, s->boxes); }
            { KScope k(ctx, KID_DENSITY);
              if (balanced) sph3_density_allpairs_bal_kernel<APB_TD><<<bal_blocks, APB_THREADS, 0, ctx->stream>>>(aos, n, tpc, lanes_d, cc, tex, sph_scratch_rp(s), cull ? s->boxes : nullptr);
              else sph3_density_allpairs_kernel<<<blocks, AP_TT * AP_L, 0, ctx->stream>>>(aos, n, cc, tex, sph_scratch_rp(s)); }
            { KScope k(ctx, KID_OTHER);
              sph3_commit_rho_pres_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(sph_scratch_rp(s), n, aos); }
        }
        if (which & 2) {
            const bool bal_force = balanced && ctx->tune.allpairs_bal == 2;
            if (cull && bal_force && !(which & 1)) {                 // (the pass dispatched alone: the boxes of the density pass are not there)
                KScope k(ctx, KID_OTHER); sph3_allpairs_boxes_kernel<<<ceil_div(n, APB_BOX), 128, 0, ctx->stream>>>(aos, n, s->boxes); }
            { KScope k(ctx, KID_FORCE);
              if (bal_force) sph3_force_allpairs_bal_kernel<APB_TF><<<bal_blocks, APB_THREADS, 0, ctx->stream>>>(aos, n, tpc, lanes_f, cc, tex, sph_scratch_force(s), cull ? s->boxes : nullptr);
              else sph3_force_allpairs_kernel<<<blocks, AP_TT * AP_L, 0, ctx->stream>>>(aos, n, cc, tex, sph_scratch_force(s)); }
            { KScope k(ctx, KID_OTHER);
              sph3_commit_force_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(sph_scratch_force(s), n, aos); }
        }
        if (which & 4) {
            KScope k(ctx, KID_INTEGRATE);
            sph3_integrate_aos_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(aos, n, cc, tex);
        }
        CWA_CUDA(cudaGetLastError());
        return 0;
    }

    // ---- grid mode
    GridObj* g = get_grid(ctx, s->grid);
    CWA_CHECK(g && g->dim == 3, "sph: the bound grid must be a 3-D grid");
    const bool full = (which == 7);
    const int cfg = nb_config(ctx);
    const bool lists = cfg >= 7;                                   // neighbour-list kernels (7: round-1 density pass, 8: flat density pass)
    const bool fused_tail = full && lists && fused_integrate(ctx);   // the force kernels finish the particle (epilogue + integrate + write-back)
    count_ahead = count_ahead && full && !fused_tail;                 // (slab frames too: the unpack kernel of the next frame counts the arrivals)
    CWA_CHECK(slab == nullptr || (full && !fused_tail), "slab pack: needs a full step with the separate integrate kernel (fused_integrate = 0)");
    if (!(which & 1)) CWA_TRY(wave_sampling_copy(ctx, s->wave, s->wave_image, &tex));
    if (which & 1) {
        // (slab frames that are followed by another frame of the call: the next build finds counter + scan state cleared, off its critical path)
        CWA_TRY(sph_snapshot(ctx, s, count_ahead, slab != nullptr && !count_ahead));   // positions changed since the last frame
        if (s->wait_before_sampling) {                             // the wave level this frame samples may still be in flight on the side stream
            CWA_CUDA(cudaStreamWaitEvent(ctx->stream, s->wait_before_sampling, 0));
            s->wait_before_sampling = nullptr;
        }
        CWA_TRY(wave_sampling_copy(ctx, s->wave, s->wave_image, &tex));   // transposed copy of the sampled level (rebuilt only when it changed)
        switch (cfg) {
        case 1: CWA_TRY((launch_density<128, 2>(ctx, s, g, tex))); break;
        case 2: CWA_TRY((launch_density<64, 4>(ctx, s, g, tex))); break;
        case 3: CWA_TRY((launch_density<256, 1>(ctx, s, g, tex))); break;
        case 4: CWA_TRY((launch_density<128, 1>(ctx, s, g, tex))); break;
        case 5: CWA_TRY((launch_density<64, 2>(ctx, s, g, tex))); break;
        case 6: CWA_TRY((launch_density<256, 2>(ctx, s, g, tex))); break;
        case 7: CWA_TRY(launch_density_list(ctx, s, g, tex, 0)); break;
        case 8: CWA_TRY(launch_density_list(ctx, s, g, tex, 1)); break;
        default: CWA_TRY((launch_density<128, 4>(ctx, s, g, tex))); break;
        }
        if (!full) {
            KScope k(ctx, KID_OTHER);
            sph3_scatter_rho_pres_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(s->pack, g->index_list, g->offset + g->view.num_cells, aos);
        }
    }
    if (which & 2) {
        CWA_CHECK(s->snapshot_valid, "force pass dispatched before a density pass built the cell-ordered snapshot");
        switch (cfg) {
        case 1: CWA_TRY((launch_force<128, 2>(ctx, s, g))); break;
        case 2: CWA_TRY((launch_force<64, 4>(ctx, s, g))); break;
        case 3: CWA_TRY((launch_force<256, 1>(ctx, s, g))); break;
        case 4: CWA_TRY((launch_force<128, 1>(ctx, s, g))); break;
        case 5: CWA_TRY((launch_force<64, 2>(ctx, s, g))); break;
        case 6: CWA_TRY((launch_force<256, 2>(ctx, s, g))); break;
        case 7: case 8:
            CWA_CHECK(s->nbr_lists_valid, "force pass: the neighbour lists of the density pass are missing");
            CWA_TRY(launch_force_list(ctx, s, g, fused_tail, aos, tex)); break;
        default: CWA_TRY((launch_force<128, 4>(ctx, s, g))); break;
        }
        s->pair_sums_valid = !fused_tail;
        if (!full) {
            KScope k(ctx, KID_OTHER);
            sph3_finalize_force_sorted_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(
                s->pack, s->forceS, s->pairP, s->pairV, g->index_list, g->offset + g->view.num_cells, aos, cc, tex);
        }
    }
    if ((which & 4) && fused_tail) {
        s->snapshot_valid = false;                                 // positions moved (inside the force kernels)
        s->pair_sums_valid = false;
    } else if (which & 4) {
        if (count_ahead) CWA_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_pipe[1], 0));   // counter cleared (side stream) after this frame's scan
        KScope k(ctx, KID_INTEGRATE);
        if (full) {
            const bool local = tex_view_is_local(tex);
            const SlabPackArgs spa = slab ? *slab : SlabPackArgs();
            const int mode = slab ? (count_ahead ? 3 : 2) : (count_ahead ? 1 : 0);
#define CWA_FIN_INT(L, M) cwa_launch(ctx, PDL_INTEGRATE, sph3_finalize_integrate_sorted_kernel<L, M>, dim3(ceil_div(n, 128)), dim3(128), 0, \
                    s->pack, s->forceS, s->miscS, s->pairP, s->pairV, g->index_list, g->offset + g->view.num_cells, aos, cc, tex, \
                    g->view, n, g->counter, s->cell_next, s->rank_next, spa)
            if (local) { if (mode == 3) CWA_FIN_INT(true, 3); else if (mode == 2) CWA_FIN_INT(true, 2); else if (mode == 1) CWA_FIN_INT(true, 1); else CWA_FIN_INT(true, 0); }
            else       { if (mode == 3) CWA_FIN_INT(false, 3); else if (mode == 2) CWA_FIN_INT(false, 2); else if (mode == 1) CWA_FIN_INT(false, 1); else CWA_FIN_INT(false, 0); }
#undef CWA_FIN_INT
            s->counts_ahead = count_ahead;
        } else {
            sph3_integrate_aos_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(aos, n, cc, tex);
        }
        s->snapshot_valid = false;                                 // positions moved
        s->pair_sums_valid = false;
    }
    CWA_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int cwa_sph_create(cwa_ctx* ctx, cwa_buf particles, int n, cwa_grid grid, cwa_sph* out)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && out, "null argument");
    *out = -1;
    BufferObj* pb = get_buffer(ctx, particles);
    CWA_CHECK(pb, "invalid particle buffer handle %d", particles);
    CWA_CHECK(n >= 0 && (size_t)n * sizeof(cwa_particle) <= pb->bytes, "cwa_sph_create: buffer holds fewer than %d particles", n);
    if (grid >= 0) {
        GridObj* g = get_grid(ctx, grid);
        CWA_CHECK(g && g->dim == 3, "cwa_sph_create: grid handle %d is not a 3-D grid", grid);
        CWA_CHECK(g->max_particles >= n, "cwa_sph_create: grid capacity %d < %d particles", g->max_particles, n);
    }
    SphObj s;
    s.live = true; s.particles = particles; s.n = n; s.capacity = n; s.grid = grid;
    const size_t bytes = (size_t)(n > 0 ? n : 1) * 16 + 64;          // + 4 slots: the neighbour loops read up to 3 slots past a row
    CWA_CUDA(cudaMalloc(&s.consts, sizeof(Sph3Const)));
    if (grid >= 0) CWA_CUDA(cudaMalloc(&s.pack, 2 * bytes));
    else {                                                        // pass results + the tile boxes of the all-pairs kernels
        const size_t box_off = (bytes + 31) / 32 * 32;
        CWA_CUDA(cudaMalloc(&s.scratch, box_off + ((size_t)(n > 0 ? n : 1) + APB_BOX - 1) / APB_BOX * 32));
        s.boxes = reinterpret_cast<float4*>(reinterpret_cast<char*>(s.scratch) + box_off);
    }
    if (grid >= 0) {
        CWA_CUDA(cudaMalloc(&s.posS, bytes));
        CWA_CUDA(cudaMemsetAsync(s.posS, 0, bytes, ctx->stream));     // the 4 pad slots are read (and masked out) by the neighbour loops
        s.xyz_stride = ((size_t)(n > 0 ? n : 1) + 4 + 31) / 32 * 32;    // x | y | z streams, each padded like posS and 128-byte aligned
        CWA_CUDA(cudaMalloc(&s.xyzS, 3 * s.xyz_stride * sizeof(float)));
        CWA_CUDA(cudaMemsetAsync(s.xyzS, 0, 3 * s.xyz_stride * sizeof(float), ctx->stream));
        CWA_CUDA(cudaMalloc(&s.velS, bytes));
        CWA_CUDA(cudaMalloc(&s.forceS, bytes));
        CWA_CUDA(cudaMalloc(&s.miscS, bytes));
        CWA_CUDA(cudaMalloc(&s.pairP, bytes));
        CWA_CUDA(cudaMalloc(&s.pairV, bytes / 2));
        s.nbr_k_alloc = nbr_k(ctx);
        CWA_CUDA(cudaMalloc(&s.nbr_list, nbr_bytes(n, s.nbr_k_alloc)));
        CWA_CUDA(cudaMalloc(&s.nbr_count, (size_t)(n > 0 ? n : 1) * 4));
        CWA_CUDA(cudaMalloc(&s.heavy_queue, (size_t)(n > 0 ? n : 1) * 2 * 4));
        CWA_CUDA(cudaMalloc(&s.heavy_cnt, 16));
        CWA_CUDA(cudaMemsetAsync(s.heavy_cnt, 0, 16, ctx->stream));
        CWA_CUDA(cudaMalloc(&s.cell_next, (size_t)(n > 0 ? n : 1) * 4));
        CWA_CUDA(cudaMalloc(&s.rank_next, (size_t)(n > 0 ? n : 1) * 4));
    }
    ctx->sphs.push_back(s);
    *out = (int)ctx->sphs.size() - 1;
    return 0;
}

extern "C" int cwa_sph_destroy(cwa_ctx* ctx, cwa_sph h)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, h);
    CWA_CHECK(s, "invalid sph handle %d", h);
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(s->pack); cudaFree(s->scratch); cudaFree(s->posS); cudaFree(s->xyzS); cudaFree(s->velS); cudaFree(s->forceS); cudaFree(s->miscS);
    cudaFree(s->pairP); cudaFree(s->pairV); cudaFree(s->consts); cudaFree(s->nbr_list); cudaFree(s->nbr_count); cudaFree(s->heavy_queue);
    cudaFree(s->heavy_cnt); cudaFree(s->cell_next); cudaFree(s->rank_next);
    for (auto& g : s->frame_graph) if (g.valid) { cudaGraphExecDestroy((cudaGraphExec_t)g.exec); g.valid = false; }
    s->live = false;
    if (ctx->bound_sph == h) ctx->bound_sph = -1;
    return 0;
}

extern "C" int cwa_sph_bind_wave(cwa_ctx* ctx, cwa_sph h, cwa_wave w, int image)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, h);
    CWA_CHECK(s, "invalid sph handle %d", h);
    CWA_CHECK(w == -1 || get_wave(ctx, w), "invalid wave handle %d", w);
    CWA_CHECK(image >= -1 && image < 3, "image index %d out of range", image);
    s->wave = w; s->wave_image = (w == -1) ? -1 : image;
    return 0;
}

static TexView sph_current_tex(cwa_ctx* ctx, SphObj* s)
{
    return wave_tex_view(ctx, s->wave, s->wave_image);
}

extern "C" int cwa_sph_rho_pres(cwa_ctx* ctx, cwa_sph h)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, h);
    CWA_CHECK(s, "invalid sph handle %d", h);
    return sph_passes_internal(ctx, s, sph_current_tex(ctx, s), 1);
}

extern "C" int cwa_sph_force(cwa_ctx* ctx, cwa_sph h)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, h);
    CWA_CHECK(s, "invalid sph handle %d", h);
    return sph_passes_internal(ctx, s, sph_current_tex(ctx, s), 2);
}

extern "C" int cwa_sph_integrate(cwa_ctx* ctx, cwa_sph h)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, h);
    CWA_CHECK(s, "invalid sph handle %d", h);
    return sph_passes_internal(ctx, s, sph_current_tex(ctx, s), 4);
}

extern "C" int cwa_sph_step(cwa_ctx* ctx, cwa_sph h, int nsteps)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, h);
    CWA_CHECK(s, "invalid sph handle %d", h);
    for (int k = 0; k < nsteps; k++) CWA_TRY(sph_passes_internal(ctx, s, sph_current_tex(ctx, s), 7));
    return 0;
}

extern "C" int cwa_sph_neighbour_count(cwa_ctx* ctx, cwa_sph h, int* host)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, h);
    CWA_CHECK(s && host, "invalid sph handle %d", h);
    BufferObj* pb = get_buffer(ctx, s->particles);
    CWA_CHECK(pb, "sph: particle buffer vanished");
    if (s->n == 0) return 0;
    int* dout = nullptr;
    CWA_CUDA(cudaMalloc(&dout, (size_t)s->n * 4));
    CWA_CUDA(cudaMemsetAsync(dout, 0, (size_t)s->n * 4, ctx->stream));   // NaN particles are in no list: 0 neighbours
    CWA_TRY(sph_prepare(ctx, s));
    const Sph3Const* cc = (const Sph3Const*)s->consts;
    if (s->grid >= 0) {
        GridObj* g = get_grid(ctx, s->grid);
        CWA_TRY(sph_snapshot(ctx, s));
        KScope k(ctx, KID_OTHER);
        sph3_neighbour_count_grid_kernel<<<ceil_div(s->n, 256), 256, 0, ctx->stream>>>(s->posS, g->index_list, g->offset + g->view.num_cells, g->view, g->offset, cc, dout);
    } else {
        KScope k(ctx, KID_OTHER);
        sph3_neighbour_count_allpairs_kernel<<<ceil_div(s->n, 256), 256, 0, ctx->stream>>>((const float4*)pb->ptr, s->n, cc, dout);
    }
    CWA_CUDA(cudaGetLastError());
    CWA_CUDA(cudaMemcpyAsync(host, dout, (size_t)s->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    CWA_CUDA(cudaFree(dout));
    return 0;
}

extern "C" int cwa_sph_init_cube(cwa_ctx* ctx, cwa_sph h, int nx, int ny, int nz)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, h);
    CWA_CHECK(s, "invalid sph handle %d", h);
    CWA_CHECK(nx > 0 && ny > 0 && nz > 0 && (long long)nx * ny * nz == s->n, "cwa_sph_init_cube: %dx%dx%d != %d particles", nx, ny, nz, s->n);
    BufferObj* pb = get_buffer(ctx, s->particles);
    CWA_CHECK(pb, "sph: particle buffer vanished");
    { KScope k(ctx, KID_OTHER);
      sph3_init_cube_kernel<<<ceil_div(s->n, 256), 256, 0, ctx->stream>>>((float4*)pb->ptr, nx, ny, nz, current_params(ctx)); }
    CWA_CUDA(cudaGetLastError());
    sph_invalidate_for_buffer(ctx, s->particles);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// coupled frame: idle() of Main.cpp:530-562 followed by the display() bind of Main.cpp:413
// ---------------------------------------------------------------------------------------------
static bool graph_mode(cwa_ctx* c) { if (c->tune.graph < 0) c->tune.graph = env_int("CWA_GRAPH", 1, 0, 1); return c->tune.graph != 0; }

// The all-pairs frame (the shipped scene: 20 480 particles, 64^2 field) as ONE graph launch: rho_pres + commit, force + commit,
// integrate, wave stencil.  Everything the kernels read is addressed through pointers that stay valid between frames (particle
// SSBO, parameter blocks, the three wave images); what changes from frame to frame is WHICH image plays which role and which one
// the sampler is bound to -- that is the key, and it cycles with period 3, so at most a handful of graphs are ever captured.
// Returns 1 when the frame was not run through a graph (the caller runs the plain sequence), 0 on success, < 0 on error.
static int coupled_frame_graph(cwa_ctx* ctx, SphObj* s, WaveObj* w, cwa_wave hw, int image)
{
    BufferObj* pb = get_buffer(ctx, s->particles);
    if (!pb || s->n == 0) return 1;
    CWA_TRY(sph_prepare(ctx, s));                                    // derived constants: outside the graph (runs only when a block changed)
    long long key[16] = {};
    key[0] = (long long)(intptr_t)pb->ptr; key[1] = s->n; key[2] = image; key[3] = w->evolve ? 1 : 0;
    key[4] = wave_image_with_unit(w, 0); key[5] = wave_image_with_unit(w, 1); key[6] = wave_image_with_unit(w, 2);
    key[7] = (long long)(intptr_t)w->image[0]; key[8] = w->w; key[9] = w->h; key[10] = w->ch;
    for (int b = 1; b <= 4; b++) key[10 + b] = ctx->ubo_binding[b];
    key[15] = hw;
    SphObj::FrameGraph* slot = nullptr;
    for (auto& g : s->frame_graph)
        if (g.valid && memcmp(g.key, key, sizeof(key)) == 0) { slot = &g; break; }
    const int outi = wave_image_with_unit(w, 2);
    if (slot == nullptr) {
        for (auto& g : s->frame_graph) if (!g.valid) { slot = &g; break; }
        if (slot == nullptr) {                                       // table full (objects were rebound many times): recycle entry 0
            slot = &s->frame_graph[0];
            cudaGraphExecDestroy((cudaGraphExec_t)slot->exec);
            slot->valid = false;
        }
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { (void)cudaGetLastError(); return 1; }
        const unsigned long long l0 = ctx->launches;
        int rc = sph_passes_internal(ctx, s, wave_tex_view(ctx, hw, image), 7);
        if (rc == 0 && w->evolve) rc = wave_dispatch_mode(ctx, w, CWA_MODE_EVOLVE);
        cudaGraph_t graph = nullptr;
        const cudaError_t ee = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc != 0 || ee != cudaSuccess || graph == nullptr) {
            (void)cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            ctx->launches = l0;
            return rc != 0 ? rc : 1;
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { (void)cudaGetLastError(); ctx->launches = l0; return 1; }
        slot->valid = true; memcpy(slot->key, key, sizeof(key)); slot->exec = exec; slot->nodes = (unsigned)(ctx->launches - l0);
        ctx->launches = l0;                                          // counted below, like every replay
    } else if (w->evolve) {
        w->version[outi]++;                                          // what wave_dispatch_mode does on the host besides launching
    }
    CWA_CUDA(cudaGraphLaunch((cudaGraphExec_t)slot->exec, ctx->stream));
    ctx->launches += slot->nodes;
    if (w->evolve) wave_pingpong_internal(w);
    return 0;
}

extern "C" int cwa_coupled_step(cwa_ctx* ctx, cwa_sph hs, cwa_wave hw, int nframes, int coupling)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, hs);
    WaveObj* w = get_wave(ctx, hw);
    CWA_CHECK(s && w, "cwa_coupled_step: invalid sph (%d) or wave (%d) handle", hs, hw);
    CWA_CHECK(coupling == CWA_COUPLING_AS_SHIPPED || coupling == CWA_COUPLING_LATEST, "unknown coupling mode %d", coupling);
    // Frame pipelining (results identical to the plain sequence, tested both ways): inside one call nothing but this loop touches
    // the state, so (a) the wave stencil of frame f -- which no SPH pass of frame f may overtake, but which is independent of the
    // grid build of frame f+1 -- runs on a side stream and the first sampling kernel of frame f+1 waits for it, and (b) the
    // integrate pass of frame f counts the particles into the cells of frame f+1 (count-ahead).  The last frame of the call runs the
    // plain sequence, so every array an application can read afterwards is in the state the sequential code leaves.
    const int pipe = (nframes > 1) ? pipeline_mode(ctx) : 0;
    bool wave_in_flight = false;
    s->counts_ahead = false;
    // whichever way the call ends (an error included): the main stream is behind everything a side stream was given, and no
    // count-ahead state outlives the call
    struct Leave {
        cwa_ctx* ctx; SphObj* s; bool* in_flight;
        ~Leave()
        {
            if (*in_flight) cudaStreamWaitEvent(ctx->stream, ctx->ev_pipe[3], 0);
            s->counts_ahead = false;
            s->wait_before_sampling = nullptr;
        }
    } leave{ctx, s, &wave_in_flight};
    for (int f = 0; f < nframes; f++) {
        int image;
        if (coupling == CWA_COUPLING_LATEST) {
            image = -1;
            for (int i = 0; i < 3; i++) if (w->unit[i] == 0) image = i;      // newest level = image unit 0
        } else {
            image = w->tex_unit0;                                            // whatever display() last bound to texture unit 0
        }
        s->wave = hw; s->wave_image = image;
        if (s->grid < 0 && !ctx->profiling && graph_mode(ctx)) {                 // all-pairs frame through a CUDA graph
            const int grc = coupled_frame_graph(ctx, s, w, hw, image);
            if (grc < 0) return grc;
            if (grc == 0) {
                CWA_TRY(cwa_wave_bind_texture_unit(ctx, hw));                        // display() :413
                continue;
            }
        }
        const bool more = f + 1 < nframes;
        // the previous frame's wave stencil may still be running on the side stream: in grid mode the wait is placed behind the grid
        // build (the first kernel that samples the field is the density pass); the all-pairs passes sample from their first kernel on
        s->wait_before_sampling = nullptr;
        if (wave_in_flight) {
            if (s->grid >= 0 && s->n > 0) s->wait_before_sampling = ctx->ev_pipe[3];
            else CWA_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_pipe[3], 0));
        }
        const int rc = sph_passes_internal(ctx, s, wave_tex_view(ctx, hw, image), 7, more && (pipe & 2) && s->grid >= 0);   // :549-557
        if (s->wait_before_sampling) {                                       // not consumed (the call failed early): wait here
            cudaStreamWaitEvent(ctx->stream, s->wait_before_sampling, 0);
            s->wait_before_sampling = nullptr;
        }
        wave_in_flight = false;
        if (rc != 0) return rc;
        if (w->evolve) {                                                         // Module::sComputeAll :560
            if (more && (pipe & 1)) {
                CWA_CUDA(cudaEventRecord(ctx->ev_pipe[2], ctx->stream));         // this frame's SPH passes are done with the field
                CWA_CUDA(cudaStreamWaitEvent(ctx->side_stream[0], ctx->ev_pipe[2], 0));
                int wrc;
                { StreamScope ss(ctx, ctx->side_stream[0]); wrc = wave_step_internal(ctx, w); }
                CWA_CUDA(cudaEventRecord(ctx->ev_pipe[3], ctx->side_stream[0]));
                wave_in_flight = true;
                CWA_TRY(wrc);
            } else {
                CWA_TRY(wave_step_internal(ctx, w));
            }
        }
        CWA_TRY(cwa_wave_bind_texture_unit(ctx, hw));                            // display() :413
    }
    return 0;                                                                    // (Leave joins the side stream)
}
