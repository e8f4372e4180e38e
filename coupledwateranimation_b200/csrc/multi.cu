// multi.cu -- stand-alone device primitives of the multi-GPU decomposition; the per-frame protocol built on the same
// message layout lives in slab.cu.
//   cwa_particles_copy_if / cwa_slab_compact: STABLE copy_if on one coordinate of the particle position (flags -> the single-pass
//     look-back scan the grid uses -> 64-byte record copy): results do not depend on scheduling.
//   cwa_slab_pack / the pack fused into the integrate pass: message slots are handed out with atomicAdd, so the ORDER of the
//     migrants and ghosts inside a message depends on scheduling.  The set does not.  Adopted particles get buffer slots in that
//     order and the grid's canonical order is ascending slot inside a cell, so the FP summation order of the neighbour loops --
//     and with it the last bits of a multi-GPU run -- can differ from run to run; single-GPU runs and checkpoints stay bit-exact.
#include "internal.cuh"
#include <math_constants.h>

// predicate kinds on the coordinate x = pos[axis]
//   0: a <= x < b        (band: ghost layer of a slab face)
//   1: x <  a            (migrated below the slab)
//   2: x >= a            (migrated above the slab)
//   3: !(x < a) && !(x >= b)   (stays; NaN coordinates stay where they are)
__device__ __forceinline__ bool multi_pred(int kind, float x, float a, float b)
{
    switch (kind) {
    case 0: return x >= a && x < b;
    case 1: return x < a;
    case 2: return x >= a;
    default: return !(x < a) && !(x >= b);
    }
}

__global__ void __launch_bounds__(256)
multi_flag_kernel(const float4* __restrict__ aos, int n, int axis, int kind, float a, float b, int* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(aos + (size_t)i * 4);
    const float x = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
    flags[i] = multi_pred(kind, x, a, b) ? 1 : 0;
}

// 4 lanes per particle move whole 64-byte records
__global__ void __launch_bounds__(256)
multi_scatter_kernel(const float4* __restrict__ aos, int n, const int* __restrict__ flags, const int* __restrict__ pos,
                     float4* __restrict__ out, int out_capacity)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 2, q = t & 3;
    if (i >= n) return;
    if (!__ldg(flags + i)) return;
    const int d = __ldg(pos + i);
    if (d < out_capacity) out[(size_t)d * 4 + q] = __ldg(aos + (size_t)i * 4 + q);
}

static MultiScratch* multi_scratch(cwa_ctx* ctx, int n)
{
    MultiScratch* s = &ctx->multi;
    if ((size_t)n <= s->cap) return s;
    cudaStreamSynchronize(ctx->stream);                      // nothing in flight still uses the old arrays
    cudaFree(s->flags); cudaFree(s->pos); cudaFree(s->ticket);
    *s = MultiScratch();
    MultiScratch t;
    const size_t cap = (size_t)n + (size_t)n / 4 + 1024;
    t.tiles = scan_num_tiles((int)cap);
    if (cudaMalloc(&t.flags, cap * 4) != cudaSuccess || cudaMalloc(&t.pos, (cap + 1) * 4) != cudaSuccess ||
        cudaMalloc(&t.ticket, 16 + t.tiles * 8) != cudaSuccess) {
        (void)cudaGetLastError();
        cudaFree(t.flags); cudaFree(t.pos); cudaFree(t.ticket);
        return nullptr;
    }
    t.state = (unsigned long long*)((char*)t.ticket + 16);
    t.cap = cap;                                             // only now: every array exists
    *s = t;
    return s;
}

extern "C" int cwa_particles_copy_if(cwa_ctx* ctx, cwa_buf src, int n, int axis, int kind, float a, float b,
                                     cwa_buf dst, int dst_offset, int* count_out)
{
    DeviceGuard _dg(ctx);
    BufferObj* s = get_buffer(ctx, src);
    BufferObj* d = get_buffer(ctx, dst);
    CWA_CHECK(s && d && count_out, "cwa_particles_copy_if: invalid buffer handle or null count");
    CWA_CHECK(axis >= 0 && axis < 3 && kind >= 0 && kind <= 3, "cwa_particles_copy_if: bad axis/kind");
    CWA_CHECK(n >= 0 && (size_t)n * 64 <= s->bytes, "cwa_particles_copy_if: source holds fewer than %d particles", n);
    CWA_CHECK(dst_offset >= 0 && (size_t)dst_offset * 64 <= d->bytes, "cwa_particles_copy_if: destination offset outside the buffer");
    *count_out = 0;
    if (n == 0) return 0;
    MultiScratch* sc = multi_scratch(ctx, n);
    CWA_CHECK(sc, "cwa_particles_copy_if: out of device memory for scratch");
    const int cap = (int)(d->bytes / 64) - dst_offset;
    CWA_CUDA(cudaMemsetAsync(sc->ticket, 0, 16 + scan_num_tiles(n) * 8, ctx->stream));
    { KScope k(ctx, KID_OTHER);
      multi_flag_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>((const float4*)s->ptr, n, axis, kind, a, b, sc->flags); }
    CWA_TRY(scan_exclusive_launch(ctx, sc->flags, sc->pos, n, sc->ticket, sc->state));
    { KScope k(ctx, KID_OTHER);
      multi_scatter_kernel<<<ceil_div((long long)n * 4, 256), 256, 0, ctx->stream>>>(
          (const float4*)s->ptr, n, sc->flags, sc->pos, (float4*)d->ptr + (size_t)dst_offset * 4, cap); }
    CWA_CUDA(cudaGetLastError());
    int total = 0;
    CWA_CUDA(cudaMemcpyAsync(&total, sc->pos + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    CWA_CHECK(total <= cap, "cwa_particles_copy_if: %d selected particles exceed the destination capacity %d", total, cap);
    *count_out = total;
    return 0;
}

extern "C" int cwa_sph_set_count(cwa_ctx* ctx, cwa_sph h, int n)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, h);
    CWA_CHECK(s, "invalid sph handle %d", h);
    BufferObj* pb = get_buffer(ctx, s->particles);
    CWA_CHECK(pb && n >= 0 && (size_t)n * 64 <= pb->bytes, "cwa_sph_set_count: %d particles exceed the particle buffer", n);
    CWA_CHECK(n <= s->capacity, "cwa_sph_set_count: %d particles exceed the capacity %d the object was created with", n, s->capacity);
    if (s->grid >= 0) {
        GridObj* g = get_grid(ctx, s->grid);
        CWA_CHECK(g && n <= g->max_particles, "cwa_sph_set_count: %d particles exceed the grid capacity", n);
    }
    s->n = n;
    s->snapshot_valid = false;
    return 0;                                              // (a pending slab pack stays valid: it only covers the owned range)
}

// ---------------------------------------------------------------------------------------------
// lean per-frame exchange (one message per neighbour and frame, counts travel in the header)
//
// message layout (floats / ints, 64-byte records):
//   record 0            header: int mig_count, int ghost_count, int overflow
//   records 1 .. cap_mig            migrants  (left the sender's slab through this face)
//   records 1+cap_mig .. +cap_ghost ghosts    (within `band` of the face, still owned by the sender)
// A migrant is marked DEAD in the sender's SSBO (pos = NaN, pos.w = CWA_DEAD_W): the grid build
// leaves NaN positions out, so dead slots cost nothing and no compaction pass is needed per frame.
// ---------------------------------------------------------------------------------------------
#define CWA_DEAD_W (-1.0f)

__device__ __forceinline__ bool slot_dead(const float4 p) { return p.w == CWA_DEAD_W && !(p.x == p.x); }

__device__ __forceinline__ void copy_record(float4* __restrict__ dst, const float4* __restrict__ src)
{
    const float4 a = src[0], b = src[1], c = src[2], d = src[3];
    dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
}

__global__ void __launch_bounds__(256)
slab_pack_kernel(float4* __restrict__ aos, int n_owned, float z_lo, float z_hi, float band, int has_left, int has_right,
                 float4* __restrict__ msg_l, float4* __restrict__ msg_r, int cap_mig, int cap_ghost)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned) return;
    const float4 p = aos[(size_t)i * 4];
    if (slot_dead(p)) return;
    const float z = p.z;                                   // NaN z: every test below is false -> stays
    float4* msg = nullptr;
    bool migrate = false;
    if (has_left && z < z_lo + band) { msg = msg_l; migrate = z < z_lo; }
    else if (has_right && z >= z_hi - band) { msg = msg_r; migrate = z >= z_hi; }
    if (msg == nullptr) return;
    int* hdr = reinterpret_cast<int*>(msg);
    const int slot = atomicAdd(hdr + (migrate ? 0 : 1), 1);
    const int cap = migrate ? cap_mig : cap_ghost;
    if (slot >= cap) { atomicExch(hdr + 2, 1); return; }    // overflow: reported to the host, particle stays put
    copy_record(msg + 4 * (size_t)(1 + (migrate ? 0 : cap_mig) + slot), aos + (size_t)i * 4);
    if (migrate) aos[(size_t)i * 4] = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CWA_DEAD_W);
}

// append the migrants of both received messages behind the owned range, then the ghosts behind them:
// the received ghosts AND the particles this rank itself just sent away as migrants (the new owner
// packed its message before adopting them, so they are not in its ghost list this frame, but they
// still sit within 2h of the face and are needed as neighbours here).
__global__ void __launch_bounds__(256)
slab_unpack_kernel(float4* __restrict__ aos, int n_owned, int capacity, const float4* __restrict__ rcv_l,
                   const float4* __restrict__ rcv_r, const float4* __restrict__ snd_l, const float4* __restrict__ snd_r,
                   int cap_mig, int cap_ghost, int* __restrict__ counts)
{
    const int* hl = reinterpret_cast<const int*>(rcv_l);
    const int* hr = reinterpret_cast<const int*>(rcv_r);
    const int ml = rcv_l ? min(hl[0], cap_mig) : 0, gl = rcv_l ? min(hl[1], cap_ghost) : 0;
    const int mr = rcv_r ? min(hr[0], cap_mig) : 0, gr = rcv_r ? min(hr[1], cap_ghost) : 0;
    const int sl = snd_l ? min(reinterpret_cast<const int*>(snd_l)[0], cap_mig) : 0;
    const int sr = snd_r ? min(reinterpret_cast<const int*>(snd_r)[0], cap_mig) : 0;
    const int total = ml + mr + gl + gr + sl + sr;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        counts[0] = n_owned + ml + mr;                                     // owned range after adoption
        counts[1] = n_owned + total;                                       // owned + ghosts
        counts[2] = (rcv_l ? hl[2] : 0) | (rcv_r ? hr[2] : 0) | ((n_owned + total > capacity) ? 2 : 0);
        counts[3] = ml + mr;
    }
    if (t >= total || n_owned + total > capacity) return;
    const float4* src;
    int u = t;
    if (u < ml) src = rcv_l + 4 * (size_t)(1 + u);
    else if ((u -= ml) < mr) src = rcv_r + 4 * (size_t)(1 + u);
    else if ((u -= mr) < gl) src = rcv_l + 4 * (size_t)(1 + cap_mig + u);
    else if ((u -= gl) < gr) src = rcv_r + 4 * (size_t)(1 + cap_mig + u);
    else if ((u -= gr) < sl) src = snd_l + 4 * (size_t)(1 + u);
    else src = snd_r + 4 * (size_t)(1 + (u - sl));
    copy_record(aos + 4 * (size_t)(n_owned + t), src);
}

extern "C" int cwa_slab_pack(cwa_ctx* ctx, cwa_buf particles, int n_owned, float z_lo, float z_hi, float band,
                             cwa_buf msg_left, cwa_buf msg_right, int cap_mig, int cap_ghost)
{
    DeviceGuard _dg(ctx);
    BufferObj* p = get_buffer(ctx, particles);
    BufferObj* ml = get_buffer(ctx, msg_left);
    BufferObj* mr = get_buffer(ctx, msg_right);
    CWA_CHECK(p && n_owned >= 0 && (size_t)n_owned * 64 <= p->bytes, "cwa_slab_pack: bad particle buffer / count");
    const size_t need = (size_t)(1 + cap_mig + cap_ghost) * 64;
    CWA_CHECK((msg_left == -1 || (ml && ml->bytes >= need)) && (msg_right == -1 || (mr && mr->bytes >= need)),
              "cwa_slab_pack: message buffers smaller than %zu bytes", need);
    // the integrate pass of cwa_sph_step_slab may already have packed exactly these messages
    for (auto& s : ctx->sphs) {
        if (!s.live || s.particles != particles || !s.slab_packed.valid) continue;
        const SphObj::SlabPacked& k = s.slab_packed;
        if (k.msg_left == msg_left && k.msg_right == msg_right && k.n_owned == n_owned && k.cap_mig == cap_mig && k.cap_ghost == cap_ghost &&
            k.z_lo == z_lo && k.z_hi == z_hi && k.band == band) {
            s.slab_packed.valid = false;                     // consumed: the messages are about to be sent
            s.snapshot_valid = false;
            return 0;
        }
    }
    if (ml) CWA_CUDA(cudaMemsetAsync(ml->ptr, 0, 64, ctx->stream));
    if (mr) CWA_CUDA(cudaMemsetAsync(mr->ptr, 0, 64, ctx->stream));
    if (n_owned == 0 || (!ml && !mr)) return 0;
    KScope k(ctx, KID_OTHER);
    slab_pack_kernel<<<ceil_div(n_owned, 256), 256, 0, ctx->stream>>>((float4*)p->ptr, n_owned, z_lo, z_hi, band, ml != nullptr, mr != nullptr,
                                                                     ml ? (float4*)ml->ptr : nullptr, mr ? (float4*)mr->ptr : nullptr, cap_mig, cap_ghost);
    CWA_CUDA(cudaGetLastError());
    sph_invalidate_for_buffer(ctx, particles);
    return 0;
}

// One SPH frame (rho_pres, force, integrate on `n_total` = owned + ghost particles) whose integrate pass also packs the messages of
// the NEXT exchange: owned particles (original slot < n_owned) that end the frame within `band` of a face or beyond it.  The next
// cwa_slab_pack call with the same arguments finds its work done -- one pass over the particle SSBO less per frame.
extern "C" int cwa_sph_step_slab(cwa_ctx* ctx, cwa_sph h, int n_owned, float z_lo, float z_hi, float band,
                                 cwa_buf msg_left, cwa_buf msg_right, int cap_mig, int cap_ghost)
{
    DeviceGuard _dg(ctx);
    SphObj* s = get_sph(ctx, h);
    CWA_CHECK(s, "invalid sph handle %d", h);
    CWA_CHECK(s->grid >= 0, "cwa_sph_step_slab: the SPH object needs a uniform grid");
    BufferObj* ml = get_buffer(ctx, msg_left);
    BufferObj* mr = get_buffer(ctx, msg_right);
    const size_t need = (size_t)(1 + cap_mig + cap_ghost) * 64;
    CWA_CHECK(n_owned >= 0 && n_owned <= s->n, "cwa_sph_step_slab: owned range %d exceeds the particle count %d", n_owned, s->n);
    CWA_CHECK((msg_left == -1 || (ml && ml->bytes >= need)) && (msg_right == -1 || (mr && mr->bytes >= need)),
              "cwa_sph_step_slab: message buffers smaller than %zu bytes", need);
    if (ml) CWA_CUDA(cudaMemsetAsync(ml->ptr, 0, 64, ctx->stream));
    if (mr) CWA_CUDA(cudaMemsetAsync(mr->ptr, 0, 64, ctx->stream));
    SlabPackArgs a;
    a.n_owned = n_owned; a.z_lo = z_lo; a.z_hi = z_hi; a.band = band;
    a.msg_l = ml ? (float4*)ml->ptr : nullptr; a.msg_r = mr ? (float4*)mr->ptr : nullptr;
    a.cap_mig = cap_mig; a.cap_ghost = cap_ghost;
    CWA_TRY(sph_passes_internal(ctx, s, wave_tex_view(ctx, s->wave, s->wave_image), 7, false, &a));
    if (s->n > 0 && (ml || mr)) {
        SphObj::SlabPacked& k = s->slab_packed;
        k.valid = true; k.particles = s->particles; k.msg_left = msg_left; k.msg_right = msg_right;
        k.n_owned = n_owned; k.cap_mig = cap_mig; k.cap_ghost = cap_ghost; k.z_lo = z_lo; k.z_hi = z_hi; k.band = band;
    }
    return 0;
}

// counts_host[4] = {owned range, owned + ghosts, flags (1: a sender overflowed its message, 2: capacity exceeded), migrants adopted}
extern "C" int cwa_slab_unpack(cwa_ctx* ctx, cwa_buf particles, int n_owned, cwa_buf rcv_left, cwa_buf rcv_right,
                               cwa_buf sent_left, cwa_buf sent_right, int cap_mig, int cap_ghost, int* counts_host)
{
    DeviceGuard _dg(ctx);
    BufferObj* p = get_buffer(ctx, particles);
    BufferObj* rl = get_buffer(ctx, rcv_left);
    BufferObj* rr = get_buffer(ctx, rcv_right);
    BufferObj* sl = get_buffer(ctx, sent_left);
    BufferObj* sr = get_buffer(ctx, sent_right);
    CWA_CHECK(p && counts_host, "cwa_slab_unpack: bad particle buffer");
    CWA_CHECK((rcv_left == -1 || rl) && (rcv_right == -1 || rr), "cwa_slab_unpack: invalid message buffer handle");
    MultiScratch* sc = multi_scratch(ctx, 1024);
    CWA_CHECK(sc, "cwa_slab_unpack: out of device memory");
    const int capacity = (int)(p->bytes / 64);
    const int max_in = 2 * (2 * cap_mig + cap_ghost);
    { KScope k(ctx, KID_OTHER);
      slab_unpack_kernel<<<ceil_div(max_in > 0 ? max_in : 1, 256), 256, 0, ctx->stream>>>(
          (float4*)p->ptr, n_owned, capacity, rl ? (const float4*)rl->ptr : nullptr, rr ? (const float4*)rr->ptr : nullptr,
          sl ? (const float4*)sl->ptr : nullptr, sr ? (const float4*)sr->ptr : nullptr, cap_mig, cap_ghost, sc->flags); }
    CWA_CUDA(cudaGetLastError());
    CWA_CUDA(cudaMemcpyAsync(counts_host, sc->flags, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));          // the one host synchronisation of a distributed frame
    sph_invalidate_for_buffer(ctx, particles);
    return 0;
}

// drop dead slots from the owned range (rare: called when the buffer runs full); stable; synchronises
__global__ void __launch_bounds__(256)
multi_flag_live_kernel(const float4* __restrict__ aos, int n, int* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = slot_dead(__ldg(aos + (size_t)i * 4)) ? 0 : 1;
}

extern "C" int cwa_slab_compact(cwa_ctx* ctx, cwa_buf particles, int n_owned, cwa_buf scratch, int* n_live)
{
    DeviceGuard _dg(ctx);
    BufferObj* p = get_buffer(ctx, particles);
    BufferObj* s = get_buffer(ctx, scratch);
    CWA_CHECK(p && s && n_live && (size_t)n_owned * 64 <= p->bytes && (size_t)n_owned * 64 <= s->bytes, "cwa_slab_compact: bad buffers");
    *n_live = 0;
    for (auto& so : ctx->sphs)
        CWA_CHECK(!(so.live && so.particles == particles && so.slab_packed.valid),
                  "cwa_slab_compact: the integrate pass already packed the next exchange's messages from this buffer (cwa_sph_step_slab): "
                  "compacting now would drop the migrants it marked; compact before the step or after cwa_slab_pack");
    if (n_owned == 0) return 0;
    MultiScratch* sc = multi_scratch(ctx, n_owned);
    CWA_CHECK(sc, "cwa_slab_compact: out of device memory");
    CWA_CUDA(cudaMemsetAsync(sc->ticket, 0, 16 + scan_num_tiles(n_owned) * 8, ctx->stream));
    { KScope k(ctx, KID_OTHER);
      multi_flag_live_kernel<<<ceil_div(n_owned, 256), 256, 0, ctx->stream>>>((const float4*)p->ptr, n_owned, sc->flags); }
    CWA_TRY(scan_exclusive_launch(ctx, sc->flags, sc->pos, n_owned, sc->ticket, sc->state));
    { KScope k(ctx, KID_OTHER);
      multi_scatter_kernel<<<ceil_div((long long)n_owned * 4, 256), 256, 0, ctx->stream>>>(
          (const float4*)p->ptr, n_owned, sc->flags, sc->pos, (float4*)s->ptr, (int)(s->bytes / 64)); }
    int total = 0;
    CWA_CUDA(cudaMemcpyAsync(&total, sc->pos + n_owned, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    if (total > 0) CWA_CUDA(cudaMemcpyAsync(p->ptr, s->ptr, (size_t)total * 64, cudaMemcpyDeviceToDevice, ctx->stream));
    sph_invalidate_for_buffer(ctx, particles);
    *n_live = total;
    return 0;
}
