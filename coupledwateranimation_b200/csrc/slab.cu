// slab.cu -- the multi-GPU coupled frame behind the C ABI (the reference is single-GPU, so the contract is
// "N ranks reproduce the 1-rank state of idle(), CoupledWaterAnimation/Main.cpp:540-561").
//
// Decomposition: particles in z slabs aligned with row blocks of the wave field (texture t = uv_scale_z * z, rho_pres_comp.glsl:72).
// Every rank owns one cwa_slab object; ranks may be processes (one per GPU, mailboxes shared through CUDA IPC) or contexts of one
// process (a C++ host driving several devices: cwa_slab_group_step).  There is NO host synchronisation, NO collective library and
// NO size negotiation on the per-frame path:
//
//   * every rank exports ONE device allocation, its MAILBOX: arrival flags, two particle-message slots per neighbour (frame
//     parity), two wave-halo slots per neighbour and a ring of global-last-row slots;
//   * the integrate pass packs the migrant / ghost messages of the next frame into local send buffers (records still in
//     registers), a small push kernel copies exactly the live records into the neighbour's mailbox with peer stores over NVLink
//     and releases the slot's flag (st.release.sys) -- message sizes are whatever the header says, nothing is padded to capacity;
//   * the receiver's stream holds a one-warp wait kernel (ld.acquire.sys on the flag, bounded by a timeout that raises an error
//     bit instead of hanging the GPU), then the unpack kernel: migrants go into slots freed by earlier emigrants (device-side free
//     list, so the owned range does not grow and no compaction pass exists), ghosts behind the owned range;
//   * particle counts are DEVICE-RESIDENT (SlabState): every kernel of the SPH frame is launched for the buffer capacity and reads
//     the live count from memory, so the host never needs to know them;
//   * the wave stencil, its halo push / receive and the global last row (WaveNormal's uv + (0,1) tap clamps to it from everywhere,
//     force_comp.glsl:136) run on a side stream next to the following frame's unpack + grid build.
//
// Slot reuse needs no acknowledgements: message k+2 of a rank is produced by its integrate pass of frame k+1, which runs after its
// unpack of frame k+1, which waited for the neighbour's message k+1, which the neighbour produced after consuming message k (same
// slot as k+2).  The last-row ring has more slots than the largest skew between the last rank and any other (world - 1 frames).
#include "internal.cuh"
#include <math_constants.h>
#include <cmath>

constexpr int    SLAB_RING = 32;
constexpr int    SLAB_MAX_WORLD = 16;
constexpr size_t SLAB_FLAG_BYTES = 4096;
constexpr int    SLAB_PUSH_CTAS = 64;

enum { SLAB_ERR_SENDER_OVERFLOW = 1, SLAB_ERR_CAPACITY = 2, SLAB_ERR_TIMEOUT_PARTICLES = 4, SLAB_ERR_TIMEOUT_WAVE = 8 };

struct SlabFlags {                    // first bytes of a mailbox; written by the PEERS, read by the owner
    unsigned part[2][2];              // [side: 0 = from the left neighbour, 1 = from the right][frame parity] = message number + 1
    unsigned wave[2][2];
    unsigned last[SLAB_RING];
};

struct SlabState { int n_owned, n_total, free_count, pad; };

struct SlabObj {
    bool live = false;
    cwa_slab_desc d{};
    cwa_sph sph = -1;
    cwa_wave wave = -1;
    char*  mail = nullptr;            // exported allocation
    size_t mail_bytes = 0, part_bytes = 0, wave_bytes = 0, row_bytes = 0;
    char*  peer[SLAB_MAX_WORLD] = {};
    bool   peer_ipc[SLAB_MAX_WORLD] = {};
    float4* send[2] = {nullptr, nullptr};   // local send buffers: to the left / to the right neighbour
    SlabState* state = nullptr;       // [2], alternating with the particle message number
    int*   free_list = nullptr;       // slots of the owned range vacated by emigrants (stack)
    int*   arrivals = nullptr;        // [max arrivals + 1]: ids of the particles the last unpack placed, [max] = their number (count-ahead across the exchange)
    int    arrivals_max = 0;
    int*   err = nullptr;             // sticky error bits
    unsigned* done = nullptr;         // last-block-done counters of the push kernels
    unsigned long long* stats = nullptr;   // [0] migrants adopted
    unsigned pseq = 0, wseq = 0;      // next particle message to consume; wave messages consumed so far
    bool   wave_synced = false, wave_in_flight = false;
    bool   push_pending = false;      // a particle push on the side stream still reads the send buffers
    cudaEvent_t ev_pack = nullptr, ev_int = nullptr, ev_wave = nullptr, ev_push = nullptr;
};

static SlabObj* get_slab(cwa_ctx* ctx, cwa_slab h)
{
    if (!ctx || h < 0 || h >= (int)ctx->slabs.size() || !ctx->slabs[h] || !ctx->slabs[h]->live) return nullptr;
    return ctx->slabs[h];
}

static size_t slab_part_off(const SlabObj* s, int side, int par) { return SLAB_FLAG_BYTES + (size_t)(side * 2 + par) * s->part_bytes; }
static size_t slab_wave_off(const SlabObj* s, int side, int par) { return SLAB_FLAG_BYTES + 4 * s->part_bytes + (size_t)(side * 2 + par) * s->wave_bytes; }
static size_t slab_last_off(const SlabObj* s, int slot) { return SLAB_FLAG_BYTES + 4 * s->part_bytes + 4 * s->wave_bytes + (size_t)slot * s->row_bytes; }

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned slab_ld_acquire(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void slab_st_release(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long slab_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#define SLAB_DEAD_W (-1.0f)
__device__ __forceinline__ bool slab_slot_dead(const float4 p) { return p.w == SLAB_DEAD_W && !(p.x == p.x); }

// One warp; lane t < n waits until flag[t] reaches expect[t].  A timeout raises an error bit and lets the stream go on (the run
// is then invalid, but no GPU hangs); once any error bit is set every later wait returns at once.
struct SlabWaitArgs { const unsigned* flag[4]; unsigned expect[4]; int n; };

__global__ void slab_wait_kernel(SlabWaitArgs a, int* err, int err_bit, unsigned long long timeout_ns)
{
    const int t = threadIdx.x;
    if (t >= a.n || a.flag[t] == nullptr) return;
    if (*reinterpret_cast<volatile int*>(err) != 0) return;
    const unsigned long long t0 = slab_globaltimer();
    while ((int)(slab_ld_acquire(a.flag[t]) - a.expect[t]) < 0) {
        if (slab_globaltimer() - t0 > timeout_ns) { atomicOr(err, err_bit); break; }
        __nanosleep(200);
    }
}

// explicit pack (first frame of a call; later frames are packed by the integrate pass, sph3.cu MODE 2): owned particles that lie
// beyond a slab face become MIGRANTS (copied into the message, slot marked dead and pushed on the free list), owned particles
// within `band` of a face become GHOSTS.  Message layout as in multi.cu: record 0 = header {migrants, ghosts, overflow}.
__global__ void __launch_bounds__(256)
slab2_pack_kernel(float4* __restrict__ aos, SlabState* __restrict__ st, float z_lo, float z_hi, float band,
                  float4* __restrict__ msg_l, float4* __restrict__ msg_r, int cap_mig, int cap_ghost, int* __restrict__ free_list)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st->n_owned) return;
    const float4 p = aos[(size_t)i * 4];
    if (slab_slot_dead(p)) return;
    const float z = p.z;                                   // NaN z: every test below is false -> stays
    float4* msg = nullptr;
    bool migrate = false;
    if (msg_l != nullptr && z < z_lo + band) { msg = msg_l; migrate = z < z_lo; }
    else if (msg_r != nullptr && z >= z_hi - band) { msg = msg_r; migrate = z >= z_hi; }
    if (msg == nullptr) return;
    int* hdr = reinterpret_cast<int*>(msg);
    const int slot = atomicAdd(hdr + (migrate ? 0 : 1), 1);
    const int cap = migrate ? cap_mig : cap_ghost;
    if (slot >= cap) { atomicExch(hdr + 2, 1); return; }    // overflow: reported through the error bits, particle stays put
    float4* d = msg + 4 * (size_t)(1 + (migrate ? 0 : cap_mig) + slot);
    const float4* src = aos + (size_t)i * 4;
    const float4 b = src[1], c = src[2], e = src[3];
    d[0] = p; d[1] = b; d[2] = c; d[3] = e;
    if (migrate) {
        aos[(size_t)i * 4] = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, SLAB_DEAD_W);
        free_list[atomicAdd(&st->free_count, 1)] = i;
    }
}

// Copy the live records of the two local send buffers (header + migrants, ghosts) into the neighbours' mailbox slots with peer
// stores, then release the slots' flags.  blockIdx.y = side (0: to the left neighbour, 1: to the right).
__global__ void __launch_bounds__(256)
slab_push_particles_kernel(const float4* __restrict__ send_l, const float4* __restrict__ send_r, float4* __restrict__ dst_l,
                           float4* __restrict__ dst_r, unsigned* flag_l, unsigned* flag_r, unsigned value, int cap_mig, int cap_ghost,
                           unsigned* __restrict__ done)
{
    const int side = blockIdx.y;
    const float4* src = side ? send_r : send_l;
    float4* dst = side ? dst_r : dst_l;
    unsigned* flag = side ? flag_r : flag_l;
    if (src == nullptr || dst == nullptr) return;
    const int* hdr = reinterpret_cast<const int*>(src);
    const int m = min(hdr[0], cap_mig), g = min(hdr[1], cap_ghost);
    const long long n4 = 4ll * (1 + m + g);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        const long long rec = t >> 2;
        const long long pos = (rec <= m) ? rec : (1 + cap_mig + (rec - 1 - m));
        dst[pos * 4 + (t & 3)] = src[pos * 4 + (t & 3)];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&done[side], 1u);
        if (prev == gridDim.x - 1) {                       // the last block of this side: every record is on its way
            done[side] = 0;
            __threadfence_system();
            slab_st_release(flag, value);
        }
    }
}

// Received migrants go into free slots of the owned range (or behind it when the free list is empty), ghosts behind the owned
// range: the received ghosts AND the particles this rank itself just sent away as migrants (the new owner packed its ghost list
// before adopting them, so they are not in it this frame, but they still sit within 2h of the face and are neighbours here).
// 4 lanes per record.  `sin` is read only; lane 0 writes the new counts to `sout` (the SPH frame reads them from there).
// count-ahead across the exchange (all null: off)
struct SlabAheadArgs {
    GridView g;
    int* counter = nullptr;           // cell counters, already holding the particles that stayed
    int* cell_of = nullptr;           // per particle id
    int* rank = nullptr;              // arrival rank inside the cell, per particle id
    int* arrivals = nullptr;          // ids placed by this unpack; [arrivals_max] = their number
    int  arrivals_max = 0;
};

__global__ void __launch_bounds__(256)
slab2_unpack_kernel(float4* __restrict__ aos, int capacity, const float4* __restrict__ rcv_l, const float4* __restrict__ rcv_r,
                    const float4* __restrict__ snd_l, const float4* __restrict__ snd_r, int cap_mig, int cap_ghost,
                    const SlabState* __restrict__ sin, SlabState* __restrict__ sout, const int* __restrict__ free_list,
                    int* __restrict__ err, unsigned long long* __restrict__ stats,
                    SlabAheadArgs ah)
{
    const int* hl = reinterpret_cast<const int*>(rcv_l);
    const int* hr = reinterpret_cast<const int*>(rcv_r);
    const int ml = rcv_l ? min(hl[0], cap_mig) : 0, gl = rcv_l ? min(hl[1], cap_ghost) : 0;
    const int mr = rcv_r ? min(hr[0], cap_mig) : 0, gr = rcv_r ? min(hr[1], cap_ghost) : 0;
    const int sl = snd_l ? min(reinterpret_cast<const int*>(snd_l)[0], cap_mig) : 0;
    const int sr = snd_r ? min(reinterpret_cast<const int*>(snd_r)[0], cap_mig) : 0;
    const int n_owned = sin->n_owned, F = sin->free_count;
    const int m_in = ml + mr, G = gl + gr + sl + sr;
    const int k = min(m_in, F);                            // migrants that find a free slot
    const int n_owned2 = n_owned + (m_in - k);
    const bool over = (long long)n_owned2 + G > (long long)capacity;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        int e = 0;
        if ((rcv_l && hl[2]) || (rcv_r && hr[2])) e |= SLAB_ERR_SENDER_OVERFLOW;
        if (over) e |= SLAB_ERR_CAPACITY;
        if (e) atomicOr(err, e);
        sout->n_owned = over ? n_owned : n_owned2;
        sout->n_total = over ? n_owned : n_owned2 + G;
        sout->free_count = over ? F : F - k;
        sout->pad = 0;
        if (!over) stats[0] += (unsigned long long)m_in;
        if (ah.arrivals != nullptr) ah.arrivals[ah.arrivals_max] = over ? 0 : min(m_in + G, ah.arrivals_max);   // how many ids follow
    }
    if (over) return;
    const int u = (int)(t >> 2), q = (int)(t & 3);
    if (u >= m_in + G) return;
    const float4* src;
    int dest;
    if (u < m_in) {
        src = (u < ml) ? rcv_l + 4 * (size_t)(1 + u) : rcv_r + 4 * (size_t)(1 + (u - ml));
        dest = (u < k) ? __ldg(free_list + (F - 1 - u)) : n_owned + (u - k);
    } else {
        int v = u - m_in;
        if (v < gl) src = rcv_l + 4 * (size_t)(1 + cap_mig + v);
        else if ((v -= gl) < gr) src = rcv_r + 4 * (size_t)(1 + cap_mig + v);
        else if ((v -= gr) < sl) src = snd_l + 4 * (size_t)(1 + v);
        else src = snd_r + 4 * (size_t)(1 + (v - sl));
        dest = n_owned2 + (u - m_in);
    }
    const float4 v = src[q];
    aos[(size_t)dest * 4 + q] = v;
    // count-ahead across the exchange: the integrate pass of the last frame counted the particles that stayed; an arrival is hashed and
    // counted here (the same cwa_cell3 on the same stored floats as grid_hash_count_kernel<3>), so the grid build starts at the scan
    if (ah.counter != nullptr && q == 0 && u < ah.arrivals_max) {
        int cell = -1, rank = 0;
        if (v.x == v.x && v.y == v.y && v.z == v.z) {
            int ci, cj, ck;
            cwa_cell3(ah.g, v.x, v.y, v.z, ci, cj, ck);
            cell = (ci * ah.g.n[1] + cj) * ah.g.kstride + ck;
            rank = atomicAdd(ah.counter + cell, 1);
        }
        ah.cell_of[dest] = cell;
        ah.rank[dest] = rank;
        ah.arrivals[u] = dest;
    }
}

// generic "copy rows to a peer, then release a flag" (wave halos, global last row)
struct SlabPushSeg { const float4* src; float4* dst; unsigned long long n16; unsigned* flag; unsigned value; unsigned pad; };
struct SlabPushArgs { SlabPushSeg seg[2 + SLAB_MAX_WORLD]; };

__global__ void __launch_bounds__(256)
slab_push_rows_kernel(SlabPushArgs a, unsigned* __restrict__ done)
{
    const SlabPushSeg s = a.seg[blockIdx.y];
    if (s.dst == nullptr) return;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < s.n16; t += (unsigned long long)gridDim.x * blockDim.x)
        s.dst[t] = s.src[t];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&done[blockIdx.y], 1u);
        if (prev == gridDim.x - 1) {
            done[blockIdx.y] = 0;
            __threadfence_system();
            if (s.flag != nullptr) slab_st_release(s.flag, s.value);
        }
    }
}

struct SlabCopySeg { const float4* src; float4* dst; unsigned long long n16; };
struct SlabCopyArgs { SlabCopySeg seg[3]; };

__global__ void __launch_bounds__(256)
slab_copy_rows_kernel(SlabCopyArgs a)
{
    const SlabCopySeg s = a.seg[blockIdx.y];
    if (s.dst == nullptr || s.src == nullptr) return;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < s.n16; t += (unsigned long long)gridDim.x * blockDim.x)
        s.dst[t] = s.src[t];
}

__global__ void slab_set_state_kernel(SlabState* st, int n_owned)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) { st->n_owned = n_owned; st->n_total = n_owned; st->free_count = 0; st->pad = 0; }
}

// ---------------------------------------------------------------------------------------------
// plan (pure host logic; mirrors coupledwateranimation_b200/distributed.py SlabPlan.make)
// ---------------------------------------------------------------------------------------------
extern "C" int cwa_slab_plan(int world, int rank, int wave_w, int wave_h, int wave_ch, double uv_scale_z, double h, const int* row_bounds,
                             cwa_slab_desc* out)
{
    CWA_CHECK(out != nullptr, "cwa_slab_plan: out is null");
    CWA_CHECK(world >= 1 && world <= SLAB_MAX_WORLD && rank >= 0 && rank < world, "cwa_slab_plan: rank %d / world %d (at most %d ranks)", rank, world, SLAB_MAX_WORLD);
    CWA_CHECK(wave_w >= 1 && wave_h >= world && (wave_ch == 1 || wave_ch == 4), "cwa_slab_plan: bad wave field %dx%dx%d", wave_w, wave_h, wave_ch);
    CWA_CHECK(uv_scale_z > 0.0 && h > 0.0, "cwa_slab_plan: uv_scale_z and h must be positive");
    cwa_slab_desc d;
    memset(&d, 0, sizeof(d));
    d.rank = rank; d.world = world; d.wave_w = wave_w; d.wave_h = wave_h; d.wave_ch = wave_ch;
    auto lo_of = [&](int r) -> int {
        if (row_bounds) return row_bounds[r];
        const int base = wave_h / world, rem = wave_h % world;
        return r * base + (r < rem ? r : rem);
    };
    if (row_bounds) {
        CWA_CHECK(row_bounds[0] == 0 && row_bounds[world] == wave_h, "cwa_slab_plan: row_bounds must run from 0 to %d", wave_h);
        for (int r = 0; r < world; r++) CWA_CHECK(row_bounds[r + 1] > row_bounds[r], "cwa_slab_plan: row_bounds must ascend");
    }
    d.row_lo = lo_of(rank);
    d.row_hi = (rank == world - 1) ? wave_h : lo_of(rank + 1);
    // texture t = uv_scale_z * z, texel row = t * H - 0.5: row boundary b <-> z = b / (H * uv_scale_z)
    const double scale = (double)wave_h * uv_scale_z;
    d.z_lo = (rank == 0) ? -3.0e38f : (float)((double)d.row_lo / scale);
    d.z_hi = (rank == world - 1) ? 3.0e38f : (float)((double)d.row_hi / scale);
    d.band = (float)(2.0 * h);                             // ghost layer 2h: ghost densities are recomputed locally
    const double ghost_rows = 2.0 * h * scale;
    const int halo_lo = (int)std::ceil(ghost_rows) + 2;                                    // ghost reach + bilinear footprint
    const int halo_hi = (int)std::ceil(ghost_rows + 0.01 * (double)wave_h) + 3;            // + WaveVelocity's uv + 0.01 tap (force_comp.glsl:117-128)
    d.store_lo = (rank == 0) ? d.row_lo : (d.row_lo - halo_lo > 0 ? d.row_lo - halo_lo : 0);
    d.store_hi = (rank == world - 1) ? d.row_hi : (d.row_hi + halo_hi < wave_h ? d.row_hi + halo_hi : wave_h);
    d.left_store_hi = (rank == 0) ? 0 : (d.row_lo + halo_hi < wave_h ? d.row_lo + halo_hi : wave_h);       // rows [row_lo, left_store_hi) = the left neighbour's upper halo
    d.right_store_lo = (rank == world - 1) ? 0 : (d.row_hi - halo_lo > 0 ? d.row_hi - halo_lo : 0);         // rows [right_store_lo, row_hi) = the right neighbour's lower halo
    d.halo_rows_max = halo_hi > halo_lo ? halo_hi : halo_lo;
    d.timeout_ms = 10000;
    if (world > 1) {
        const double width = (double)(d.row_hi - d.row_lo) / scale;
        CWA_CHECK(width > 4.0 * h, "cwa_slab_plan: slab of width %g is thinner than two ghost layers (4h = %g)", width, 4.0 * h);
        CWA_CHECK(d.row_hi - d.row_lo >= d.halo_rows_max, "cwa_slab_plan: row block of %d rows is smaller than its neighbours' halos (%d rows)",
                  d.row_hi - d.row_lo, d.halo_rows_max);
    }
    *out = d;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// object
// ---------------------------------------------------------------------------------------------
extern "C" int cwa_slab_create(cwa_ctx* ctx, const cwa_slab_desc* desc, cwa_sph hs, cwa_wave hw, cwa_slab* out)
{
    CWA_CHECK(ctx && desc && out, "null argument");
    DeviceGuard dg(ctx);
    *out = -1;
    SphObj* s = get_sph(ctx, hs);
    WaveObj* w = get_wave(ctx, hw);
    CWA_CHECK(s && w, "cwa_slab_create: invalid sph (%d) or wave (%d) handle", hs, hw);
    CWA_CHECK(s->grid >= 0, "cwa_slab_create: the SPH object needs a uniform grid");
    const cwa_slab_desc& d = *desc;
    CWA_CHECK(d.world >= 1 && d.world <= SLAB_MAX_WORLD && d.rank >= 0 && d.rank < d.world, "cwa_slab_create: rank %d / world %d", d.rank, d.world);
    CWA_CHECK(d.capacity > 0 && d.capacity <= s->capacity, "cwa_slab_create: capacity %d exceeds the SPH object's %d", d.capacity, s->capacity);
    BufferObj* pb = get_buffer(ctx, s->particles);
    CWA_CHECK(pb && (size_t)d.capacity * 64 <= pb->bytes, "cwa_slab_create: particle buffer smaller than the capacity %d", d.capacity);
    CWA_CHECK(d.cap_mig > 0 && d.cap_ghost > 0, "cwa_slab_create: message capacities must be positive");
    CWA_CHECK(w->w == d.wave_w && w->h_global == d.wave_h && w->ch == d.wave_ch && w->row0 == d.store_lo && w->h == d.store_hi - d.store_lo,
              "cwa_slab_create: the wave object must be the row block [%d, %d) of a %dx%d field", d.store_lo, d.store_hi, d.wave_w, d.wave_h);
    CWA_CHECK(((size_t)d.wave_w * d.wave_ch * 4) % 16 == 0, "cwa_slab_create: wave rows must be a multiple of 16 bytes");
    CWA_CHECK(d.world == 1 || w->last_row[0] != nullptr, "cwa_slab_create: the wave object must come from cwa_wave_create_block");
    SlabObj* sl = new SlabObj();
    sl->d = d; sl->sph = hs; sl->wave = hw;
    if (sl->d.timeout_ms <= 0) sl->d.timeout_ms = 10000;
    sl->part_bytes = (size_t)(1 + d.cap_mig + d.cap_ghost) * 64;
    sl->row_bytes = (size_t)d.wave_w * d.wave_ch * 4;
    sl->wave_bytes = (size_t)(d.halo_rows_max > 0 ? d.halo_rows_max : 1) * sl->row_bytes;
    sl->mail_bytes = SLAB_FLAG_BYTES + 4 * sl->part_bytes + 4 * sl->wave_bytes + (size_t)SLAB_RING * sl->row_bytes;
    auto fail = [&](const char* what) { cwa_set_error("cwa_slab_create: out of device memory (%s)", what); delete sl; return -2; };
    if (cudaMalloc(&sl->mail, sl->mail_bytes) != cudaSuccess) return fail("mailbox");
    CWA_CUDA(cudaMemsetAsync(sl->mail, 0, sl->mail_bytes, ctx->stream));
    for (int k = 0; k < 2; k++) {
        if (cudaMalloc(&sl->send[k], sl->part_bytes) != cudaSuccess) return fail("send buffers");
        CWA_CUDA(cudaMemsetAsync(sl->send[k], 0, 64, ctx->stream));
    }
    char* blk = nullptr;
    const size_t small = 2 * sizeof(SlabState) + 64 + 64 + (2 + 2 + SLAB_MAX_WORLD + 4) * sizeof(unsigned);
    if (cudaMalloc(&blk, small + 256) != cudaSuccess) return fail("state");
    CWA_CUDA(cudaMemsetAsync(blk, 0, small + 256, ctx->stream));
    sl->state = reinterpret_cast<SlabState*>(blk);
    sl->err = reinterpret_cast<int*>(blk + 64);
    sl->stats = reinterpret_cast<unsigned long long*>(blk + 128);
    sl->done = reinterpret_cast<unsigned*>(blk + 192);
    if (cudaMalloc(&sl->free_list, (size_t)d.capacity * 4) != cudaSuccess) return fail("free list");
    sl->arrivals_max = 2 * (2 * d.cap_mig + d.cap_ghost);
    if (cudaMalloc(&sl->arrivals, ((size_t)sl->arrivals_max + 1) * 4) != cudaSuccess) return fail("arrival list");
    cudaMemset(sl->arrivals, 0, ((size_t)sl->arrivals_max + 1) * 4);
    CWA_CUDA(cudaEventCreateWithFlags(&sl->ev_pack, cudaEventDisableTiming));
    CWA_CUDA(cudaEventCreateWithFlags(&sl->ev_int, cudaEventDisableTiming));
    CWA_CUDA(cudaEventCreateWithFlags(&sl->ev_wave, cudaEventDisableTiming));
    CWA_CUDA(cudaEventCreateWithFlags(&sl->ev_push, cudaEventDisableTiming));
    sl->peer[d.rank] = sl->mail;
    sl->live = true;
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));           // the mailbox is zero before any peer can be told about it
    ctx->slabs.push_back(sl);
    *out = (int)ctx->slabs.size() - 1;
    return 0;
}

static void slab_free(SlabObj* sl)
{
    if (!sl) return;
    for (int r = 0; r < SLAB_MAX_WORLD; r++)
        if (sl->peer_ipc[r] && sl->peer[r]) cudaIpcCloseMemHandle(sl->peer[r]);
    cudaFree(sl->mail); cudaFree(sl->send[0]); cudaFree(sl->send[1]); cudaFree(sl->state); cudaFree(sl->free_list); cudaFree(sl->arrivals);
    if (sl->ev_pack) cudaEventDestroy(sl->ev_pack);
    if (sl->ev_int) cudaEventDestroy(sl->ev_int);
    if (sl->ev_wave) cudaEventDestroy(sl->ev_wave);
    if (sl->ev_push) cudaEventDestroy(sl->ev_push);
    delete sl;
}

void slab_destroy_all(cwa_ctx* ctx)
{
    for (auto& p : ctx->slabs) { if (p) slab_free(p); p = nullptr; }
    ctx->slabs.clear();
}

extern "C" int cwa_slab_destroy(cwa_ctx* ctx, cwa_slab h)
{
    SlabObj* sl = get_slab(ctx, h);
    CWA_CHECK(sl, "invalid slab handle %d", h);
    DeviceGuard dg(ctx);
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->side_stream[0]));
    if (SphObj* s = get_sph(ctx, sl->sph)) s->n_dev = nullptr;
    slab_free(sl);
    ctx->slabs[h] = nullptr;
    return 0;
}

extern "C" int cwa_slab_mailbox(cwa_ctx* ctx, cwa_slab h, void** ptr, size_t* bytes)
{
    SlabObj* sl = get_slab(ctx, h);
    CWA_CHECK(sl, "invalid slab handle %d", h);
    if (ptr) *ptr = sl->mail;
    if (bytes) *bytes = sl->mail_bytes;
    return 0;
}

// 64-byte cudaIpcMemHandle_t of the mailbox, for ranks that live in other processes
extern "C" int cwa_slab_export(cwa_ctx* ctx, cwa_slab h, void* handle64)
{
    SlabObj* sl = get_slab(ctx, h);
    CWA_CHECK(sl && handle64, "invalid slab handle %d or null output", h);
    DeviceGuard dg(ctx);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t hnd;
    CWA_CUDA(cudaIpcGetMemHandle(&hnd, sl->mail));
    memcpy(handle64, &hnd, 64);
    return 0;
}

// tell this rank where the mailbox of `peer_rank` lives: an IPC handle exported by another process, or (same process) the pointer
// cwa_slab_mailbox returned -- peer access between the two devices is enabled when they differ
extern "C" int cwa_slab_connect(cwa_ctx* ctx, cwa_slab h, int peer_rank, const void* handle64, void* direct_ptr)
{
    SlabObj* sl = get_slab(ctx, h);
    CWA_CHECK(sl, "invalid slab handle %d", h);
    CWA_CHECK(peer_rank >= 0 && peer_rank < sl->d.world, "cwa_slab_connect: peer rank %d outside the world of %d", peer_rank, sl->d.world);
    if (peer_rank == sl->d.rank) return 0;
    DeviceGuard dg(ctx);
    if (direct_ptr != nullptr) {
        cudaPointerAttributes attr;
        CWA_CUDA(cudaPointerGetAttributes(&attr, direct_ptr));
        CWA_CHECK(attr.type == cudaMemoryTypeDevice, "cwa_slab_connect: the direct pointer is not device memory");
        if (attr.device != ctx->device) {
            int can = 0;
            CWA_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, attr.device));
            CWA_CHECK(can, "cwa_slab_connect: device %d cannot access device %d", ctx->device, attr.device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            else CWA_CUDA(e);
        }
        sl->peer[peer_rank] = (char*)direct_ptr; sl->peer_ipc[peer_rank] = false;
        return 0;
    }
    CWA_CHECK(handle64 != nullptr, "cwa_slab_connect: neither an IPC handle nor a pointer");
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle64, 64);
    void* p = nullptr;
    CWA_CUDA(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
    sl->peer[peer_rank] = (char*)p; sl->peer_ipc[peer_rank] = true;
    return 0;
}

// the owned range [0, n) of the particle SSBO was (re)written by the application: no free slots, no ghosts
extern "C" int cwa_slab_set_owned(cwa_ctx* ctx, cwa_slab h, int n_owned)
{
    SlabObj* sl = get_slab(ctx, h);
    CWA_CHECK(sl, "invalid slab handle %d", h);
    CWA_CHECK(n_owned >= 0 && n_owned <= sl->d.capacity, "cwa_slab_set_owned: %d particles exceed the capacity %d", n_owned, sl->d.capacity);
    DeviceGuard dg(ctx);
    { KScope k(ctx, KID_OTHER);
      slab_set_state_kernel<<<1, 32, 0, ctx->stream>>>(sl->state + (sl->pseq & 1u), n_owned); }
    CWA_CUDA(cudaGetLastError());
    sph_invalidate_for_buffer(ctx, get_sph(ctx, sl->sph)->particles);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// per-frame schedule
// ---------------------------------------------------------------------------------------------
static int slab_check_peers(SlabObj* sl)
{
    const cwa_slab_desc& d = sl->d;
    if (d.rank > 0) CWA_CHECK(sl->peer[d.rank - 1], "slab: the left neighbour (rank %d) is not connected", d.rank - 1);
    if (d.rank < d.world - 1) CWA_CHECK(sl->peer[d.rank + 1], "slab: the right neighbour (rank %d) is not connected", d.rank + 1);
    if (d.rank == d.world - 1)
        for (int r = 0; r < d.world - 1; r++) CWA_CHECK(sl->peer[r], "slab: rank %d is not connected (the last rank sends it the global last row)", r);
    return 0;
}

static SlabFlags* slab_flags(char* mail) { return reinterpret_cast<SlabFlags*>(mail); }

// push the packed particle messages (local send buffers) as message number `msg`
static int slab_push_particles(cwa_ctx* ctx, SlabObj* sl, unsigned msg)
{
    const cwa_slab_desc& d = sl->d;
    if (d.world == 1) return 0;
    const int par = (int)(msg & 1u);
    const bool has_left = d.rank > 0, has_right = d.rank < d.world - 1;
    // my message to the LEFT neighbour lands in its "from the right" slot (side 1) and vice versa
    char* pl = has_left ? sl->peer[d.rank - 1] : nullptr;
    char* pr = has_right ? sl->peer[d.rank + 1] : nullptr;
    float4* dst_l = pl ? reinterpret_cast<float4*>(pl + slab_part_off(sl, 1, par)) : nullptr;
    float4* dst_r = pr ? reinterpret_cast<float4*>(pr + slab_part_off(sl, 0, par)) : nullptr;
    unsigned* fl = pl ? &slab_flags(pl)->part[1][par] : nullptr;
    unsigned* fr = pr ? &slab_flags(pr)->part[0][par] : nullptr;
    KScope k(ctx, KID_EXCHANGE);
    slab_push_particles_kernel<<<dim3(SLAB_PUSH_CTAS, 2), 256, 0, ctx->stream>>>(has_left ? sl->send[0] : nullptr, has_right ? sl->send[1] : nullptr,
                                                                               dst_l, dst_r, fl, fr, msg + 1u, d.cap_mig, d.cap_ghost, sl->done);
    CWA_CUDA(cudaGetLastError());
    return 0;
}

// push the halo rows of physical image `img` to the neighbours and (last rank) its last row to everybody, as wave message `msg`
static int slab_push_wave(cwa_ctx* ctx, SlabObj* sl, WaveObj* w, int img, unsigned msg)
{
    const cwa_slab_desc& d = sl->d;
    if (d.world == 1) return 0;
    const int par = (int)(msg & 1u);
    const size_t rb = sl->row_bytes;
    SlabPushArgs a;
    memset(&a, 0, sizeof(a));
    const char* base = reinterpret_cast<const char*>(w->image[img]);
    int nseg = 2;
    if (d.rank > 0) {                                      // rows [row_lo, left_store_hi): the left neighbour's upper halo -> its side 1
        char* p = sl->peer[d.rank - 1];
        a.seg[0].src = reinterpret_cast<const float4*>(base + (size_t)(d.row_lo - d.store_lo) * rb);
        a.seg[0].dst = reinterpret_cast<float4*>(p + slab_wave_off(sl, 1, par));
        a.seg[0].n16 = (unsigned long long)(d.left_store_hi - d.row_lo) * rb / 16;
        a.seg[0].flag = &slab_flags(p)->wave[1][par]; a.seg[0].value = msg + 1u;
    }
    if (d.rank < d.world - 1) {                            // rows [right_store_lo, row_hi): the right neighbour's lower halo -> its side 0
        char* p = sl->peer[d.rank + 1];
        a.seg[1].src = reinterpret_cast<const float4*>(base + (size_t)(d.right_store_lo - d.store_lo) * rb);
        a.seg[1].dst = reinterpret_cast<float4*>(p + slab_wave_off(sl, 0, par));
        a.seg[1].n16 = (unsigned long long)(d.row_hi - d.right_store_lo) * rb / 16;
        a.seg[1].flag = &slab_flags(p)->wave[0][par]; a.seg[1].value = msg + 1u;
    }
    if (d.rank == d.world - 1) {                           // global last row -> ring slot of every other rank
        const int slot = (int)(msg % SLAB_RING);
        for (int r = 0; r < d.world - 1; r++) {
            char* p = sl->peer[r];
            SlabPushSeg& s = a.seg[nseg++];
            s.src = reinterpret_cast<const float4*>(base + (size_t)(d.wave_h - 1 - d.store_lo) * rb);
            s.dst = reinterpret_cast<float4*>(p + slab_last_off(sl, slot));
            s.n16 = rb / 16;
            s.flag = &slab_flags(p)->last[slot]; s.value = msg + 1u;
        }
    }
    KScope k(ctx, KID_EXCHANGE);
    slab_push_rows_kernel<<<dim3(SLAB_PUSH_CTAS, nseg), 256, 0, ctx->stream>>>(a, sl->done + 2);
    CWA_CUDA(cudaGetLastError());
    return 0;
}

// wait for wave message `msg` and copy it into the halo rows / last-row copy of physical image `img`
static int slab_recv_wave(cwa_ctx* ctx, SlabObj* sl, WaveObj* w, int img, unsigned msg)
{
    const cwa_slab_desc& d = sl->d;
    const int par = (int)(msg & 1u);
    const size_t rb = sl->row_bytes;
    const int slot = (int)(msg % SLAB_RING);
    SlabFlags* f = slab_flags(sl->mail);
    if (d.world > 1) {
        SlabWaitArgs wa;
        memset(&wa, 0, sizeof(wa));
        wa.n = 3;
        if (d.rank > 0) { wa.flag[0] = &f->wave[0][par]; wa.expect[0] = msg + 1u; }
        if (d.rank < d.world - 1) { wa.flag[1] = &f->wave[1][par]; wa.expect[1] = msg + 1u; }
        if (d.rank != d.world - 1) { wa.flag[2] = &f->last[slot]; wa.expect[2] = msg + 1u; }
        KScope k(ctx, KID_EXCHANGE);
        slab_wait_kernel<<<1, 32, 0, ctx->stream>>>(wa, sl->err, SLAB_ERR_TIMEOUT_WAVE, (unsigned long long)sl->d.timeout_ms * 1000000ull);
        CWA_CUDA(cudaGetLastError());
    }
    if (w->last_row[img] == nullptr) return 0;             // whole-field object in a world of one: nothing to copy
    SlabCopyArgs c;
    memset(&c, 0, sizeof(c));
    char* base = reinterpret_cast<char*>(w->image[img]);
    if (d.rank > 0) {                                      // lower halo [store_lo, row_lo) <- left neighbour
        c.seg[0].src = reinterpret_cast<const float4*>(sl->mail + slab_wave_off(sl, 0, par));
        c.seg[0].dst = reinterpret_cast<float4*>(base);
        c.seg[0].n16 = (unsigned long long)(d.row_lo - d.store_lo) * rb / 16;
    }
    if (d.rank < d.world - 1) {                            // upper halo [row_hi, store_hi) <- right neighbour
        c.seg[1].src = reinterpret_cast<const float4*>(sl->mail + slab_wave_off(sl, 1, par));
        c.seg[1].dst = reinterpret_cast<float4*>(base + (size_t)(d.row_hi - d.store_lo) * rb);
        c.seg[1].n16 = (unsigned long long)(d.store_hi - d.row_hi) * rb / 16;
    }
    c.seg[2].src = (d.rank == d.world - 1) ? reinterpret_cast<const float4*>(base + (size_t)(d.wave_h - 1 - d.store_lo) * rb)
                                           : reinterpret_cast<const float4*>(sl->mail + slab_last_off(sl, slot));
    c.seg[2].dst = reinterpret_cast<float4*>(w->last_row[img]);
    c.seg[2].n16 = rb / 16;
    { KScope k(ctx, KID_EXCHANGE);
      slab_copy_rows_kernel<<<dim3(32, 3), 256, 0, ctx->stream>>>(c); }
    CWA_CUDA(cudaGetLastError());
    w->version[img]++;                                     // derived copies (transposed sampling copy) of this image are stale
    return 0;
}

// Wave message exchange for the CURRENT newest level (first use, or after the application rewrote the field): push part
static int slab_sync_wave_push(cwa_ctx* ctx, SlabObj* sl)
{
    WaveObj* w = get_wave(ctx, sl->wave);
    CWA_CHECK(w, "slab: wave object vanished");
    CWA_TRY(slab_check_peers(sl));
    return slab_push_wave(ctx, sl, w, wave_image_with_unit(w, 0), sl->wseq);
}
static int slab_sync_wave_recv(cwa_ctx* ctx, SlabObj* sl)
{
    WaveObj* w = get_wave(ctx, sl->wave);
    CWA_CHECK(w, "slab: wave object vanished");
    CWA_TRY(slab_recv_wave(ctx, sl, w, wave_image_with_unit(w, 0), sl->wseq));
    sl->wseq++;
    sl->wave_synced = true;
    return 0;
}

// The send buffers are about to be rewritten on the main stream: a push of the previous message (side stream) must be done with them
static int slab_send_buffers_free(cwa_ctx* ctx, SlabObj* sl)
{
    if (sl->push_pending) CWA_CUDA(cudaStreamWaitEvent(ctx->stream, sl->ev_push, 0));
    sl->push_pending = false;
    return 0;
}

// first frame of a call: the call starts from a complete state (every particle in its owner's buffer), so its messages are
// packed by an explicit pass over the owned range and pushed
static int slab_frame_pack_first(cwa_ctx* ctx, SlabObj* sl)
{
    const cwa_slab_desc& d = sl->d;
    if (d.world == 1) return 0;
    SphObj* s = get_sph(ctx, sl->sph);
    BufferObj* pb = s ? get_buffer(ctx, s->particles) : nullptr;
    CWA_CHECK(pb, "slab: sph object or particle buffer vanished");
    cudaStream_t M = ctx->stream, W = ctx->side_stream[0];
    float4* snd_l = d.rank > 0 ? sl->send[0] : nullptr;
    float4* snd_r = d.rank < d.world - 1 ? sl->send[1] : nullptr;
    CWA_TRY(slab_send_buffers_free(ctx, sl));
    if (snd_l) CWA_CUDA(cudaMemsetAsync(snd_l, 0, 64, M));
    if (snd_r) CWA_CUDA(cudaMemsetAsync(snd_r, 0, 64, M));
    { KScope k(ctx, KID_EXCHANGE);
      slab2_pack_kernel<<<ceil_div(d.capacity, 256), 256, 0, M>>>((float4*)pb->ptr, sl->state + (sl->pseq & 1u), d.z_lo, d.z_hi, d.band, snd_l, snd_r,
                                                                d.cap_mig, d.cap_ghost, sl->free_list); }
    CWA_CUDA(cudaGetLastError());
    CWA_CUDA(cudaEventRecord(sl->ev_pack, M));
    CWA_CUDA(cudaStreamWaitEvent(W, sl->ev_pack, 0));
    { StreamScope ss(ctx, W); CWA_TRY(slab_push_particles(ctx, sl, sl->pseq)); }
    CWA_CUDA(cudaEventRecord(sl->ev_push, W));
    sl->push_pending = true;
    return 0;
}

// tuning "slab_ahead" / CWA_SLAB_AHEAD (default 1): count-ahead across the exchange
static bool slab_count_ahead(cwa_ctx* c)
{
    if (c->tune.slab_ahead < 0) { const char* e = getenv("CWA_SLAB_AHEAD"); c->tune.slab_ahead = (e && atoi(e) == 0) ? 0 : 1; }
    return c->tune.slab_ahead != 0;
}

// part 1 of a frame: everything up to and including this rank's pushes.  Nothing in it waits for a push that another rank enqueues
// in the SAME part of the SAME frame, so ranks that share a host thread (cwa_slab_group_step) can be enqueued one after the other.
static int slab_frame_begin(cwa_ctx* ctx, SlabObj* sl, bool last, int coupling)
{
    SphObj* s = get_sph(ctx, sl->sph);
    WaveObj* w = get_wave(ctx, sl->wave);
    CWA_CHECK(s && w, "slab: sph or wave object vanished");
    BufferObj* pb = get_buffer(ctx, s->particles);
    CWA_CHECK(pb, "slab: particle buffer vanished");
    const cwa_slab_desc& d = sl->d;
    const bool has_left = d.rank > 0, has_right = d.rank < d.world - 1;
    cudaStream_t M = ctx->stream, W = ctx->side_stream[0];
    const unsigned msg = sl->pseq;
    const int par = (int)(msg & 1u);
    SlabState* sin = sl->state + par;
    SlabState* sout = sl->state + (par ^ 1);
    float4* aos = (float4*)pb->ptr;
    float4* snd_l = has_left ? sl->send[0] : nullptr;
    float4* snd_r = has_right ? sl->send[1] : nullptr;

    if (d.world > 1) {
        SlabFlags* f = slab_flags(sl->mail);
        SlabWaitArgs wa;
        memset(&wa, 0, sizeof(wa));
        wa.n = 2;
        if (has_left) { wa.flag[0] = &f->part[0][par]; wa.expect[0] = msg + 1u; }
        if (has_right) { wa.flag[1] = &f->part[1][par]; wa.expect[1] = msg + 1u; }
        KScope k(ctx, KID_EXCHANGE);
        slab_wait_kernel<<<1, 32, 0, M>>>(wa, sl->err, SLAB_ERR_TIMEOUT_PARTICLES, (unsigned long long)d.timeout_ms * 1000000ull);
        CWA_CUDA(cudaGetLastError());
    }
    {
        const float4* rcv_l = has_left ? reinterpret_cast<const float4*>(sl->mail + slab_part_off(sl, 0, par)) : nullptr;
        const float4* rcv_r = has_right ? reinterpret_cast<const float4*>(sl->mail + slab_part_off(sl, 1, par)) : nullptr;
        const long long max_in = 4ll * (2ll * (2ll * d.cap_mig + d.cap_ghost));
        KScope k(ctx, KID_EXCHANGE);
        SlabAheadArgs ah;
        memset(&ah, 0, sizeof(ah));
        GridObj* g = get_grid(ctx, s->grid);
        if (s->counts_ahead && g != nullptr && d.world > 1) {          // the last frame's integrate pass counted the particles that stayed
            ah.g = g->view; ah.counter = g->counter; ah.cell_of = g->cell_of; ah.rank = g->rank;
            ah.arrivals = sl->arrivals; ah.arrivals_max = sl->arrivals_max;
            s->arrivals = sl->arrivals; s->arrivals_count = sl->arrivals + sl->arrivals_max; s->arrivals_max = sl->arrivals_max;
        } else {
            s->arrivals = nullptr; s->arrivals_count = nullptr; s->arrivals_max = 0;
        }
        slab2_unpack_kernel<<<d.world > 1 ? ceil_div(max_in, 256) : 1, 256, 0, M>>>(aos, d.capacity, rcv_l, rcv_r, snd_l, snd_r, d.cap_mig, d.cap_ghost,
                                                                                  sin, sout, sl->free_list, sl->err, sl->stats, ah);
        CWA_CUDA(cudaGetLastError());
    }
    sl->pseq++;

    // ---- idle(): rho_pres, force, integrate (Main.cpp:549-557) on owned + ghosts; counts are read from `sout` on the device
    s->n = d.capacity;
    s->n_dev = &sout->n_total;
    s->snapshot_valid = false;
    int image;
    if (coupling == CWA_COUPLING_LATEST) image = wave_image_with_unit(w, 0);
    else image = w->tex_unit0;                             // whatever display() last bound to texture unit 0
    s->wave = sl->wave; s->wave_image = image;
    s->wait_before_sampling = sl->wave_in_flight ? sl->ev_wave : nullptr;   // the previous frame's stencil + halo receive (side stream)
    const bool pack_next = !last && d.world > 1;
    SlabPackArgs pa;
    if (pack_next) {
        CWA_TRY(slab_send_buffers_free(ctx, sl));
        if (snd_l) CWA_CUDA(cudaMemsetAsync(snd_l, 0, 64, M));
        if (snd_r) CWA_CUDA(cudaMemsetAsync(snd_r, 0, 64, M));
        pa.n_owned = d.capacity; pa.n_owned_dev = &sout->n_owned; pa.z_lo = d.z_lo; pa.z_hi = d.z_hi; pa.band = d.band;
        pa.msg_l = snd_l; pa.msg_r = snd_r; pa.cap_mig = d.cap_mig; pa.cap_ghost = d.cap_ghost;
        pa.free_list = sl->free_list; pa.free_count = &sout->free_count;
    }
    const bool ahead = pack_next && slab_count_ahead(ctx);      // the integrate pass counts the particles that stay, the next unpack the arrivals
    const int rc = sph_passes_internal(ctx, s, wave_tex_view(ctx, sl->wave, image), 7, ahead, pack_next ? &pa : nullptr);
    if (s->wait_before_sampling) {                         // not consumed (the call failed early)
        cudaStreamWaitEvent(M, s->wait_before_sampling, 0);
        s->wait_before_sampling = nullptr;
    }
    sl->wave_in_flight = false;
    CWA_TRY(rc);
    s->slab_packed.valid = false;
    CWA_CUDA(cudaEventRecord(sl->ev_int, M));
    CWA_CUDA(cudaStreamWaitEvent(W, sl->ev_int, 0));       // the stencil overwrites a level the SPH passes may have sampled
    {
        StreamScope ss(ctx, W);
        if (pack_next) {                                   // first: the neighbours' next frame waits for it
            CWA_TRY(slab_push_particles(ctx, sl, sl->pseq));
            CWA_CUDA(cudaEventRecord(sl->ev_push, W));
            sl->push_pending = true;
        }
        if (w->evolve) {                                   // Module::sComputeAll (Main.cpp:560)
            CWA_TRY(wave_step_internal(ctx, w));
            CWA_TRY(slab_push_wave(ctx, sl, w, wave_image_with_unit(w, 0), sl->wseq));
        }
    }
    return 0;
}

// part 2: receive the neighbours' wave rows of this frame; display(): GetReadImage(0).BindTextureUnit() (Main.cpp:413)
static int slab_frame_end(cwa_ctx* ctx, SlabObj* sl)
{
    WaveObj* w = get_wave(ctx, sl->wave);
    CWA_CHECK(w, "slab: wave object vanished");
    if (w->evolve) {
        StreamScope ss(ctx, ctx->side_stream[0]);
        CWA_TRY(slab_recv_wave(ctx, sl, w, wave_image_with_unit(w, 0), sl->wseq));
        sl->wseq++;
    }
    CWA_CUDA(cudaEventRecord(sl->ev_wave, ctx->side_stream[0]));
    sl->wave_in_flight = true;
    CWA_TRY(cwa_wave_bind_texture_unit(ctx, sl->wave));
    return 0;
}

static int slab_call_end(cwa_ctx* ctx, SlabObj* sl)
{
    if (sl->wave_in_flight) CWA_CUDA(cudaStreamWaitEvent(ctx->stream, sl->ev_wave, 0));   // the main stream is behind everything the side stream was given
    sl->wave_in_flight = false;
    if (SphObj* s = get_sph(ctx, sl->sph)) { s->counts_ahead = false; s->arrivals = nullptr; s->arrivals_count = nullptr; s->arrivals_max = 0; }
    return 0;
}

// nframes coupled frames of this rank's slab (every rank of the world must make the same calls)
extern "C" int cwa_slab_step(cwa_ctx* ctx, cwa_slab h, int nframes, int coupling)
{
    SlabObj* sl = get_slab(ctx, h);
    CWA_CHECK(sl, "invalid slab handle %d", h);
    CWA_CHECK(coupling == CWA_COUPLING_AS_SHIPPED || coupling == CWA_COUPLING_LATEST, "unknown coupling mode %d", coupling);
    DeviceGuard dg(ctx);
    CWA_TRY(slab_check_peers(sl));
    if (!sl->wave_synced) {                               // first use: halo rows + last row of the newest level as it stands
        CWA_CUDA(cudaEventRecord(sl->ev_pack, ctx->stream));
        CWA_CUDA(cudaStreamWaitEvent(ctx->side_stream[0], sl->ev_pack, 0));
        StreamScope ss(ctx, ctx->side_stream[0]);
        CWA_TRY(slab_sync_wave_push(ctx, sl));
        CWA_TRY(slab_sync_wave_recv(ctx, sl));
        CWA_CUDA(cudaEventRecord(sl->ev_wave, ctx->stream));
        sl->wave_in_flight = true;
    }
    for (int f = 0; f < nframes; f++) {
        if (f == 0) CWA_TRY(slab_frame_pack_first(ctx, sl));
        CWA_TRY(slab_frame_begin(ctx, sl, f + 1 == nframes, coupling));
        CWA_TRY(slab_frame_end(ctx, sl));
    }
    return slab_call_end(ctx, sl);
}

// the same for ranks that are contexts of ONE process and host thread (e.g. a C++ host driving all GPUs of the box): every phase
// that waits for peers is enqueued only after every rank's pushes of that phase
extern "C" int cwa_slab_group_step(cwa_ctx* const* ctxs, const cwa_slab* slabs, int n, int nframes, int coupling)
{
    CWA_CHECK(ctxs && slabs && n >= 1 && n <= SLAB_MAX_WORLD, "cwa_slab_group_step: bad arguments");
    CWA_CHECK(coupling == CWA_COUPLING_AS_SHIPPED || coupling == CWA_COUPLING_LATEST, "unknown coupling mode %d", coupling);
    SlabObj* sl[SLAB_MAX_WORLD];
    for (int r = 0; r < n; r++) {
        sl[r] = get_slab(ctxs[r], slabs[r]);
        CWA_CHECK(sl[r], "cwa_slab_group_step: invalid slab handle of rank %d", r);
        CWA_TRY(slab_check_peers(sl[r]));
    }
    bool need_sync = false;
    for (int r = 0; r < n; r++) need_sync = need_sync || !sl[r]->wave_synced;
    if (need_sync) {
        for (int r = 0; r < n; r++) {
            cwa_ctx* c = ctxs[r];
            DeviceGuard dg(c);
            CWA_CUDA(cudaEventRecord(sl[r]->ev_pack, c->stream));
            CWA_CUDA(cudaStreamWaitEvent(c->side_stream[0], sl[r]->ev_pack, 0));
            StreamScope ss(c, c->side_stream[0]);
            CWA_TRY(slab_sync_wave_push(c, sl[r]));
        }
        for (int r = 0; r < n; r++) {
            cwa_ctx* c = ctxs[r];
            DeviceGuard dg(c);
            StreamScope ss(c, c->side_stream[0]);
            CWA_TRY(slab_sync_wave_recv(c, sl[r]));
            CWA_CUDA(cudaEventRecord(sl[r]->ev_wave, c->stream));
            sl[r]->wave_in_flight = true;
        }
    }
    for (int f = 0; f < nframes; f++) {
        // the explicit pack + push of a call's first frame is enqueued on EVERY rank before any rank's wait
        if (f == 0)
            for (int r = 0; r < n; r++) { DeviceGuard dg(ctxs[r]); CWA_TRY(slab_frame_pack_first(ctxs[r], sl[r])); }
        for (int r = 0; r < n; r++) { DeviceGuard dg(ctxs[r]); CWA_TRY(slab_frame_begin(ctxs[r], sl[r], f + 1 == nframes, coupling)); }
        for (int r = 0; r < n; r++) { DeviceGuard dg(ctxs[r]); CWA_TRY(slab_frame_end(ctxs[r], sl[r])); }
    }
    for (int r = 0; r < n; r++) { DeviceGuard dg(ctxs[r]); CWA_TRY(slab_call_end(ctxs[r], sl[r])); }
    return 0;
}

// counts[8] = {owned range (dead slots included), owned + ghosts, free slots, error bits, migrants adopted (low, high 32 bits),
// particle messages consumed, wave messages consumed}; synchronises
extern "C" int cwa_slab_counts(cwa_ctx* ctx, cwa_slab h, int* counts)
{
    SlabObj* sl = get_slab(ctx, h);
    CWA_CHECK(sl && counts, "invalid slab handle %d or null output", h);
    DeviceGuard dg(ctx);
    SlabState st;
    int err = 0;
    unsigned long long mig = 0;
    CWA_CUDA(cudaMemcpyAsync(&st, sl->state + (sl->pseq & 1u), sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaMemcpyAsync(&err, sl->err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaMemcpyAsync(&mig, sl->stats, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    counts[0] = st.n_owned; counts[1] = st.n_total; counts[2] = st.free_count; counts[3] = err;
    counts[4] = (int)(mig & 0xffffffffull); counts[5] = (int)(mig >> 32);
    counts[6] = (int)sl->pseq; counts[7] = (int)sl->wseq;
    return 0;
}

// enqueue only: `pinned_counts8` (pinned host memory) holds {owned range, owned + ghosts, free slots, error bits} after the next
// cwa_synchronize -- for hosts that read the owned range back every step without a second blocking call
extern "C" int cwa_slab_counts_async(cwa_ctx* ctx, cwa_slab h, int* pinned_counts8)
{
    SlabObj* sl = get_slab(ctx, h);
    CWA_CHECK(sl && pinned_counts8, "invalid slab handle %d or null output", h);
    DeviceGuard dg(ctx);
    CWA_CUDA(cudaMemcpyAsync(pinned_counts8, sl->state + (sl->pseq & 1u), 12, cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaMemcpyAsync(pinned_counts8 + 3, sl->err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}
