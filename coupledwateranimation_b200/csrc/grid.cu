// grid.cu -- uniform-grid neighbour build for sm_100a:
//   (1) cell hash + warp-aggregated atomic count   <- ComputeCounts   (uniform_grid_sph_cs.glsl:112-125,
//                                                                      ugrid_particles_cs.glsl:90-95)
//   (2) single-pass decoupled-look-back scan       <- ParallelScan::Compute (ParallelScan.cpp:43-95,
//                                                                      prefix_sum_cs.glsl:18-43)
//   (3) insert + canonical per-cell ordering       <- InsertPoint / InsertParticle
//                                                     (uniform_grid_sph_cs.glsl:154-165, ugrid_particles_cs.glsl:110-117)
// The reference clears the counter twice with glCopyNamedBufferSubData and runs 2*log2(C)
// dispatches for the scan; here one memset + three kernels (+ a tiny per-cell ordering pass that
// makes the index list deterministic: ascending particle id inside a cell).
#include "internal.cuh"

// ---------------------------------------------------------------------------------------------
// (1) hash + count
// ---------------------------------------------------------------------------------------------
// One thread per particle.  Lanes of a warp that fall in the same cell are grouped with
// __match_any_sync; the group leader issues ONE atomicAdd for the whole group and every lane
// derives its arrival rank from the returned base, so the later insert needs no second atomic
// pass (the reference runs the atomics twice: count, clear, insert).
template <int DIM>
__global__ void __launch_bounds__(256)
grid_hash_count_kernel(const char* __restrict__ particles, int stride_bytes, int n, const int* __restrict__ n_dev, GridView g,
                       int* __restrict__ counter, int* __restrict__ cell_of, int* __restrict__ rank)
{
    CWA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev != nullptr) n = min(n, __ldg(n_dev));     // device-resident particle count (slab decomposition): `n` is only the launch bound
    int cell = -1;
    if (i < n) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(particles + (size_t)i * stride_bytes));
        if (DIM == 2) {
            // point_in_aabb: strict on both axes (uniform_grid_sph_cs.glsl:19-23,115)
            const bool inside = (p.x > g.min[0]) && (p.y > g.min[1]) && (p.x < g.max[0]) && (p.y < g.max[1]);
            if (inside) {
                int ci, cj;
                cwa_cell2(g, p.x, p.y, ci, cj);
                cell = ci * g.n[1] + cj;                       // Index(i,j) :149-152
            }
        } else if (p.x == p.x && p.y == p.y && p.z == p.z) {
            // ivec3(floor(NaN)) is undefined in GLSL; a NaN particle can never pass a distance test
            // again, so the canonical choice (shared with the oracle) is: not inserted.
            int ci, cj, ck;
            cwa_cell3(g, p.x, p.y, p.z, ci, cj, ck);
            cell = (ci * g.n[1] + cj) * g.kstride + ck;        // Index(i,j,k) ugrid_particles_cs.glsl:105-108: kstride = Nx (sic)
        }
        cell_of[i] = cell;
    }
    // warp aggregation (all 32 lanes participate; lanes without a cell use key -1 and skip)
    const unsigned lane = threadIdx.x & 31u;
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    const int leader = __ffs(peers) - 1;
    const int my_rank = __popc(peers & ((1u << lane) - 1u));
    int base = 0;
    if (cell >= 0 && (int)lane == leader) base = atomicAdd(&counter[cell], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < n && cell >= 0) rank[i] = base + my_rank;
}

// ---------------------------------------------------------------------------------------------
// (2) decoupled look-back exclusive scan (single pass, any n)
// ---------------------------------------------------------------------------------------------
// Tile shapes (cwa_set_tuning "scan_config"): the look-back is a serial chain over the tiles in front of a tile, 32 per step, and with
// a few million cells every tile of the grid is resident at once -- so the chain length, not the bandwidth, sets the time.  Large tiles
// (16 K items: 279 tiles for the 4.6 M cells of C4 instead of 1116) shorten it 4x.
//   0: 256 threads x 16 items, blocked (thread t owns 16 consecutive items)
//   1: 512 x 32 blocked
//   2: 512 x 32 warp-striped (a warp owns 1024 consecutive items; lane l loads the int4 at 4*(32 q + l), q = 0..7: every load
//      instruction of a warp covers 512 contiguous bytes)
//   3: 1024 x 16 warp-striped
constexpr int SCAN_MIN_TILE = 4096;                     // smallest tile of any configuration: sizes the tile-state arrays
static int scan_config(cwa_ctx* ctx)
{
    if (ctx->tune.scan_config < 0) {
        const char* e = getenv("CWA_SCAN_CONFIG");
        const int v = e ? atoi(e) : 2;
        ctx->tune.scan_config = (v < 0 || v > 3) ? 2 : v;
    }
    return ctx->tune.scan_config;
}

size_t scan_num_tiles(int n) { return (size_t)((n + SCAN_MIN_TILE - 1) / SCAN_MIN_TILE); }

// tile_state word: [63:62] flag (0 invalid, 1 aggregate, 2 inclusive prefix) | [31:0] value
__device__ __forceinline__ unsigned long long scan_pack(unsigned flag, int v)
{
    return ((unsigned long long)flag << 62) | (unsigned long long)(unsigned)v;
}

template <int THREADS, int ITEMS, bool STRIPED>
__global__ void __launch_bounds__(THREADS)
scan_lookback_kernel(const int* __restrict__ in, int* __restrict__ out, int n, int write_total,
                     int* __restrict__ ticket, volatile unsigned long long* __restrict__ tile_state)
{
    CWA_PDL_ENTER();
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    constexpr int Q = ITEMS / 4;
    __shared__ int s_tile;
    __shared__ int s_warp[WARPS];
    __shared__ int s_excl[WARPS];
    __shared__ int s_flag[WARPS];
    __shared__ int s_prefix, s_done, s_aggregate, s_prefix_final;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1);        // tiles are handed out in start order
    __syncthreads();
    const int tile = s_tile;
    const long long tile_base = (long long)tile * TILE;
    // first item of this thread's q-th int4: blocked = consecutive per thread, striped = consecutive per warp instruction
    const long long base = STRIPED ? tile_base + (long long)wid * (32 * ITEMS) + 4 * lane
                                   : tile_base + (long long)tid * ITEMS;
    constexpr int QSTRIDE = STRIPED ? 128 : 4;
    const bool full_tile = tile_base + TILE <= n;

    int v[ITEMS];
    if (full_tile) {
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int4 t = __ldg(reinterpret_cast<const int4*>(in + base + q * QSTRIDE));
            v[4 * q + 0] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < Q; q++)
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const long long idx = base + q * QSTRIDE + e;
                v[4 * q + e] = (idx < n) ? __ldg(in + idx) : 0;
            }
    }

    // exclusive offset of each of the thread's int4 groups inside the warp's segment (striped) / of the thread's run (blocked)
    int gexcl[Q];
    int incl;            // inclusive scan of the per-thread totals across the warp, in item order
    int tsum = 0;
    if (STRIPED) {
        int carry = 0;   // items of the warp's strips 0..q-1
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int g4 = v[4 * q] + v[4 * q + 1] + v[4 * q + 2] + v[4 * q + 3];
            int sc = g4;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, sc, d);
                if (lane >= d) sc += t;
            }
            gexcl[q] = carry + sc - g4;
            carry += __shfl_sync(0xffffffffu, sc, 31);
        }
        tsum = carry;    // the whole warp's total (same on every lane)
        incl = carry;
    } else {
#pragma unroll
        for (int q = 0; q < Q; q++) {
            gexcl[q] = tsum;
            tsum += v[4 * q] + v[4 * q + 1] + v[4 * q + 2] + v[4 * q + 3];
        }
        incl = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
    }
    if (lane == 31) s_warp[wid] = incl;                 // warp total
    __syncthreads();
    int warp_excl = 0;
    if (wid == 0) {
        const int w = (lane < WARPS) ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < WARPS; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < WARPS) s_excl[lane] = wi - w;          // exclusive warp offsets
        const int aggregate = __shfl_sync(0xffffffffu, wi, WARPS - 1);
        if (lane == 0) {
            s_aggregate = aggregate;
            tile_state[tile] = scan_pack(tile == 0 ? 2u : 1u, aggregate);      // publish first, then look back
        }
    }
    __syncthreads();
    warp_excl = s_excl[wid];

    // Look-back over the predecessors, THREADS tiles per step (one tile state per thread): with every tile of a few-million-cell
    // grid resident at once the look-back is a pure latency chain, so its length is what counts -- 279 tiles of 16 K items are
    // covered in ONE step.  Per step: every thread waits for its predecessor's word, each warp reduces (values up to and including
    // its nearest inclusive prefix | all values), thread 0 walks the warps from the nearest tile backwards.
    if (tile > 0) {
        int exclusive = 0;
        int look = tile - 1;
        while (true) {
            const int idx = look - tid;
            unsigned long long st = scan_pack(2u, 0);                  // tiles before 0: prefix 0
            if (idx >= 0) {
                do { st = tile_state[idx]; } while ((st >> 62) == 0ull);
            }
            const unsigned flag = (unsigned)(st >> 62);
            const int val = (int)(unsigned)(st & 0xffffffffull);
            const unsigned has_prefix = __ballot_sync(0xffffffffu, flag == 2u);
            const int first = has_prefix ? (__ffs(has_prefix) - 1) : 31;
            int contrib = (lane <= first) ? val : 0;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
            __syncthreads();                                           // s_warp / s_flag of the previous step are consumed
            if (lane == 0) { s_warp[wid] = contrib; s_flag[wid] = has_prefix != 0u; }
            __syncthreads();
            if (tid == 0) {
                int acc = 0, done = 0;
                for (int wq = 0; wq < WARPS; wq++) { acc += s_warp[wq]; if (s_flag[wq]) { done = 1; break; } }
                s_prefix = acc; s_done = done;
            }
            __syncthreads();
            exclusive += s_prefix;
            if (s_done) break;
            look -= THREADS;
        }
        if (tid == 0) tile_state[tile] = scan_pack(2u, exclusive + s_aggregate);
        s_prefix_final = exclusive;      // every thread holds the same value; written by all, read by all after the barrier below
    } else if (tid == 0) {
        s_prefix_final = 0;
    }
    __syncthreads();

    const int thread_base = s_prefix_final + warp_excl + (STRIPED ? 0 : (incl - tsum));
#pragma unroll
    for (int q = 0; q < Q; q++) {
        int run = thread_base + gexcl[q];
        int4 t;
        t.x = run; run += v[4 * q + 0];
        t.y = run; run += v[4 * q + 1];
        t.z = run; run += v[4 * q + 2];
        t.w = run; run += v[4 * q + 3];
        const long long idx = base + q * QSTRIDE;
        if (full_tile) {
            *reinterpret_cast<int4*>(out + idx) = t;
        } else {
            if (idx + 0 < n) out[idx + 0] = t.x;
            if (idx + 1 < n) out[idx + 1] = t.y;
            if (idx + 2 < n) out[idx + 2] = t.z;
            if (idx + 3 < n) out[idx + 3] = t.w;
        }
        // the thread whose group holds element n-1 also writes the grand total to out[n] (elements past n are zeros)
        if (write_total && idx <= (long long)n - 1 && (long long)n - 1 < idx + 4) out[n] = run;
    }
}

int scan_exclusive_launch(cwa_ctx* ctx, const int* in, int* out, int n, int* ticket, unsigned long long* tile_state)
{
    // n < 0 encodes "do not write the total at out[|n|]"
    const int write_total = n > 0;
    if (n < 0) n = -n;
    KScope k(ctx, KID_SCAN);
    switch (scan_config(ctx)) {
    case 0: cwa_launch(ctx, PDL_SCAN, scan_lookback_kernel<256, 16, false>, dim3(ceil_div(n, 256 * 16)), dim3(256), 0, in, out, n, write_total, ticket, tile_state); break;
    case 1: cwa_launch(ctx, PDL_SCAN, scan_lookback_kernel<512, 32, false>, dim3(ceil_div(n, 512 * 32)), dim3(512), 0, in, out, n, write_total, ticket, tile_state); break;
    case 3: cwa_launch(ctx, PDL_SCAN, scan_lookback_kernel<1024, 16, true>, dim3(ceil_div(n, 1024 * 16)), dim3(1024), 0, in, out, n, write_total, ticket, tile_state); break;
    default: cwa_launch(ctx, PDL_SCAN, scan_lookback_kernel<512, 32, true>, dim3(ceil_div(n, 512 * 32)), dim3(512), 0, in, out, n, write_total, ticket, tile_state); break;
    }
    CWA_CUDA(cudaGetLastError());
    return 0;
}

// One level of the reference's multi-dispatch Blelloch scan (prefix_sum_cs.glsl:18-43), kept so a
// host that still drives ParallelScan::Compute level by level through ComputeShader::Dispatch gets
// identical results.  The fused path never uses it.
__global__ void prefix_sum_level_kernel(int* x, int n, int phase, int stride, int nthreads)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nthreads) return;
    const long long ka = (long long)(stride / 2 - 1) + (long long)gid * stride;
    const long long kb = ka + stride / 2;
    if (kb >= n) return;
    if (phase == 0) {
        x[kb] = x[ka] + x[kb];
    } else {
        if (stride == n) x[n - 1] = 0;
        int t = x[ka];
        x[ka] = x[kb];
        x[kb] = t + x[kb];
    }
}

int prefix_sum_level_launch(cwa_ctx* ctx, int* x, int n, int phase, int stride, int nthreads)
{
    CWA_CHECK(n >= 2 && stride >= 2 && nthreads >= 1, "prefix_sum_cs: bad uniforms n=%d stride=%d", n, stride);
    { KScope k(ctx, KID_OTHER);
      prefix_sum_level_kernel<<<ceil_div(nthreads, 256), 256, 0, ctx->stream>>>(x, n, phase, stride, nthreads); }
    CWA_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// (3) insert + canonical ordering
// ---------------------------------------------------------------------------------------------
// arrival[offset + arrival rank] = gid: the reference's InsertPoint with the atomic already folded
// into the count pass.  Order inside a cell is the (scheduling-dependent) arrival order.
__global__ void __launch_bounds__(256)
grid_insert_kernel(const int* __restrict__ cell_of, const int* __restrict__ rank, const int* __restrict__ offset,
                   int n, const int* __restrict__ n_dev, int* __restrict__ arrival)
{
    CWA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev != nullptr) n = min(n, __ldg(n_dev));
    if (i >= n) return;
    const int c = __ldg(cell_of + i);
    if (c < 0) return;
    arrival[__ldg(offset + c) + __ldg(rank + i)] = i;             // mIndexList[offset+count] = gid
}

// Canonical order (ascending particle id inside a cell == stable counting sort == the CPU twin,
// SURVEY F7) by rank-by-counting: particle i lands at offset[c] + #{ids in its cell smaller than i}.
// One thread per particle, reads only (no in-place sorting, no write hazards); the cell's arrival
// list is a handful of consecutive ints that stay in L1 for the threads of the same cell.
__global__ void __launch_bounds__(256)
grid_cell_order_kernel(const int* __restrict__ cell_of, const int* __restrict__ offset, const int* __restrict__ arrival,
                       int n, const int* __restrict__ n_dev, int* __restrict__ index_list)
{
    CWA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev != nullptr) n = min(n, __ldg(n_dev));
    if (i >= n) return;
    const int c = __ldg(cell_of + i);
    if (c < 0) return;
    const int b = __ldg(offset + c), e = __ldg(offset + c + 1);
    int smaller = 0;
    for (int q = b; q < e; q++) smaller += (__ldg(arrival + q) < i) ? 1 : 0;
    index_list[b + smaller] = i;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Count-ahead insert: the previous frame's integrate pass already hashed and counted the new positions (slot-indexed cell id and
// arrival rank); after the scan the particle of old slot s goes to arrival[offset[cell] + rank].  Coalesced reads, and the
// public cell_of[] array (original particle order) is kept up to date for the reorder pass and cwa_grid_read.
__global__ void __launch_bounds__(256)
grid_insert_ahead_kernel(const int* __restrict__ cell_s, const int* __restrict__ rank_s, const int* __restrict__ old_index_list,
                         const int* __restrict__ offset, int n, int* __restrict__ cell_of, int* __restrict__ arrival, bool keep_vanished)
{
    CWA_PDL_ENTER();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int c = __ldg(cell_s + s);
    if (c == -2) return;                                           // slot beyond the previous build's inserted count
    const int id = __ldg(old_index_list + s);
    if (c >= 0 || !keep_vanished) cell_of[id] = c;                 // (slab frames: the slot of a particle that left may already hold an arrival)
    if (c >= 0) arrival[__ldg(offset + c) + __ldg(rank_s + s)] = id;
}

// arrivals of a slab exchange (hashed and counted by the unpack kernel): same insert, indexed by particle id
__global__ void __launch_bounds__(256)
grid_insert_arrivals_kernel(const int* __restrict__ ids, const int* __restrict__ count, int max_count, const int* __restrict__ cell_of,
                            const int* __restrict__ rank, const int* __restrict__ offset, int* __restrict__ arrival)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= min(__ldg(count), max_count)) return;
    const int id = __ldg(ids + u);
    const int c = __ldg(cell_of + id);
    if (c >= 0) arrival[__ldg(offset + c) + __ldg(rank + id)] = id;
}

int grid_build_internal(cwa_ctx* ctx, GridObj* g, const void* particles, int stride_bytes, int n, const GridBuildOpts& opts)
{
    CWA_CHECK(n >= 0 && n <= g->max_particles, "grid build: %d particles exceed the grid's capacity %d", n, g->max_particles);
    CWA_CHECK(stride_bytes >= 16 && stride_bytes % 16 == 0, "grid build: particle stride must be a multiple of 16 bytes");
    const bool ahead = opts.ahead_cell != nullptr && opts.ahead_rank != nullptr;
    g->n_built = n;
    const int C = g->view.num_cells;
    if (!ahead) {
        if (g->cleared_ahead) {                                                // the previous build of this grid cleared them behind its scan (side stream)
            CWA_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_pipe[1], 0));
        } else {
            KScope k(ctx, KID_CLEAR);                                          // ClearCounter + scan state, one memset
            CWA_CUDA(cudaMemsetAsync(g->counter, 0, g->clear_bytes, ctx->stream));
        }
        if (n > 0) {
            KScope k(ctx, KID_HASH_COUNT);
            if (g->dim == 2)
                cwa_launch(ctx, PDL_GRID2, grid_hash_count_kernel<2>, dim3(ceil_div(n, 256)), dim3(256), 0, (const char*)particles, stride_bytes, n, opts.n_dev, g->view, g->counter, g->cell_of, g->rank);
            else
                grid_hash_count_kernel<3><<<ceil_div(n, 256), 256, 0, ctx->stream>>>((const char*)particles, stride_bytes, n, opts.n_dev, g->view, g->counter, g->cell_of, g->rank);
            CWA_CUDA(cudaGetLastError());
        }
    }
    CWA_TRY(scan_exclusive_launch(ctx, g->counter, g->offset, C, g->ticket, g->tile_state));
    if (opts.clear_after_scan) {
        // the NEXT build is counted ahead by this frame's integrate pass: counter and scan state are free once the scan is done,
        // so they are cleared on a side stream while the insert / reorder passes run (ev_pipe[1] = cleared)
        CWA_CUDA(cudaEventRecord(ctx->ev_pipe[0], ctx->stream));
        CWA_CUDA(cudaStreamWaitEvent(ctx->side_stream[1], ctx->ev_pipe[0], 0));
        { StreamScope ss(ctx, ctx->side_stream[1]);
          KScope k(ctx, KID_CLEAR);
          CWA_CUDA(cudaMemsetAsync(g->counter, 0, g->clear_bytes, ctx->stream)); }
        CWA_CUDA(cudaEventRecord(ctx->ev_pipe[1], ctx->side_stream[1]));
    }
    g->cleared_ahead = opts.clear_after_scan && opts.clear_is_for_next_build;
    if (n > 0) {
        { KScope k(ctx, KID_INSERT);
          if (ahead)
          {
              cwa_launch(ctx, PDL_INSERT, grid_insert_ahead_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, opts.ahead_cell, opts.ahead_rank, g->index_list, g->offset, n, g->cell_of, g->arrival,
                                                                                  opts.arrivals != nullptr);
              if (opts.arrivals != nullptr && opts.arrivals_max > 0)
                  grid_insert_arrivals_kernel<<<ceil_div(opts.arrivals_max, 256), 256, 0, ctx->stream>>>(opts.arrivals, opts.arrivals_count, opts.arrivals_max,
                                                                                                          g->cell_of, g->rank, g->offset, g->arrival);
          }
          else
              cwa_launch(ctx, g->dim == 2 ? PDL_GRID2 : 0, grid_insert_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, g->cell_of, g->rank, g->offset, n, opts.n_dev, g->arrival); }
        CWA_CUDA(cudaGetLastError());
        if (opts.canonical_order) {
            KScope k(ctx, KID_CELL_ORDER);
            cwa_launch(ctx, g->dim == 2 ? PDL_GRID2 : 0, grid_cell_order_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, g->cell_of, g->offset, g->arrival, n, opts.n_dev, g->index_list);
        }
        CWA_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int cwa_grid_create(cwa_ctx* ctx, int dim, const float* mn, const float* mx, const int* num_cells,
                               int max_particles, cwa_grid* out)
{
    DeviceGuard _dg(ctx);
    CWA_CHECK(ctx && mn && mx && num_cells && out, "null argument");
    CWA_CHECK((dim & 15) == 2 || (dim & 15) == 3, "cwa_grid_create: dim must be 2 or 3");
    CWA_CHECK(max_particles > 0, "cwa_grid_create: max_particles must be positive");
    *out = -1;
    const bool compact = (dim >= 16);            // dim | CWA_GRID_COMPACT_INDEX: k-stride Nz instead of the reference's Nx
    dim &= 15;
    GridObj g;
    g.live = true; g.dim = dim; g.max_particles = max_particles;
    long long C = 1;
    for (int a = 0; a < 4; a++) { g.info.min[a] = g.info.max[a] = 0.0f; g.info.num_cells[a] = 1; g.info.cell_size[a] = 0.0f; }
    for (int a = 0; a < dim; a++) {
        CWA_CHECK(num_cells[a] >= 1 && mx[a] > mn[a], "cwa_grid_create: bad extent/cell count on axis %d", a);
        g.info.min[a] = mn[a]; g.info.max[a] = mx[a]; g.info.num_cells[a] = num_cells[a];
        // mCellSize = (mMax - mMin) / vec(mNumCells), FP32 on the host (UniformGridGpu2D.cpp:160)
        g.info.cell_size[a] = (mx[a] - mn[a]) / (float)num_cells[a];
    }
    int kstride = 1;
    if (dim == 3) {
        if (compact) {
            kstride = num_cells[2];
        } else {
            // The reference's 3-D index (i*Ny + j)*Nx + k aliases cells when Nz > Nx; refuse loudly.
            CWA_CHECK(num_cells[2] <= num_cells[0],
                      "cwa_grid_create: Nz (%d) > Nx (%d) aliases cells under the reference index formula (i*Ny+j)*Nx+k",
                      num_cells[2], num_cells[0]);
            kstride = num_cells[0];
        }
        C = (long long)num_cells[0] * num_cells[1] * kstride;        // covers every index the formula can produce
    } else {
        C = (long long)num_cells[0] * num_cells[1];
    }
    CWA_CHECK(C < (1ll << 30), "cwa_grid_create: too many cells (%lld)", C);
    for (int a = 0; a < 3; a++) {
        g.view.min[a] = g.info.min[a]; g.view.max[a] = g.info.max[a];
        g.view.cell[a] = (a < dim) ? g.info.cell_size[a] : 1.0f;
        g.view.n[a] = (a < dim) ? g.info.num_cells[a] : 1;
        g.view.inv_cell[a] = 1.0f / g.view.cell[a];
    }
    g.view.dim = dim; g.view.num_cells = (int)C; g.view.kstride = kstride;

    const size_t tiles = scan_num_tiles((int)C);
    // [counter C ints (padded to 16 B)][4 ints: scan ticket, -, queue counters of the heavy SPH kernels][tile_state tiles x 8 B]
    const size_t counter_bytes = (((size_t)C * 4 + 15) / 16) * 16;
    g.clear_bytes = counter_bytes + 16 + tiles * 8;
    char* blk = nullptr;
    CWA_CUDA(cudaMalloc(&blk, g.clear_bytes));
    g.counter = (int*)blk;
    g.ticket = (int*)(blk + counter_bytes);
    g.tile_state = (unsigned long long*)(blk + counter_bytes + 16);
    CWA_CUDA(cudaMalloc(&g.offset, ((size_t)C + 1) * 4));
    CWA_CUDA(cudaMalloc(&g.cell_of, (size_t)max_particles * 4));
    CWA_CUDA(cudaMalloc(&g.rank, (size_t)max_particles * 4));
    CWA_CUDA(cudaMalloc(&g.arrival, (size_t)max_particles * 4));
    CWA_CUDA(cudaMalloc(&g.index_list, (size_t)max_particles * 4));
    CWA_CUDA(cudaMemsetAsync(blk, 0, g.clear_bytes, ctx->stream));
    CWA_CUDA(cudaMemsetAsync(g.offset, 0, ((size_t)C + 1) * 4, ctx->stream));
    CWA_CUDA(cudaMemsetAsync(g.index_list, 0xff, (size_t)max_particles * 4, ctx->stream));
    g.buf_counter = new_buffer(ctx, g.counter, (size_t)C * 4, false);
    g.buf_offset = new_buffer(ctx, g.offset, ((size_t)C + 1) * 4, false);
    g.buf_index = new_buffer(ctx, g.index_list, (size_t)max_particles * 4, false);
    g.buf_cell_of = new_buffer(ctx, g.cell_of, (size_t)max_particles * 4, false);
    ctx->grids.push_back(g);
    *out = (int)ctx->grids.size() - 1;
    return 0;
}

extern "C" int cwa_grid_destroy(cwa_ctx* ctx, cwa_grid h)
{
    DeviceGuard _dg(ctx);
    GridObj* g = get_grid(ctx, h);
    CWA_CHECK(g, "invalid grid handle %d", h);
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(g->counter); cudaFree(g->offset); cudaFree(g->cell_of); cudaFree(g->rank); cudaFree(g->arrival); cudaFree(g->index_list);
    for (cwa_buf b : {g->buf_counter, g->buf_offset, g->buf_index, g->buf_cell_of})
        if (BufferObj* o = get_buffer(ctx, b)) o->live = false;
    g->live = false;
    return 0;
}

extern "C" int cwa_grid_get_info(cwa_ctx* ctx, cwa_grid h, cwa_grid_info* out, int* num_cells_total)
{
    DeviceGuard _dg(ctx);
    GridObj* g = get_grid(ctx, h);
    CWA_CHECK(g, "invalid grid handle %d", h);
    if (out) *out = g->info;
    if (num_cells_total) *num_cells_total = g->view.num_cells;
    return 0;
}

extern "C" int cwa_grid_build(cwa_ctx* ctx, cwa_grid h, cwa_buf particles, int stride_bytes, int n)
{
    DeviceGuard _dg(ctx);
    GridObj* g = get_grid(ctx, h);
    CWA_CHECK(g, "invalid grid handle %d", h);
    BufferObj* p = get_buffer(ctx, particles);
    CWA_CHECK(p, "invalid particle buffer handle %d", particles);
    CWA_CHECK((size_t)n * stride_bytes <= p->bytes, "cwa_grid_build: %d particles of %d bytes exceed the buffer", n, stride_bytes);
    return grid_build_internal(ctx, g, p->ptr, stride_bytes, n);
}

extern "C" int cwa_grid_read(cwa_ctx* ctx, cwa_grid h, int which, int* host, int count)
{
    DeviceGuard _dg(ctx);
    GridObj* g = get_grid(ctx, h);
    CWA_CHECK(g && host, "invalid grid handle %d", h);
    const int* src = nullptr; int avail = 0;
    switch (which) {
    case CWA_GRID_COUNTER: src = g->counter; avail = g->view.num_cells; break;
    case CWA_GRID_OFFSET: src = g->offset; avail = g->view.num_cells + 1; break;
    case CWA_GRID_INDEX_LIST: src = g->index_list; avail = g->max_particles; break;
    case CWA_GRID_CELL_OF: src = g->cell_of; avail = g->max_particles; break;
    default: CWA_CHECK(false, "cwa_grid_read: unknown array %d", which);
    }
    CWA_CHECK(count >= 0 && count <= avail, "cwa_grid_read: count %d exceeds %d", count, avail);
    CWA_CUDA(cudaMemcpyAsync(host, src, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CWA_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cwa_grid_buffer(cwa_ctx* ctx, cwa_grid h, int which, cwa_buf* out)
{
    DeviceGuard _dg(ctx);
    GridObj* g = get_grid(ctx, h);
    CWA_CHECK(g && out, "invalid grid handle %d", h);
    switch (which) {
    case CWA_GRID_COUNTER: *out = g->buf_counter; break;
    case CWA_GRID_OFFSET: *out = g->buf_offset; break;
    case CWA_GRID_INDEX_LIST: *out = g->buf_index; break;
    case CWA_GRID_CELL_OF: *out = g->buf_cell_of; break;
    default: CWA_CHECK(false, "cwa_grid_buffer: unknown array %d", which);
    }
    return 0;
}
