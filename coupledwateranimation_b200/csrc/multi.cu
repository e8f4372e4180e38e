// multi.cu -- device primitives of the multi-GPU decomposition: selecting ghost
// particles near a slab face and splitting off particles that migrated out of the slab.
// Both are a STABLE copy_if on one coordinate of the particle position (flags -> the same
// single-pass look-back scan the grid uses -> 64-byte record copy), so results do not depend on
// scheduling and the packed send buffers are contiguous (one NCCL send per neighbour).
#include "internal.cuh"

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

struct MultiScratch {
    int* flags = nullptr;
    int* pos = nullptr;           // n + 1 entries
    int* ticket = nullptr;
    unsigned long long* state = nullptr;
    size_t cap = 0, tiles = 0;
};

static MultiScratch* multi_scratch(cwa_ctx* ctx, int n)
{
    static MultiScratch table[64];                 // one per device
    MultiScratch* s = &table[ctx->device & 63];
    if ((size_t)n <= s->cap) return s;
    if (s->flags) { cudaFree(s->flags); cudaFree(s->pos); cudaFree(s->ticket); }
    s->cap = (size_t)n + (size_t)n / 4 + 1024;
    s->tiles = scan_num_tiles((int)s->cap);
    if (cudaMalloc(&s->flags, s->cap * 4) != cudaSuccess) return nullptr;
    if (cudaMalloc(&s->pos, (s->cap + 1) * 4) != cudaSuccess) return nullptr;
    if (cudaMalloc(&s->ticket, 16 + s->tiles * 8) != cudaSuccess) return nullptr;
    s->state = (unsigned long long*)((char*)s->ticket + 16);
    return s;
}

extern "C" int cwa_particles_copy_if(cwa_ctx* ctx, cwa_buf src, int n, int axis, int kind, float a, float b,
                                     cwa_buf dst, int dst_offset, int* count_out)
{
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
    return 0;
}
