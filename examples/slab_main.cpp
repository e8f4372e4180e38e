// slab_main.cpp -- a C++ host driving N ranks of the slab-decomposed coupled frame through the C ABI (include/cwa_b200.h,
// csrc/slab.cu): what CoupledWaterAnimation/Main.cpp's idle() (:530-562) becomes on several GPUs.  All ranks live in THIS process
// (rank r on device r % device count; on one GPU they share it), are wired with direct mailbox pointers and stepped with
// cwa_slab_group_step; the same scene then runs through the single-GPU path (cwa_coupled_step) and the two states are compared:
// every particle exactly once, wave rows bit-equal, positions to rounding.  No Python, no torch, no collective library.
//   build: g++ -std=c++17 -Iinclude examples/slab_main.cpp -L<pkg> -lcwa_b200        usage: slab_main [ranks=3] [frames=8]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cwa_b200.h"

struct Particle { float pos[4], vel[4], force[4], extras[4]; };

#define CHECK(call)                                                                                  \
    do { if ((call) != 0) { std::fprintf(stderr, "%s failed: %s\n", #call, cwa_last_error()); std::exit(2); } } while (0)

// scene: a 64 x 5 x 160 sheet of the shipped lattice in a long tank, particles kicked along z so that they cross the slab faces
static const int NX = 64, NY = 5, NZ = 160, WAVE_W = 128, WAVE_H = 512;
static const float BOX_X = 0.6f, BOX_Z = 1.45f, UV = 0.68f, CELL = 0.0101f, GRID_Y = 0.30f, H_SUPPORT = 0.01f;   // cells of ~h: the row-mask kernels
static const int CELLS_Y = 31;
static const int CAP_MIG = 4096, CAP_GHOST = 16384;

static std::vector<Particle> make_scene()
{
    std::vector<Particle> p((size_t)NX * NY * NZ);
    const float sp = 0.0085f;
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); };   // [0, 1)
    size_t n = 0;
    for (int i = 0; i < NX; i++)
        for (int j = 0; j < NY; j++)
            for (int k = 0; k < NZ; k++, n++) {
                Particle& q = p[n];
                std::memset(&q, 0, sizeof(q));
                q.pos[0] = i * sp + (rnd() - 0.5f) * 0.2f * sp; q.pos[1] = j * sp + (rnd() - 0.5f) * 0.2f * sp;
                q.pos[2] = k * sp + (rnd() - 0.5f) * 0.2f * sp; q.pos[3] = 1.0f;
                q.vel[2] = (rnd() - 0.5f) * 60.0f;                           // 8 frames x 5e-5 x 30 = 0.012: crosses the faces
                q.extras[0] = 1000.0f; q.extras[2] = 500.0f; q.extras[3] = (float)n;   // id rides in the unused extras.w
            }
    return p;
}

static void set_params(cwa_ctx* ctx)
{
    CHECK(cwa_param_set(ctx, "upper.x", BOX_X)); CHECK(cwa_param_set(ctx, "upper.z", BOX_Z));
    CHECK(cwa_param_set(ctx, "uv_scale", UV));
}

struct Rank { cwa_ctx* ctx = nullptr; cwa_slab_desc d{}; cwa_buf buf = -1; cwa_grid grid = -1; cwa_sph sph = -1; cwa_wave wave = -1; cwa_slab slab = -1; };

int main(int argc, char** argv)
{
    const int world = argc > 1 ? std::atoi(argv[1]) : 3;
    const int frames = argc > 2 ? std::atoi(argv[2]) : 8;
    const std::vector<Particle> scene = make_scene();
    int ndev = 1;
    { cwa_ctx* probe = nullptr; CHECK(cwa_create(0, &probe)); cwa_destroy(probe); }
    if (const char* e = std::getenv("CWA_EXAMPLE_DEVICES")) ndev = std::max(1, std::atoi(e));

    // ---- the ranks
    std::vector<Rank> rk(world);
    for (int r = 0; r < world; r++) {
        Rank& k = rk[r];
        CHECK(cwa_create(r % ndev, &k.ctx));
        set_params(k.ctx);
        CHECK(cwa_slab_plan(world, r, WAVE_W, WAVE_H, 1, UV, H_SUPPORT, nullptr, &k.d));
        std::vector<Particle> mine;
        for (const Particle& q : scene) if (q.pos[2] >= k.d.z_lo && q.pos[2] < k.d.z_hi) mine.push_back(q);
        k.d.cap_mig = CAP_MIG; k.d.cap_ghost = CAP_GHOST;
        k.d.capacity = (int)(mine.size() * 3 / 2) + 2 * (2 * CAP_MIG + CAP_GHOST) + 1024;
        k.d.timeout_ms = 10000;
        const float zl = r > 0 ? std::max(0.0f, k.d.z_lo - 0.06f) : 0.0f, zh = r < world - 1 ? std::min(BOX_Z, k.d.z_hi + 0.06f) : BOX_Z;
        const float mn[3] = {0.0f, -0.02f, zl}, mx[3] = {BOX_X, GRID_Y, zh};
        const int cells[3] = {(int)(BOX_X / CELL), CELLS_Y, std::max(4, (int)std::floor((zh - zl) / CELL + 1e-6f))};
        CHECK(cwa_buffer_create(k.ctx, (size_t)k.d.capacity * sizeof(Particle), nullptr, &k.buf));
        CHECK(cwa_grid_create(k.ctx, 3 | CWA_GRID_COMPACT_INDEX, mn, mx, cells, k.d.capacity, &k.grid));
        CHECK(cwa_sph_create(k.ctx, k.buf, k.d.capacity, k.grid, &k.sph));
        CHECK(cwa_wave_create_block(k.ctx, WAVE_W, WAVE_H, k.d.store_lo, k.d.store_hi - k.d.store_lo, 1, CWA_WAVE_COUPLED, &k.wave));
        CHECK(cwa_slab_create(k.ctx, &k.d, k.sph, k.wave, &k.slab));
        if (!mine.empty()) CHECK(cwa_buffer_sub_data(k.ctx, k.buf, 0, mine.size() * sizeof(Particle), mine.data()));
        CHECK(cwa_slab_set_owned(k.ctx, k.slab, (int)mine.size()));
    }
    for (int r = 0; r < world; r++)                       // wiring, once: every rank learns the others' mailboxes (same process: plain pointers;
        for (int o = 0; o < world; o++) {                 // processes would exchange the 64-byte handles of cwa_slab_export instead)
            if (o == r) continue;
            void* ptr = nullptr; size_t bytes = 0;
            CHECK(cwa_slab_mailbox(rk[o].ctx, rk[o].slab, &ptr, &bytes));
            CHECK(cwa_slab_connect(rk[r].ctx, rk[r].slab, o, nullptr, ptr));
        }

    // ---- idle() x frames, in two calls (the second starts from a complete state: covers the explicit pack at a call's start)
    std::vector<cwa_ctx*> ctxs(world); std::vector<cwa_slab> slabs(world);
    for (int r = 0; r < world; r++) { ctxs[r] = rk[r].ctx; slabs[r] = rk[r].slab; }
    CHECK(cwa_slab_group_step(ctxs.data(), slabs.data(), world, frames / 2, CWA_COUPLING_AS_SHIPPED));
    CHECK(cwa_slab_group_step(ctxs.data(), slabs.data(), world, frames - frames / 2, CWA_COUPLING_AS_SHIPPED));

    // ---- gather: owned particles (dead slots = emigrated, skipped) and owned wave rows of every rank
    std::vector<Particle> got;
    std::vector<float> wave_got((size_t)WAVE_W * WAVE_H);
    long long migrated = 0;
    for (int r = 0; r < world; r++) {
        Rank& k = rk[r];
        int c[8];
        CHECK(cwa_slab_counts(k.ctx, k.slab, c));
        if (c[3] != 0) { std::fprintf(stderr, "rank %d: slab error bits %d\n", r, c[3]); return 3; }
        migrated += (long long)(unsigned)c[4];
        std::vector<Particle> own((size_t)c[0]);
        if (c[0] > 0) CHECK(cwa_buffer_read(k.ctx, k.buf, 0, own.size() * sizeof(Particle), own.data()));
        for (const Particle& q : own) if (!(q.pos[3] == -1.0f && std::isnan(q.pos[0]))) got.push_back(q);
        int img = 0;
        CHECK(cwa_wave_role_image(k.ctx, k.wave, 0, &img));
        std::vector<float> rows((size_t)WAVE_W * (k.d.store_hi - k.d.store_lo));
        CHECK(cwa_wave_read_image(k.ctx, k.wave, img, rows.data()));
        std::memcpy(&wave_got[(size_t)k.d.row_lo * WAVE_W], &rows[(size_t)(k.d.row_lo - k.d.store_lo) * WAVE_W], (size_t)(k.d.row_hi - k.d.row_lo) * WAVE_W * sizeof(float));
    }
    std::sort(got.begin(), got.end(), [](const Particle& a, const Particle& b) { return a.extras[3] < b.extras[3]; });

    // ---- the same scene on one GPU
    cwa_ctx* one = nullptr;
    CHECK(cwa_create(0, &one));
    set_params(one);
    cwa_buf buf; cwa_grid grid; cwa_sph sph; cwa_wave wave;
    const float mn[3] = {0.0f, -0.02f, 0.0f}, mx[3] = {BOX_X, GRID_Y, BOX_Z};
    const int cells[3] = {(int)(BOX_X / CELL), CELLS_Y, (int)(BOX_Z / CELL)};
    CHECK(cwa_buffer_create(one, scene.size() * sizeof(Particle), scene.data(), &buf));
    CHECK(cwa_grid_create(one, 3 | CWA_GRID_COMPACT_INDEX, mn, mx, cells, (int)scene.size(), &grid));
    CHECK(cwa_sph_create(one, buf, (int)scene.size(), grid, &sph));
    CHECK(cwa_wave_create(one, WAVE_W, WAVE_H, 1, CWA_WAVE_COUPLED, &wave));
    CHECK(cwa_coupled_step(one, sph, wave, frames, CWA_COUPLING_AS_SHIPPED));
    std::vector<Particle> ref(scene.size());
    CHECK(cwa_buffer_read(one, buf, 0, ref.size() * sizeof(Particle), ref.data()));
    std::vector<float> wave_ref((size_t)WAVE_W * WAVE_H);
    int img = 0;
    CHECK(cwa_wave_role_image(one, wave, 0, &img));
    CHECK(cwa_wave_read_image(one, wave, img, wave_ref.data()));

    // ---- compare
    bool ids_ok = got.size() == ref.size();
    double max_pos = 0.0, max_vel = 0.0, vel_rms = 0.0;
    int nan_diff = 0;
    if (ids_ok) {
        for (size_t i = 0; i < ref.size(); i++) {
            if (got[i].extras[3] != ref[i].extras[3]) { ids_ok = false; break; }
            const bool ng = std::isnan(got[i].pos[0]) || std::isnan(got[i].pos[1]) || std::isnan(got[i].pos[2]);
            const bool nr = std::isnan(ref[i].pos[0]) || std::isnan(ref[i].pos[1]) || std::isnan(ref[i].pos[2]);
            if (ng != nr) nan_diff++;
            if (ng || nr) continue;
            for (int a = 0; a < 3; a++) {
                max_pos = std::max(max_pos, (double)std::fabs(got[i].pos[a] - ref[i].pos[a]));
                max_vel = std::max(max_vel, (double)std::fabs(got[i].vel[a] - ref[i].vel[a]));
                vel_rms += (double)ref[i].vel[a] * ref[i].vel[a];
            }
        }
        vel_rms = std::sqrt(vel_rms / (3.0 * ref.size()));
    }
    const bool wave_ok = std::memcmp(wave_got.data(), wave_ref.data(), wave_ref.size() * sizeof(float)) == 0;
    std::printf("ranks=%d frames=%d particles=%zu ids_ok=%d nan_diff=%d migrated=%lld wave_bit_exact=%d max_pos_diff=%.3e max_vel_rel=%.3e\n",
                world, frames, got.size(), ids_ok ? 1 : 0, nan_diff, migrated, wave_ok ? 1 : 0, max_pos, vel_rms > 0 ? max_vel / vel_rms : 0.0);
    for (Rank& k : rk) cwa_destroy(k.ctx);
    cwa_destroy(one);
    return (ids_ok && wave_ok && nan_diff == 0) ? 0 : 1;
}
