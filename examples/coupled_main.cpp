// coupled_main.cpp -- the simulation half of CoupledWaterAnimation/Main.cpp written against the
// mirrored classes of include/cwa/ (no GL): globals as in Main.cpp:74-75, initOpenGL() (:860-985)
// without the render state, idle() (:530-562) and the display() uploads/binds that the next step
// depends on (:366-382, :413).  Prints a short state summary so a test can compare it with the
// oracle.   build: g++ -std=c++17 -Iinclude examples/coupled_main.cpp -L<pkg> -lcwa_b200
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cwa/Buffer.h"
#include "cwa/ComputeShader.h"
#include "cwa/ParallelScan.h"
#include "cwa/StencilImage2DTripleBuffered.h"

#define NUM_PARTICLES 20480
#define PARTICLE_RADIUS 0.005f
#define WORK_GROUP_SIZE 1024
#define PART_WORK_GROUPS 20

struct Particle { cwa::vec4 pos, vel, force, extras; };

struct ConstantsUniform { float mass = 0.02f, smoothing_coeff = 2.0f, visc = 3000.0f, resting_rho = 1000.0f; } ConstantsData;
struct BoundaryUniform { cwa::vec4 upper = cwa::vec4(0.48f, 1.0f, 0.48f, 500.0f), lower = cwa::vec4(0.0f, -0.02f, 0.0f, 50.0f); } BoundaryData;
struct WaveUniforms { cwa::vec4 attributes = cwa::vec4(0.01f, 0.985f, 0.001f, 1.0f), mesh_ws_pos = cwa::vec4(2.0f, 0.35f, -1.0f, 0.0f); } WaveData;

// file-scope objects, constructed before main() like the reference's: constructors touch no device
ComputeShader waveCS("wave_comp.glsl");
StencilImage2DTripleBuffered wave2d;
ComputeShader compute_programs[3] = {ComputeShader("rho_pres_comp.glsl"), ComputeShader("force_comp.glsl"), ComputeShader("integrate_comp.glsl")};
Buffer particles_ssbo(cwa::SHADER_STORAGE_BUFFER, 0);
Buffer constants_ubo(cwa::UNIFORM_BUFFER, 1), boundary_ubo(cwa::UNIFORM_BUFFER, 2), wave_ubo(cwa::UNIFORM_BUFFER, 3);
cwa_sph sph = -1;
bool simulate = true;

static std::vector<cwa::vec4> make_cube()
{
    std::vector<cwa::vec4> positions;
    const float spacing = ConstantsData.smoothing_coeff * 0.85f * PARTICLE_RADIUS;
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 5; j++)
            for (int k = 0; k < 64; k++) positions.push_back(cwa::vec4(i * spacing, j * spacing, k * spacing, 1.0f));
    return positions;
}

static void init_particles()
{
    std::vector<Particle> particles(NUM_PARTICLES);
    std::vector<cwa::vec4> grid_positions = make_cube();
    for (int i = 0; i < NUM_PARTICLES; i++) {
        particles[i].pos = grid_positions[i];
        particles[i].vel = cwa::vec4(0.0f);
        particles[i].force = cwa::vec4(0.0f);
        particles[i].extras = cwa::vec4(ConstantsData.resting_rho, 0.0f, 500.0f, 50.0f);
    }
    particles_ssbo.Init((int)(sizeof(Particle) * NUM_PARTICLES), particles.data());
    particles_ssbo.BindBufferBase();
}

static void display_uploads_and_binds()
{
    // display(): the UBO uploads of Main.cpp:369-380 and the wave texture bind of :413
    constants_ubo.BufferSubData(0, sizeof(ConstantsUniform), &ConstantsData);
    boundary_ubo.BufferSubData(0, sizeof(BoundaryUniform), &BoundaryData);
    wave_ubo.BufferSubData(0, sizeof(WaveUniforms), &WaveData);
    wave2d.GetReadImage(0).BindTextureUnit();
}

static void initSimulation()
{
    init_particles();
    constants_ubo.Init(sizeof(ConstantsUniform), &ConstantsData); constants_ubo.BindBufferBase();
    boundary_ubo.Init(sizeof(BoundaryUniform), &BoundaryData); boundary_ubo.BindBufferBase();
    wave_ubo.Init(sizeof(WaveUniforms), &WaveData); wave_ubo.BindBufferBase();
    waveCS.SetMaxWorkGroupSize(cwa::ivec3(32, 32, 1));
    wave2d.SetShader(waveCS);
    Module::sInitAll();                                   // StencilImage2DTripleBuffered::Init -> Reinit
    // the three SPH programs act on the SSBO bound at binding 0 (all-pairs, as shipped: no grid)
    if (!cwa::Ok(cwa_sph_create(cwa::Ctx(), (cwa_buf)particles_ssbo.mBuffer, NUM_PARTICLES, -1, &sph), "cwa_sph_create")) std::exit(2);
    for (ComputeShader& cs : compute_programs) {
        cs.SetGridSize(cwa::ivec3(NUM_PARTICLES, 1, 1));
        cs.Init();
        cs.BindObject(sph);
    }
}

static void idle()
{
    if (!simulate) return;
    // sampler unit 0 of the SPH programs = whatever display() last left on texture unit 0
    int tex0 = -1;
    cwa_wave_state(cwa::Ctx(), wave2d.Handle(), nullptr, nullptr, nullptr, &tex0);
    cwa_sph_bind_wave(cwa::Ctx(), sph, wave2d.Handle(), tex0);
    for (ComputeShader& cs : compute_programs) { cs.UseProgram(); cs.Dispatch(); }   // Main.cpp:549-557
    Module::sComputeAll();                                                          // :560
}

int main(int argc, char** argv)
{
    const int frames = argc > 1 ? std::atoi(argv[1]) : 3;
    if (!ParallelScanTest()) return 3;                    // the reference's own KAT, run on the device
    initSimulation();
    display_uploads_and_binds();                          // simulate starts false: one display() precedes the first step
    for (int f = 0; f < frames; f++) { idle(); display_uploads_and_binds(); }
    particles_ssbo.mEnableDebug = true;
    std::vector<float> raw = particles_ssbo.DebugReadFloat();
    double sum_rho = 0.0, sum_y = 0.0; int nan = 0;
    for (int i = 0; i < NUM_PARTICLES; i++) {
        const float y = raw[16 * i + 1], rho = raw[16 * i + 12];
        if (std::isnan(y)) { nan++; continue; }
        sum_rho += rho; sum_y += y;
    }
    std::printf("frames=%d particles=%d nan=%d mean_rho=%.3f mean_y=%.7f\n", frames, NUM_PARTICLES, nan, sum_rho / (NUM_PARTICLES - nan), sum_y / (NUM_PARTICLES - nan));
    cwa::DestroyContext();
    return 0;
}
