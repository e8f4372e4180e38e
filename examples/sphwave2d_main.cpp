// sphwave2d_main.cpp -- the simulation half of SphWave2D/Main.cpp written against the mirrored classes of include/cwa/ (no GL):
// the globals of Main.cpp:42-61, InitShallowWaterEquation() (:77-97), the compute part of initOpenGL() (:268-284) and idle()
// (:213-241): Module::sComputeAll() = SphUgrid::Compute (2 substeps of grid build + density + forces on the 2-D Koschier SPH)
// followed by ImageStencil::Compute (both Lax-Wendroff phases of the 1-D shallow-water wave), then the wave's newest image is
// bound as the SPH sampler.  Prints a short state summary so a test can compare it with the oracle.
//   build: g++ -std=c++17 -Iinclude examples/sphwave2d_main.cpp -L<pkg> -lcwa_b200
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cwa/ComputeShader.h"
#include "cwa/StencilBuffer.h"
#include "cwa/StencilImage2D.h"

struct Particle { cwa::vec4 pos, vel, acc; };          // SphWave2D/Main.cpp:35-40

int num_particles = 32 * 128;                          // :42

// file-scope objects, constructed before main() like the reference's (module order = construction order: sph2d, then wave1d)
ComputeShader SphCS("SphWaveKoschier2D_grid_cs.glsl");
SphUgrid sph2d;
ComputeShader WaveCS("Wave1D_cs.glsl");
ComputeShader ShallowCS("Shallow1D_cs.glsl");
ImageStencil wave1d;

static void InitShallowWaterEquation()
{
    wave1d.SetShader(ShallowCS);
    wave1d.SetNumBuffers(2);                           // double buffered
    wave1d.MODE_ITERATE_FIRST = 2;                     // two phase update for Lax Wendroff
    wave1d.MODE_ITERATE_LAST = 3;
    wave1d.SetGridSize(cwa::ivec3(128, 1, 1));         // (the GL_LINEAR / GL_CLAMP_TO_EDGE calls have no counterpart: the sampler is fixed)
}

static void initSimulation()
{
    sph2d.SetSubsteps(2);
    sph2d.SetShader(SphCS);
    sph2d.mElementSize = sizeof(Particle);
    sph2d.mNumElements = num_particles;
    InitShallowWaterEquation();
    Module::sInitAll();
}

static void idle()
{
    Module::sComputeAll();                             // :237
    // SphCS.UseProgram(); wave1d.GetReadImage(0).BindTextureUnit();   :239-240
    cwa_sph2_bind_wave1d(cwa::Ctx(), sph2d.Handle(), wave1d.GetReadImageBuffer(0), 128);
}

int main(int argc, char** argv)
{
    const int frames = argc > 1 ? std::atoi(argv[1]) : 3;
    initSimulation();
    for (int f = 0; f < frames; f++) idle();
    std::vector<Particle> p(num_particles);
    if (cwa_sph2_read(cwa::Ctx(), sph2d.Handle(), reinterpret_cast<cwa_particle2d*>(p.data())) != 0) return 2;
    std::vector<float> w(128 * 4);
    int n = 0, r[2] = {0, 0}, wi = 0;
    cwa_stencil1d_state(cwa::Ctx(), wave1d.Handle(), &n, r, &wi, nullptr);
    if (cwa_stencil1d_read_image(cwa::Ctx(), wave1d.Handle(), r[0], w.data()) != 0) return 2;
    double sx = 0, sy = 0, srho = 0, sh = 0; int enabled = 0;
    for (int i = 0; i < num_particles; i++) { sx += p[i].pos.x; sy += p[i].pos.y; srho += p[i].acc.w; enabled += p[i].pos.w > 0.5f ? 1 : 0; }
    for (int i = 0; i < 128; i++) sh += w[4 * i];
    std::printf("frames=%d particles=%d enabled=%d mean_x=%.6f mean_y=%.6f mean_rho=%.3f wave_h_sum=%.5f\n", frames, num_particles, enabled,
                sx / num_particles, sy / num_particles, srho / num_particles, sh);
    cwa::DestroyContext();
    return 0;
}
