// cwa/StencilImage2D.h -- mirror of ImageStencil (SphWave2D/StencilImage2D.h:10-66, .cpp:4-164) for the 1-D images of the 2-D app:
// Shallow1D_cs.glsl (SetNumBuffers(2), MODE_ITERATE_FIRST/LAST = 2/3) and Wave1D_cs.glsl (SetNumBuffers(3), SetSubsteps(10)).
// Host code written like InitShallowWaterEquation() / InitWaveEquation() (SphWave2D/Main.cpp:63-97) compiles against it unchanged,
// minus the GL filter / wrap calls.
#pragma once

#include <string>

#include "Buffer.h"
#include "ComputeShader.h"
#include "Module.h"

struct ImageStencil : public Module {
    int mMODE_INIT_FIRST = 0;
    int mMODE_INIT_LAST = 1;
    int MODE_ITERATE_FIRST = 2;
    int MODE_ITERATE_LAST = 2;
    int mMODE_INIT_FROM_TEXTURE = -1;

    void SetShader(ComputeShader& cs) { pShader = &cs; }
    ComputeShader* GetpShader() { return pShader; }
    void SetGridSize(cwa::ivec3 grid_size) { mGridSize = cwa::ivec3(grid_size.x < 1 ? 1 : grid_size.x, grid_size.y < 1 ? 1 : grid_size.y, grid_size.z < 1 ? 1 : grid_size.z); }
    cwa::ivec3 GetGridSize() { return mGridSize; }
    void SetSubsteps(int s) { mSubsteps = s; if (mStencil >= 0) cwa_stencil1d_set_substeps(cwa::Ctx(), mStencil, s); }
    void SetNumBuffers(int n) { mNumImages = n < 1 ? 1 : n; mNumReadImages = mNumImages == 1 ? 1 : mNumImages - 1; }
    int GetNumReadImages() { return mNumReadImages; }

    void Init() override
    {
        if (pShader == nullptr) return;
        const bool shallow = pShader->GetName().find("Shallow1D") != std::string::npos;
        const bool wave = pShader->GetName().find("Wave1D") != std::string::npos;
        if (!(shallow || wave) || mGridSize.y != 1 || mGridSize.z != 1 || mNumImages != (shallow ? 2 : 3)) {
            cwa::Ok(-1, "ImageStencil::Init: only the 1-D Shallow1D_cs (2 buffers) / Wave1D_cs (3 buffers) configurations of the reference are built");
            return;
        }
        if (mStencil >= 0) cwa_stencil1d_destroy(cwa::Ctx(), mStencil);
        cwa_stencil1d s = -1;
        if (!cwa::Ok(cwa_stencil1d_create(cwa::Ctx(), shallow ? CWA_STENCIL1D_SHALLOW : CWA_STENCIL1D_WAVE, mGridSize.x, &s), "ImageStencil::Init")) return;
        mStencil = s;                                    // cwa_stencil1d_create already ran Reinit(), like Init() :35
        cwa_stencil1d_set_substeps(cwa::Ctx(), mStencil, mSubsteps);
    }
    void Reinit() override { if (pShader && mStencil >= 0) cwa::Ok(cwa_stencil1d_reinit(cwa::Ctx(), mStencil), "ImageStencil::Reinit"); }
    void ReinitFromTexture(const float* rgba, int width) { if (pShader && mStencil >= 0) cwa::Ok(cwa_stencil1d_reinit_from_texture(cwa::Ctx(), mStencil, rgba, width), "ImageStencil::ReinitFromTexture"); }
    void ComputeFunc(int mode) { if (mStencil >= 0) cwa::Ok(cwa_stencil1d_compute_func(cwa::Ctx(), mStencil, mode), "ImageStencil::ComputeFunc"); }
    void Compute() override
    {
        if (!mIterate || pShader == nullptr || mStencil < 0) return;      // silent no-ops like the reference :144-145
        cwa::Ok(cwa_stencil1d_compute(cwa::Ctx(), mStencil, 1), "ImageStencil::Compute");
    }
    void SetIterate(bool it) { mIterate = it; if (mStencil >= 0) cwa_stencil1d_set_iterate(cwa::Ctx(), mStencil, it ? 1 : 0); }
    // uniforms at locations 2..5 of the two shaders + the BC constant
    void SetUniforms(float lambda, float dx_or_atten, float beta, float boundary0 = 0.0f, float boundary1 = 0.0f, int bc = CWA_BC_FREE)
    {
        if (mStencil >= 0) cwa::Ok(cwa_stencil1d_set_params(cwa::Ctx(), mStencil, lambda, dx_or_atten, beta, boundary0, boundary1, bc), "ImageStencil::SetUniforms");
    }
    // GetReadImage(i) as the Buffer handle that SphUgrid's sampler takes (wave1d.GetReadImage(0).BindTextureUnit(), Main.cpp:240)
    cwa_buf GetReadImageBuffer(int i)
    {
        int n = 0, r[2] = {0, 0}, w = 0;
        cwa_buf b = -1;
        if (mStencil >= 0 && cwa_stencil1d_state(cwa::Ctx(), mStencil, &n, r, &w, nullptr) == 0) cwa_stencil1d_image_buffer(cwa::Ctx(), mStencil, r[i], &b);
        return b;
    }
    cwa_stencil1d Handle() const { return mStencil; }

private:
    ComputeShader* pShader = nullptr;                    // non-owning, like the reference
    bool mIterate = true;
    int mNumImages = 1, mNumReadImages = 1, mSubsteps = 1;
    cwa::ivec3 mGridSize = cwa::ivec3(1, 1, 1);
    cwa_stencil1d mStencil = -1;
};
