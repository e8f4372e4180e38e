// cwa/StencilBuffer.h -- mirror of SphUgrid (SphWave2D/StencilBuffer.h:64-79, StencilBuffer.cpp:138-179):
// the double-buffered 2-D particle SSBO + uniform grid + the Koschier SPH passes.
#pragma once

#include "ComputeShader.h"
#include "Module.h"
#include "UniformGrid.h"

class SphUgrid : public Module {
public:
    int mNumElements = 0;
    int mElementSize = 48;
    UniformGridSph2D mGrid;

    // ctor of the reference: 32x32 cells over [0, 9.6]^2 (StencilBuffer.cpp:138)
    SphUgrid(cwa::ivec2 cells = cwa::ivec2(32, 32), aabb2D ext = aabb2D(cwa::vec2(0.0f), cwa::vec2(2.0f * 4.8f))) : mGrid(cells, ext) {}

    void SetShader(ComputeShader& cs) { pShader = &cs; }
    void SetSubsteps(int s) { mSubsteps = s; if (mSph >= 0) cwa_sph2_set_substeps(cwa::Ctx(), mSph, s); }
    void Init() override
    {
        if (pShader == nullptr) return;
        const int variant = (pShader->GetName().find("SphWave") != std::string::npos) ? CWA_SPH2_WAVE : CWA_SPH2_KOSCHIER;
        mGrid.Init(unsigned(-1), 0, mNumElements, mElementSize);
        if (!cwa::Ok(cwa_sph2_create(cwa::Ctx(), mNumElements, variant, mGrid.Handle(), &mSph), "SphUgrid::Init")) return;   // runs MODE_INIT + PingPong
        cwa_sph2_set_substeps(cwa::Ctx(), mSph, mSubsteps);
    }
    void Reinit() override { if (mSph >= 0) cwa::Ok(cwa_sph2_reinit(cwa::Ctx(), mSph), "SphUgrid::Reinit"); }
    void Compute() override
    {
        if (!mEvolve || pShader == nullptr || mSph < 0) return;
        cwa::Ok(cwa_sph2_compute(cwa::Ctx(), mSph, 1), "SphUgrid::Compute");
    }
    unsigned GetReadBuffer() { cwa_buf b = -1; if (mSph >= 0) cwa_sph2_read_buffer(cwa::Ctx(), mSph, &b); return (unsigned)b; }
    cwa_sph2 Handle() const { return mSph; }

protected:
    ComputeShader* pShader = nullptr;
    bool mEvolve = true;
    int mSubsteps = 1;
    cwa_sph2 mSph = -1;
};
