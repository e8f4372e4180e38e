// cwa/StencilImage2DTripleBuffered.h -- mirror of StencilImage2DTripleBuffered
// (CoupledWaterAnimation/StencilImage2DTripleBuffered.h:8-41, .cpp:4-95): the triple-buffered wave
// height field as three device arrays rotated by role.
#pragma once

#include "ComputeShader.h"
#include "ImageTexture.h"
#include "Module.h"

struct StencilImage2DTripleBuffered : public Module {
    const int MODE_INIT = CWA_MODE_INIT;
    const int MODE_INIT_FROM_TEXTURE = CWA_MODE_INIT_FROM_TEXTURE;
    const int MODE_EVOLVE = CWA_MODE_EVOLVE;
    const int MODE_TEST = CWA_MODE_TEST;

    // extension: scalar (1) or RGBA32F-compatible (4) storage; only .x is ever live
    void SetChannels(int ch) { mChannels = ch; }
    void SetGridSize(cwa::ivec2 size) { mGridSize = size; }

    void Init() override
    {
        if (pShader == nullptr) return;
        const int variant = (pShader->GetName().find("Wave2D_cs") != std::string::npos) ? CWA_WAVE_SIMP : CWA_WAVE_COUPLED;
        if (mWave >= 0) cwa_wave_destroy(cwa::Ctx(), mWave);
        cwa_wave w = -1;
        if (!cwa::Ok(cwa_wave_create(cwa::Ctx(), mGridSize.x, mGridSize.y, mChannels, variant, &w), "StencilImage2DTripleBuffered::Init")) return;
        mWave = w;                                       // cwa_wave_create already ran Reinit(), like Init() :30
        for (int i = 0; i < 3; i++) mImage[i].Attach(mWave, i);
        pShader->SetGridSize(cwa::ivec3(mGridSize, 1));
        pShader->SetMaxWorkGroupSize(cwa::ivec3(32, 32, 1));
        pShader->Init();
        pShader->BindObject(mWave);
    }
    void Reinit() override { if (pShader && mWave >= 0) cwa::Ok(cwa_wave_reinit(cwa::Ctx(), mWave), "Reinit"); }
    void ReinitFromTexture(const float* rgba, int tw, int th) { if (pShader && mWave >= 0) cwa::Ok(cwa_wave_reinit_from_texture(cwa::Ctx(), mWave, rgba, tw, th), "ReinitFromTexture"); }
    void Compute() override
    {
        if (!mEvolve || pShader == nullptr || mWave < 0) return;     // silent no-op like the reference :81-82
        cwa::Ok(cwa_wave_compute(cwa::Ctx(), mWave, 1), "StencilImage2DTripleBuffered::Compute");
    }
    void PingPong() { if (mWave >= 0) cwa_wave_pingpong(cwa::Ctx(), mWave); }
    void SetEvolve(bool e) { mEvolve = e; if (mWave >= 0) cwa_wave_set_evolve(cwa::Ctx(), mWave, e ? 1 : 0); }

    ImageTexture& GetWriteImage() { int r[2], w = 2; cwa_wave_state(cwa::Ctx(), mWave, r, &w, nullptr, nullptr); return mImage[w]; }
    ImageTexture& GetReadImage(int i) { int r[2] = {0, 1}, w; cwa_wave_state(cwa::Ctx(), mWave, r, &w, nullptr, nullptr); return mImage[r[i]]; }
    void SetShader(ComputeShader& cs) { pShader = &cs; pShader->SetGridSize(cwa::ivec3(mGridSize, 1)); }
    ComputeShader* GetpShader() { return pShader; }
    cwa::ivec2 GetGridSize() { return mGridSize; }
    cwa_wave Handle() const { return mWave; }

private:
    ComputeShader* pShader = nullptr;                    // non-owning, like the reference (.h:34)
    bool mEvolve = true;
    int mChannels = 4;                                   // GL_RGBA32F as shipped (.cpp:21)
    cwa::ivec2 mGridSize = cwa::ivec2(64, 64);           // .h:39
    cwa_wave mWave = -1;
    ImageTexture mImage[3];
};
