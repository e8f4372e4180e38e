// cwa/ImageTexture.h -- mirror of ImageTexture (CoupledWaterAnimation/ImageTexture.h:6-45) as far as
// the simulation path uses it: one level of the wave field with its image/texture unit.  The
// storage belongs to the StencilImage2DTripleBuffered object (three device arrays).
#pragma once

#include "Common.h"

class ImageTexture {
public:
    ImageTexture() {}
    void Attach(cwa_wave wave, int image) { mWave = wave; mImage = image; }
    int  GetUnit() const
    {
        int unit[3] = {0, 1, 2};
        if (mWave >= 0) cwa_wave_state(cwa::Ctx(), mWave, nullptr, nullptr, unit, nullptr);
        return (mImage >= 0) ? unit[mImage] : -1;
    }
    // glBindImageTexture: the CUDA kernels address the three levels by role; nothing to bind
    void BindImageTexture(unsigned /*access*/) {}
    // glBindTextureUnit(mUnit, mTexture): only a bind on unit 0 is visible to the SPH samplers
    void BindTextureUnit() { if (mWave >= 0) cwa_wave_bind_texture_unit(cwa::Ctx(), mWave); }
    // the level as a Buffer handle (CUDA-GL interop hand-off)
    unsigned GetTexture() const
    {
        cwa_buf b = -1;
        if (mWave >= 0) cwa_wave_image_buffer(cwa::Ctx(), mWave, mImage, &b);
        return (unsigned)b;
    }
    cwa::ivec3 GetSize() const { int w = 0, h = 0; if (mWave >= 0) cwa_wave_size(cwa::Ctx(), mWave, &w, &h, nullptr); return cwa::ivec3(w, h, 1); }
    int  Index() const { return mImage; }

private:
    cwa_wave mWave = -1;
    int mImage = -1;
};
