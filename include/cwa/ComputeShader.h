// cwa/ComputeShader.h -- mirror of ComputeShader (CoupledWaterAnimation/ComputeShader.h:7-39,
// ComputeShader.cpp:9-56).  Init() resolves the GLSL file NAME to the CUDA kernel set that replaces
// it; Dispatch() launches it on the currently bound buffers / object, like glDispatchCompute.
#pragma once

#include <string>

#include "Common.h"

class ComputeShader {
public:
    ComputeShader(const std::string& filename = "") : mFilename(filename) {}

    void Init()
    {
        cwa_shader s = -1;
        // an unknown shader leaves mShader == -1, exactly what InitShader() hands back on failure
        mShader = cwa::Ok(cwa_shader_create(cwa::Ctx(), mFilename.c_str(), &s), "ComputeShader::Init") ? (unsigned)s : unsigned(-1);
        SetMode(mMode);
        if (mObject >= 0) BindObject(mObject);
    }
    void UseProgram() {}                                   // kernels need no "current program"
    void Dispatch()
    {
        if (mShader == unsigned(-1)) return;
        cwa::Ok(cwa_shader_dispatch(cwa::Ctx(), (cwa_shader)mShader, mNumWorkgroups.x, mNumWorkgroups.y, mNumWorkgroups.z), "ComputeShader::Dispatch");
    }
    unsigned GetShader() { return mShader; }
    std::string GetName() { return mFilename; }
    cwa::ivec3 GetMaxWorkGroupSize() { return mMaxWorkGroupSize; }
    void SetMaxWorkGroupSize(cwa::ivec3 max_size) { mMaxWorkGroupSize = max_size; UpdateWorkgroups(); }
    void SetGridSize(cwa::ivec3 grid_size) { mGridSize = grid_size; UpdateWorkgroups(); }
    cwa::ivec3 GetGridSize() { return mGridSize; }
    int GetMode() { return mMode; }
    void SetMode(int mode)
    {
        mMode = mode;
        if (mShader != unsigned(-1)) cwa_shader_set_mode(cwa::Ctx(), (cwa_shader)mShader, mode);
    }
    int GetUniformLocation(const char*) { return -1; }      // uniforms are parameter blocks here
    void SetUniform(int location, int v) { if (mShader != unsigned(-1)) cwa_shader_set_uniform_i(cwa::Ctx(), (cwa_shader)mShader, location, v); }
    void SetUniform(int location, float v) { if (mShader != unsigned(-1)) cwa_shader_set_uniform_f(cwa::Ctx(), (cwa_shader)mShader, location, v); }
    // the simulation object (cwa_sph / cwa_wave handle) this program drives
    void BindObject(int handle) { mObject = handle; if (mShader != unsigned(-1)) cwa_shader_bind_object(cwa::Ctx(), (cwa_shader)mShader, handle); }

private:
    void UpdateWorkgroups()
    {
        auto cdiv = [](int a, int b) { return b > 0 ? (a + b - 1) / b : 0; };
        mNumWorkgroups = cwa::ivec3(cdiv(mGridSize.x, mMaxWorkGroupSize.x), cdiv(mGridSize.y, mMaxWorkGroupSize.y), cdiv(mGridSize.z, mMaxWorkGroupSize.z));
        if (mNumWorkgroups.x < 1) mNumWorkgroups.x = 1;
        if (mNumWorkgroups.y < 1) mNumWorkgroups.y = 1;
        if (mNumWorkgroups.z < 1) mNumWorkgroups.z = 1;
    }
    std::string mFilename;
    unsigned mShader = unsigned(-1);
    cwa::ivec3 mMaxWorkGroupSize = cwa::ivec3(1024, 1, 1);
    cwa::ivec3 mNumWorkgroups = cwa::ivec3(0);
    cwa::ivec3 mGridSize = cwa::ivec3(0);
    int mMode = 0;
    int mObject = -1;
};
