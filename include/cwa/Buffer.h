// cwa/Buffer.h -- mirror of Buffer (SphWave2D/Buffer.h:5-24, Buffer.cpp:5-83): an SSBO/UBO with a
// binding index, backed by device memory.  mBuffer holds the cwa_buf handle instead of a GL name.
#pragma once

#include <algorithm>
#include <vector>

#include "Common.h"

class Buffer {
public:
    unsigned mBuffer = unsigned(-1);      // cwa_buf handle (GLuint(-1) == not created)
    unsigned mBinding = unsigned(-1);
    unsigned mTarget = unsigned(-1);
    unsigned mFlags = unsigned(-1);
    unsigned mSize = 0;
    bool mEnableDebug = false;

    Buffer(unsigned target = cwa::SHADER_STORAGE_BUFFER, unsigned binding = unsigned(-1)) : mBinding(binding), mTarget(target) {}

    void Init(int size, void* data = nullptr, unsigned flags = 0)
    {
        mSize = (unsigned)size; mFlags = flags;
        if (mBuffer != unsigned(-1)) cwa_buffer_destroy(cwa::Ctx(), (cwa_buf)mBuffer);   // glDeleteBuffers on re-Init
        cwa_buf b = -1;
        if (cwa::Ok(cwa_buffer_create(cwa::Ctx(), (size_t)size, data, &b), "Buffer::Init")) mBuffer = (unsigned)b;
        if (data) cwa_synchronize(cwa::Ctx());      // glNamedBufferStorage copies before returning
    }
    // adopt memory owned elsewhere (a mapped CUDA-GL interop pointer)
    void InitFromDevicePointer(void* device_ptr, size_t bytes)
    {
        cwa_buf b = -1;
        if (cwa::Ok(cwa_buffer_wrap(cwa::Ctx(), device_ptr, bytes, &b), "Buffer::InitFromDevicePointer")) { mBuffer = (unsigned)b; mSize = (unsigned)bytes; }
    }
    void BufferSubData(int offset, int size, void* data)
    {
        cwa::Ok(cwa_buffer_sub_data(cwa::Ctx(), (cwa_buf)mBuffer, (size_t)offset, (size_t)size, data), "Buffer::BufferSubData");
        cwa_synchronize(cwa::Ctx());                // caller may reuse `data` immediately (GL semantics)
    }
    void BindBufferBase() { cwa::Ok(cwa_buffer_bind_base(cwa::Ctx(), (int)mTarget, (int)mBinding, (cwa_buf)mBuffer), "Buffer::BindBufferBase"); }

    // DebugRead*: blocking read-back into a throw-away vector when mEnableDebug is set; the
    // vector is returned here so tests can look at it.
    std::vector<float> DebugReadFloat()
    {
        std::vector<float> v;
        if (mEnableDebug) { v.resize(mSize / sizeof(float)); cwa::Ok(cwa_buffer_read(cwa::Ctx(), (cwa_buf)mBuffer, 0, v.size() * sizeof(float), v.data()), "Buffer::DebugReadFloat"); }
        return v;
    }
    std::vector<int> DebugReadInt()
    {
        std::vector<int> v;
        if (mEnableDebug) { v.resize(mSize / sizeof(int)); cwa::Ok(cwa_buffer_read(cwa::Ctx(), (cwa_buf)mBuffer, 0, v.size() * sizeof(int), v.data()), "Buffer::DebugReadInt"); }
        return v;
    }
    friend void SwapBindings(Buffer& b0, Buffer& b1) { std::swap(b0.mBinding, b1.mBinding); }
};
