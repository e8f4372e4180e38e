// cwa/Common.h -- the process-wide context the mirrored classes talk to (stands in for the GL
// context the reference creates in main(), CoupledWaterAnimation/Main.cpp:990-1019) and the few
// vector types the reference takes from glm.  Define CWA_USE_GLM before including to use glm's.
#pragma once

#include <cstdio>
#include <cstdlib>

#include "../cwa_b200.h"

#ifdef CWA_USE_GLM
#include <glm/glm.hpp>
namespace cwa { using glm::ivec2; using glm::ivec3; using glm::ivec4; using glm::vec2; using glm::vec3; using glm::vec4; }
#else
namespace cwa {
struct ivec2 { int x = 0, y = 0; ivec2() {} ivec2(int a, int b) : x(a), y(b) {} explicit ivec2(int a) : x(a), y(a) {} };
struct ivec3 { int x = 0, y = 0, z = 0; ivec3() {} ivec3(int a, int b, int c) : x(a), y(b), z(c) {} explicit ivec3(int a) : x(a), y(a), z(a) {} ivec3(ivec2 v, int c) : x(v.x), y(v.y), z(c) {} };
struct ivec4 { int x = 0, y = 0, z = 0, w = 0; ivec4() {} ivec4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {} };
struct vec2 { float x = 0, y = 0; vec2() {} vec2(float a, float b) : x(a), y(b) {} explicit vec2(float a) : x(a), y(a) {} };
struct vec3 { float x = 0, y = 0, z = 0; vec3() {} vec3(float a, float b, float c) : x(a), y(b), z(c) {} explicit vec3(float a) : x(a), y(a), z(a) {} };
struct vec4 { float x = 0, y = 0, z = 0, w = 0; vec4() {} vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} explicit vec4(float a) : x(a), y(a), z(a), w(a) {} };
}  // namespace cwa
#endif

namespace cwa {

// GL enums the mirrored signatures keep (values are irrelevant to the CUDA back end)
enum : unsigned { SHADER_STORAGE_BUFFER = CWA_TARGET_SSBO, UNIFORM_BUFFER = CWA_TARGET_UBO, READ_ONLY = 0x88B8, WRITE_ONLY = 0x88B9, READ_WRITE = 0x88BA };

inline cwa_ctx*& ContextSlot() { static cwa_ctx* c = nullptr; return c; }

// Create the context once (like glfwMakeContextCurrent + glewInit).  No device -> hard failure: the
// simulation step has no CPU fallback.
inline cwa_ctx* Ctx(int device = 0)
{
    cwa_ctx*& c = ContextSlot();
    if (!c && cwa_create(device, &c) != 0) {
        std::fprintf(stderr, "cwa: cannot create context: %s\n", cwa_last_error());
        std::abort();
    }
    return c;
}
inline void DestroyContext() { cwa_ctx*& c = ContextSlot(); if (c) { cwa_destroy(c); c = nullptr; } }

// The reference has no error returns; failures surface on stderr like its GL debug callback
// (DebugCallback.cpp:5-25) and the call becomes a no-op.
inline bool Ok(int rc, const char* what)
{
    if (rc != 0) std::fprintf(stderr, "cwa: %s failed: %s\n", what, cwa_last_error());
    return rc == 0;
}

}  // namespace cwa
