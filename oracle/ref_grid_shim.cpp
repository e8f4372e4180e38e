// ref_grid_shim.cpp -- C entry points over the REFERENCE's own CPU uniform grid
// (UniformGrid2D::Build, /root/reference/UniformGrid2D/UniformGrid2D.cpp:33-92), compiled from
// the reference sources where they lie (see oracle/Makefile, target _ref/libref_grid.so).
// TEST INFRASTRUCTURE ONLY: validates oracle/cwa_oracle.c's grid restatement.
//
// The reference translation unit also holds interactive GL/ImGui demo classes; their GL and
// ImGui symbols are satisfied by the abort()-ing stubs below and are never called.
#include "UniformGrid2D.h"
#include "imgui.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

extern "C" {

// Build the reference grid over n point particles (degenerate boxes mMin == mMax == pos), which
// is the single-cell insert the SPH grid shader performs (uniform_grid_sph_cs.glsl:120-124).
// Outputs: counter[C], offset[C], index_list[n] (as left by Build), cell[2] = mCellSize.
int ref_grid2d_build(const float* xy, int n, const float mn[2], const float mx[2],
                     const int ncells[2], int* counter, int* offset, int* index_list,
                     float* cell_size)
{
    UniformGrid2D grid(glm::ivec2(ncells[0], ncells[1]),
                       aabb2D(glm::vec2(mn[0], mn[1]), glm::vec2(mx[0], mx[1])));
    std::vector<aabb2D> boxes(n);
    for (int i = 0; i < n; i++) {
        glm::vec2 p(xy[2 * i], xy[2 * i + 1]);
        boxes[i] = aabb2D(p, p);
    }
    grid.Build(boxes);
    const int C = ncells[0] * ncells[1];
    std::memcpy(counter, grid.mGridCounter.data(), sizeof(int) * C);
    std::memcpy(offset, grid.mGridOffset.data(), sizeof(int) * C);
    if ((int)grid.mIndexList.size() != n) return -1;
    std::memcpy(index_list, grid.mIndexList.data(), sizeof(int) * n);
    cell_size[0] = grid.mCellSize.x; cell_size[1] = grid.mCellSize.y;
    return 0;
}

// ComputeCellIndex of the reference (UniformGrid2D.cpp:25-31)
void ref_grid2d_cell(const float mn[2], const float mx[2], const int ncells[2], float x, float y, int* out)
{
    UniformGrid2D grid(glm::ivec2(ncells[0], ncells[1]),
                       aabb2D(glm::vec2(mn[0], mn[1]), glm::vec2(mx[0], mx[1])));
    glm::ivec2 c = grid.ComputeCellIndex(glm::vec2(x, y));
    out[0] = c.x; out[1] = c.y;
}

// Query(aabb) with the reference's "home cell" de-dup rule (UniformGrid2D.cpp:141-172): returns
// the number of boxes found, indices into out (capacity cap).
int ref_grid2d_query(const float* xy, int n, const float mn[2], const float mx[2], const int ncells[2],
                     const float qmin[2], const float qmax[2], int* out, int cap)
{
    UniformGrid2D grid(glm::ivec2(ncells[0], ncells[1]),
                       aabb2D(glm::vec2(mn[0], mn[1]), glm::vec2(mx[0], mx[1])));
    std::vector<aabb2D> boxes(n);
    for (int i = 0; i < n; i++) { glm::vec2 p(xy[2 * i], xy[2 * i + 1]); boxes[i] = aabb2D(p, p); }
    grid.Build(boxes);
    std::vector<int> r = grid.Query(aabb2D(glm::vec2(qmin[0], qmin[1]), glm::vec2(qmax[0], qmax[1])));
    int m = (int)r.size() < cap ? (int)r.size() : cap;
    for (int i = 0; i < m; i++) out[i] = r[i];
    return (int)r.size();
}

} // extern "C"

// ---------------------------------------------------------------------------------------------
// Link-time stubs for the GL / GLEW / ImGui / shader-loader symbols referenced by the demo
// classes that share UniformGrid2D.cpp with the grid.  None is reachable from the entry points
// above; each aborts loudly if it ever is.
// ---------------------------------------------------------------------------------------------
static void ref_stub_abort(const char* what)
{
    std::fprintf(stderr, "oracle/_ref: GL stub '%s' called -- not available headless\n", what);
    std::abort();
}
#define REF_GLEW_STUB(type, name) type name = nullptr;
extern "C" {
REF_GLEW_STUB(PFNGLBINDBUFFERPROC, __glewBindBuffer)
REF_GLEW_STUB(PFNGLBINDBUFFERBASEPROC, __glewBindBufferBase)
REF_GLEW_STUB(PFNGLBINDVERTEXARRAYPROC, __glewBindVertexArray)
REF_GLEW_STUB(PFNGLDELETEBUFFERSPROC, __glewDeleteBuffers)
REF_GLEW_STUB(PFNGLDELETEPROGRAMPROC, __glewDeleteProgram)
REF_GLEW_STUB(PFNGLDRAWARRAYSINSTANCEDPROC, __glewDrawArraysInstanced)
REF_GLEW_STUB(PFNGLGENBUFFERSPROC, __glewGenBuffers)
REF_GLEW_STUB(PFNGLGENVERTEXARRAYSPROC, __glewGenVertexArrays)
REF_GLEW_STUB(PFNGLNAMEDBUFFERSTORAGEPROC, __glewNamedBufferStorage)
REF_GLEW_STUB(PFNGLNAMEDBUFFERSUBDATAPROC, __glewNamedBufferSubData)
REF_GLEW_STUB(PFNGLPROGRAMUNIFORM4FVPROC, __glewProgramUniform4fv)
REF_GLEW_STUB(PFNGLUSEPROGRAMPROC, __glewUseProgram)
void glDisable(GLenum) { ref_stub_abort("glDisable"); }
void glEnable(GLenum) { ref_stub_abort("glEnable"); }
}
GLuint InitShader(const char*, const char*) { ref_stub_abort("InitShader"); return 0; }
namespace ImGui {
bool SliderFloat(const char*, float*, float, float, const char*, ImGuiSliderFlags) { ref_stub_abort("ImGui"); return false; }
bool SliderFloat2(const char*, float*, float, float, const char*, ImGuiSliderFlags) { ref_stub_abort("ImGui"); return false; }
bool SliderInt(const char*, int*, int, int, const char*, ImGuiSliderFlags) { ref_stub_abort("ImGui"); return false; }
bool Button(const char*, const ImVec2&) { ref_stub_abort("ImGui"); return false; }
bool Begin(const char*, bool*, ImGuiWindowFlags) { ref_stub_abort("ImGui"); return false; }
void End() { ref_stub_abort("ImGui"); }
}
