// cwa/UniformGrid.h -- mirrors of UniformGridSph2D (SphWave2D/UniformGridGpu2D.h:82-142,
// UniformGridGpu2D.cpp:156-265) and Ugrid3D / UgridParticles3D
// (UniformGrid2D/UniformGridParticles3D.h:50-130, .cpp:92-211).
#pragma once

#include "Common.h"

struct aabb2D { cwa::vec2 mMin, mMax; aabb2D() {} aabb2D(cwa::vec2 a, cwa::vec2 b) : mMin(a), mMax(b) {} };
struct aabb3D { cwa::vec4 mMin, mMax; aabb3D() {} aabb3D(cwa::vec3 a, cwa::vec3 b) : mMin(a.x, a.y, a.z, 0), mMax(b.x, b.y, b.z, 0) {} };

class UniformGridSph2D {
public:
    struct UniformGridInfo { aabb2D mExtents; cwa::ivec2 mNumCells; cwa::vec2 mCellSize; } mUniformGridInfo;
    struct Ssbo { unsigned mGridCounter = unsigned(-1), mGridOffset = unsigned(-1), mIndexList = unsigned(-1), mParticles = unsigned(-1); } mSsbo;

    UniformGridSph2D(cwa::ivec2 num_cells, aabb2D extents)
    {
        mUniformGridInfo.mNumCells = num_cells; mUniformGridInfo.mExtents = extents;
        mUniformGridInfo.mCellSize = cwa::vec2((extents.mMax.x - extents.mMin.x) / float(num_cells.x), (extents.mMax.y - extents.mMin.y) / float(num_cells.y));
    }
    void Init(unsigned particles_ssbo, unsigned /*particles_binding*/, int num_particles, int stride_bytes = 48)
    {
        mSsbo.mParticles = particles_ssbo; mNumParticles = num_particles; mStride = stride_bytes;
        const float mn[2] = {mUniformGridInfo.mExtents.mMin.x, mUniformGridInfo.mExtents.mMin.y};
        const float mx[2] = {mUniformGridInfo.mExtents.mMax.x, mUniformGridInfo.mExtents.mMax.y};
        const int nc[2] = {mUniformGridInfo.mNumCells.x, mUniformGridInfo.mNumCells.y};
        if (!cwa::Ok(cwa_grid_create(cwa::Ctx(), 2, mn, mx, nc, num_particles, &mGrid), "UniformGridSph2D::Init")) return;
        cwa_buf b;
        cwa_grid_buffer(cwa::Ctx(), mGrid, CWA_GRID_COUNTER, &b); mSsbo.mGridCounter = (unsigned)b;
        cwa_grid_buffer(cwa::Ctx(), mGrid, CWA_GRID_OFFSET, &b); mSsbo.mGridOffset = (unsigned)b;
        cwa_grid_buffer(cwa::Ctx(), mGrid, CWA_GRID_INDEX_LIST, &b); mSsbo.mIndexList = (unsigned)b;
    }
    // clear + count + scan + insert (+ canonical ordering): the whole body of CollisionQuery() :220-258
    void CollisionQuery() { cwa::Ok(cwa_grid_build(cwa::Ctx(), mGrid, (cwa_buf)mSsbo.mParticles, mStride, mNumParticles), "UniformGridSph2D::CollisionQuery"); }
    void ClearCounter() {}
    void ClearOffset() {}
    cwa_grid Handle() const { return mGrid; }

private:
    cwa_grid mGrid = -1;
    int mNumParticles = 0, mStride = 48;
};

class UgridParticles3D {
public:
    struct UniformGridInfo { aabb3D mExtents; cwa::ivec4 mNumCells; cwa::vec4 mCellSize; } mUniformGridInfo;
    unsigned mGridCounterSsbo = unsigned(-1), mGridOffsetSsbo = unsigned(-1), mIndexListSsbo = unsigned(-1), mParticleSsbo = unsigned(-1);

    UgridParticles3D(cwa::ivec4 num_cells, aabb3D extents)
    {
        mUniformGridInfo.mNumCells = num_cells; mUniformGridInfo.mExtents = extents;
        mUniformGridInfo.mCellSize = cwa::vec4((extents.mMax.x - extents.mMin.x) / float(num_cells.x), (extents.mMax.y - extents.mMin.y) / float(num_cells.y),
                                               (extents.mMax.z - extents.mMin.z) / float(num_cells.z), 0.0f);
    }
    void Init(unsigned particleSsbo, int num_particles, int stride_bytes = 64)
    {
        mParticleSsbo = particleSsbo; mNumParticles = num_particles; mStride = stride_bytes;
        const float mn[3] = {mUniformGridInfo.mExtents.mMin.x, mUniformGridInfo.mExtents.mMin.y, mUniformGridInfo.mExtents.mMin.z};
        const float mx[3] = {mUniformGridInfo.mExtents.mMax.x, mUniformGridInfo.mExtents.mMax.y, mUniformGridInfo.mExtents.mMax.z};
        const int nc[3] = {mUniformGridInfo.mNumCells.x, mUniformGridInfo.mNumCells.y, mUniformGridInfo.mNumCells.z};
        if (!cwa::Ok(cwa_grid_create(cwa::Ctx(), 3, mn, mx, nc, num_particles, &mGrid), "UgridParticles3D::Init")) return;
        cwa_buf b;
        cwa_grid_buffer(cwa::Ctx(), mGrid, CWA_GRID_COUNTER, &b); mGridCounterSsbo = (unsigned)b;
        cwa_grid_buffer(cwa::Ctx(), mGrid, CWA_GRID_OFFSET, &b); mGridOffsetSsbo = (unsigned)b;
        cwa_grid_buffer(cwa::Ctx(), mGrid, CWA_GRID_INDEX_LIST, &b); mIndexListSsbo = (unsigned)b;
    }
    void BuildGrid() { cwa::Ok(cwa_grid_build(cwa::Ctx(), mGrid, (cwa_buf)mParticleSsbo, mStride, mNumParticles), "UgridParticles3D::BuildGrid"); }   // :170-211
    cwa_grid Handle() const { return mGrid; }

private:
    cwa_grid mGrid = -1;
    int mNumParticles = 0, mStride = 64;
};
