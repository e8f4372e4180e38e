// cwa/ParallelScan.h -- mirror of ParallelScan (SphWave2D/ParallelScan.h, ParallelScan.cpp:11-121).
// Compute() is ONE single-pass decoupled-look-back kernel instead of 2*log2(n) dispatches, and n
// no longer has to be a power of two (the reference asserts it, ParallelScan.cpp:15-16).
#pragma once

#include <vector>

#include "Common.h"

class ParallelScan {
public:
    unsigned mInputSsbo = unsigned(-1);
    unsigned mOutputSsbo = unsigned(-1);
    int mNumElements = 0;
    bool mEnableReadback = false;

    void Init(unsigned input, int n, bool readback = false)
    {
        mEnableReadback = readback; mNumElements = n; mInputSsbo = input;
        if (mOutputSsbo != unsigned(-1)) cwa_buffer_destroy(cwa::Ctx(), (cwa_buf)mOutputSsbo);
        cwa_buf out = -1;
        if (cwa::Ok(cwa_buffer_create(cwa::Ctx(), sizeof(int) * (size_t)(n + 1), nullptr, &out), "ParallelScan::Init")) mOutputSsbo = (unsigned)out;
    }
    void Compute() { cwa::Ok(cwa_scan_exclusive(cwa::Ctx(), (cwa_buf)mInputSsbo, (cwa_buf)mOutputSsbo, mNumElements), "ParallelScan::Compute"); }
    std::vector<int> Readback()
    {
        std::vector<int> r;
        if (mEnableReadback) { r.resize(mNumElements); cwa::Ok(cwa_buffer_read(cwa::Ctx(), (cwa_buf)mOutputSsbo, 0, sizeof(int) * r.size(), r.data()), "ParallelScan::Readback"); }
        return r;
    }
};

// ParallelScanTest (ParallelScan.cpp:124-155): the reference's known-answer test, returns true on success
inline bool ParallelScanTest()
{
    std::vector<int> x = {1, 0, 1, 0, 1, 2, 1, 2, 1, 2, 0, 1, 0, 2, 1, 0};
    std::vector<int> truth(x.size());
    int sum = 0;
    for (size_t i = 0; i < x.size(); i++) { truth[i] = sum; sum += x[i]; }
    cwa_buf in = -1;
    if (!cwa::Ok(cwa_buffer_create(cwa::Ctx(), sizeof(int) * x.size(), x.data(), &in), "ParallelScanTest")) return false;
    ParallelScan scan;
    scan.Init((unsigned)in, (int)x.size(), true);
    scan.Compute();
    const bool ok = scan.Readback() == truth;
    if (!ok) std::printf("Failed\n");
    return ok;
}
