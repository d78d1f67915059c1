// slb_program_host.h -- host-side hazard analysis of step programs (slb_program.cuh): which recorded ops need a grid
// barrier in front of them.  Pure C++ (no CUDA): used by slb_api.cu and, through lib/libslb200_hosttest.so, by the CPU
// tests (tests/test_host_logic.py).
#pragma once
#include <stddef.h>

#include <vector>

struct ProgRange {
    const char* p;
    size_t bytes;
    bool b0;   // the access is made by ONE block only (accesses of one block are ordered without a grid barrier)
};

struct ProgAccess {
    std::vector<ProgRange> reads, writes;
};

static inline ProgRange prog_range(const void* p, size_t bytes, bool b0 = false) { return ProgRange{(const char*)p, bytes, b0}; }

static inline bool prog_overlap(const ProgRange& a, const ProgRange& b)
{
    return !(a.b0 && b.b0) && a.p < b.p + b.bytes && b.p < a.p + a.bytes;
}

// does `late` have to wait for `early` (issued before it, no barrier in between)?
static inline bool prog_conflict(const ProgAccess& early, const ProgAccess& late)
{
    for (const auto& w : early.writes) {
        for (const auto& x : late.reads)
            if (prog_overlap(w, x)) return true;  // read after write
        for (const auto& x : late.writes)
            if (prog_overlap(w, x)) return true;  // write after write
    }
    for (const auto& rd : early.reads)
        for (const auto& x : late.writes)
            if (prog_overlap(rd, x)) return true;  // write after read
    return false;
}

// barrier_before[k] for a program that REPEATS its op list: walk the list twice and cut wherever an op conflicts with
// one issued since the last cut; the flags of the second walk are valid for every repetition (in the first one fewer
// ops are pending at any point, never more).
static inline std::vector<int> prog_place_barriers(const std::vector<ProgAccess>& ops)
{
    const int n = (int)ops.size();
    std::vector<int> flags((size_t)n, 0), pending;
    for (int pass = 0; pass < 2; ++pass) {
        for (int k = 0; k < n; ++k) {
            bool cut = false;
            for (int j : pending)
                if (prog_conflict(ops[j], ops[k])) cut = true;
            if (pass == 1) flags[k] = cut ? 1 : 0;
            if (cut) pending.clear();
            pending.push_back(k);
        }
    }
    return flags;
}
