// gf_knobs.h -- developer / test knobs of the launch paths, WITHOUT environment look-ups per launch.
//
// Every knob is a named non-negative integer in a small process-wide table.  A launch path reads its knobs
// through GF_KNOB("NAME", default): the slot pointer is resolved once per call site (function-local static),
// afterwards a read is one load.  A slot is created on first use and seeded ONCE from the environment variable
// of the same name (so `GF_WS_K=8 python ...` still works for experiments); gf_set_option(name, value) (C ABI,
// include/gf_b200.h) overrides it at any time, value < 0 restores the default.  No knob changes results -- they
// pick kernels, band heights and residency for tests and timing experiments.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

struct GfKnobSlot {
    char name[40];
    std::atomic<int> value;      // < 0: unset
};

inline GfKnobSlot* gf_knob_slot(const char* name)
{
    static GfKnobSlot slots[96];
    static std::atomic<int> count{0};
    static std::mutex mu;
    const int n = count.load(std::memory_order_acquire);
    for (int i = 0; i < n; ++i)
        if (std::strcmp(slots[i].name, name) == 0) return &slots[i];
    std::lock_guard<std::mutex> lock(mu);
    const int m = count.load(std::memory_order_relaxed);
    for (int i = n; i < m; ++i)
        if (std::strcmp(slots[i].name, name) == 0) return &slots[i];
    if (m >= 96) return nullptr;
    std::strncpy(slots[m].name, name, sizeof(slots[m].name) - 1);
    const char* e = std::getenv(name);            // once per knob and process
    slots[m].value.store(e ? std::atoi(e) : -1, std::memory_order_relaxed);
    count.store(m + 1, std::memory_order_release);
    return &slots[m];
}

inline int gf_knob_read(GfKnobSlot* s, int dflt)
{
    if (!s) return dflt;
    const int v = s->value.load(std::memory_order_relaxed);
    return v < 0 ? dflt : v;
}

#define GF_KNOB(NAME, DFLT) ([&]() -> int { static GfKnobSlot* s_ = gf_knob_slot(NAME); return gf_knob_read(s_, (DFLT)); }())
#define GF_KNOB_SET(NAME) (GF_KNOB(NAME, -1) >= 0)
