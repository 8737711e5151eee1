// prof.cu — launch counter and optional per-kernel CUDA-event timers of liblrb200.so.
//
// Every kernel launch of the library sits in a ProfScope (common.h).  The scope always bumps the launch counter
// (bench.py's gpu_launches is this number, not an estimate); when timing is enabled (lrb_prof_enable) it also records a
// CUDA event before and after the launch ON THE STREAM THE KERNEL IS LAUNCHED ON, so bench.py can quote every single
// kernel's duration inside its timed region without a profiler attached.  Disabled, a scope costs one relaxed atomic add.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.h"

namespace {
std::atomic<uint64_t> g_launches{0};
std::atomic<bool> g_timing{false};
std::mutex g_mu;
struct Rec { const char* name; cudaEvent_t a, b; };
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

ProfScope::ProfScope(const char* name, void* stream) : name_(name), stream_(stream), a_(nullptr) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_timing.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lock(g_mu);
    a_ = get_event();
    if (a_) cudaEventRecord((cudaEvent_t)a_, (cudaStream_t)stream_);
}

ProfScope::~ProfScope() {
    if (!a_) return;
    std::lock_guard<std::mutex> lock(g_mu);
    cudaEvent_t b = get_event();
    if (b) cudaEventRecord(b, (cudaStream_t)stream_);
    g_recs.push_back({name_, (cudaEvent_t)a_, b});
}

extern "C" uint64_t lrb_prof_launches(void) { return g_launches.load(); }

extern "C" int lrb_prof_enable(int on) {
    const bool was = g_timing.exchange(on != 0);
    if (on) {
        std::lock_guard<std::mutex> lock(g_mu);
        for (auto& r : g_recs) { g_pool.push_back(r.a); if (r.b) g_pool.push_back(r.b); }
        g_recs.clear();
    }
    return was ? 1 : 0;
}

// "name launches total_ms\n" per kernel, in order of first launch.  The caller has synchronised the device(s).
extern "C" int lrb_prof_report(char* buf, size_t cap) {
    if (!buf || !cap) return lrb_set_error(LRB_EINVAL, "lrb_prof_report: null buffer");
    std::lock_guard<std::mutex> lock(g_mu);
    std::vector<std::string> order;
    std::map<std::string, std::pair<uint64_t, double>> acc;
    for (auto& r : g_recs) {
        float ms = 0.f;
        if (!r.b || cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
        auto it = acc.find(r.name);
        if (it == acc.end()) { order.push_back(r.name); it = acc.emplace(r.name, std::make_pair(0ull, 0.0)).first; }
        it->second.first += 1;
        it->second.second += ms;
    }
    size_t off = 0;
    buf[0] = 0;
    for (auto& nm : order) {
        char line[160];
        const int n = snprintf(line, sizeof line, "%s %llu %.6f\n", nm.c_str(), (unsigned long long)acc[nm].first, acc[nm].second);
        if (off + (size_t)n + 1 > cap) return lrb_set_error(LRB_ENOMEM, "lrb_prof_report: buffer too small");
        memcpy(buf + off, line, (size_t)n + 1);
        off += (size_t)n;
    }
    return LRB_OK;
}
