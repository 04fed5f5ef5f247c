// Bump-allocating operator new for the oracle/_ref builds of reference sources (include in exactly one translation unit of a
// library linked with -Bsymbolic, so that the replacement stays private to it).  TEST INFRASTRUCTURE ONLY.
//
// Why: DistributeOctTree sorts (size, ExtractorNode*) pairs (ORBextractor.cc:684), so among nodes with equally many keypoints the
// expansion order depends on where malloc put the list nodes.  The oracle and the CUDA path define that order as CREATION
// order; under this allocator a later allocation always has the larger address, so the reference's code follows the same
// definition.  Handles retain / release the arena; it rewinds when the last one is gone.
//
// -DORBREF_SYSTEM_ALLOCATOR (the drop-in build, liborbmatcher_adapter.so): no replacement at all.  The adapters keep handles in
// static tables that must outlive a rewind, and nothing in that build depends on pointer order.
#pragma once
#ifdef ORBREF_SYSTEM_ALLOCATOR
inline void orbref_arena_retain() {}
inline void orbref_arena_release() {}
#else
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <sys/mman.h>

namespace {
const size_t kArenaBytes = size_t(4) << 30;        // virtual reservation; pages are touched only as they are used
char *g_arena = nullptr;
std::atomic<size_t> g_used(0);
int g_live = 0;
void *arena_alloc(size_t n) {
    if (!g_arena) {
        void *p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { std::fprintf(stderr, "orbref: cannot reserve the arena\n"); std::abort(); }
        g_arena = static_cast<char *>(p);
    }
    n = (n + 15) & ~size_t(15);
    const size_t at = g_used.fetch_add(n);
    if (at + n > kArenaBytes) { std::fprintf(stderr, "orbref: arena exhausted\n"); std::abort(); }
    return g_arena + at;
}
bool in_arena(void *p) { return g_arena && p >= g_arena && p < g_arena + kArenaBytes; }
}  // namespace
inline void orbref_arena_retain() { g_live++; }
inline void orbref_arena_release() {
    if (--g_live == 0 && g_arena) {                // nothing allocated through operator new is alive any more
        madvise(g_arena, (g_used.load() + 4095) & ~size_t(4095), MADV_DONTNEED);
        g_used = 0;
    }
}
// hidden: a library loaded as a dependency of this one (liborbx.so in the drop-in build) must keep libstdc++'s operators
#define ORBREF_LOCAL __attribute__((visibility("hidden")))
ORBREF_LOCAL void *operator new(size_t n) { return arena_alloc(n ? n : 1); }
ORBREF_LOCAL void *operator new[](size_t n) { return arena_alloc(n ? n : 1); }
ORBREF_LOCAL void operator delete(void *p) noexcept { if (p && !in_arena(p)) std::free(p); }
ORBREF_LOCAL void operator delete[](void *p) noexcept { if (p && !in_arena(p)) std::free(p); }
ORBREF_LOCAL void operator delete(void *p, size_t) noexcept { if (p && !in_arena(p)) std::free(p); }
ORBREF_LOCAL void operator delete[](void *p, size_t) noexcept { if (p && !in_arena(p)) std::free(p); }
#endif
