// ORACLE -- TEST INFRASTRUCTURE ONLY.  operator new / delete for oracle/_ref/libpilotguru_ref.so (linked -Bsymbolic, so
// only code inside that library uses them): a bump allocator over one lazily committed virtual reservation that never
// reuses memory.  Why: ORBextractor::DistributeOctTree orders nodes of equal size by their HEAP ADDRESS
// (ORBextractor.cc:684 sorts pair<int, ExtractorNode*>), so with a general-purpose malloc the reference's own output
// changes from run to run with the allocator's reuse pattern (observed here: 505 / 509 keypoints on the same frame).
// With monotonically increasing addresses that tie-break becomes "the node created later first", which is the rule
// the oracle and the CUDA kernel document -- and the comparison becomes deterministic.
#include <sys/mman.h>

#include <cstddef>
#include <cstdlib>
#include <new>

namespace {
constexpr size_t kArenaBytes = (size_t)24 << 30;  // virtual; pages are committed when touched
char* g_base = nullptr;
size_t g_off = 0;

void* bump(size_t n) {
  if (!g_base) {
    void* p = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) std::abort();
    g_base = static_cast<char*>(p);
  }
  n = (n + 15) & ~(size_t)15;
  if (g_off + n > kArenaBytes) std::abort();
  void* r = g_base + g_off;
  g_off += n;
  return r;
}
}  // namespace

void* operator new(size_t n) { return bump(n ? n : 1); }
void* operator new[](size_t n) { return bump(n ? n : 1); }
void operator delete(void*) noexcept {}
void operator delete[](void*) noexcept {}
void operator delete(void*, size_t) noexcept {}
void operator delete[](void*, size_t) noexcept {}

extern "C" size_t pgr_arena_bytes_used(void) { return g_off; }
