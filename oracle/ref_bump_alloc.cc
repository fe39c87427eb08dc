// ORACLE -- TEST INFRASTRUCTURE ONLY.  operator new / delete for oracle/_ref/libpilotguru_ref.so (linked -Bsymbolic, so
// only code inside that library uses them): while the reference's extractor runs, a bump allocator over one lazily
// committed virtual reservation that never reuses memory.  Why: ORBextractor::DistributeOctTree orders nodes of equal size by their HEAP ADDRESS
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
size_t g_off = 0, g_cap = 0;
bool g_monotonic = false;

void* bump(size_t n) {
  if (!g_base) {
    void* p = MAP_FAILED;
    for (g_cap = kArenaBytes; g_cap >= ((size_t)1 << 30); g_cap >>= 1) {   // smaller reservations where address space is capped
      p = mmap(nullptr, g_cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (p != MAP_FAILED) break;
    }
    if (p == MAP_FAILED) std::abort();
    g_base = static_cast<char*>(p);
  }
  n = (n + 15) & ~(size_t)15;
  if (g_off + n > g_cap) std::abort();
  void* r = g_base + g_off;
  g_off += n;
  return r;
}
}  // namespace

// Monotonic mode is switched on by pgr_orb_extract() for the duration of one ORBextractor::operator() call; everything
// else in the library (the calibration objective allocates vectors in every evaluation) uses malloc / free as usual.
// Arena blocks are recognised by address on delete and never recycled.
extern "C" void pgr_set_monotonic_alloc(int on) { g_monotonic = on != 0; }

static void* alloc_any(size_t n) {
  if (g_monotonic) return bump(n ? n : 1);
  void* p = std::malloc(n ? n : 1);
  if (!p) std::abort();
  return p;
}
static void free_any(void* p) noexcept {
  if (!p) return;
  if (g_base && static_cast<char*>(p) >= g_base && static_cast<char*>(p) < g_base + g_cap) return;
  std::free(p);
}
void* operator new(size_t n) { return alloc_any(n); }
void* operator new[](size_t n) { return alloc_any(n); }
void operator delete(void* p) noexcept { free_any(p); }
void operator delete[](void* p) noexcept { free_any(p); }
void operator delete(void* p, size_t) noexcept { free_any(p); }
void operator delete[](void* p, size_t) noexcept { free_any(p); }

extern "C" size_t pgr_arena_bytes_used(void) { return g_off; }
