// Step tracer ring buffer: lock-free-ish fixed-capacity event store for per-op timing records.
// TF's tracer is C++ (StepStatsCollector); the Python StepTracer uses this when the runtime is built so that
// tracing a hot Session.run adds ~100 ns per node instead of a Python dict allocation.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
struct Event {
  int64_t node_id;
  int64_t start_ns;
  int64_t end_ns;
  int64_t bytes;
  int32_t thread;
  int32_t pad;
};
struct Tracer {
  std::vector<Event> ring;
  std::atomic<int64_t> head{0};
};
}  // namespace

extern "C" {
void* dtf_tracer_create(int64_t capacity) {
  auto* t = new Tracer();
  t->ring.resize((size_t)(capacity > 0 ? capacity : 4096));
  return t;
}
void dtf_tracer_destroy(void* h) { delete static_cast<Tracer*>(h); }
int64_t dtf_tracer_now_ns() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void dtf_tracer_record(void* h, int64_t node_id, int64_t start_ns, int64_t end_ns, int64_t bytes, int32_t thread) {
  auto* t = static_cast<Tracer*>(h);
  const int64_t i = t->head.fetch_add(1);
  Event& e = t->ring[(size_t)(i % (int64_t)t->ring.size())];
  e.node_id = node_id; e.start_ns = start_ns; e.end_ns = end_ns; e.bytes = bytes; e.thread = thread; e.pad = 0;
}
// copies up to `max` most recent events (oldest first) into out[max*6] int64 (thread in slot 4, pad 0); returns count
int64_t dtf_tracer_drain(void* h, int64_t* out, int64_t max) {
  auto* t = static_cast<Tracer*>(h);
  const int64_t head = t->head.load();
  const int64_t cap = (int64_t)t->ring.size();
  int64_t n = head < cap ? head : cap;
  if (n > max) n = max;
  const int64_t first = head - n;
  for (int64_t k = 0; k < n; ++k) {
    const Event& e = t->ring[(size_t)((first + k) % cap)];
    out[k * 6 + 0] = e.node_id; out[k * 6 + 1] = e.start_ns; out[k * 6 + 2] = e.end_ns;
    out[k * 6 + 3] = e.bytes; out[k * 6 + 4] = e.thread; out[k * 6 + 5] = 0;
  }
  t->head.store(0);
  return n;
}
}  // extern "C"
