// Native parameter-server state: conditional gradient accumulator + FIFO token queue.
// TF keeps these in C++ (ConditionalAccumulator, FIFOQueue); so does this framework.  Plain C ABI for ctypes.
//   * accumulator: apply_grad(g, local_step) drops stale gradients (local_step < global_step), otherwise adds
//     into a double-precision-free fp32 running sum; take_grad(n) blocks (condition variable, cancellable,
//     optional deadline) until >= n fresh gradients arrived, then returns their MEAN and resets.
//   * queue: unbounded FIFO of int64 tokens, blocking dequeue with cancel/deadline/close.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <vector>

namespace {

struct Accumulator {
  std::mutex mu;
  std::condition_variable cv;
  std::vector<float> sum;
  int64_t count = 0;
  int64_t global_step = 0;
  int64_t dropped = 0, applied = 0;
  bool closed = false;
};

struct Queue {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<int64_t> q;
  bool closed = false;
};

// return codes for blocking calls
enum { OK = 0, CANCELLED = 1, DEADLINE = 2, CLOSED = 3, BAD = 4 };

template <class Pred>
int wait_until(std::unique_lock<std::mutex>& lk, std::condition_variable& cv, Pred pred, const bool& closed,
               const volatile int32_t* cancel, double timeout_s) {
  const auto start = std::chrono::steady_clock::now();
  while (!pred()) {
    if (closed) return CLOSED;
    if (cancel && *cancel) return CANCELLED;
    if (timeout_s >= 0) {
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
      if (el > timeout_s) return DEADLINE;
    }
    cv.wait_for(lk, std::chrono::milliseconds(20));
  }
  return OK;
}

}  // namespace

extern "C" {

void* dtf_acc_create() { return new Accumulator(); }
void dtf_acc_destroy(void* h) { delete static_cast<Accumulator*>(h); }

// returns 1 when accepted, 0 when dropped as stale, -1 on shape mismatch
int dtf_acc_apply_grad(void* h, const float* g, int64_t n, int64_t local_step) {
  auto* a = static_cast<Accumulator*>(h);
  std::lock_guard<std::mutex> lk(a->mu);
  if (local_step < a->global_step) {
    a->dropped++;
    return 0;
  }
  if (a->count == 0) {
    a->sum.assign(g, g + n);
  } else {
    if ((int64_t)a->sum.size() != n) return -1;
    float* s = a->sum.data();
    for (int64_t i = 0; i < n; ++i) s[i] += g[i];
  }
  a->count++;
  a->applied++;
  a->cv.notify_all();
  return 1;
}

// writes the mean into out[n]; returns OK / CANCELLED / DEADLINE / CLOSED / BAD
int dtf_acc_take_grad(void* h, int64_t num_required, float* out, int64_t n, const volatile int32_t* cancel,
                      double timeout_s) {
  auto* a = static_cast<Accumulator*>(h);
  std::unique_lock<std::mutex> lk(a->mu);
  int rc = wait_until(lk, a->cv, [&] { return a->count >= num_required; }, a->closed, cancel, timeout_s);
  if (rc != OK) return rc;
  if ((int64_t)a->sum.size() != n) return BAD;
  const float inv = 1.0f / (float)a->count;
  for (int64_t i = 0; i < n; ++i) out[i] = a->sum[i] * inv;
  a->sum.clear();
  a->count = 0;
  a->global_step += 1;
  return OK;
}

int64_t dtf_acc_size(void* h) {          // elements of the pending sum (0 when empty): lets the caller size `out`
  auto* a = static_cast<Accumulator*>(h);
  std::lock_guard<std::mutex> lk(a->mu);
  return (int64_t)a->sum.size();
}
int dtf_acc_wait_count(void* h, int64_t num_required, const volatile int32_t* cancel, double timeout_s) {
  auto* a = static_cast<Accumulator*>(h);
  std::unique_lock<std::mutex> lk(a->mu);
  return wait_until(lk, a->cv, [&] { return a->count >= num_required; }, a->closed, cancel, timeout_s);
}
void dtf_acc_set_global_step(void* h, int64_t s) {
  auto* a = static_cast<Accumulator*>(h);
  std::lock_guard<std::mutex> lk(a->mu);
  if (s > a->global_step) a->global_step = s;
}
int64_t dtf_acc_num_accumulated(void* h) {
  auto* a = static_cast<Accumulator*>(h);
  std::lock_guard<std::mutex> lk(a->mu);
  return a->count;
}
int64_t dtf_acc_global_step(void* h) {
  auto* a = static_cast<Accumulator*>(h);
  std::lock_guard<std::mutex> lk(a->mu);
  return a->global_step;
}
int64_t dtf_acc_dropped(void* h) {
  auto* a = static_cast<Accumulator*>(h);
  std::lock_guard<std::mutex> lk(a->mu);
  return a->dropped;
}
void dtf_acc_close(void* h) {
  auto* a = static_cast<Accumulator*>(h);
  std::lock_guard<std::mutex> lk(a->mu);
  a->closed = true;
  a->cv.notify_all();
}

void* dtf_queue_create() { return new Queue(); }
void dtf_queue_destroy(void* h) { delete static_cast<Queue*>(h); }
int dtf_queue_enqueue_many(void* h, int64_t value, int64_t count) {
  auto* q = static_cast<Queue*>(h);
  std::lock_guard<std::mutex> lk(q->mu);
  if (q->closed) return CLOSED;
  for (int64_t i = 0; i < count; ++i) q->q.push_back(value);
  q->cv.notify_all();
  return OK;
}
int dtf_queue_enqueue_values(void* h, const int64_t* values, int64_t count) {
  auto* q = static_cast<Queue*>(h);
  std::lock_guard<std::mutex> lk(q->mu);
  if (q->closed) return CLOSED;
  for (int64_t i = 0; i < count; ++i) q->q.push_back(values[i]);
  q->cv.notify_all();
  return OK;
}
int dtf_queue_dequeue(void* h, int64_t* out, const volatile int32_t* cancel, double timeout_s) {
  auto* q = static_cast<Queue*>(h);
  std::unique_lock<std::mutex> lk(q->mu);
  // a closed queue still drains what it holds
  const auto start = std::chrono::steady_clock::now();
  while (q->q.empty()) {
    if (q->closed) return CLOSED;
    if (cancel && *cancel) return CANCELLED;
    if (timeout_s >= 0 &&
        std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count() > timeout_s)
      return DEADLINE;
    q->cv.wait_for(lk, std::chrono::milliseconds(20));
  }
  *out = q->q.front();
  q->q.pop_front();
  return OK;
}
int64_t dtf_queue_size(void* h) {
  auto* q = static_cast<Queue*>(h);
  std::lock_guard<std::mutex> lk(q->mu);
  return (int64_t)q->q.size();
}
void dtf_queue_close(void* h) {
  auto* q = static_cast<Queue*>(h);
  std::lock_guard<std::mutex> lk(q->mu);
  q->closed = true;
  q->cv.notify_all();
}

}  // extern "C"
