// Error reporting, version and launch accounting for libribca_b200.so.
#include "common.cuh"

#include <atomic>
#include <map>
#include <mutex>
#include <stdlib.h>
#include <set>
#include <utility>
#include <vector>

namespace ribca {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Opt a kernel in to more than 48 KB of dynamic shared memory, once per (kernel, device): the attribute
// belongs to the device's context, and several host threads may drive different streams.
int ensure_dynamic_smem(const void* func, int bytes, const char* name) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  RIBCA_TRY(check_cuda(cudaGetDevice(&dev), "cudaGetDevice"));
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({func, dev})) return RIBCA_OK;
  RIBCA_TRY(check_cuda(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), name));
  done.insert({func, dev});
  return RIBCA_OK;
}

// ---- fork / join side stream of the two-way interleave (stage4_networks.cu) -----------------------------
// One non-blocking side stream + two events per (host thread, device): a forward pass forks half of its cells onto the
// side stream and joins back before it returns, so the caller's stream ordering is unchanged.
static std::atomic<int> g_interleave{-1};      // -1 = read RIBCA_INTERLEAVE on first use

bool interleave_enabled() {
  int v = g_interleave.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("RIBCA_INTERLEAVE");
    v = (e && e[0] == '1') ? 1 : 0;            // opt-in: no gain under the 1000 W power cap (profiles/r02_interleave.md)
    g_interleave.store(v, std::memory_order_relaxed);
  }
  return v != 0 && !profiling();               // the per-kernel timing leg wants one kernel at a time
}

int side_stream(SideStream* out) {
  static thread_local std::map<int, SideStream> per_device;
  int dev = 0;
  RIBCA_TRY(check_cuda(cudaGetDevice(&dev), "cudaGetDevice"));
  auto it = per_device.find(dev);
  if (it == per_device.end()) {
    SideStream s{};
    RIBCA_TRY(check_cuda(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreateWithFlags"));
    RIBCA_TRY(check_cuda(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming), "cudaEventCreateWithFlags"));
    RIBCA_TRY(check_cuda(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming), "cudaEventCreateWithFlags"));
    it = per_device.emplace(dev, s).first;
  }
  *out = it->second;
  return RIBCA_OK;
}

// ---- optional per-kernel-class timing (bench.py's roofline leg) ----------------------------------
// When enabled, the launch helpers of the dominant kernels bracket each launch with CUDA events on
// the launching stream; ribca_profile_end() synchronises and sums the elapsed times per class.
struct ProfSpan { int cls; cudaEvent_t a, b; double work; };
static std::atomic<bool> g_prof_on{false};
static std::vector<ProfSpan> g_spans;
static std::mutex g_prof_mu;

bool profiling() { return g_prof_on.load(std::memory_order_relaxed); }

// (profiling is a single-stream measurement aid: spans do not nest and are closed in issue order)
void prof_begin_span(int cls, double work, cudaStream_t st) {
  ProfSpan s{cls, nullptr, nullptr, work};
  if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
  cudaEventRecord(s.a, st);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_spans.push_back(s);
}
void prof_end_span(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!g_spans.empty()) cudaEventRecord(g_spans.back().b, st);
}

}  // namespace ribca

extern "C" {

const char* ribca_last_error(void) { return ribca::g_err; }
int ribca_version(void) { return 100; }
long long ribca_launch_count(void) { return ribca::g_launches.load(std::memory_order_relaxed); }

int ribca_set_interleave(int on) {
  ribca::g_interleave.store(on ? 1 : 0, std::memory_order_relaxed);
  return RIBCA_OK;
}

int ribca_profile_begin(void) {
  { std::lock_guard<std::mutex> lock(ribca::g_prof_mu); ribca::g_spans.clear(); }
  ribca::g_prof_on = true;
  return RIBCA_OK;
}

int ribca_profile_end(double* ms, long long* launches, double* work, int n_classes) {
  ribca::g_prof_on = false;
  std::lock_guard<std::mutex> lock(ribca::g_prof_mu);
  for (int c = 0; c < n_classes; ++c) { ms[c] = 0.0; launches[c] = 0; work[c] = 0.0; }
  int rc = RIBCA_OK;
  for (auto& s : ribca::g_spans) {
    float t = 0.f;
    if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) {
      if (s.cls >= 0 && s.cls < n_classes) { ms[s.cls] += t; launches[s.cls] += 1; work[s.cls] += s.work; }
    } else {
      rc = RIBCA_ECUDA;
    }
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  ribca::g_spans.clear();
  if (rc != RIBCA_OK) ribca::set_error("ribca_profile_end: event timing failed");
  return rc;
}

}  // extern "C"
