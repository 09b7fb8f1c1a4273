#include "common.cuh"

#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
static thread_local double g_next_work = 0.0;
static thread_local char g_next_name[96] = "";

int cvb_fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int cvb_fail_cuda(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return CV_ERR_CUDA;
}

void cvb_reset_launches() { g_launches = 0; }
void cvb_count_launch() { g_launches++; }
void cvb_next_work(double w) { g_next_work = w; }
void cvb_next_name(const char* name) { snprintf(g_next_name, sizeof(g_next_name), "%s", name); }
double cvb_take_work() {
  double w = g_next_work;
  g_next_work = 0.0;
  return w;
}

// ---------------------------------------------------------------- per-kernel event timing
namespace {
struct Span {
  std::string name;
  cudaEvent_t e0, e1;
  double work;
};
struct Agg {
  long long count = 0;
  double ms = 0.0, work = 0.0;
};
std::mutex g_mu;
bool g_prof = false;
std::vector<Span> g_spans;
std::vector<cudaEvent_t> g_free_events;
std::map<std::string, Agg> g_agg;
std::vector<std::string> g_order;
thread_local Span g_open;

cudaEvent_t take_event() {
  if (!g_free_events.empty()) {
    cudaEvent_t e = g_free_events.back();
    g_free_events.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool cvb_profile_on() { return g_prof; }

void cvb_profile_begin(const char* name, cudaStream_t st, double work) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_open.name = g_next_name[0] ? g_next_name : name;
  g_next_name[0] = 0;
  g_open.work = work;
  g_open.e0 = take_event();
  g_open.e1 = take_event();
  cudaEventRecord(g_open.e0, st);
}

void cvb_profile_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_open.e1, st);
  g_spans.push_back(g_open);
}

extern "C" int cv_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_prof = on != 0;
  return CV_OK;
}

// Folds all finished spans into the per-kernel table (synchronises on their end events).
static void fold_spans() {
  for (auto& s : g_spans) {
    cudaEventSynchronize(s.e1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.e0, s.e1) == cudaSuccess) {
      // strip template brackets / parentheses the macro stringified
      std::string n = s.name;
      while (!n.empty() && (n.front() == '(' || n.front() == ' ')) n.erase(n.begin());
      while (!n.empty() && (n.back() == ')' || n.back() == ' ')) n.pop_back();
      auto it = g_agg.find(n);
      if (it == g_agg.end()) {
        g_order.push_back(n);
        it = g_agg.emplace(n, Agg()).first;
      }
      it->second.count++;
      it->second.ms += ms;
      it->second.work += s.work;
    }
    g_free_events.push_back(s.e0);
    g_free_events.push_back(s.e1);
  }
  g_spans.clear();
}

extern "C" int cv_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  fold_spans();
  g_agg.clear();
  g_order.clear();
  return CV_OK;
}

extern "C" int cv_profile_count(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  fold_spans();
  return (int)g_order.size();
}

extern "C" int cv_profile_get(int i, char* name, int name_cap, long long* launches, double* total_ms, double* work) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (i < 0 || i >= (int)g_order.size() || !name || name_cap <= 0) return cvb_fail(CV_ERR_INVALID, "cv_profile_get: bad index");
  const std::string& n = g_order[i];
  snprintf(name, name_cap, "%s", n.c_str());
  const Agg& a = g_agg[n];
  if (launches) *launches = a.count;
  if (total_ms) *total_ms = a.ms;
  if (work) *work = a.work;
  return CV_OK;
}

extern "C" const char* cv_last_error(void) { return g_err; }
extern "C" const char* cv_version(void) { return "circuitvision_b200 0.1.0 sm_100a"; }
extern "C" int cv_last_launch_count(void) { return g_launches; }

extern "C" int cv_device_is_sm100(int device) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}
