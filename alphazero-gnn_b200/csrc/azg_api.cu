// azg_api.cu -- library-level entry points: error string, ABI version, device check.
#include "azg_common.cuh"

#include <stdarg.h>

static thread_local char g_azg_error[512] = "";

void azg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_azg_error, sizeof(g_azg_error), fmt, ap);
  va_end(ap);
}

unsigned long long g_azg_launches = 0;

// ---- phase timing: a fixed pool of event pairs per phase, recorded only while enabled ----------
namespace {
constexpr int kMaxRecords = 512;
struct PhaseSlots {
  cudaEvent_t beg[kMaxRecords], end[kMaxRecords];
  int used = 0, created = 0;
};
PhaseSlots g_slots[AZG_NUM_PHASES];
bool g_timing = false;
}  // namespace

void azg_phase_begin(int phase, cudaStream_t st) {
  if (!g_timing) return;
  PhaseSlots& s = g_slots[phase];
  if (s.used >= kMaxRecords) return;
  if (s.used >= s.created) {
    if (cudaEventCreate(&s.beg[s.created]) != cudaSuccess || cudaEventCreate(&s.end[s.created]) != cudaSuccess) return;
    ++s.created;
  }
  cudaEventRecord(s.beg[s.used], st);
}

void azg_phase_end(int phase, cudaStream_t st) {
  if (!g_timing) return;
  PhaseSlots& s = g_slots[phase];
  if (s.used >= s.created || s.used >= kMaxRecords) return;
  cudaEventRecord(s.end[s.used], st);
  ++s.used;
}

extern "C" {

unsigned long long azg_launch_count(void) { return g_azg_launches; }

int azg_timing_enable(int on) {
  g_timing = on != 0;
  for (int p = 0; p < AZG_NUM_PHASES; ++p) g_slots[p].used = 0;
  return AZG_OK;
}

int azg_timing_read(int phase, double* total_ms, int* records) {
  AZG_REQUIRE(phase >= 0 && phase < AZG_NUM_PHASES && total_ms && records, "azg_timing_read: bad argument");
  PhaseSlots& s = g_slots[phase];
  double tot = 0.0;
  for (int i = 0; i < s.used; ++i) {
    AZG_CUDA_CHECK(cudaEventSynchronize(s.end[i]));
    float ms = 0.f;
    AZG_CUDA_CHECK(cudaEventElapsedTime(&ms, s.beg[i], s.end[i]));
    tot += ms;
  }
  *total_ms = tot;
  *records = s.used;
  return AZG_OK;
}

const char* azg_last_error(void) { return g_azg_error; }

int azg_abi_version(void) { return AZG_ABI_VERSION; }

int azg_device_info(int* cc_major, int* cc_minor, int* sm_count) {
  int dev = 0;
  AZG_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  AZG_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    azg_set_error("libazgnn_b200 is built for sm_100a only; device %d is compute %d.%d", dev, prop.major, prop.minor);
    return AZG_ERR_DEVICE;
  }
  return AZG_OK;
}

}  // extern "C"
