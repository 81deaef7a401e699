// Library-level entry points: version, last-error string, and the per-stream launch settings.
#include <stdio.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

static thread_local char g_last_error[512] = "";

void mmemo_set_error(const char* what, const char* file, int line) {
  snprintf(g_last_error, sizeof(g_last_error), "%s (%s:%d)", what ? what : "?", file, line);
}

// Launch settings are attributes of the caller's STREAM, never of the process: every entry point
// takes the stream it enqueues on, looks its settings up here (a copy is taken under the lock), and
// two host threads driving two streams share nothing.  Unregistered streams get the defaults.
namespace {
std::mutex g_cfg_mu;
std::unordered_map<cudaStream_t, MmStreamCfg>& cfg_table() {
  static std::unordered_map<cudaStream_t, MmStreamCfg> t;
  return t;
}
template <typename F>
int cfg_update(mmemo_stream_t s, F f) {
  std::lock_guard<std::mutex> lk(g_cfg_mu);
  f(cfg_table()[mm_stream(s)]);      // value-initialised to the defaults on first use
  return MMEMO_OK;
}
}  // namespace

MmStreamCfg mm_stream_cfg(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_cfg_mu);
  auto& t = cfg_table();
  auto it = t.find(st);
  return it == t.end() ? MmStreamCfg{} : it->second;
}

extern "C" {
int mmemo_version(void) { return 101; }
int mmemo_stream_set_workspace(mmemo_stream_t stream, void* ptr, int64_t bytes) {
  if (bytes < 0) return MMEMO_ERR_ARG;
  return cfg_update(stream, [&](MmStreamCfg& c) {
    c.ws = static_cast<float*>(ptr);
    c.ws_bytes = ptr ? (size_t)bytes : 0;
  });
}
int mmemo_stream_set_sm_budget(mmemo_stream_t stream, int n_sms) {
  if (n_sms < 0) return MMEMO_ERR_ARG;
  return cfg_update(stream, [&](MmStreamCfg& c) { c.sm_budget = n_sms; });
}
int mmemo_stream_set_pdl(mmemo_stream_t stream, int enabled) {
  return cfg_update(stream, [&](MmStreamCfg& c) { c.pdl = enabled ? 1 : 0; });
}
int mmemo_stream_reset(mmemo_stream_t stream) {
  std::lock_guard<std::mutex> lk(g_cfg_mu);
  cfg_table().erase(mm_stream(stream));
  return MMEMO_OK;
}
const char* mmemo_last_error(void) { return g_last_error; }
}
