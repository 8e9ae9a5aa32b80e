// vnl_xla_status.h -- failure reporting of the XLA custom calls (status-returning legacy ABI, API_VERSION_STATUS_RETURNING):
//     void f(cudaStream_t, void** buffers, const char* opaque, size_t opaque_len, XlaCustomCallStatus* status)
// XlaCustomCallStatusSetFailure lives in the process that hosts the call (jaxlib's xla_extension); it is resolved with
// dlsym per failing call, not linked: the library builds and runs without XLA, and keeps no state.
#pragma once
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>

namespace vnl {
inline void xla_report(void* status, const char* what, int rc) {
  if (!rc || !status) return;
  typedef void (*SetFailureFn)(void*, const char*, size_t);
  SetFailureFn fn = reinterpret_cast<SetFailureFn>(dlsym(RTLD_DEFAULT, "XlaCustomCallStatusSetFailure"));
  if (!fn) return;
  char msg[96];
  const int n = snprintf(msg, sizeof(msg), "%s failed with code %d (see include/vnl_b200.h)", what, rc);
  fn(status, msg, (size_t)(n < 0 ? 0 : n));
}
}  // namespace vnl
