// Error plumbing + version for the C ABI (include/ssdsb200.h).
#include <stdarg.h>

#include "common.cuh"

namespace ssdsb {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

}  // namespace ssdsb

extern "C" int ssdsb_version(void) { return SSDSB_ABI_VERSION; }

extern "C" int ssdsb_abi_info(int* out4) {
  if (!out4) return ssdsb::fail(SSDSB_ERR_INVALID_ARGUMENT, "abi_info: NULL argument");
  out4[0] = SSDSB_ABI_VERSION;
  out4[1] = (int)sizeof(ssdsb_conv_desc);
  out4[2] = (int)sizeof(ssdsb_level);
  out4[3] = SSDSB_MAX_LEVELS;
  return SSDSB_OK;
}

extern "C" const char* ssdsb_last_error_string(void) { return ssdsb::last_error().c_str(); }
