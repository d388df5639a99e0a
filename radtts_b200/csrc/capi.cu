// Library-wide pieces of the C ABI (include/radtts_b200.h).
#include "common.cuh"

#include <cstdlib>

namespace rb {
long long g_launches = 0;
int g_gemm_tile_select = [] { const char* e = std::getenv("RADTTS_GEMM_TILE_SELECT"); return (e && e[0] == '0') ? 0 : 1; }();
}

extern "C" int radtts_abi_version(void) { return 1; }
extern "C" long long radtts_launch_count(void) { return rb::g_launches; }
extern "C" int radtts_set_gemm_tile_select(int enabled) {
  const int old = rb::g_gemm_tile_select;
  rb::g_gemm_tile_select = enabled ? 1 : 0;
  return old;
}
extern "C" const char* radtts_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case RADTTS_ERR_INVALID_ARG: return "radtts_b200: invalid argument";
    case RADTTS_ERR_UNSUPPORTED: return "radtts_b200: unsupported shape or mode";
    case RADTTS_ERR_WORKSPACE: return "radtts_b200: workspace too small";
    case RADTTS_ERR_NOT_SM100: return "radtts_b200: device is not sm_100";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "radtts_b200: unknown error";
}
