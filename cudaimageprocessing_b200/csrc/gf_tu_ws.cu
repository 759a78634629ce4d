// gf_tu_ws.cu -- translation unit of the warp-specialised gray kernels (gf_ws.cuh)
#define GF_S8_NO_TRY
#define GF_WP_NO_TRY
#define GF_FAST_NO_TRY
#include "gf_ws.cuh"

const char* gf_ws_try_x(const Job& j, bool* done, const char** name) { return gf_ws_try(j, done, name); }
