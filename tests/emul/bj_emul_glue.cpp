// bj_emul_glue.cpp -- TEST INFRASTRUCTURE ONLY: what bj_factor.cu / bj_solve.cu expect from ctx.cu, for the CPU emulation.
#define PCU_EMUL 1
#include <stdarg.h>

#include "../../prealps_b200/csrc/common.cuh"

namespace pcu {
static thread_local char g_err[1024] = "no error";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
}  // namespace pcu

extern "C" {
const char* pcu_last_error(void) { return pcu::g_err; }
pcu_ctx* emul_ctx_create(void) { return new pcu_ctx(); }
long long emul_launch_count(pcu_ctx* c) { return c->launches; }
long long emul_graph_replay_count(void) { return emul_graph_replays; }
}
