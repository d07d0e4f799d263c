// unrolled_none.cu — dispatch table used when no per-lens unrolled kernels are built.
#include "unrolled_dispatch.h"
namespace lb {
#ifndef LB_HAVE_UNROLLED
FwLauncher unrolled_fw_launcher(int) { return nullptr; }
BwLauncher unrolled_bw_launcher(int) { return nullptr; }
#endif
}  // namespace lb
